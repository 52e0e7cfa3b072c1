/* ref_convf.c -- the reference's convolution/convf.c compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#include "pconvolution.h"
/* g_object_set(obj, "out", image, NULL) is the only property write on this path */
#define g_object_set(OBJ, NAME, VAL, END) (((VipsConvolution *) (OBJ))->out = (VAL))
#define vips_convf vips_convf__via_call_split
#include "convf.c"
#undef vips_convf

int
vips_convf(VipsImage *in, VipsImage **out, VipsImage *mask, ...)
{
	VipsConvf *convf = (VipsConvf *) vips__shim_object_new(vips_convf_get_type());
	VipsConvolution *convolution = (VipsConvolution *) convf;

	convolution->in = in;
	convolution->mask = mask;
	/* vips_convolution_build, convolution.c:99-102 */
	if (vips_check_matrix("convf", mask, &convolution->M))
		return -1;
	if (vips_convf_build((VipsObject *) convf))
		return -1;
	*out = convolution->out;
	return 0;
}
