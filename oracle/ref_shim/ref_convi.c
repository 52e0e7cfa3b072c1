/* ref_convi.c -- the reference's convolution/convi.c (C path: no HAVE_HWY) compiled in place.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#include "pconvolution.h"
#define g_object_set(OBJ, NAME, VAL, END) (((VipsConvolution *) (OBJ))->out = (VAL))
#define vips_convi vips_convi__via_call_split
#include "convi.c"
#undef vips_convi

int
vips_convi(VipsImage *in, VipsImage **out, VipsImage *mask, ...)
{
	VipsConvi *convi = (VipsConvi *) vips__shim_object_new(vips_convi_get_type());
	VipsConvolution *convolution = (VipsConvolution *) convi;

	convolution->in = in;
	convolution->mask = mask;
	if (vips_check_matrix("convi", mask, &convolution->M))
		return -1;
	if (vips_convi_build((VipsObject *) convi))
		return -1;
	*out = convolution->out;
	return 0;
}
