/* ref_conv.c -- the reference's convolution/conv.c (the precision dispatch) compiled in place.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#include "pconvolution.h"
int vips_convf(VipsImage *in, VipsImage **out, VipsImage *mask, ...);
int vips_convi(VipsImage *in, VipsImage **out, VipsImage *mask, ...);
int vips_conva(VipsImage *in, VipsImage **out, VipsImage *mask, ...); /* approximate: not built, aborts */
#define g_object_set(OBJ, NAME, VAL, END) (((VipsConvolution *) (OBJ))->out = (VAL))
#define vips_conv vips_conv__via_call_split
#include "conv.c"
#undef vips_conv

int
vips_conv(VipsImage *in, VipsImage **out, VipsImage *mask, ...)
{
	VipsConv *conv = (VipsConv *) vips__shim_object_new(vips_conv_get_type());
	VipsConvolution *convolution = (VipsConvolution *) conv;
	va_list ap;
	const char *name;

	conv->precision = VIPS_PRECISION_FLOAT; /* conv.c class default */
	conv->layers = 5;
	conv->cluster = 1;
	va_start(ap, mask);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "precision") == 0)
			conv->precision = (VipsPrecision) va_arg(ap, int);
		else if (strcmp(name, "layers") == 0)
			conv->layers = va_arg(ap, int);
		else if (strcmp(name, "cluster") == 0)
			conv->cluster = va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	convolution->in = in;
	convolution->mask = mask;
	if (vips_check_matrix("conv", mask, &convolution->M)) /* vips_convolution_build, convolution.c:99-102 */
		return -1;
	if (vips_conv_build((VipsObject *) conv))
		return -1;
	*out = convolution->out;
	return 0;
}
