/* ref_convsep.c -- the reference's convolution/convsep.c compiled in place: conv(M) then
 * conv(rot90(M), offset 0).  TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#include "pconvolution.h"
enum { VIPS_ANGLE_D0, VIPS_ANGLE_D90, VIPS_ANGLE_D180, VIPS_ANGLE_D270 };
int vips_rot(VipsImage *in, VipsImage **out, int angle, ...);
int vips_conv(VipsImage *in, VipsImage **out, VipsImage *mask, ...);
int vips_convasep(VipsImage *in, VipsImage **out, VipsImage *mask, ...); /* approximate: not built, aborts */
#define g_object_set(OBJ, NAME, VAL, END) (((VipsConvolution *) (OBJ))->out = (VAL))
#define vips_convsep vips_convsep__via_call_split
#include "convsep.c"
#undef vips_convsep

int
vips_convsep(VipsImage *in, VipsImage **out, VipsImage *mask, ...)
{
	VipsConvsep *convsep = (VipsConvsep *) vips__shim_object_new(vips_convsep_get_type());
	VipsConvolution *convolution = (VipsConvolution *) convsep;
	va_list ap;
	const char *name;

	convsep->precision = VIPS_PRECISION_FLOAT; /* convsep.c class default */
	convsep->layers = 5;
	convsep->cluster = 1;
	va_start(ap, mask);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "precision") == 0)
			convsep->precision = (VipsPrecision) va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	convolution->in = in;
	convolution->mask = mask;
	if (vips_check_matrix("convsep", mask, &convolution->M))
		return -1;
	if (vips_convsep_build((VipsObject *) convsep))
		return -1;
	*out = convolution->out;
	return 0;
}
