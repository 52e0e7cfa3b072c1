/* ref_gaussmat.c -- the reference's create/gaussmat.c compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#include "pcreate.h"
#define vips_gaussmat vips_gaussmat__via_call_split
#include "gaussmat.c"
#undef vips_gaussmat

int
vips_gaussmat(VipsImage **out, double sigma, double min_ampl, ...)
{
	static const char *set_precision[] = { "precision", NULL };
	VipsGaussmat *g = (VipsGaussmat *) vips__shim_object_new(vips_gaussmat_get_type());
	VipsCreate *create = (VipsCreate *) g;
	va_list ap;
	const char *name;

	g->sigma = sigma;
	g->min_ampl = min_ampl;
	g->separable = FALSE;
	g->precision = VIPS_PRECISION_FLOAT; /* gaussmat.c class default */
	va_start(ap, min_ampl);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "separable") == 0)
			g->separable = va_arg(ap, int);
		else if (strcmp(name, "precision") == 0) {
			g->precision = (VipsPrecision) va_arg(ap, int);
			((VipsObject *) g)->set_args = set_precision;
		}
		else
			return -1;
	}
	va_end(ap);
	create->out = vips_image_new();
	if (vips_gaussmat_build((VipsObject *) g))
		return -1;
	*out = create->out;
	return 0;
}
