/* ref_unpremultiply.c -- the reference's conversion/unpremultiply.c compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_unpremultiply vips_unpremultiply__via_call_split
#include "unpremultiply.c"
#undef vips_unpremultiply

int
vips_unpremultiply(VipsImage *in, VipsImage **out, ...)
{
	static const char *set_max_alpha[] = { "max_alpha", NULL };
	VipsUnpremultiply *unpre = (VipsUnpremultiply *) vips__shim_object_new(vips_unpremultiply_get_type());
	VipsConversion *conversion = (VipsConversion *) unpre;
	va_list ap;
	const char *name;

	unpre->in = in;
	unpre->max_alpha = 255.0;
	unpre->alpha_band = 3;
	unpre->uchar = FALSE;
	va_start(ap, out);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "max_alpha") == 0) {
			unpre->max_alpha = va_arg(ap, double);
			((VipsObject *) unpre)->set_args = set_max_alpha;
		}
		else if (strcmp(name, "uchar") == 0)
			unpre->uchar = va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	conversion->out = vips_image_new();
	if (vips_unpremultiply_build((VipsObject *) unpre))
		return -1;
	*out = conversion->out;
	return 0;
}
