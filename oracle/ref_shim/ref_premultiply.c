/* ref_premultiply.c -- the reference's conversion/premultiply.c compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_premultiply vips_premultiply__via_call_split
#include "premultiply.c"
#undef vips_premultiply

int
vips_premultiply(VipsImage *in, VipsImage **out, ...)
{
	static const char *set_max_alpha[] = { "max_alpha", NULL };
	VipsPremultiply *pre = (VipsPremultiply *) vips__shim_object_new(vips_premultiply_get_type());
	VipsConversion *conversion = (VipsConversion *) pre;
	va_list ap;
	const char *name;

	pre->in = in;
	pre->max_alpha = 255.0;
	pre->uchar = FALSE;
	va_start(ap, out);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "max_alpha") == 0) {
			pre->max_alpha = va_arg(ap, double);
			((VipsObject *) pre)->set_args = set_max_alpha;
		}
		else if (strcmp(name, "uchar") == 0)
			pre->uchar = va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	conversion->out = vips_image_new(); /* conversion.c:313 */
	if (vips_premultiply_build((VipsObject *) pre))
		return -1;
	*out = conversion->out;
	return 0;
}
