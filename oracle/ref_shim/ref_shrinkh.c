/* ref_shrinkh.c -- the reference's resample/shrinkh.c compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_shrinkh vips_shrinkh__via_call_split
#include "shrinkh.c"
#undef vips_shrinkh

int
vips_shrinkh(VipsImage *in, VipsImage **out, int hshrink, ...)
{
	VipsShrinkh *shrink = (VipsShrinkh *) vips__shim_object_new(vips_shrinkh_get_type());
	VipsResample *resample = (VipsResample *) shrink;
	va_list ap;
	const char *name;

	shrink->hshrink = hshrink;
	shrink->ceil = FALSE;
	va_start(ap, hshrink);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "ceil") == 0)
			shrink->ceil = va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	resample->in = in;
	resample->out = vips_image_new();
	if (vips_shrinkh_build((VipsObject *) shrink))
		return -1;
	*out = resample->out;
	return 0;
}
