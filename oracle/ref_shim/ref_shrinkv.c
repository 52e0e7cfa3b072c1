/* ref_shrinkv.c -- the reference's resample/shrinkv.c compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_shrinkv vips_shrinkv__via_call_split
#include "shrinkv.c"
#undef vips_shrinkv

int
vips_shrinkv(VipsImage *in, VipsImage **out, int vshrink, ...)
{
	VipsShrinkv *shrink = (VipsShrinkv *) vips__shim_object_new(vips_shrinkv_get_type());
	VipsResample *resample = (VipsResample *) shrink;
	va_list ap;
	const char *name;

	shrink->vshrink = vshrink;
	shrink->ceil = FALSE;
	va_start(ap, vshrink);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "ceil") == 0)
			shrink->ceil = va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	resample->in = in;
	resample->out = vips_image_new();
	if (vips_shrinkv_build((VipsObject *) shrink))
		return -1;
	*out = resample->out;
	return 0;
}
