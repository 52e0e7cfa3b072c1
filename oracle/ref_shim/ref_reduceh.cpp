/* ref_reduceh.cpp -- the reference's resample/reduceh.cpp compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <cstdarg>
#include <cstring>
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_reduceh vips_reduceh__via_call_split
#include "reduceh.cpp"
#undef vips_reduceh

extern "C" int
vips_reduceh(VipsImage *in, VipsImage **out, double hshrink, ...)
{
	VipsReduceh *reduceh = (VipsReduceh *) vips__shim_object_new(vips_reduceh_get_type());
	VipsResample *resample = (VipsResample *) reduceh;
	va_list ap;
	const char *name;

	reduceh->hshrink = hshrink;
	reduceh->kernel = VIPS_KERNEL_LANCZOS3;
	reduceh->gap = 0.0;
	va_start(ap, hshrink);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "kernel") == 0)
			reduceh->kernel = (VipsKernel) va_arg(ap, int);
		else if (strcmp(name, "gap") == 0)
			reduceh->gap = va_arg(ap, double);
		else
			return -1;
	}
	va_end(ap);
	resample->in = in;
	resample->out = vips_image_new();
	if (vips_reduceh_build((VipsObject *) reduceh))
		return -1;
	*out = resample->out;
	return 0;
}
