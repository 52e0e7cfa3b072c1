/* ref_reducev.cpp -- compiles the reference's resample/reducev.cpp IN PLACE
 * (it is #included from $(REF), not copied) and gives its class a C
 * constructor, standing in for vips_call_split + the GObject property system.
 * TEST INFRASTRUCTURE ONLY.
 */
#include <cstdarg>
#include <cstring>
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_reducev vips_reducev__via_call_split
#include "reducev.cpp"
#undef vips_reducev /* -I$(REF)/libvips/resample */

extern "C" int
vips_reducev(VipsImage *in, VipsImage **out, double vshrink, ...)
{
	VipsReducev *reducev = (VipsReducev *) vips__shim_object_new(vips_reducev_get_type());
	VipsResample *resample = (VipsResample *) reducev;
	va_list ap;
	const char *name;

	/* class defaults: reducev.cpp class_init/init */
	reducev->vshrink = vshrink;
	reducev->kernel = VIPS_KERNEL_LANCZOS3;
	reducev->gap = 0.0;
	va_start(ap, vshrink);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "kernel") == 0)
			reducev->kernel = (VipsKernel) va_arg(ap, int);
		else if (strcmp(name, "gap") == 0)
			reducev->gap = va_arg(ap, double);
		else
			return -1;
	}
	va_end(ap);
	resample->in = in;
	resample->out = vips_image_new(); /* resample.c:82 */
	if (vips_reducev_build((VipsObject *) reducev))
		return -1;
	*out = resample->out;
	return 0;
}

extern "C" void
ref_reduce_make_mask(double *c, int kernel, int n_points, double shrink, double x)
{
	vips_reduce_make_mask(c, (VipsKernel) kernel, n_points, shrink, x);
}
