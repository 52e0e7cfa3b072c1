/* ref_resize.c -- the reference's resample/resize.c compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
typedef struct _VipsInterpolate VipsInterpolate;
VipsInterpolate *vips_interpolate_new(const char *nickname);
int vips_subsample(VipsImage *in, VipsImage **out, int xfac, int yfac, ...);
int vips_zoom(VipsImage *in, VipsImage **out, int xfac, int yfac, ...);
int vips_affine(VipsImage *in, VipsImage **out, double a, double b, double c, double d, ...);
int vips_reducev(VipsImage *in, VipsImage **out, double vshrink, ...);
int vips_reduceh(VipsImage *in, VipsImage **out, double hshrink, ...);
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_resize vips_resize__via_call_split
#include "resize.c"
#undef vips_resize

int
vips_resize(VipsImage *in, VipsImage **out, double scale, ...)
{
	static const char *set_vscale[] = { "vscale", NULL };
	VipsResize *resize = (VipsResize *) vips__shim_object_new(vips_resize_get_type());
	VipsResample *resample = (VipsResample *) resize;
	va_list ap;
	const char *name;

	resize->scale = scale;
	resize->kernel = VIPS_KERNEL_LANCZOS3;
	resize->gap = 2.0; /* resize.c:352-357 */
	va_start(ap, scale);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "vscale") == 0) {
			resize->vscale = va_arg(ap, double);
			((VipsObject *) resize)->set_args = set_vscale;
		}
		else if (strcmp(name, "kernel") == 0)
			resize->kernel = (VipsKernel) va_arg(ap, int);
		else if (strcmp(name, "gap") == 0)
			resize->gap = va_arg(ap, double);
		else
			return -1;
	}
	va_end(ap);
	resample->in = in;
	resample->out = vips_image_new();
	if (vips_resize_build((VipsObject *) resize))
		return -1;
	*out = resample->out;
	return 0;
}
