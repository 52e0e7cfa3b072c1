/* ref_affine.c -- the reference's resample/{transform,interpolate,affine}.c compiled in place.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <float.h>
#include <vips/vips.h>
#include "presample.h"
#include "transform.c"
/* vips_interpolate_new() goes through the GType registry by nickname: ours is below */
#define vips_interpolate_new vips_interpolate_new__via_type_find
#include "interpolate.c"
#undef vips_interpolate_new
#define vips_affine vips_affine__via_call_split
#include "affine.c"
#undef vips_affine

GType vips_interpolate_bicubic_get_type(void);

VipsInterpolate *
vips_interpolate_new(const char *nickname)
{
	if (strcmp(nickname, "nearest") == 0)
		return (VipsInterpolate *) vips__shim_object_new(vips_interpolate_nearest_get_type());
	if (strcmp(nickname, "bilinear") == 0)
		return (VipsInterpolate *) vips__shim_object_new(vips_interpolate_bilinear_get_type());
	if (strcmp(nickname, "bicubic") == 0)
		return (VipsInterpolate *) vips__shim_object_new(vips_interpolate_bicubic_get_type());
	return NULL;
}

/* vips_affine(in, &out, a, b, c, d, "interpolate", i, "idx", .., "idy", .., "odx", .., "ody", ..,
 * "extend", VIPS_EXTEND_COPY, "premultiplied", TRUE, NULL): the options vips_resize passes.
 */
int
vips_affine(VipsImage *in, VipsImage **out, double a, double b, double c, double d, ...)
{
	VipsAffine *affine = (VipsAffine *) vips__shim_object_new(vips_affine_get_type());
	VipsResample *resample = (VipsResample *) affine;
	static VipsArea matrix_area, background_area;
	double *matrix = (double *) calloc(4, sizeof(double));
	double *background = (double *) calloc(1, sizeof(double));
	const char **set = (const char **) calloc(16, sizeof(char *));
	int n_set = 0;
	va_list ap;
	const char *name;
	VipsArea *ma = (VipsArea *) calloc(1, sizeof(VipsArea));
	VipsArea *ba = (VipsArea *) calloc(1, sizeof(VipsArea));

	(void) matrix_area;
	(void) background_area;
	matrix[0] = a;
	matrix[1] = b;
	matrix[2] = c;
	matrix[3] = d;
	ma->data = matrix;
	ma->n = 4;
	ba->data = background;
	ba->n = 1;
	affine->matrix = ma;
	affine->background = (VipsArrayDouble *) ba;
	affine->extend = VIPS_EXTEND_BACKGROUND;
	va_start(ap, d);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "interpolate") == 0)
			affine->interpolate = va_arg(ap, VipsInterpolate *);
		else if (strcmp(name, "idx") == 0) {
			affine->idx = va_arg(ap, double);
			set[n_set++] = "idx";
		}
		else if (strcmp(name, "idy") == 0) {
			affine->idy = va_arg(ap, double);
			set[n_set++] = "idy";
		}
		else if (strcmp(name, "odx") == 0) {
			affine->odx = va_arg(ap, double);
			set[n_set++] = "odx";
		}
		else if (strcmp(name, "ody") == 0) {
			affine->ody = va_arg(ap, double);
			set[n_set++] = "ody";
		}
		else if (strcmp(name, "extend") == 0)
			affine->extend = (VipsExtend) va_arg(ap, int);
		else if (strcmp(name, "premultiplied") == 0)
			affine->premultiplied = va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	((VipsObject *) affine)->set_args = set;
	if (affine->extend != VIPS_EXTEND_COPY)
		return -1; /* the shim's embed only knows EXTEND_COPY */
	resample->in = in;
	resample->out = vips_image_new();
	if (vips_affine_build((VipsObject *) affine))
		return -1;
	*out = resample->out;
	return 0;
}

void *
ref_affine(void *in, double a, double b, double c, double d, const char *interpolate, double idx, double idy,
	double odx, double ody)
{
	VipsImage *out = NULL;
	VipsInterpolate *i = vips_interpolate_new(interpolate);
	if (!i)
		return NULL;
	if (vips_affine((VipsImage *) in, &out, a, b, c, d, "interpolate", i, "idx", idx, "idy", idy, "odx", odx, "ody",
			ody, "extend", VIPS_EXTEND_COPY, "premultiplied", TRUE, NULL))
		return NULL;
	return out;
}
