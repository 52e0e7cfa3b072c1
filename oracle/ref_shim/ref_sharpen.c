/* ref_sharpen.c -- the reference's convolution/sharpen.c compiled in place: its build() (sigma ->
 * integer gaussmat, the 65536-entry LUT, L-band extraction, convsep) and vips_sharpen_generate run as
 * written.  The graph ops around it that are not on the hot path (vips_colourspace, vips_cast_short,
 * vips_extract_band, vips_bandjoin2) are eager stand-ins in shim_ops.c; vips_colourspace there runs
 * the reference's own colour *_line functions (ref_colour.c).  TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
int vips_colourspace(VipsImage *in, VipsImage **out, VipsInterpretation space, ...);
int vips_gaussmat(VipsImage **out, double sigma, double min_ampl, ...);
int vips_cast_short(VipsImage *in, VipsImage **out, ...);
int vips_extract_band(VipsImage *in, VipsImage **out, int band, ...);
int vips_bandjoin2(VipsImage *in1, VipsImage *in2, VipsImage **out, ...);
int vips_convsep(VipsImage *in, VipsImage **out, VipsImage *mask, ...);
int vips_image_pipeline_array(VipsImage *image, VipsDemandStyle hint, VipsImage **in);
int vips_reorder_prepare_many(VipsImage *image, VipsRegion **regions, VipsRect *r);
/* g_object_set(object, "out", vips_image_new(), NULL) is the only property write */
#define g_object_set(OBJ, NAME, VAL, END) (((VipsSharpen *) (OBJ))->out = (VAL))
#define vips_sharpen vips_sharpen__via_call_split
#include "sharpen.c"
#undef vips_sharpen

int
vips_sharpen(VipsImage *in, VipsImage **out, ...)
{
	static const char *set_sigma[] = { "sigma", NULL };
	VipsSharpen *sharpen = (VipsSharpen *) vips__shim_object_new(vips_sharpen_get_type());
	va_list ap;
	const char *name;

	/* vips_sharpen_init, sharpen.c (class defaults) */
	sharpen->sigma = 0.5;
	sharpen->x1 = 2.0;
	sharpen->y2 = 10.0;
	sharpen->y3 = 20.0;
	sharpen->m1 = 0.0;
	sharpen->m2 = 3.0;
	((VipsObject *) sharpen)->set_args = set_sigma;
	va_start(ap, out);
	while ((name = va_arg(ap, const char *))) {
		double v = va_arg(ap, double);
		if (strcmp(name, "sigma") == 0) sharpen->sigma = v;
		else if (strcmp(name, "x1") == 0) sharpen->x1 = v;
		else if (strcmp(name, "y2") == 0) sharpen->y2 = v;
		else if (strcmp(name, "y3") == 0) sharpen->y3 = v;
		else if (strcmp(name, "m1") == 0) sharpen->m1 = v;
		else if (strcmp(name, "m2") == 0) sharpen->m2 = v;
		else return -1;
	}
	va_end(ap);
	sharpen->in = in;
	if (vips_sharpen_build((VipsObject *) sharpen))
		return -1;
	*out = sharpen->out;
	return 0;
}

void *ref_sharpen(void *in, double sigma, double x1, double y2, double y3, double m1, double m2)
{
	VipsImage *out = NULL;
	return vips_sharpen((VipsImage *) in, &out, "sigma", sigma, "x1", x1, "y2", y2, "y3", y3, "m1", m1, "m2", m2, NULL)
		? NULL : out;
}

/* the LUT the build() made, for a direct comparison (65536 ints) */
int ref_sharpen_lut(double sigma, double x1, double y2, double y3, double m1, double m2, int *lut_out)
{
	/* a 1 x 1 LABS image is enough to run build() */
	static short px[3] = { 100, 0, 0 };
	VipsImage *im = vips__shim_image_from_memory(px, 1, 1, 3, VIPS_FORMAT_SHORT, VIPS_INTERPRETATION_LABS);
	VipsSharpen *sharpen = (VipsSharpen *) vips__shim_object_new(vips_sharpen_get_type());
	static const char *set_sigma[] = { "sigma", NULL };
	sharpen->sigma = sigma; sharpen->x1 = x1; sharpen->y2 = y2; sharpen->y3 = y3; sharpen->m1 = m1; sharpen->m2 = m2;
	((VipsObject *) sharpen)->set_args = set_sigma;
	sharpen->in = im;
	if (vips_sharpen_build((VipsObject *) sharpen))
		return -1;
	memcpy(lut_out, sharpen->lut, 65536 * sizeof(int));
	return 0;
}
