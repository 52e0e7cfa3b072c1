/* ref_api.c -- flat C entry points (for ctypes) over the reference's lazily
 * evaluated pipeline objects.  TEST INFRASTRUCTURE ONLY.
 */
#include <vips/vips.h>

int vips_shrinkv(VipsImage *in, VipsImage **out, int vshrink, ...);
int vips_shrinkh(VipsImage *in, VipsImage **out, int hshrink, ...);
int vips_reducev(VipsImage *in, VipsImage **out, double vshrink, ...);
int vips_reduceh(VipsImage *in, VipsImage **out, double hshrink, ...);
int vips_resize(VipsImage *in, VipsImage **out, double scale, ...);
int vips_premultiply(VipsImage *in, VipsImage **out, ...);
int vips_unpremultiply(VipsImage *in, VipsImage **out, ...);

void *ref_image_new_from_memory(const void *data, int w, int h, int bands, int fmt, int type)
{
	return vips__shim_image_from_memory(data, w, h, bands, (VipsBandFormat) fmt, (VipsInterpretation) type);
}
int ref_image_width(void *im) { return ((VipsImage *) im)->Xsize; }
int ref_image_height(void *im) { return ((VipsImage *) im)->Ysize; }
int ref_image_bands(void *im) { return ((VipsImage *) im)->Bands; }
int ref_image_format(void *im) { return ((VipsImage *) im)->BandFmt; }
int ref_image_dhint(void *im) { return ((VipsImage *) im)->dhint; }
int ref_image_write_to_memory(void *im, void *out, int tile_w, int tile_h)
{
	return vips__shim_write_to_memory((VipsImage *) im, out, tile_w, tile_h);
}
const char *ref_error(void) { return vips__shim_error(); }

void *ref_shrinkv(void *in, int vshrink, int ceil_mode)
{
	VipsImage *out = NULL;
	return vips_shrinkv((VipsImage *) in, &out, vshrink, "ceil", ceil_mode, NULL) ? NULL : out;
}
void *ref_shrinkh(void *in, int hshrink, int ceil_mode)
{
	VipsImage *out = NULL;
	return vips_shrinkh((VipsImage *) in, &out, hshrink, "ceil", ceil_mode, NULL) ? NULL : out;
}
void *ref_reducev(void *in, double vshrink, int kernel, double gap)
{
	VipsImage *out = NULL;
	return vips_reducev((VipsImage *) in, &out, vshrink, "kernel", kernel, "gap", gap, NULL) ? NULL : out;
}
void *ref_reduceh(void *in, double hshrink, int kernel, double gap)
{
	VipsImage *out = NULL;
	return vips_reduceh((VipsImage *) in, &out, hshrink, "kernel", kernel, "gap", gap, NULL) ? NULL : out;
}
void *ref_resize(void *in, double scale, double vscale, int kernel, double gap)
{
	VipsImage *out = NULL;
	return vips_resize((VipsImage *) in, &out, scale, "vscale", vscale, "kernel", kernel, "gap", gap, NULL) ? NULL : out;
}
void *ref_premultiply(void *in, double max_alpha, int uchar_mode)
{
	VipsImage *out = NULL;
	int rc = max_alpha > 0 ? vips_premultiply((VipsImage *) in, &out, "max_alpha", max_alpha, "uchar", uchar_mode, NULL)
						   : vips_premultiply((VipsImage *) in, &out, "uchar", uchar_mode, NULL);
	return rc ? NULL : out;
}
void *ref_unpremultiply(void *in, double max_alpha, int uchar_mode)
{
	VipsImage *out = NULL;
	int rc = max_alpha > 0 ? vips_unpremultiply((VipsImage *) in, &out, "max_alpha", max_alpha, "uchar", uchar_mode, NULL)
						   : vips_unpremultiply((VipsImage *) in, &out, "uchar", uchar_mode, NULL);
	return rc ? NULL : out;
}

/* ---------------------------------------------------------------- convolution */
int vips_convf(VipsImage *in, VipsImage **out, VipsImage *mask, ...);
int vips_convi(VipsImage *in, VipsImage **out, VipsImage *mask, ...);
int ref_convi_vector(VipsImage *in, VipsImage **out, VipsImage *mask);
int vips_gaussmat(VipsImage **out, double sigma, double min_ampl, ...);

void *ref_matrix(const double *coeff, int w, int h, double scale, double offset)
{
	VipsImage *m = vips_image_new_matrix(w, h);
	memcpy(m->data, coeff, sizeof(double) * w * h);
	vips_image_set_double(m, "scale", scale);
	vips_image_set_double(m, "offset", offset);
	return m;
}
double ref_matrix_scale(void *m) { return vips_image_get_scale((VipsImage *) m); }
double ref_matrix_offset(void *m) { return vips_image_get_offset((VipsImage *) m); }
const double *ref_matrix_data(void *m) { return (const double *) ((VipsImage *) m)->data; }

/* vips_conv's dispatch, conv.c:84-112; vector: 0 = C path, 1 = the Highway arithmetic */
void *ref_conv(void *in, void *mask, int precision, int vector)
{
	VipsImage *out = NULL;
	int rc;
	if (precision == VIPS_PRECISION_FLOAT)
		rc = vips_convf((VipsImage *) in, &out, (VipsImage *) mask, NULL);
	else if (vector)
		rc = ref_convi_vector((VipsImage *) in, &out, (VipsImage *) mask);
	else
		rc = vips_convi((VipsImage *) in, &out, (VipsImage *) mask, NULL);
	return rc ? NULL : out;
}

void *ref_gaussmat(double sigma, double min_ampl, int separable, int precision)
{
	VipsImage *out = NULL;
	return vips_gaussmat(&out, sigma, min_ampl, "separable", separable, "precision", precision, NULL) ? NULL : out;
}

int vips_convsep(VipsImage *in, VipsImage **out, VipsImage *mask, ...);
/* vips_convsep through the reference's own convsep.c + conv.c + rot.c (C-path convi / convf) */
void *ref_convsep(void *in, void *mask, int precision)
{
	VipsImage *out = NULL;
	return vips_convsep((VipsImage *) in, &out, (VipsImage *) mask, "precision", precision, NULL) ? NULL : out;
}
