/* shim_ops.c -- eager stand-ins for the graph ops vips_sharpen_build / vips_convsep_build call that
 * are NOT on the hot path (band plumbing, copies, colourspace dispatch).  Each materialises its input
 * with the shim's sink and returns a memory image.  vips_colourspace, vips_cast* and vips_linear1 are NOT here: the
 * reference's own colourspace.c / colour.c / cast.c / linear.c serve them (ref_colourspace.c, ref_colourbuild.c,
 * ref_cast.c, ref_linear.c).  TEST INFRASTRUCTURE ONLY.
 */
#include <stdarg.h>
#include <vips/vips.h>

VipsImage *
vips__shim_new_memory(int w, int h, int bands, VipsBandFormat fmt, VipsInterpretation type)
{
	VipsImage *im = vips_image_new();
	vips_image_init_fields(im, w, h, bands, fmt, VIPS_CODING_NONE, type, 1.0, 1.0);
	im->dhint = VIPS_DEMAND_STYLE_ANY;
	im->data = (VipsPel *) calloc((size_t) w * h * bands > 0 ? (size_t) w * h * bands : 1, vips__shim_sizeof(fmt));
	return im;
}

VipsImage *
vips__shim_materialise(VipsImage *in)
{
	VipsImage *im;
	if (in->data && !in->generate_fn)
		return in;
	im = vips__shim_new_memory(in->Xsize, in->Ysize, in->Bands, in->BandFmt, in->Type);
	im->mat_scale = in->mat_scale;
	im->mat_offset = in->mat_offset;
	im->mat_meta_set = in->mat_meta_set;
	if (vips__shim_write_to_memory(in, im->data, 0, 0))
		return NULL;
	return im;
}

int
vips_copy(VipsImage *in, VipsImage **out, ...)
{
	VipsImage *m = vips__shim_materialise(in);
	VipsImage *im;
	if (!m)
		return -1;
	im = vips__shim_new_memory(m->Xsize, m->Ysize, m->Bands, m->BandFmt, m->Type);
	memcpy(im->data, m->data, (size_t) m->Xsize * m->Ysize * VIPS_IMAGE_SIZEOF_PEL(m));
	im->mat_scale = m->mat_scale;
	im->mat_offset = m->mat_offset;
	im->mat_meta_set = m->mat_meta_set;
	*out = im;
	return 0;
}

int
vips_extract_band(VipsImage *in, VipsImage **out, int band, ...)
{
	va_list ap;
	const char *name;
	int n = 1, x, k;
	size_t es;
	VipsImage *m, *im;
	va_start(ap, band);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "n") == 0)
			n = va_arg(ap, int);
		else
			return -1;
	}
	va_end(ap);
	if (!(m = vips__shim_materialise(in)) || band < 0 || n < 1 || band + n > m->Bands)
		return -1;
	es = VIPS_IMAGE_SIZEOF_ELEMENT(m);
	im = vips__shim_new_memory(m->Xsize, m->Ysize, n, m->BandFmt, m->Type); /* bandary keeps the interpretation */
	for (x = 0; x < m->Xsize * m->Ysize; x++)
		for (k = 0; k < n; k++)
			memcpy(im->data + ((size_t) x * n + k) * es, m->data + ((size_t) x * m->Bands + band + k) * es, es);
	*out = im;
	return 0;
}

int
vips_bandjoin2(VipsImage *in1, VipsImage *in2, VipsImage **out, ...)
{
	VipsImage *a = vips__shim_materialise(in1), *b = vips__shim_materialise(in2), *im;
	size_t es;
	int x;
	if (!a || !b || a->Xsize != b->Xsize || a->Ysize != b->Ysize || a->BandFmt != b->BandFmt)
		return -1;
	es = VIPS_IMAGE_SIZEOF_ELEMENT(a);
	/* vips_bandary_build copies the header of the first input: the interpretation is in1's */
	im = vips__shim_new_memory(a->Xsize, a->Ysize, a->Bands + b->Bands, a->BandFmt, a->Type);
	for (x = 0; x < a->Xsize * a->Ysize; x++) {
		memcpy(im->data + (size_t) x * (a->Bands + b->Bands) * es, a->data + (size_t) x * a->Bands * es, a->Bands * es);
		memcpy(im->data + ((size_t) x * (a->Bands + b->Bands) + a->Bands) * es, b->data + (size_t) x * b->Bands * es,
			b->Bands * es);
	}
	*out = im;
	return 0;
}

int
vips_image_pipeline_array(VipsImage *image, VipsDemandStyle hint, VipsImage **in)
{
	/* generate.c:397-420 over a NULL-terminated array; two inputs are all this shim needs */
	if (in[0] && in[1] && !in[2])
		return vips_image_pipelinev(image, hint, in[0], in[1], NULL);
	if (in[0] && !in[1])
		return vips_image_pipelinev(image, hint, in[0], NULL);
	return -1;
}

int
vips_reorder_prepare_many(VipsImage *image, VipsRegion **regions, VipsRect *r)
{
	int i;
	for (i = 0; regions[i]; i++)
		if (vips_region_prepare(regions[i], r))
			return -1;
	return 0;
}

/* iofuncs/error.c:1238-1249, :759-770: two one-line checks the builds make */
int
vips_check_separable(const char *domain, VipsImage *im)
{
	if (im->Xsize != 1 && im->Ysize != 1) {
		vips_error(domain, "%s", "separable matrix images must have width or height 1");
		return -1;
	}
	return 0;
}

int
vips_check_bands_atleast(const char *domain, VipsImage *im, int bands)
{
	if (im->Bands < bands) {
		vips_error(domain, "image must have at least %d bands", bands);
		return -1;
	}
	return 0;
}

/* header.c:737-758 (vips_image_pio_input): memory and partial images are both readable as they are */
int
vips_image_pio_input(VipsImage *image)
{
	return 0;
}

/* iofuncs/error.c:862-873 */
int
vips_check_coding(const char *domain, VipsImage *im, VipsCoding coding)
{
	if (im->Coding != coding) {
		vips_error(domain, "coding %d required", (int) coding);
		return -1;
	}
	return 0;
}

const char *
vips_enum_nick(GType enm, int value)
{
	static char txt[32];
	snprintf(txt, sizeof(txt), "%d", value); /* only inside error messages here */
	return txt;
}

/* vips_image_guess_interpretation (header.c:560-700) returns Type unchanged when it is sane for the image's bands and
 * format; every image the tests make is tagged with the interpretation its pixels are in
 */
VipsInterpretation
vips_image_guess_interpretation(const VipsImage *image)
{
	return image->Type;
}

/* conversion/bandjoin.c: n images band-interleaved, the first input's header */
int
vips_bandjoin(VipsImage **in, VipsImage **out, int n, ...)
{
	VipsImage *cur = in[0], *next;
	int i;
	for (i = 1; i < n; i++) {
		if (vips_bandjoin2(cur, in[i], &next, NULL))
			return -1;
		cur = next;
	}
	*out = cur;
	return 0;
}

VipsBandFormat
vips_image_get_format(const VipsImage *image)
{
	return image->BandFmt;
}

/* vips_image_guess_format, header.c:486-558, the interpretations the tests tag images with */
VipsBandFormat
vips_image_guess_format(const VipsImage *image)
{
	switch (image->Type) {
	case VIPS_INTERPRETATION_sRGB:
		return VIPS_FORMAT_UCHAR;
	case VIPS_INTERPRETATION_XYZ:
	case VIPS_INTERPRETATION_LAB:
	case VIPS_INTERPRETATION_LCH:
	case VIPS_INTERPRETATION_scRGB:
	case VIPS_INTERPRETATION_YXY:
		return VIPS_FORMAT_FLOAT;
	case VIPS_INTERPRETATION_LABS:
		return VIPS_FORMAT_SHORT;
	case VIPS_INTERPRETATION_RGB16:
		return VIPS_FORMAT_USHORT;
	default:
		return image->BandFmt;
	}
}
