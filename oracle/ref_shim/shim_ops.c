/* shim_ops.c -- eager stand-ins for the graph ops vips_sharpen_build / vips_convsep_build call that
 * are NOT on the hot path (band plumbing, copies, colourspace dispatch).  Each materialises its input
 * with the shim's sink and returns a memory image.  vips_colourspace runs the reference's own colour
 * *_line functions (ref_colour.c) in the order vips_colourspace_build's route table gives
 * (colourspace.c:317, 362-366).  TEST INFRASTRUCTURE ONLY.
 */
#include <stdarg.h>
#include <vips/vips.h>

int ref_colour_line(int step, const void *in, void *out, int n);

static VipsImage *
shim_new_memory(int w, int h, int bands, VipsBandFormat fmt, VipsInterpretation type)
{
	VipsImage *im = vips_image_new();
	vips_image_init_fields(im, w, h, bands, fmt, VIPS_CODING_NONE, type, 1.0, 1.0);
	im->dhint = VIPS_DEMAND_STYLE_ANY;
	im->data = (VipsPel *) calloc((size_t) w * h * bands > 0 ? (size_t) w * h * bands : 1, vips__shim_sizeof(fmt));
	return im;
}

static VipsImage *
shim_materialise(VipsImage *in)
{
	VipsImage *im;
	if (in->data && !in->generate_fn)
		return in;
	im = shim_new_memory(in->Xsize, in->Ysize, in->Bands, in->BandFmt, in->Type);
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
	VipsImage *m = shim_materialise(in);
	VipsImage *im;
	if (!m)
		return -1;
	im = shim_new_memory(m->Xsize, m->Ysize, m->Bands, m->BandFmt, m->Type);
	memcpy(im->data, m->data, (size_t) m->Xsize * m->Ysize * VIPS_IMAGE_SIZEOF_PEL(m));
	im->mat_scale = m->mat_scale;
	im->mat_offset = m->mat_offset;
	im->mat_meta_set = m->mat_meta_set;
	*out = im;
	return 0;
}

int
vips_cast_short(VipsImage *in, VipsImage **out, ...)
{
	if (in->BandFmt != VIPS_FORMAT_SHORT) {
		vips_error("shim", "vips_cast_short: only short input is modelled");
		return -1;
	}
	return vips_copy(in, out, NULL);
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
	if (!(m = shim_materialise(in)) || band < 0 || n < 1 || band + n > m->Bands)
		return -1;
	es = VIPS_IMAGE_SIZEOF_ELEMENT(m);
	im = shim_new_memory(m->Xsize, m->Ysize, n, m->BandFmt, m->Type); /* bandary keeps the interpretation */
	for (x = 0; x < m->Xsize * m->Ysize; x++)
		for (k = 0; k < n; k++)
			memcpy(im->data + ((size_t) x * n + k) * es, m->data + ((size_t) x * m->Bands + band + k) * es, es);
	*out = im;
	return 0;
}

int
vips_bandjoin2(VipsImage *in1, VipsImage *in2, VipsImage **out, ...)
{
	VipsImage *a = shim_materialise(in1), *b = shim_materialise(in2), *im;
	size_t es;
	int x;
	if (!a || !b || a->Xsize != b->Xsize || a->Ysize != b->Ysize || a->BandFmt != b->BandFmt)
		return -1;
	es = VIPS_IMAGE_SIZEOF_ELEMENT(a);
	/* vips_bandary_build copies the header of the first input: the interpretation is in1's */
	im = shim_new_memory(a->Xsize, a->Ysize, a->Bands + b->Bands, a->BandFmt, a->Type);
	for (x = 0; x < a->Xsize * a->Ysize; x++) {
		memcpy(im->data + (size_t) x * (a->Bands + b->Bands) * es, a->data + (size_t) x * a->Bands * es, a->Bands * es);
		memcpy(im->data + ((size_t) x * (a->Bands + b->Bands) + a->Bands) * es, b->data + (size_t) x * b->Bands * es,
			b->Bands * es);
	}
	*out = im;
	return 0;
}

/* step numbers of ref_colour_line */
static int
shim_route(VipsInterpretation from, VipsInterpretation to, int *steps)
{
	int n = 0;
	if (from == to)
		return 0;
	/* into LAB */
	if (from == VIPS_INTERPRETATION_sRGB) { steps[n++] = 1; steps[n++] = 2; steps[n++] = 3; }
	else if (from == VIPS_INTERPRETATION_LABS) steps[n++] = 5;
	else if (from != VIPS_INTERPRETATION_LAB) return -1;
	/* out of LAB */
	if (to == VIPS_INTERPRETATION_sRGB) { steps[n++] = 6; steps[n++] = 7; steps[n++] = 8; }
	else if (to == VIPS_INTERPRETATION_LABS) steps[n++] = 4;
	else if (to != VIPS_INTERPRETATION_LAB) return -1;
	/* LABS -> sRGB is LabS2Lab, Lab2XYZ, XYZ2scRGB, scRGB2sRGB; sRGB -> LABS ends Lab2LabS (colourspace.c:317, 366) */
	return n;
}

int
vips_colourspace(VipsImage *in, VipsImage **out, VipsInterpretation space, ...)
{
	int steps[8], n, i;
	VipsImage *m = shim_materialise(in), *cur;
	if (!m)
		return -1;
	if (m->Bands != 3) {
		vips_error("shim", "vips_colourspace: only 3-band images are modelled");
		return -1;
	}
	if ((n = shim_route(m->Type, space, steps)) < 0) {
		vips_error("shim", "vips_colourspace: route %d -> %d not modelled", m->Type, space);
		return -1;
	}
	cur = m;
	if (n == 0)
		return vips_copy(m, out, NULL);
	for (i = 0; i < n; i++) {
		static const VipsBandFormat ofmt[11] = { 0, VIPS_FORMAT_FLOAT, VIPS_FORMAT_FLOAT, VIPS_FORMAT_FLOAT, VIPS_FORMAT_SHORT,
			VIPS_FORMAT_FLOAT, VIPS_FORMAT_FLOAT, VIPS_FORMAT_FLOAT, VIPS_FORMAT_UCHAR, VIPS_FORMAT_USHORT, VIPS_FORMAT_FLOAT };
		VipsImage *next = shim_new_memory(m->Xsize, m->Ysize, 3, ofmt[steps[i]], space);
		if (ref_colour_line(steps[i], cur->data, next->data, m->Xsize * m->Ysize))
			return -1;
		cur = next;
	}
	*out = cur;
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

/* vips_cast to the format the image already has is a copy; morph.c:868 asks for uchar */
int
vips_cast(VipsImage *in, VipsImage **out, VipsBandFormat format, ...)
{
	if (in->BandFmt != format) {
		vips_error("shim", "vips_cast: only the identity cast is modelled");
		return -1;
	}
	*out = in;
	return 0;
}
