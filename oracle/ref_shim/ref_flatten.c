/* ref_flatten.c -- the reference's conversion/flatten.c (uchar LUT paths, integer and double loops, the
 * cast-to-double detour for max_alpha below the format's range) compiled in place.  TEST INFRASTRUCTURE ONLY.
 *
 * vips__vector_to_ink (conversion/insert.c:244-359) builds the background pixel as black -> vips_linear(1, bg) ->
 * vips_cast(format): the linear step on a black uchar pixel is q = (float) (1.0 * 0 + bg) (linear.c LOOPN into its
 * float output, format table :425-429); the cast is the reference's own cast.c (ref_cast.c).
 */
#include <stdarg.h>
#include <vips/vips.h>
VipsImage *vips__shim_materialise(VipsImage *in);
int vips_cast(VipsImage *in, VipsImage **out, VipsBandFormat format, ...);

static VipsPel *
ref__flatten_ink(const char *domain, VipsImage *im, double *real, double *imag, int n)
{
	float px[64];
	VipsImage *f, *c, *m;
	int i;

	if (im->Bands > 64 || (n != 1 && n != im->Bands)) {
		vips_error(domain, "%s", "vector must have 1 or %d elements"); /* vips_linear's check, arithmetic/linear.c:265 */
		return NULL;
	}
	for (i = 0; i < im->Bands; i++)
		px[i] = (float) (1.0 * 0 + real[n == 1 ? 0 : i]);
	f = vips__shim_image_from_memory(px, 1, 1, im->Bands, VIPS_FORMAT_FLOAT, im->Type);
	if (vips_cast(f, &c, im->BandFmt, NULL) || !(m = vips__shim_materialise(c)))
		return NULL;
	return m->data;
}
#define vips__vector_to_ink ref__flatten_ink
#define vips_flatten vips_flatten__via_call_split
#include "flatten.c"
#undef vips_flatten

/* max_alpha <= 0: unset (the interpretation's default, flatten.c:454-455) */
void *ref_flatten(void *in, const double *background, int n, double max_alpha)
{
	static const char *set_max_alpha[] = { "max_alpha", NULL };
	VipsFlatten *flatten = (VipsFlatten *) vips__shim_object_new(vips_flatten_get_type());
	VipsConversion *conversion = (VipsConversion *) flatten;
	VipsArea *area;

	flatten->in = (VipsImage *) in;
	flatten->max_alpha = 255.0;
	if (max_alpha > 0) {
		flatten->max_alpha = max_alpha;
		((VipsObject *) flatten)->set_args = set_max_alpha;
	}
	flatten->background = vips_array_double_newv(1, 0.0);
	area = VIPS_AREA(flatten->background);
	area->n = n;
	area->data = malloc(sizeof(double) * (n ? n : 1));
	memcpy(area->data, background, sizeof(double) * n);
	conversion->out = vips_image_new(); /* conversion.c:313 */
	if (vips_flatten_build((VipsObject *) flatten))
		return NULL;
	return conversion->out;
}
