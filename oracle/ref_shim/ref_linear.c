/* ref_linear.c -- the reference's arithmetic/linear.c compiled in place. TEST INFRASTRUCTURE ONLY.
 *
 * Only vips_linear1(alpha, &out, scale, 0.0) of vips_colour_build (colour.c:268-270) is served: the input is
 * materialised and the reference's own vips_linear_buffer (linear.c:324-412) runs over every line, with the output
 * format from its own format table (linear.c:425-429).  VipsArithmetic's build (bandalike, format promotion of several
 * inputs) has nothing to do for one input and two scalars, and is not compiled.
 */
#include <stdarg.h>
#include <vips/vips.h>
#define vips_linear vips_linear__via_call_split
#define vips_linear1 vips_linear1__via_call_split
#include "linear.c"
#undef vips_linear
#undef vips_linear1

VipsImage *vips__shim_materialise(VipsImage *in);
VipsImage *vips__shim_new_memory(int w, int h, int bands, VipsBandFormat fmt, VipsInterpretation type);

void
vips_arithmetic_set_format_table(VipsArithmeticClass *class, const VipsBandFormat *format_table)
{
	class->format_table = format_table; /* arithmetic.c:537-544 */
}

int
vips_linear1(VipsImage *in, VipsImage **out, double a, double b, ...)
{
	VipsImage *m = vips__shim_materialise(in);
	VipsLinear *linear;
	VipsArithmetic *arithmetic;
	VipsImage *ready[2];
	VipsImage *im;
	int i, y;

	if (!m)
		return -1;
	linear = (VipsLinear *) vips__shim_object_new(vips_linear_get_type());
	arithmetic = (VipsArithmetic *) linear;
	ready[0] = m;
	ready[1] = NULL;
	arithmetic->ready = ready;
	linear->uchar = FALSE;
	linear->single_element = TRUE;
	linear->a_ready = (double *) calloc(m->Bands, sizeof(double));
	linear->b_ready = (double *) calloc(m->Bands, sizeof(double));
	for (i = 0; i < m->Bands; i++) {
		linear->a_ready[i] = a;
		linear->b_ready[i] = b;
	}
	im = vips__shim_new_memory(m->Xsize, m->Ysize, m->Bands, vips_linear_format_table[m->BandFmt], m->Type);
	for (y = 0; y < m->Ysize; y++) {
		VipsPel *p[2];

		p[0] = m->data + (size_t) y * VIPS_IMAGE_SIZEOF_LINE(m);
		p[1] = NULL;
		vips_linear_buffer(arithmetic, im->data + (size_t) y * VIPS_IMAGE_SIZEOF_LINE(im), p, m->Xsize);
	}
	*out = im;
	return 0;
}
