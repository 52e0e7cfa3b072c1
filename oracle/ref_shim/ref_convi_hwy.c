/* ref_convi_hwy.c -- the reference's convolution/convi.c compiled AGAIN with HAVE_HWY, to get
 * its vips_convi_intize() (8-bit mantissa / shared exponent mask quantiser) and
 * vips_convi_uchar_vector_gen.  The Highway kernel itself (convi_hwy.cpp) needs libhwy, which
 * is not in this image: vips_convi_uchar_hwy below is its scalar tail, convi_hwy.cpp:265-273,
 * which the reference documents as the exact statement of the vector arithmetic.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#include "pconvolution.h"
#define HAVE_HWY 1
#define g_object_set(OBJ, NAME, VAL, END) (((VipsConvolution *) (OBJ))->out = (VAL))
/* every external-linkage name of convi.c gets a private alias in this second copy */
#define vips_convi vips_convi_hwy__via_call_split
#define vips__image_intize vips__image_intize__hwycopy
#define vips_convi_get_type vips_convi_hwy_get_type
#define vips_vector_isenabled() (1)
#include "convi.c"
#undef vips_convi

void
vips_convi_uchar_hwy(VipsRegion *out_region, VipsRegion *ir, VipsRect *r, int ne, int nnz, int offset,
	const int *restrict offsets, const short *restrict mant, int exp)
{
	int y, x, i;
	for (y = 0; y < r->height; y++) {
		VipsPel *p = VIPS_REGION_ADDR(ir, r->left, r->top + y);
		VipsPel *q = VIPS_REGION_ADDR(out_region, r->left, r->top + y);
		for (x = 0; x < ne; ++x) {
			int32_t sum = 1 << (exp - 1);
			for (i = 0; i < nnz; ++i)
				sum += p[offsets[i]] * mant[i];
			q[x] = VIPS_CLIP(0, (sum >> exp) + offset, UCHAR_MAX);
			p += 1;
		}
	}
}

int
ref_convi_vector(VipsImage *in, VipsImage **out, VipsImage *mask)
{
	VipsConvi *convi = (VipsConvi *) vips__shim_object_new(vips_convi_get_type());
	VipsConvolution *convolution = (VipsConvolution *) convi;

	convolution->in = in;
	convolution->mask = mask;
	if (vips_check_matrix("convi", mask, &convolution->M))
		return -1;
	if (vips_convi_build((VipsObject *) convi))
		return -1;
	*out = convolution->out;
	return 0;
}
