/* ref_rank.c -- the reference's morphology/rank.c (vips_rank_generate: min / max / select / uchar histogram
 * paths) compiled in place.  TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#define VIPS_TYPE_MORPHOLOGY (vips_morphology_get_type())
GType vips_morphology_get_type(void);
/* include/vips/util.h:127-134, which the shim header does not carry */
#define VIPS_SWAP(TYPE, A, B) do { TYPE t_ = (A); (A) = (B); (B) = t_; } while (0)
#define g_object_set(OBJ, NAME, VAL, END) (((VipsRank *) (OBJ))->out = (VAL))
#define vips_rank vips_rank__via_call_split
#define vips_median vips_median__via_call_split
#include "rank.c"
#undef vips_rank
#undef vips_median

void *ref_rank(void *in, int width, int height, int index)
{
	VipsRank *rank = (VipsRank *) vips__shim_object_new(vips_rank_get_type());
	((VipsMorphology *) rank)->in = (VipsImage *) in;
	rank->width = width;
	rank->height = height;
	rank->index = index;
	if (vips_rank_build((VipsObject *) rank))
		return NULL;
	return rank->out;
}
