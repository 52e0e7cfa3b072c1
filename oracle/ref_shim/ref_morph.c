/* ref_morph.c -- the reference's morphology/morph.c (C path: vips_dilate_gen / vips_erode_gen) compiled in place.
 * TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
typedef enum { VIPS_OPERATION_MORPHOLOGY_ERODE, VIPS_OPERATION_MORPHOLOGY_DILATE, VIPS_OPERATION_MORPHOLOGY_LAST } VipsOperationMorphology;
#define VIPS_TYPE_OPERATION_MORPHOLOGY 0
#define VIPS_TYPE_MORPHOLOGY (vips_morphology_get_type())
GType vips_morphology_get_type(void);
int vips__image_intize(VipsImage *in, VipsImage **out);
#define g_object_set(OBJ, NAME, VAL, END) (((VipsMorph *) (OBJ))->out = (VAL))
#define vips_morph vips_morph__via_call_split
#include "morph.c"
#undef vips_morph

int
vips_morph(VipsImage *in, VipsImage **out, VipsImage *mask, int morph, ...)
{
	VipsMorph *m = (VipsMorph *) vips__shim_object_new(vips_morph_get_type());
	((VipsMorphology *) m)->in = in;
	m->mask = mask;
	m->morph = (VipsOperationMorphology) morph;
	if (vips_morph_build((VipsObject *) m))
		return -1;
	*out = m->out;
	return 0;
}

void *ref_morph(void *in, void *mask, int morph)
{
	VipsImage *out = NULL;
	return vips_morph((VipsImage *) in, &out, (VipsImage *) mask, morph, NULL) ? NULL : out;
}
