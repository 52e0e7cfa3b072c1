/* ref_colourbuild.c -- the reference's colour/colour.c compiled in place. TEST INFRASTRUCTURE ONLY.
 *
 * VipsColour's build (extra bands detached, alpha rescaled by the max_alpha ratio, cast and re-attached:
 * colour.c:196-291), vips_colour_gen, and the VipsColourTransform / VipsColourCode builds (input casts, colour.c:325-445).
 */
#include <stdarg.h>
#include <vips/vips.h>
/* g_object_set(colour, "out", out, NULL) is the only property write */
#define g_object_set(OBJ, NAME, VAL, END) (((VipsColour *) (OBJ))->out = (VAL))
#include "colour.c"
