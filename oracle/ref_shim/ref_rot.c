/* ref_rot.c -- the reference's conversion/rot.c compiled in place (vips_convsep rotates its mask with it,
 * convsep.c:92-96).  TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
typedef enum { VIPS_ANGLE_D0, VIPS_ANGLE_D90, VIPS_ANGLE_D180, VIPS_ANGLE_D270, VIPS_ANGLE_LAST } VipsAngle;
#define VIPS_TYPE_ANGLE 0
#define VIPS_MEMCPY(D, S, N) memcpy((D), (S), (N))
static int vips_image_pio_input(VipsImage *image) { return 0; }
/* the reference's varargs front ends go through vips_call_split(): park them */
#define vips_rot vips_rot__via_call_split
#define vips_rot90 vips_rot90__via_call_split
#define vips_rot180 vips_rot180__via_call_split
#define vips_rot270 vips_rot270__via_call_split
#include "rot.c"
#undef vips_rot
#undef vips_rot90
#undef vips_rot180
#undef vips_rot270

int
vips_rot(VipsImage *in, VipsImage **out, int angle, ...)
{
	VipsRot *rot = (VipsRot *) vips__shim_object_new(vips_rot_get_type());
	VipsConversion *conversion = (VipsConversion *) rot;

	rot->in = in;
	rot->angle = (VipsAngle) angle;
	conversion->out = vips_image_new(); /* conversion.c:313 */
	if (vips_rot_build((VipsObject *) rot))
		return -1;
	*out = conversion->out;
	return 0;
}
