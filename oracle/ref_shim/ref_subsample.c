/* ref_subsample.c -- the reference's conversion/subsample.c compiled in place: vips_resize's pre-shrink for
 * VIPS_KERNEL_NEAREST (resize.c:167-205).  TEST INFRASTRUCTURE ONLY. */
#include <stdarg.h>
#include <vips/vips.h>
#define VIPS_MEMCPY(Q, P, N) memcpy((Q), (P), (N)) /* include/vips/util.h:92 */
/* iofuncs/sinkscreen.c:202: "has this render been cancelled" -- never, there is no sinkscreen here */
static gboolean vips__worker_exit(void) { return FALSE; }
#define vips_subsample vips_subsample__via_call_split
#include "subsample.c"
#undef vips_subsample

int
vips_subsample(VipsImage *in, VipsImage **out, int xfac, int yfac, ...)
{
	VipsSubsample *subsample = (VipsSubsample *) vips__shim_object_new(vips_subsample_get_type());
	VipsConversion *conversion = (VipsConversion *) subsample;

	subsample->in = in;
	subsample->xfac = xfac;
	subsample->yfac = yfac;
	subsample->point = FALSE;
	conversion->out = vips_image_new(); /* conversion.c:313 */
	if (vips_subsample_build((VipsObject *) subsample))
		return -1;
	*out = conversion->out;
	return 0;
}

void *ref_subsample(void *in, int xfac, int yfac)
{
	VipsImage *out = NULL;
	return vips_subsample((VipsImage *) in, &out, xfac, yfac, NULL) ? NULL : out;
}
