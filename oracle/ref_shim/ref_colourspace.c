/* ref_colourspace.c -- the reference's colour/colourspace.c compiled in place: the route table and
 * vips_colourspace_build.  TEST INFRASTRUCTURE ONLY.
 */
#include <stdarg.h>
#include <vips/vips.h>
#define D(N) int N(VipsImage *in, VipsImage **out, ...);
D(vips_cast_float) D(vips_cast_short) D(vips_cast_uchar) D(vips_cast_ushort) D(vips_rad2float)
D(vips_XYZ2Lab) D(vips_Lab2LabQ) D(vips_Lab2LCh) D(vips_LCh2CMC) D(vips_Lab2LabS) D(vips_XYZ2CMYK) D(vips_XYZ2scRGB)
D(vips_scRGB2sRGB) D(vips_sRGB2HSV) D(vips_scRGB2BW) D(vips_XYZ2Yxy) D(vips_XYZ2Oklab) D(vips_Oklab2Oklch) D(vips_Lab2XYZ)
D(vips_LabQ2Lab) D(vips_LabQ2LabS) D(vips_LabQ2sRGB) D(vips_LCh2Lab) D(vips_CMC2LCh) D(vips_LabS2Lab) D(vips_LabS2LabQ)
D(vips_CMYK2XYZ) D(vips_scRGB2XYZ) D(vips_sRGB2scRGB) D(vips_HSV2sRGB) D(vips_Yxy2XYZ) D(vips_Oklab2XYZ) D(vips_Oklch2Oklab)
#undef D
int vips_bandjoin(VipsImage **in, VipsImage **out, int n, ...);
/* g_object_set(colourspace, "out", vips_image_new(), NULL) is the only property write */
#define g_object_set(OBJ, NAME, VAL, END) (((VipsColourspace *) (OBJ))->out = (VAL))
/* the reference's own varargs front end goes through vips_call_split(): park it */
#define vips_colourspace vips_colourspace__via_call_split
#include "colourspace.c"
#undef vips_colourspace

int
vips_colourspace(VipsImage *in, VipsImage **out, VipsInterpretation space, ...)
{
	static const char *set_source_space[] = { "source_space", NULL };
	VipsColourspace *colourspace = (VipsColourspace *) vips__shim_object_new(vips_colourspace_get_type());
	va_list ap;
	const char *name;

	colourspace->in = in;
	colourspace->space = space;
	va_start(ap, space);
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "source_space") == 0) {
			colourspace->source_space = (VipsInterpretation) va_arg(ap, int);
			((VipsObject *) colourspace)->set_args = set_source_space;
		}
		else
			return -1;
	}
	va_end(ap);
	if (vips_colourspace_build((VipsObject *) colourspace))
		return -1;
	*out = colourspace->out;
	return 0;
}

void *
ref_colourspace_build(void *in, int space, int source_space)
{
	VipsImage *out = NULL;
	return vips_colourspace((VipsImage *) in, &out, (VipsInterpretation) space, "source_space", source_space, NULL) ? NULL : out;
}
