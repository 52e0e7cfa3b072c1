/* ref_cast.c -- the reference's conversion/cast.c compiled in place. TEST INFRASTRUCTURE ONLY.
 *
 * vips_colour_build casts the detached alpha to the output format (colour.c:276), the identity rows of the route table
 * are vips_cast_* calls (colourspace.c:224, 314, 369, 423) and sRGB <-> RGB16 is a shifting cast (colourspace.c:88-110):
 * the reference's own cast loops run for all of them.
 */
#include <stdarg.h>
#include <vips/vips.h>
#define VIPS_LSHIFT_INT(I, N) ((int) ((unsigned int) (I) << (N))) /* include/vips/internal.h:87 */
/* the reference's own varargs front ends go through vips_call_split(): park them */
#define vips_cast vips_cast__via_call_split
#define vips_cast_uchar vips_cast_uchar__via_call_split
#define vips_cast_char vips_cast_char__via_call_split
#define vips_cast_ushort vips_cast_ushort__via_call_split
#define vips_cast_short vips_cast_short__via_call_split
#define vips_cast_uint vips_cast_uint__via_call_split
#define vips_cast_int vips_cast_int__via_call_split
#define vips_cast_float vips_cast_float__via_call_split
#define vips_cast_double vips_cast_double__via_call_split
#define vips_cast_complex vips_cast_complex__via_call_split
#define vips_cast_dpcomplex vips_cast_dpcomplex__via_call_split
/* cast.c:491 calls vips_cast before defining it (float pixels tagged as an integer space, with shift): that call goes
 * through the parked front end to vips_call_split("cast", ...) (ref_colour.c), which comes back to ref__castv below
 */
int vips_cast(VipsImage *in, VipsImage **out, VipsBandFormat format, ...);
#include "cast.c"
#undef vips_cast
#undef vips_cast_uchar
#undef vips_cast_char
#undef vips_cast_ushort
#undef vips_cast_short
#undef vips_cast_uint
#undef vips_cast_int
#undef vips_cast_float
#undef vips_cast_double
#undef vips_cast_complex
#undef vips_cast_dpcomplex

int
ref__castv(VipsImage *in, VipsImage **out, VipsBandFormat format, va_list ap)
{
	VipsCast *cast = (VipsCast *) vips__shim_object_new(vips_cast_get_type());
	VipsConversion *conversion = (VipsConversion *) cast;
	const char *name;

	cast->in = in;
	cast->format = format;
	cast->shift = FALSE;
	while ((name = va_arg(ap, const char *))) {
		if (strcmp(name, "shift") == 0)
			cast->shift = va_arg(ap, int);
		else
			return -1;
	}
	conversion->out = vips_image_new(); /* conversion.c:313 */
	if (vips_cast_build((VipsObject *) cast))
		return -1;
	*out = conversion->out;
	return 0;
}

int
vips_cast(VipsImage *in, VipsImage **out, VipsBandFormat format, ...)
{
	va_list ap;
	int rc;

	va_start(ap, format);
	rc = ref__castv(in, out, format, ap);
	va_end(ap);
	return rc;
}

#define REF_CAST(NAME, FORMAT) \
	int NAME(VipsImage *in, VipsImage **out, ...) \
	{ \
		va_list ap; \
		int rc; \
		va_start(ap, out); \
		rc = ref__castv(in, out, FORMAT, ap); \
		va_end(ap); \
		return rc; \
	}
REF_CAST(vips_cast_uchar, VIPS_FORMAT_UCHAR)
REF_CAST(vips_cast_ushort, VIPS_FORMAT_USHORT)
REF_CAST(vips_cast_short, VIPS_FORMAT_SHORT)
REF_CAST(vips_cast_float, VIPS_FORMAT_FLOAT)

void *
ref_cast(void *in, int format, int shift)
{
	VipsImage *out = NULL;
	return vips_cast((VipsImage *) in, &out, (VipsBandFormat) format, "shift", shift, NULL) ? NULL : out;
}
