/* ref_colour.c -- the reference's colour/*.c line functions and table builders,
 * compiled in place, driven directly (no class dispatch).  TEST INFRASTRUCTURE ONLY.
 */
#include "LabQ2sRGB.c"
#include "sRGB2scRGB.c"
#include "scRGB2XYZ.c"
#include "XYZ2scRGB.c"
#include "XYZ2Lab.c"
#include "Lab2XYZ.c"
#include "scRGB2sRGB.c"
#include "Lab2LabS.c"
#include "LabS2Lab.c"
#include "Lab2LCh.c"
#include "LCh2Lab.c"
#include "XYZ2Yxy.c"
#include "Yxy2XYZ.c"
#include "sRGB2HSV.c"
#include "HSV2sRGB.c"
#include "scRGB2BW.c"

/* step numbers as in oracle/colour_oracle.cpp */
int
ref_colour_line(int step, const void *in, void *out, int n)
{
	VipsImage im;
	VipsImage *ins[2] = { &im, NULL };
	VipsPel *inp[2] = { (VipsPel *) in, NULL };

	memset(&im, 0, sizeof(im));
	im.Bands = 3;
	switch (step) {
	case 1:
	case 10: {
		VipssRGB2scRGB obj;
		memset(&obj, 0, sizeof(obj));
		((VipsColour *) &obj)->in = ins;
		im.BandFmt = step == 1 ? VIPS_FORMAT_UCHAR : VIPS_FORMAT_USHORT;
		vips_col_make_tables_RGB_8();
		vips_col_make_tables_RGB_16();
		vips_sRGB2scRGB_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 2: {
		VipsscRGB2XYZ obj;
		memset(&obj, 0, sizeof(obj));
		vips_scRGB2XYZ_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 3: {
		VipsXYZ2Lab obj;
		memset(&obj, 0, sizeof(obj));
		vips_XYZ2Lab_init(&obj); /* D65 */
		vips_XYZ2Lab_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 4: {
		VipsLab2LabS obj;
		memset(&obj, 0, sizeof(obj));
		vips_Lab2LabS_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 5: {
		VipsLabS2Lab obj;
		memset(&obj, 0, sizeof(obj));
		vips_LabS2Lab_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 6: {
		VipsLab2XYZ obj;
		memset(&obj, 0, sizeof(obj));
		vips_Lab2XYZ_init(&obj);
		vips_Lab2XYZ_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 7: {
		VipsXYZ2scRGB obj;
		memset(&obj, 0, sizeof(obj));
		vips_XYZ2scRGB_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 8:
	case 9: {
		VipsscRGB2sRGB obj;
		memset(&obj, 0, sizeof(obj));
		obj.depth = step == 8 ? 8 : 16;
		vips_scRGB2sRGB_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 11: {
		VipsLab2LCh obj;
		memset(&obj, 0, sizeof(obj));
		vips_Lab2LCh_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 12: {
		VipsLCh2Lab obj;
		memset(&obj, 0, sizeof(obj));
		vips_LCh2Lab_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 13: {
		VipsXYZ2Yxy obj;
		memset(&obj, 0, sizeof(obj));
		vips_XYZ2Yxy_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 14: {
		VipsYxy2XYZ obj;
		memset(&obj, 0, sizeof(obj));
		vips_Yxy2XYZ_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 17: {
		VipssRGB2HSV obj;
		memset(&obj, 0, sizeof(obj));
		vips_sRGB2HSV_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 18: {
		VipsHSV2sRGB obj;
		memset(&obj, 0, sizeof(obj));
		vips_HSV2sRGB_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	case 19:
	case 20: {
		VipsscRGB2BW obj;
		memset(&obj, 0, sizeof(obj));
		obj.depth = step == 19 ? 8 : 16;
		vips_scRGB2BW_line((VipsColour *) &obj, (VipsPel *) out, inp, n);
		return 0;
	}
	}
	return -1;
}

const void *
ref_colour_table(int which, int *n)
{
	VipsXYZ2Lab obj;
	float xyz[3] = { 1, 1, 1 }, lab[3];
	VipsPel *inp[2] = { (VipsPel *) xyz, NULL };

	vips_col_make_tables_RGB_8();
	vips_col_make_tables_RGB_16();
	memset(&obj, 0, sizeof(obj));
	vips_XYZ2Lab_init(&obj);
	vips_XYZ2Lab_line((VipsColour *) &obj, (VipsPel *) lab, inp, 1); /* builds cbrt_table */
	switch (which) {
	case 0: *n = 257; return vips_Y2v_8;
	case 1: *n = 256; return vips_v2Y_8;
	case 2: *n = 65537; return vips_Y2v_16;
	case 3: *n = 65536; return vips_v2Y_16;
	case 4: *n = QUANT_ELEMENTS; return cbrt_table;
	}
	*n = 0;
	return NULL;
}

/* vips_call_split (iofuncs/operation.c:1089-1120) for the converters above: each one's own varargs front end calls
 * vips_call_split("nickname", optional, in, out).  The object is made from the converter's own type (its class_init and
 * init run), "in" is set (VipsColourTransform and VipsColourCode both keep it first, pcolour.h:116-159), the class's
 * build runs -- the converter's, then VipsColourTransform's / VipsColourCode's, then VipsColour's (colour.c under
 * ref_colourbuild.c) -- and "out" is handed back lazy: pixels come through vips_colour_gen and the line function.
 */
int ref__castv(VipsImage *in, VipsImage **out, VipsBandFormat format, va_list ap);

int
vips_call_split(const char *operation_name, va_list optional, ...)
{
	static const struct {
		const char *nickname;
		GType (*get_type)(void);
	} ops[] = {
		{ "sRGB2scRGB", vips_sRGB2scRGB_get_type }, { "scRGB2XYZ", vips_scRGB2XYZ_get_type },
		{ "XYZ2scRGB", vips_XYZ2scRGB_get_type }, { "XYZ2Lab", vips_XYZ2Lab_get_type }, { "Lab2XYZ", vips_Lab2XYZ_get_type },
		{ "scRGB2sRGB", vips_scRGB2sRGB_get_type }, { "Lab2LabS", vips_Lab2LabS_get_type },
		{ "LabS2Lab", vips_LabS2Lab_get_type }, { "Lab2LCh", vips_Lab2LCh_get_type }, { "LCh2Lab", vips_LCh2Lab_get_type },
		{ "XYZ2Yxy", vips_XYZ2Yxy_get_type }, { "Yxy2XYZ", vips_Yxy2XYZ_get_type }, { "sRGB2HSV", vips_sRGB2HSV_get_type },
		{ "HSV2sRGB", vips_HSV2sRGB_get_type }, { "scRGB2BW", vips_scRGB2BW_get_type }
	};
	static const char *set_in[] = { "in", NULL };
	va_list required;
	VipsImage *in, **out;
	VipsColour *colour;
	const char *name;
	int i;

	va_start(required, optional);
	in = va_arg(required, VipsImage *);
	out = va_arg(required, VipsImage **);
	if (strcmp(operation_name, "cast") == 0) {
		/* cast.c's own front ends (parked in ref_cast.c): the third required argument is the format */
		VipsBandFormat format = (VipsBandFormat) va_arg(required, int);
		va_end(required);
		return ref__castv(in, out, format, optional);
	}
	va_end(required);
	for (i = 0; i < VIPS_NUMBER(ops); i++)
		if (strcmp(ops[i].nickname, operation_name) == 0)
			break;
	if (i == VIPS_NUMBER(ops)) {
		vips_error("shim", "vips_call_split: %s is not compiled into this shim", operation_name);
		return -1;
	}
	colour = (VipsColour *) vips__shim_object_new(ops[i].get_type());
	((VipsColourTransform *) colour)->in = in;
	((VipsObject *) colour)->set_args = set_in; /* vips_sRGB2scRGB_build asks vips_object_argument_isset(object, "in") */
	while ((name = va_arg(optional, const char *))) {
		if (strcmp(name, "depth") == 0 && strcmp(operation_name, "scRGB2sRGB") == 0)
			((VipsscRGB2sRGB *) colour)->depth = va_arg(optional, int);
		else if (strcmp(name, "depth") == 0 && strcmp(operation_name, "scRGB2BW") == 0)
			((VipsscRGB2BW *) colour)->depth = va_arg(optional, int);
		else {
			vips_error("shim", "vips_call_split: %s: option %s is not modelled", operation_name, name);
			return -1;
		}
	}
	if (VIPS_OBJECT_GET_CLASS(colour)->build((VipsObject *) colour))
		return -1;
	*out = colour->out;
	return 0;
}
