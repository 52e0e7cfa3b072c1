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
