/* colour_steps.cuh -- the per-pixel arithmetic of the reference's colour process_line functions as
 * device functions, shared by the colour route kernels (colour.cu), the fused sharpen (conv.cu) and the
 * fused linear-light thumbnail (thumbnail_linear.cu).  Each restates its reference function exactly --
 * same float / double mix, same evaluation order, explicit round-to-nearest intrinsics so nothing is
 * contracted into an FMA:
 *   sRGB2scRGB.c:71-107, scRGB2XYZ.c:58-79, XYZ2Lab.c:108-171, Lab2LabS.c:58-74, LabS2Lab.c:54-69,
 *   Lab2XYZ.c:83-143, XYZ2scRGB.c:72-97 (LabQ2sRGB.c:263-284), scRGB2sRGB.c:83-131 (LabQ2sRGB.c:290-361).
 */
#ifndef VB200_COLOUR_STEPS_CUH
#define VB200_COLOUR_STEPS_CUH

#include <climits>

#include <math_constants.h>

#include "vb200_internal.h"

namespace vb200 {

constexpr int kQuant = 100000; /* QUANT_ELEMENTS, XYZ2Lab.c:66 */

enum Step {
	S_sRGB2scRGB = 1, S_scRGB2XYZ, S_XYZ2Lab, S_Lab2LabS, S_LabS2Lab, S_Lab2XYZ, S_XYZ2scRGB, S_scRGB2sRGB,
	S_scRGB2RGB16, S_RGB162scRGB, S_Lab2LCh, S_LCh2Lab, S_XYZ2Yxy, S_Yxy2XYZ
};

struct ColourTables {
	const float *v2Y_8;	 /* [256] */
	const int *Y2v_8;	 /* [257] */
	const float *v2Y_16; /* [65536] */
	const int *Y2v_16;	 /* [65537] */
	const float *cbrt;	 /* [100000] */
	const float2 *cbrt2; /* [kQuant] (cbrt[i], cbrt[i + 1]) */
};

struct StepInfo {
	int step;
	int out_fmt;   /* format of the step's output image */
	float alpha_a; /* max_alpha_after / max_alpha_before, as float (vips_linear1 a1) */
	int rescale;   /* alpha scale changes on this step */
};

struct RouteParams {
	int n_steps;
	StepInfo steps[6];
	ColourTables t;
	int w, bands;
	size_t in_bpl, out_bpl;
	int in_fmt, out_fmt;
};

/* the device copies of the reference's tables (built on the host with the host libm, once per device) */
int get_tables(const char *domain, ColourTables *out);
/* the route of vips_colourspace_build among sRGB / RGB16 / scRGB / XYZ / LAB / LABS with the per-step
 * alpha handling of vips_colour_build filled in: n_steps, steps[], t.  -1: no such route.
 */
int colour_route_params(const char *domain, int source_space, int space, RouteParams *P);

namespace {

/* ------------------------------------------------------------ device steps */

/* x86 cvttss2si: out-of-range and NaN give INT_MIN (what "(int) nX" does in
 * the reference build); CUDA's cast would saturate instead.
 */
__device__ __forceinline__ int
x86_float_to_int(float v)
{
	if (!(v > -2147483904.0f && v < 2147483648.0f))
		return INT_MIN;
	return (int) v;
}

/* x / D for a compile-time constant D, correctly rounded, in 3 FP64 instructions instead of the
 * ~25 of __ddiv_rn: q0 = RN(x * RN(1 / D)), the exact remainder by FMA, one corrected
 * rounding (Markstein).  tests/test_div_const.py checks it against the hardware quotient for
 * every float mantissa (and float * 100000 products, and 2 * 10^7 random doubles) per constant
 * used here; zero remainders and infinities return q0 so that signed zeros and Inf survive.
 */
__device__ __forceinline__ double
div_const(double x, double d, double r)
{
	const double q0 = __dmul_rn(x, r);
	const double rem = __fma_rn(-q0, d, x);
	if (rem == 0.0 || !(fabs(q0) < CUDART_INF))
		return q0;
	return __fma_rn(rem, r, q0);
}
#define DIVC(x, D) div_const((x), (D), 1.0 / (D))

/* i = clip((int) nX, 0, QUANT_ELEMENTS - 2) and (float) i, XYZ2Lab.c:121-123.  For 0 <= nX < 2^23 the
 * truncation is an add with round-toward-zero against 2^23 (the integer lands in the significand, and
 * subtracting 2^23 again is exact): the fma pipe instead of two trips through the conversion unit.
 */
__device__ __forceinline__ int
cbrt_index(float nX, float *fi)
{
	if (nX >= 0.0f && nX < 8388608.0f) {
		const float t = __fadd_rz(nX, 8388608.0f);
		*fi = fminf(__fsub_rn(t, 8388608.0f), (float) (kQuant - 2));
		return min(__float_as_int(t) - 0x4B000000, kQuant - 2);
	}
	int i = x86_float_to_int(nX);
	i = max(0, min(kQuant - 2, i));
	*fi = (float) i;
	return i;
}

__device__ __forceinline__ float
cbrt_lookup(const float *__restrict__ table, float nX)
{
	float fi;
	const int i = cbrt_index(nX, &fi);
	const float f = __fsub_rn(nX, fi);
	const float t0 = __ldg(table + i), t1 = __ldg(table + i + 1);
	return __fadd_rn(t0, __fmul_rn(f, __fsub_rn(t1, t0)));
}

__device__ __forceinline__ void
step_scRGB2XYZ(float &a, float &b, float &c)
{
	/* p * VIPS_D65_Y0 is a double product rounded to float.  A 24-bit significand times 100 (7 bits) is
	 * exact in double, so that is ONE rounding of the exact product -- which is what the float multiply
	 * does (denormals, overflow and NaN included): no trip through the 64-bit conversion unit.
	 */
	const float R = __fmul_rn(a, 100.0F);
	const float G = __fmul_rn(b, 100.0F);
	const float B = __fmul_rn(c, 100.0F);
	a = __fadd_rn(__fadd_rn(__fmul_rn(0.4124F, R), __fmul_rn(0.3576F, G)), __fmul_rn(0.1805F, B));
	b = __fadd_rn(__fadd_rn(__fmul_rn(0.2126F, R), __fmul_rn(0.7152F, G)), __fmul_rn(0.0722F, B));
	c = __fadd_rn(__fadd_rn(__fmul_rn(0.0193F, R), __fmul_rn(0.1192F, G)), __fmul_rn(0.9505F, B));
}

__device__ __forceinline__ void
step_XYZ2scRGB(float &a, float &b, float &c)
{
	const float X = (float) DIVC((double) a, 100.0);
	const float Y = (float) DIVC((double) b, 100.0);
	const float Z = (float) DIVC((double) c, 100.0);
	a = __fadd_rn(__fadd_rn(__fmul_rn(3.240625F, X), __fmul_rn(-1.537208F, Y)), __fmul_rn(-0.498629F, Z));
	b = __fadd_rn(__fadd_rn(__fmul_rn(-0.968931F, X), __fmul_rn(1.875756F, Y)), __fmul_rn(0.041518F, Z));
	c = __fadd_rn(__fadd_rn(__fmul_rn(0.055710F, X), __fmul_rn(-0.204021F, Y)), __fmul_rn(1.056996F, Z));
}

__device__ __forceinline__ void
step_XYZ2Lab(const float *__restrict__ table, float &a, float &b, float &c)
{
	/* nX = QUANT_ELEMENTS * X / X0: float product, double quotient, float store */
	const float nX = (float) DIVC((double) __fmul_rn(100000.0f, a), 95.0470);
	const float nY = (float) DIVC((double) __fmul_rn(100000.0f, b), 100.0);
	const float nZ = (float) DIVC((double) __fmul_rn(100000.0f, c), 108.8827);
	const float cbx = cbrt_lookup(table, nX);
	const float cby = cbrt_lookup(table, nY);
	const float cbz = cbrt_lookup(table, nZ);
	a = __fsub_rn(__fmul_rn(116.0F, cby), 16.0F);
	b = __fmul_rn(500.0F, __fsub_rn(cbx, cby));
	c = __fmul_rn(200.0F, __fsub_rn(cby, cbz));
}

__device__ __forceinline__ void
step_Lab2XYZ(float &a, float &b, float &c)
{
	const double X0 = 95.0470, Y0 = 100.0, Z0 = 108.8827;
	const float L = a, A = b, B = c;
	double cby, tmp;
	float X, Y, Z;

	if ((double) L < 8.0) {
		Y = (float) DIVC(__dmul_rn((double) L, Y0), 903.3);
		cby = __dadd_rn(__dmul_rn(7.787, DIVC((double) Y, Y0)), 16.0 / 116.0);
	}
	else {
		cby = DIVC(__dadd_rn((double) L, 16.0), 116.0);
		Y = (float) __dmul_rn(__dmul_rn(__dmul_rn(Y0, cby), cby), cby);
	}
	tmp = __dadd_rn(DIVC((double) A, 500.0), cby);
	if (tmp < 0.2069)
		X = (float) DIVC(__dmul_rn(X0, __dsub_rn(tmp, 0.13793)), 7.787);
	else
		X = (float) __dmul_rn(__dmul_rn(__dmul_rn(X0, tmp), tmp), tmp);
	tmp = __dsub_rn(cby, DIVC((double) B, 200.0));
	if (tmp < 0.2069)
		Z = (float) DIVC(__dmul_rn(Z0, __dsub_rn(tmp, 0.13793)), 7.787);
	else
		Z = (float) __dmul_rn(__dmul_rn(__dmul_rn(Z0, tmp), tmp), tmp);
	a = X;
	b = Y;
	c = Z;
}

/* vips_col_scRGB2sRGB for one channel (LabQ2sRGB.c:323-353) with the table held as floats (its entries
 * are integers <= 65535, exact) and the three conversions done on the fma pipe: Yf is in [0, maxval], so
 * (int) Yf is a round-toward-zero add against 2^23, (float) (l1 - l0) is the exact float difference, and
 * rintf() of a value in [0, 65535] is the add / subtract of 1.5 * 2^23 in round-to-nearest-even.
 */
__device__ __forceinline__ int
scRGB2sRGB_channel_f(const float *lutf, float maxval, float R)
{
	float Yf = __fmul_rn(R, maxval);
	if (Yf < 0)
		Yf = 0;
	else if (Yf > maxval)
		Yf = maxval;
	const float t = __fadd_rz(Yf, 8388608.0f);
	const int Yi = __float_as_int(t) - 0x4B000000;
	const float fYi = __fsub_rn(t, 8388608.0f);
	const float l0 = lutf[Yi], l1 = lutf[Yi + 1];
	const float v = __fadd_rn(l0, __fmul_rn(__fsub_rn(l1, l0), __fsub_rn(Yf, fYi)));
	return __float_as_int(__fadd_rn(v, 12582912.0f)) - 0x4B400000;
}

/* vips_col_scRGB2sRGB for one channel, LabQ2sRGB.c:323-353 */
__device__ __forceinline__ int
scRGB2sRGB_channel(const int *lut, int maxval, float R)
{
	float Yf = __fmul_rn(R, (float) maxval);
	if (Yf < 0)
		Yf = 0;
	else if (Yf > maxval)
		Yf = maxval;
	const int Yi = (int) Yf;
	const int l0 = lut[Yi], l1 = lut[Yi + 1];
	const float v = __fadd_rn((float) l0, __fmul_rn((float) (l1 - l0), __fsub_rn(Yf, (float) Yi)));
	return (int) rintf(v);
}

__device__ __forceinline__ double
clipd(double lo, double v, double hi)
{
	/* VIPS_CLIP(A, V, B) = MAX(A, MIN(B, V)) with C's ?: on doubles */
	const double m = hi < v ? hi : v;
	return lo > m ? lo : m;
}

__device__ __forceinline__ double
load_elem(const void *p, int fmt, int idx)
{
	switch (fmt) {
	case VB200_FORMAT_UCHAR: return ((const uint8_t *) p)[idx];
	case VB200_FORMAT_CHAR: return ((const int8_t *) p)[idx];
	case VB200_FORMAT_USHORT: return ((const uint16_t *) p)[idx];
	case VB200_FORMAT_SHORT: return ((const int16_t *) p)[idx];
	case VB200_FORMAT_UINT: return ((const uint32_t *) p)[idx];
	case VB200_FORMAT_INT: return ((const int32_t *) p)[idx];
	default: return ((const float *) p)[idx];
	}
}

/* vips_cast of a value (conversion/cast.c:123-265): clip in double, truncate */
__device__ __forceinline__ double
cast_value(double v, int fmt)
{
	switch (fmt) {
	case VB200_FORMAT_UCHAR: return (double) (uint8_t) clipd(0, v, 255);
	case VB200_FORMAT_USHORT: return (double) (uint16_t) clipd(0, v, 65535);
	case VB200_FORMAT_SHORT: return (double) (int16_t) clipd(-32768, v, 32767);
	case VB200_FORMAT_FLOAT: return (double) (float) v;
	default: return v;
	}
}

__device__ __forceinline__ void
store_elem(void *p, int fmt, int idx, double v)
{
	switch (fmt) {
	case VB200_FORMAT_UCHAR: ((uint8_t *) p)[idx] = (uint8_t) v; break;
	case VB200_FORMAT_USHORT: ((uint16_t *) p)[idx] = (uint16_t) v; break;
	case VB200_FORMAT_SHORT: ((int16_t *) p)[idx] = (int16_t) v; break;
	default: ((float *) p)[idx] = (float) v; break;
	}
}

__device__ __forceinline__ float
cbrt_lookup2(const float2 *__restrict__ table, float nX)
{
	float fi;
	const int i = cbrt_index(nX, &fi);
	const float f = __fsub_rn(nX, fi);
	const float2 t = __ldg(table + i);
	return __fadd_rn(t.x, __fmul_rn(f, __fsub_rn(t.y, t.x)));
}

/* ---- Lab <-> LCh, XYZ <-> Yxy */

/* vips_Lab2LCh_line + vips_col_ab2h, Lab2LCh.c:61-124.  C is exact (float multiplies / add, IEEE sqrtf); the hue
 * goes through the double atan(), where CUDA's libm and glibc may differ by an ulp of the DOUBLE result -- after
 * the rounding to float that is the same float except in rare halfway cases: float results within 1 ULP.
 */
__device__ __forceinline__ void
step_Lab2LCh(float &a, float &b, float &c)
{
	const float A = b, B = c;
	const float C = __fsqrt_rn(__fadd_rn(__fmul_rn(A, A), __fmul_rn(B, B)));
	const double PI = 3.14159265358979323846;
	const double da = (double) A, db = (double) B;
	double h;
	if (da == 0) {
		if (db < 0.0)
			h = 270;
		else if (db == 0.0)
			h = 0;
		else
			h = 90;
	}
	else {
		const double t = atan(__ddiv_rn(db, da));
		if (da > 0.0)
			if (db < 0.0)
				h = __dmul_rn(__ddiv_rn(__dadd_rn(t, __dmul_rn(PI, 2.0)), __dmul_rn(2.0, PI)), 360.0);
			else
				h = __dmul_rn(__ddiv_rn(t, __dmul_rn(2.0, PI)), 360.0);
		else
			h = __dmul_rn(__ddiv_rn(__dadd_rn(t, PI), __dmul_rn(2.0, PI)), 360.0);
	}
	b = C;
	c = (float) h;
}

/* vips_LCh2Lab_line + vips_col_Ch2ab, LCh2Lab.c:70-103: cosf / sinf of (float) VIPS_RAD(h); within 1-2 ULP of glibc's */
__device__ __forceinline__ void
step_LCh2Lab(float &a, float &b, float &c)
{
	const double PI = 3.14159265358979323846;
	const float C = b, h = c;
	const float rad = (float) __dmul_rn(__dmul_rn(__ddiv_rn((double) h, 360.0), 2.0), PI);
	b = __fmul_rn(C, cosf(rad));
	c = __fmul_rn(C, sinf(rad));
}

/* vips_XYZ2Yxy_line, XYZ2Yxy.c:57-88: float sum, double quotients -- exact */
__device__ __forceinline__ void
step_XYZ2Yxy(float &a, float &b, float &c)
{
	const float X = a, Y = b, Z = c;
	const double total = (double) __fadd_rn(__fadd_rn(X, Y), Z);
	float x, y;
	if (total == 0.0) {
		x = 0;
		y = 0;
	}
	else {
		x = (float) __ddiv_rn((double) X, total);
		y = (float) __ddiv_rn((double) Y, total);
	}
	a = Y;
	b = x;
	c = y;
}

/* vips_Yxy2XYZ_line, Yxy2XYZ.c:59-93: float arithmetic as written -- exact */
__device__ __forceinline__ void
step_Yxy2XYZ(float &a, float &b, float &c)
{
	const float Y = a, x = b, y = c;
	float X, Z;
	if (x == 0.0f || y == 0.0f) {
		X = 0.0F;
		Z = 0.0F;
	}
	else {
		const float total = __fdiv_rn(Y, y);
		X = __fmul_rn(x, total);
		Z = __fdiv_rn(__fsub_rn(__fsub_rn(X, __fmul_rn(x, X)), __fmul_rn(x, Y)), x);
	}
	a = X;
	b = Y;
	c = Z;
}

/* a band beyond the third through a route, as vips_colour_build re-attaches it per step
 * (colour.c:252-291): rescale by max_alpha_after / max_alpha_before in float, then vips_cast
 */
__device__ __forceinline__ double
carry_extra_band(double v, const StepInfo *steps, int n_steps)
{
	for (int s = 0; s < n_steps; s++) {
		if (steps[s].rescale)
			v = (double) __fadd_rn(__fmul_rn(steps[s].alpha_a, (float) v), 0.0f);
		v = cast_value(v, steps[s].out_fmt);
	}
	return v;
}

} // namespace

} // namespace vb200

#endif
