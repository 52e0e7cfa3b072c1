/* colour_oracle.cpp -- CPU restatement of the reference's colour hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  -O2 -ffp-contract=off.
 *
 * Follows (reference = libvips 8.19, libvips/colour/):
 *   LabQ2sRGB.c:130-159      calcul_tables (vips_Y2v_8/16, vips_v2Y_8/16; powf on the host)
 *   LabQ2sRGB.c:224-284      vips_col_scRGB2XYZ / vips_col_XYZ2scRGB
 *   LabQ2sRGB.c:290-361      vips_col_scRGB2sRGB (LUT + lerp + rintf)
 *   sRGB2scRGB.c:71-107      vips_sRGB2scRGB_line
 *   scRGB2XYZ.c:58-79        vips_scRGB2XYZ_line
 *   XYZ2scRGB.c:72-97        vips_XYZ2scRGB_line
 *   XYZ2Lab.c:91-171         table_init (cbrtf on the host), vips_col_XYZ2Lab_helper
 *   Lab2XYZ.c:83-143         vips_col_Lab2XYZ_helper
 *   scRGB2sRGB.c:83-131      vips_scRGB2sRGB_line
 *   Lab2LabS.c:58-74, LabS2Lab.c:54-69
 *   colour.c:159-296         vips_colour_build: extra bands detached, rescaled
 *                            (vips_linear1, arithmetic/linear.c:213-223), cast
 *                            (conversion/cast.c:123-265) and re-attached
 *   colourspace.c:223-497    the route table
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <mutex>
#include <vector>

#include "oracle.h"

#define D65_X0 (95.0470)
#define D65_Y0 (100.0)
#define D65_Z0 (108.8827)
#define QUANT_ELEMENTS (100000)

static int Y2v_8[256 + 1];
static float v2Y_8[256];
static int Y2v_16[65536 + 1];
static float v2Y_16[65536];
static float cbrt_table[QUANT_ELEMENTS];
static std::once_flag tables_once;

/* LabQ2sRGB.c:130-159 */
static void
calcul_tables(int range, int *Y2v, float *v2Y)
{
	for (int i = 0; i < range; i++) {
		float f = (float) i / (range - 1);
		float v;

		if (f <= 0.0031308)
			v = 12.92F * f;
		else
			v = (1.0F + 0.055F) * powf(f, 1.0F / 2.4F) - 0.055F;

		Y2v[i] = rintf((range - 1) * v);
	}
	Y2v[range] = Y2v[range - 1];

	for (int i = 0; i < range; i++) {
		float f = (float) i / (range - 1);

		if (f <= 0.04045)
			v2Y[i] = f / 12.92F;
		else
			v2Y[i] = powf((f + 0.055F) / (1 + 0.055F), 2.4F);
	}
}

static void
make_tables()
{
	std::call_once(tables_once, []() {
		calcul_tables(256, Y2v_8, v2Y_8);
		calcul_tables(65536, Y2v_16, v2Y_16);
		/* XYZ2Lab.c:91-106 */
		for (int i = 0; i < QUANT_ELEMENTS; i++) {
			float Y = (double) i / QUANT_ELEMENTS;

			if (Y < 0.008856)
				cbrt_table[i] = 7.787F * Y + (16.0F / 116.0F);
			else
				cbrt_table[i] = cbrtf(Y);
		}
	});
}

extern "C" const void *
orc_colour_table(int which, int *n)
{
	make_tables();
	switch (which) {
	case 0: *n = 257; return Y2v_8;
	case 1: *n = 256; return v2Y_8;
	case 2: *n = 65537; return Y2v_16;
	case 3: *n = 65536; return v2Y_16;
	case 4: *n = QUANT_ELEMENTS; return cbrt_table;
	}
	*n = 0;
	return nullptr;
}

/* ------------------------------------------------------------ line functions
 * Each takes n 3-band pixels.
 */

static void
sRGB2scRGB_line(const void *in, int in_fmt, float *q, size_t n)
{
	if (in_fmt == ORC_FORMAT_UCHAR) {
		const uint8_t *p = (const uint8_t *) in;
		for (size_t i = 0; i < 3 * n; i++)
			q[i] = v2Y_8[p[i]];
	}
	else {
		const uint16_t *p = (const uint16_t *) in;
		for (size_t i = 0; i < 3 * n; i++)
			q[i] = v2Y_16[p[i]];
	}
}

static void
scRGB2XYZ_line(const float *p, float *q, size_t n)
{
	for (size_t i = 0; i < n; i++) {
		const float R = p[0] * D65_Y0;
		const float G = p[1] * D65_Y0;
		const float B = p[2] * D65_Y0;

		q[0] = 0.4124F * R + 0.3576F * G + 0.1805F * B;
		q[1] = 0.2126F * R + 0.7152F * G + 0.0722F * B;
		q[2] = 0.0193F * R + 0.1192F * G + 0.9505F * B;
		p += 3;
		q += 3;
	}
}

static void
XYZ2scRGB_line(const float *p, float *q, size_t n)
{
	for (size_t i = 0; i < n; i++) {
		float X = p[0], Y = p[1], Z = p[2];

		X /= D65_Y0;
		Y /= D65_Y0;
		Z /= D65_Y0;
		q[0] = 3.240625F * X + -1.537208F * Y + -0.498629F * Z;
		q[1] = -0.968931F * X + 1.875756F * Y + 0.041518F * Z;
		q[2] = 0.055710F * X + -0.204021F * Y + 1.056996F * Z;
		p += 3;
		q += 3;
	}
}

/* the reference converts float to int with x86 cvttss2si semantics: anything
 * outside int range (and NaN) becomes INT_MIN
 */
static inline int
x86_float_to_int(float v)
{
	if (!(v > -2147483904.0f && v < 2147483648.0f))
		return INT_MIN;
	return (int) v;
}

static inline float
cbrt_lookup(float nX)
{
	int i = x86_float_to_int(nX);
	i = std::max(0, std::min(QUANT_ELEMENTS - 2, i));
	float f = nX - i;
	return cbrt_table[i] + f * (cbrt_table[i + 1] - cbrt_table[i]);
}

static void
XYZ2Lab_line(const float *p, float *q, size_t n, double X0, double Y0, double Z0)
{
	for (size_t x = 0; x < n; x++) {
		const float X = p[0], Y = p[1], Z = p[2];
		float nX, nY, nZ;
		float cbx, cby, cbz;

		nX = QUANT_ELEMENTS * X / X0;
		nY = QUANT_ELEMENTS * Y / Y0;
		nZ = QUANT_ELEMENTS * Z / Z0;
		cbx = cbrt_lookup(nX);
		cby = cbrt_lookup(nY);
		cbz = cbrt_lookup(nZ);
		q[0] = 116.0F * cby - 16.0F;
		q[1] = 500.0F * (cbx - cby);
		q[2] = 200.0F * (cby - cbz);
		p += 3;
		q += 3;
	}
}

static void
Lab2XYZ_line(const float *p, float *q, size_t n, double X0, double Y0, double Z0)
{
	for (size_t x = 0; x < n; x++) {
		const float L = p[0], a = p[1], b = p[2];
		float X, Y, Z;
		double cby, tmp;

		if (L < 8.0) {
			Y = (L * Y0) / 903.3;
			cby = 7.787 * (Y / Y0) + 16.0 / 116.0;
		}
		else {
			cby = (L + 16.0) / 116.0;
			Y = Y0 * cby * cby * cby;
		}
		tmp = a / 500.0 + cby;
		if (tmp < 0.2069)
			X = X0 * (tmp - 0.13793) / 7.787;
		else
			X = X0 * tmp * tmp * tmp;
		tmp = cby - b / 200.0;
		if (tmp < 0.2069)
			Z = Z0 * (tmp - 0.13793) / 7.787;
		else
			Z = Z0 * tmp * tmp * tmp;
		q[0] = X;
		q[1] = Y;
		q[2] = Z;
		p += 3;
		q += 3;
	}
}

/* LabQ2sRGB.c:290-361 for one channel */
static inline int
scRGB2sRGB_channel(float R, int maxval, const int *lut)
{
	float Yf = R * maxval;
	if (Yf < 0)
		Yf = 0;
	else if (Yf > maxval)
		Yf = maxval;
	int Yi = (int) Yf;
	float v = lut[Yi] + (lut[Yi + 1] - lut[Yi]) * (Yf - Yi);
	return rintf(v);
}

static void
scRGB2sRGB_line(const float *p, void *out, int depth, size_t n)
{
	const int maxval = depth == 16 ? 65535 : 255;
	const int *lut = depth == 16 ? Y2v_16 : Y2v_8;

	for (size_t i = 0; i < n; i++) {
		const float R = p[0], G = p[1], B = p[2];
		int r, g, b;

		if (std::isnan(R) || std::isnan(G) || std::isnan(B))
			r = g = b = 0;
		else {
			r = scRGB2sRGB_channel(R, maxval, lut);
			g = scRGB2sRGB_channel(G, maxval, lut);
			b = scRGB2sRGB_channel(B, maxval, lut);
		}
		if (depth == 16) {
			uint16_t *q = (uint16_t *) out + 3 * i;
			q[0] = r;
			q[1] = g;
			q[2] = b;
		}
		else {
			uint8_t *q = (uint8_t *) out + 3 * i;
			q[0] = r;
			q[1] = g;
			q[2] = b;
		}
		p += 3;
	}
}

/* VIPS_CLIP's ?: forms (a NaN falls through both and converts to 0 on x86) */
static inline double
clipd_macro(double lo, double v, double hi)
{
	const double m = hi < v ? hi : v;
	return lo > m ? lo : m;
}
#define CLIPD(A, V, B) clipd_macro((A), (V), (B))

static void
Lab2LabS_line(const float *p, int16_t *q, size_t n)
{
	for (size_t i = 0; i < n; i++) {
		q[0] = CLIPD(0, p[0] * (32767.0 / 100.0), SHRT_MAX);
		q[1] = CLIPD(SHRT_MIN, p[1] * (32768.0 / 128.0), SHRT_MAX);
		q[2] = CLIPD(SHRT_MIN, p[2] * (32768.0 / 128.0), SHRT_MAX);
		q += 3;
		p += 3;
	}
}

static void
LabS2Lab_line(const int16_t *p, float *q, size_t n)
{
	for (size_t i = 0; i < n; i++) {
		q[0] = p[0] / (32767.0 / 100.0);
		q[1] = p[1] / (32768.0 / 128.0);
		q[2] = p[2] / (32768.0 / 128.0);
		p += 3;
		q += 3;
	}
}

/* ----------------------------------------------------------------- steps */

enum { S_sRGB2scRGB = 1, S_scRGB2XYZ, S_XYZ2Lab, S_Lab2LabS, S_LabS2Lab, S_Lab2XYZ, S_XYZ2scRGB, S_scRGB2sRGB,
	S_scRGB2RGB16, S_RGB162scRGB, S_Lab2LCh, S_LCh2Lab, S_XYZ2Yxy, S_Yxy2XYZ, S_sRGB2RGB16, S_RGB162sRGB,
	S_sRGB2HSV, S_HSV2sRGB, S_scRGB2BW, S_scRGB2BW16, S_BW2sRGB, S_GREY162RGB16 };

/* ---- the next VipsColour converters: Lab <-> LCh, XYZ <-> Yxy */

/* vips_col_ab2h, Lab2LCh.c:61-89 */
static double
col_ab2h(double a, double b)
{
	const double PI = 3.14159265358979323846;
	double h;
	if (a == 0) {
		if (b < 0.0)
			h = 270;
		else if (b == 0.0)
			h = 0;
		else
			h = 90;
	}
	else {
		double t = atan(b / a);
		if (a > 0.0)
			if (b < 0.0)
				h = ((t + PI * 2.0) / (2.0 * PI)) * 360.0;
			else
				h = (t / (2.0 * PI)) * 360.0;
		else
			h = ((t + PI) / (2.0 * PI)) * 360.0;
	}
	return h;
}

/* vips_Lab2LCh_line, Lab2LCh.c:98-124 */
static void
Lab2LCh_line(const float *p, float *q, size_t n)
{
	for (size_t x = 0; x < n; x++) {
		float L = p[0], a = p[1], b = p[2];
		float C = sqrtf(a * a + b * b);
		float h = col_ab2h(a, b);
		q[0] = L;
		q[1] = C;
		q[2] = h;
		p += 3;
		q += 3;
	}
}

/* vips_LCh2Lab_line + vips_col_Ch2ab, LCh2Lab.c:70-103 */
static void
LCh2Lab_line(const float *p, float *q, size_t n)
{
	const double PI = 3.14159265358979323846;
	for (size_t x = 0; x < n; x++) {
		float L = p[0], C = p[1], h = p[2];
		float a = C * cosf(((h) / 360.0) * 2.0 * PI);
		float b = C * sinf(((h) / 360.0) * 2.0 * PI);
		q[0] = L;
		q[1] = a;
		q[2] = b;
		p += 3;
		q += 3;
	}
}

/* vips_XYZ2Yxy_line, XYZ2Yxy.c:57-88 */
static void
XYZ2Yxy_line(const float *p, float *q, size_t n)
{
	for (size_t i = 0; i < n; i++) {
		float X = p[0], Y = p[1], Z = p[2];
		double total = X + Y + Z;
		float x, y;
		if (total == 0.0) {
			x = 0;
			y = 0;
		}
		else {
			x = X / total;
			y = Y / total;
		}
		q[0] = Y;
		q[1] = x;
		q[2] = y;
		p += 3;
		q += 3;
	}
}

/* vips_Yxy2XYZ_line, Yxy2XYZ.c:59-93 */
static void
Yxy2XYZ_line(const float *p, float *q, size_t n)
{
	for (size_t i = 0; i < n; i++) {
		float Y = p[0], x = p[1], y = p[2];
		float X, Z;
		if (x == 0.0 || y == 0.0) {
			X = 0.0F;
			Z = 0.0F;
		}
		else {
			float total = Y / y;
			X = x * total;
			Z = (X - x * X - x * Y) / x;
		}
		q[0] = X;
		q[1] = Y;
		q[2] = Z;
		p += 3;
		q += 3;
	}
}


/* ---- SURVEY 8f rank 3, second lot: sRGB <-> HSV, scRGB -> B_W / GREY16, B_W -> sRGB, GREY16 -> RGB16 */

/* vips_sRGB2HSV_line, sRGB2HSV.c:50-126: uchar in, uchar out; the double expressions are stored by C truncation */
static void
sRGB2HSV_line(const uint8_t *p, uint8_t *q, size_t n)
{
	for (size_t i = 0; i < n; i++, p += 3, q += 3) {
		unsigned char c_max, c_min;
		float secondary_diff, wrap_around_hue;
		if (p[1] < p[2]) {
			if (p[2] < p[0]) {
				c_max = p[0];
				c_min = p[1];
				secondary_diff = p[1] - p[2];
				wrap_around_hue = 255.0F;
			}
			else {
				c_max = p[2];
				c_min = p[1] < p[0] ? p[1] : p[0];
				secondary_diff = p[0] - p[1];
				wrap_around_hue = 170.0F;
			}
		}
		else {
			if (p[1] < p[0]) {
				c_max = p[0];
				c_min = p[2];
				secondary_diff = p[1] - p[2];
				wrap_around_hue = 0.0F;
			}
			else {
				c_max = p[1];
				c_min = p[2] < p[0] ? p[2] : p[0];
				secondary_diff = p[2] - p[0];
				wrap_around_hue = 85.0F;
			}
		}
		if (c_max == 0)
			q[0] = q[1] = q[2] = 0;
		else {
			q[2] = c_max;
			const unsigned char delta = c_max - c_min;
			if (delta == 0)
				q[0] = 0;
			else
				q[0] = 42.5 * (secondary_diff / (float) delta) + wrap_around_hue;
			q[1] = delta * 255.0 / (float) c_max;
		}
	}
}

/* vips_HSV2sRGB_line, HSV2sRGB.c:54-108 (SIXTH_OF_CHAR 42.5, a double) */
static void
HSV2sRGB_line(const uint8_t *p, uint8_t *q, size_t n)
{
	for (size_t i = 0; i < n; i++, p += 3, q += 3) {
		float c, x, m;
		c = p[2] * p[1] / 255.0;
		x = c * (1 - fabsf(fmodf(p[0] / 42.5, 2) - 1));
		m = p[2] - c;
		if (p[0] < (int) 42.5) {
			q[0] = c + m;
			q[1] = x + m;
			q[2] = 0 + m;
		}
		else if (p[0] < (int) (2 * 42.5)) {
			q[0] = x + m;
			q[1] = c + m;
			q[2] = 0 + m;
		}
		else if (p[0] < (int) (3 * 42.5)) {
			q[0] = 0 + m;
			q[1] = c + m;
			q[2] = x + m;
		}
		else if (p[0] < (int) (4 * 42.5)) {
			q[0] = 0 + m;
			q[1] = x + m;
			q[2] = c + m;
		}
		else if (p[0] < (int) (5 * 42.5)) {
			q[0] = x + m;
			q[1] = 0 + m;
			q[2] = c + m;
		}
		else {
			q[0] = c + m;
			q[1] = 0 + m;
			q[2] = x + m;
		}
	}
}

/* vips_scRGB2BW_line, scRGB2BW.c:58-105, over vips_col_scRGB2BW, LabQ2sRGB.c:385-429: the CIE luminance of the linear
 * pixel through the same interpolated gamma table as scRGB -> sRGB; three floats in, one uchar / ushort out
 */
static void
scRGB2BW_line(const float *p, uint8_t *q, int depth, size_t n)
{
	const int maxval = depth == 16 ? 65535 : 255;
	const int *lut = depth == 16 ? Y2v_16 : Y2v_8;
	for (size_t i = 0; i < n; i++, p += 3) {
		const float R = p[0], G = p[1], B = p[2];
		const float Y = 0.2126F * R + 0.7152F * G + 0.0722F * B;
		int g;
		if (std::isnan(Y))
			g = 0;
		else {
			float Yf = Y * maxval;
			if (Yf < 0)
				Yf = 0;
			else if (Yf > maxval)
				Yf = maxval;
			const int Yi = (int) Yf;
			const float v = lut[Yi] + (lut[Yi + 1] - lut[Yi]) * (Yf - Yi);
			g = rintf(v);
		}
		if (depth == 16)
			((uint16_t *) q)[i] = g;
		else
			q[i] = g;
	}
}

struct Img {
	int w = 0, h = 0, bands = 0, fmt = 0, type = 0;
	std::vector<uint8_t> data;
	size_t npix() const { return (size_t) w * h; }
};

static double
max_alpha_of(int type)
{
	/* iofuncs/header.c:195-206 */
	if (type == 26 || type == 25)
		return 65535.0;
	if (type == 28)
		return 1.0;
	return 255.0;
}

static double
elem_as_double(const uint8_t *p, int fmt)
{
	switch (fmt) {
	case ORC_FORMAT_UCHAR: return *(const uint8_t *) p;
	case ORC_FORMAT_CHAR: return *(const int8_t *) p;
	case ORC_FORMAT_USHORT: return *(const uint16_t *) p;
	case ORC_FORMAT_SHORT: return *(const int16_t *) p;
	case ORC_FORMAT_UINT: return *(const uint32_t *) p;
	case ORC_FORMAT_INT: return *(const int32_t *) p;
	case ORC_FORMAT_FLOAT: return *(const float *) p;
	case ORC_FORMAT_DOUBLE: return *(const double *) p;
	}
	return 0;
}

/* vips_cast of one element held as float (or an int format held exactly in a
 * double): conversion/cast.c:123-265.  float -> int formats clip in double
 * then truncate; int -> int clip in int.
 */
static void
cast_store(double v, bool from_float, int ofmt, uint8_t *q)
{
	switch (ofmt) {
	case ORC_FORMAT_UCHAR: *(uint8_t *) q = (uint8_t) CLIPD(0, v, UCHAR_MAX); break;
	case ORC_FORMAT_CHAR: *(int8_t *) q = (int8_t) CLIPD(SCHAR_MIN, v, SCHAR_MAX); break;
	case ORC_FORMAT_USHORT: *(uint16_t *) q = (uint16_t) CLIPD(0, v, USHRT_MAX); break;
	case ORC_FORMAT_SHORT: *(int16_t *) q = (int16_t) CLIPD(SHRT_MIN, v, SHRT_MAX); break;
	case ORC_FORMAT_FLOAT: *(float *) q = (float) v; break;
	case ORC_FORMAT_DOUBLE: *(double *) q = v; break;
	default: break;
	}
	(void) from_float;
}

/* vips_sRGB2RGB16 / vips_RGB162sRGB, colourspace.c:85-110: not colour objects at all but vips_cast(..., "shift", TRUE)
 * over EVERY band ("we can short-circuit the extra band processing"), then the interpretation is re-tagged.  The
 * integer loops are cast.c:137-164: a right shift by the width difference going down; going up, a left shift with the
 * bottom bit copied into the new bits.  A cast to the format the image already has is a copy (cast.c:476-477).
 */
static int
shift_cast_step(int step, const Img &in, Img &out)
{
	const int ofmt = step == S_sRGB2RGB16 ? ORC_FORMAT_USHORT : ORC_FORMAT_UCHAR;
	if (in.fmt != ORC_FORMAT_UCHAR && in.fmt != ORC_FORMAT_USHORT)
		return -1; /* float pixels tagged as an integer space (cast.c:483-495): not restated */
	const size_t cnt = in.npix() * in.bands;
	out.w = in.w;
	out.h = in.h;
	out.bands = in.bands;
	out.fmt = ofmt;
	out.type = step == S_sRGB2RGB16 ? 25 : 22;
	out.data.resize(cnt * orc_sizeof_format(ofmt));
	if (in.fmt == ofmt) {
		out.data = in.data;
		return 0;
	}
	if (ofmt == ORC_FORMAT_USHORT) {
		const uint8_t *p = in.data.data();
		uint16_t *q = (uint16_t *) out.data.data();
		for (size_t i = 0; i < cnt; i++)
			q[i] = (uint16_t) ((p[i] << 8) | (((p[i] & 1) << 8) - (p[i] & 1)));
	}
	else {
		const uint16_t *p = (const uint16_t *) in.data.data();
		uint8_t *q = out.data.data();
		for (size_t i = 0; i < cnt; i++)
			q[i] = (uint8_t) (p[i] >> 8);
	}
	return 0;
}

/* vips_BW2sRGB / vips_GREY162RGB16, colourspace.c:152-188: not colour objects but vips__colourspace_process_n(in, 1,
 * bandjoin(in, in, in)): the first band three times, the other bands cast to the same format (a copy) and re-attached;
 * the format stays whatever it was, only Type changes.
 */
static int
replicate_step(int step, const Img &in, Img &out)
{
	if (in.bands < 1)
		return -1;
	const size_t es = orc_sizeof_format(in.fmt), n = in.npix();
	out.w = in.w;
	out.h = in.h;
	out.bands = in.bands + 2;
	out.fmt = in.fmt;
	out.type = step == S_BW2sRGB ? 22 : 25;
	out.data.resize(n * out.bands * es);
	for (size_t i = 0; i < n; i++) {
		const uint8_t *p = &in.data[i * in.bands * es];
		uint8_t *q = &out.data[i * out.bands * es];
		for (int k = 0; k < 3; k++)
			memcpy(q + k * es, p, es);
		memcpy(q + 3 * es, p + es, (in.bands - 1) * es);
	}
	return 0;
}

/* One colour op on an image: first 3 bands through the line function, extra
 * bands through colour.c:252-291.
 */
static int
run_step(int step, const Img &in, Img &out)
{
	make_tables();
	if (step == S_sRGB2RGB16 || step == S_RGB162sRGB)
		return shift_cast_step(step, in, out);
	if (step == S_BW2sRGB || step == S_GREY162RGB16)
		return replicate_step(step, in, out);
	int in_fmt_wanted, out_fmt, out_type;
	int out_main = 3; /* bands the converter makes from its three input bands (colour->bands) */
	switch (step) {
	case S_sRGB2HSV: in_fmt_wanted = ORC_FORMAT_UCHAR; out_fmt = ORC_FORMAT_UCHAR; out_type = 29; break;
	case S_HSV2sRGB: in_fmt_wanted = ORC_FORMAT_UCHAR; out_fmt = ORC_FORMAT_UCHAR; out_type = 22; break;
	case S_scRGB2BW: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_UCHAR; out_type = 1; out_main = 1; break;
	case S_scRGB2BW16: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_USHORT; out_type = 26; out_main = 1; break;
	case S_sRGB2scRGB: in_fmt_wanted = ORC_FORMAT_UCHAR; out_fmt = ORC_FORMAT_FLOAT; out_type = 28; break;
	case S_RGB162scRGB: in_fmt_wanted = ORC_FORMAT_USHORT; out_fmt = ORC_FORMAT_FLOAT; out_type = 28; break;
	case S_scRGB2XYZ: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 12; break;
	case S_XYZ2Lab: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 13; break;
	case S_Lab2LabS: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_SHORT; out_type = 21; break;
	case S_LabS2Lab: in_fmt_wanted = ORC_FORMAT_SHORT; out_fmt = ORC_FORMAT_FLOAT; out_type = 13; break;
	case S_Lab2XYZ: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 12; break;
	case S_XYZ2scRGB: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 28; break;
	case S_scRGB2sRGB: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_UCHAR; out_type = 22; break;
	case S_scRGB2RGB16: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_USHORT; out_type = 25; break;
	case S_Lab2LCh: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 19; break;
	case S_LCh2Lab: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 13; break;
	case S_XYZ2Yxy: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 23; break;
	case S_Yxy2XYZ: in_fmt_wanted = ORC_FORMAT_FLOAT; out_fmt = ORC_FORMAT_FLOAT; out_type = 12; break;
	default: return -1;
	}
	if (in.bands < 3)
		return -1;

	/* the op first casts the WHOLE image to its input format
	 * (colour.c vips_colour_code_build / vips_colour_transform_build)
	 */
	const size_t n = in.npix();
	const size_t ies = orc_sizeof_format(in.fmt);
	const size_t wes = orc_sizeof_format(in_fmt_wanted);
	std::vector<uint8_t> cast;
	const uint8_t *src = in.data.data();
	if (in.fmt != in_fmt_wanted) {
		cast.resize(n * in.bands * wes);
		for (size_t i = 0; i < n * in.bands; i++)
			cast_store(elem_as_double(src + i * ies, in.fmt), in.fmt == ORC_FORMAT_FLOAT || in.fmt == ORC_FORMAT_DOUBLE,
				in_fmt_wanted, cast.data() + i * wes);
		src = cast.data();
	}

	/* split off the first three bands */
	std::vector<uint8_t> rgb(n * 3 * wes);
	for (size_t i = 0; i < n; i++)
		memcpy(&rgb[i * 3 * wes], src + i * in.bands * wes, 3 * wes);

	const size_t oes = orc_sizeof_format(out_fmt);
	std::vector<uint8_t> res(n * out_main * oes);
	switch (step) {
	case S_sRGB2HSV: sRGB2HSV_line(rgb.data(), res.data(), n); break;
	case S_HSV2sRGB: HSV2sRGB_line(rgb.data(), res.data(), n); break;
	case S_scRGB2BW: scRGB2BW_line((const float *) rgb.data(), res.data(), 8, n); break;
	case S_scRGB2BW16: scRGB2BW_line((const float *) rgb.data(), res.data(), 16, n); break;
	case S_sRGB2scRGB:
	case S_RGB162scRGB:
		sRGB2scRGB_line(rgb.data(), in_fmt_wanted, (float *) res.data(), n);
		break;
	case S_scRGB2XYZ: scRGB2XYZ_line((const float *) rgb.data(), (float *) res.data(), n); break;
	case S_XYZ2Lab: XYZ2Lab_line((const float *) rgb.data(), (float *) res.data(), n, D65_X0, D65_Y0, D65_Z0); break;
	case S_Lab2LabS: Lab2LabS_line((const float *) rgb.data(), (int16_t *) res.data(), n); break;
	case S_LabS2Lab: LabS2Lab_line((const int16_t *) rgb.data(), (float *) res.data(), n); break;
	case S_Lab2XYZ: Lab2XYZ_line((const float *) rgb.data(), (float *) res.data(), n, D65_X0, D65_Y0, D65_Z0); break;
	case S_XYZ2scRGB: XYZ2scRGB_line((const float *) rgb.data(), (float *) res.data(), n); break;
	case S_scRGB2sRGB: scRGB2sRGB_line((const float *) rgb.data(), res.data(), 8, n); break;
	case S_scRGB2RGB16: scRGB2sRGB_line((const float *) rgb.data(), res.data(), 16, n); break;
	case S_Lab2LCh: Lab2LCh_line((const float *) rgb.data(), (float *) res.data(), n); break;
	case S_LCh2Lab: LCh2Lab_line((const float *) rgb.data(), (float *) res.data(), n); break;
	case S_XYZ2Yxy: XYZ2Yxy_line((const float *) rgb.data(), (float *) res.data(), n); break;
	case S_Yxy2XYZ: Yxy2XYZ_line((const float *) rgb.data(), (float *) res.data(), n); break;
	}

	const int out_bands = in.bands - 3 + out_main;
	out.w = in.w;
	out.h = in.h;
	out.bands = out_bands;
	out.fmt = out_fmt;
	out.type = out_type;
	out.data.resize(n * out_bands * oes);

	const int extra = in.bands - 3;
	const double before = max_alpha_of(in.type);
	const double after = max_alpha_of(out_type);
	const bool rescale = before != after;
	/* vips_linear1 single element: OUT a1 = a[0] with OUT float (double for
	 * double input); q = a1 * (OUT) p + b1
	 */
	const double a_d = after / before;
	for (size_t i = 0; i < n; i++) {
		uint8_t *q = &out.data[i * out_bands * oes];
		memcpy(q, &res[i * out_main * oes], out_main * oes);
		for (int e = 0; e < extra; e++) {
			const uint8_t *p = src + (i * in.bands + 3 + e) * wes;
			double v = elem_as_double(p, in_fmt_wanted);
			bool is_float = in_fmt_wanted == ORC_FORMAT_FLOAT;
			if (rescale) {
				const float a1 = a_d;
				const float b1 = 0.0;
				v = a1 * (float) v + b1;
				is_float = true;
			}
			cast_store(v, is_float, out_fmt, q + (out_main + e) * oes);
		}
	}
	return 0;
}

/* colourspace.c:223-497, the rows among sRGB / scRGB / XYZ / LAB / LABS / RGB16 */
static int
route_for(int from, int to, int steps[8])
{
	enum { BW = 1, XYZ = 12, LAB = 13, LCH = 19, LABS = 21, sRGB = 22, YXY = 23, RGB16 = 25, GREY16 = 26, scRGB = 28, HSV = 29 };
	int n = 0;
	auto push = [&](std::initializer_list<int> l) { for (int s : l) steps[n++] = s; };
	if (from == to)
		return 0;
	/* B_W, GREY16 and HSV: every row of the table that starts there opens with BW2sRGB / GREY162RGB16 / HSV2sRGB and goes
	 * on as the sRGB (RGB16) row does (colourspace.c:386-403, 367-384, 424-441); every row that ends there is the row to
	 * scRGB + scRGB2BW[16], or the row to sRGB + sRGB2HSV (:234-236, 351-353, 372 ...)
	 */
	if (from == BW || from == GREY16 || from == HSV) {
		const int first = from == BW ? S_BW2sRGB : (from == GREY16 ? S_GREY162RGB16 : S_HSV2sRGB);
		const int hub = from == GREY16 ? RGB16 : sRGB;
		int rest[8];
		const int m = to == hub ? 0 : route_for(hub, to, rest);
		if (m < 0 || m > 6)
			return -1;
		steps[0] = first;
		for (int i = 0; i < m; i++)
			steps[1 + i] = rest[i];
		return m + 1;
	}
	if (to == BW || to == GREY16) {
		int m = 0;
		if (from != scRGB) {
			m = route_for(from, scRGB, steps);
			if (m < 0)
				return -1;
		}
		steps[m++] = to == BW ? S_scRGB2BW : S_scRGB2BW16;
		return m;
	}
	if (to == HSV) {
		int m = 0;
		if (from != sRGB) {
			m = route_for(from, sRGB, steps);
			if (m < 0)
				return -1;
		}
		steps[m++] = S_sRGB2HSV;
		return m;
	}
	/* colourspace.c:372, 420: the two rows that are not colour conversions */
	if (from == sRGB && to == RGB16) {
		steps[0] = S_sRGB2RGB16;
		return 1;
	}
	if (from == RGB16 && to == sRGB) {
		steps[0] = S_RGB162sRGB;
		return 1;
	}
	/* LCH hangs off LAB and YXY off XYZ in every row of the table (colourspace.c:226, 236, 242, 252, 275-290 ...):
	 * route to the hub, then one more step
	 */
	if (to == LCH || to == YXY) {
		const int hub = to == LCH ? LAB : XYZ;
		int m = 0;
		if (from != hub) {
			m = route_for(from, hub, steps);
			if (m < 0)
				return -1;
		}
		steps[m++] = to == LCH ? S_Lab2LCh : S_XYZ2Yxy;
		return m;
	}
	/* everything goes up to a hub and down again, exactly as the table rows spell out */
	switch (from) {
	case LCH: push({ S_LCh2Lab }); from = LAB; break;
	case YXY: push({ S_Yxy2XYZ }); from = XYZ; break;
	}
	switch (from) {
	case sRGB: push({ S_sRGB2scRGB }); from = scRGB; break;
	case RGB16: push({ S_RGB162scRGB }); from = scRGB; break;
	case LABS: push({ S_LabS2Lab }); from = LAB; break;
	}
	if (from == to)
		return n;
	if (from == scRGB && (to == XYZ || to == LAB || to == LABS)) {
		push({ S_scRGB2XYZ });
		from = XYZ;
	}
	if (from == LAB && (to == XYZ || to == scRGB || to == sRGB || to == RGB16)) {
		push({ S_Lab2XYZ });
		from = XYZ;
	}
	if (from == to)
		return n;
	if (from == XYZ && (to == LAB || to == LABS)) {
		push({ S_XYZ2Lab });
		from = LAB;
	}
	if (from == XYZ && (to == scRGB || to == sRGB || to == RGB16)) {
		push({ S_XYZ2scRGB });
		from = scRGB;
	}
	if (from == to)
		return n;
	if (from == LAB && to == LABS) {
		push({ S_Lab2LabS });
		return n;
	}
	if (from == scRGB && to == sRGB) {
		push({ S_scRGB2sRGB });
		return n;
	}
	if (from == scRGB && to == RGB16) {
		push({ S_scRGB2RGB16 });
		return n;
	}
	return -1;
}

extern "C" int
orc_colourspace_route(int from, int to, int *steps)
{
	return route_for(from, to, steps);
}

static int
fmt_of_space(int space)
{
	switch (space) {
	case 1:
	case 29:
	case 22: return ORC_FORMAT_UCHAR;
	case 26:
	case 25: return ORC_FORMAT_USHORT;
	case 21: return ORC_FORMAT_SHORT;
	default: return ORC_FORMAT_FLOAT;
	}
}

extern "C" int
orc_colourspace_format(int space)
{
	return fmt_of_space(space);
}

/* bands of vips_colourspace's output: B_W / GREY16 sources gain two bands, B_W / GREY16 targets lose two */
extern "C" int
orc_colourspace_bands(int from, int to, int bands)
{
	const bool grey_from = from == 1 || from == 26, grey_to = to == 1 || to == 26;
	if (from == to)
		return bands;
	return bands + (grey_from ? 2 : 0) - (grey_to ? 2 : 0);
}

/* the format of the result: the space's own, except for the one-step rows B_W -> sRGB and GREY16 -> RGB16, which keep
 * the image's (a bandjoin)
 */
extern "C" int
orc_colourspace_out_format(int from, int to, int fmt)
{
	if ((from == 1 && to == 22) || (from == 26 && to == 25))
		return fmt;
	return fmt_of_space(to);
}

/* vips_colourspace(in, &out, space) with source_space = from.  out must hold
 * w * h * bands elements of orc_colourspace_format(to).
 */
extern "C" int
orc_colourspace(const void *in, int w, int h, int bands, int fmt, int from, int to, void *out)
{
	int steps[8];
	const int n = route_for(from, to, steps);
	if (n < 0)
		return -1;

	Img cur;
	cur.w = w;
	cur.h = h;
	cur.bands = bands;
	cur.fmt = fmt;
	cur.type = from;
	cur.data.assign((const uint8_t *) in, (const uint8_t *) in + (size_t) w * h * bands * orc_sizeof_format(fmt));

	if (n == 0) {
		/* identity routes are a cast to the space's format (colourspace.c:242,260,...) */
		const int ofmt = fmt_of_space(to);
		const size_t cnt = (size_t) w * h * bands;
		for (size_t i = 0; i < cnt; i++)
			cast_store(elem_as_double(cur.data.data() + i * orc_sizeof_format(fmt), fmt),
				fmt == ORC_FORMAT_FLOAT || fmt == ORC_FORMAT_DOUBLE, ofmt,
				(uint8_t *) out + i * orc_sizeof_format(ofmt));
		return 0;
	}
	for (int s = 0; s < n; s++) {
		Img next;
		if (run_step(steps[s], cur, next))
			return -1;
		cur = std::move(next);
	}
	memcpy(out, cur.data.data(), cur.data.size());
	return 0;
}

/* One named step (for pinning each line function against the reference). */
extern "C" int
orc_colour_step(int step, const void *in, int w, int h, int bands, int fmt, int type, void *out)
{
	Img cur, next;
	cur.w = w;
	cur.h = h;
	cur.bands = bands;
	cur.fmt = fmt;
	cur.type = type;
	cur.data.assign((const uint8_t *) in, (const uint8_t *) in + (size_t) w * h * bands * orc_sizeof_format(fmt));
	if (run_step(step, cur, next))
		return -1;
	memcpy(out, next.data.data(), next.data.size());
	return 0;
}
