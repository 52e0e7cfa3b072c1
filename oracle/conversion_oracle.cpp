/* conversion_oracle.cpp -- CPU restatement of vips_flatten.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows conversion/flatten.c:
 *   :421-529  vips_flatten_build: 1 band = copy; max_alpha defaults to the interpretation's (:454-455); integer images
 *             with max_alpha below the format's range go through double and are cast back (:463-470, :519-523); an
 *             all-zero background picks the "black" loops (:479-497); otherwise the background becomes `ink` in the
 *             working format (vips__vector_to_ink, insert.c:244-359: (float) bg, then vips_cast: clip in double, truncate)
 *   :170-225  vips_flatten_black_gen_uchar (float LUT i / max_alpha), :293-354 vips_flatten_gen_uchar (two LUTs)
 *   :88-166   the per-format loops: integer arithmetic in int for char, double for the wider formats
 * Declined (-2), because the reference's own arithmetic converts an out-of-range double to an integer type there
 * (undefined in C): the non-uchar integer loops outside the double detour except ushort with max_alpha 65535.
 */
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <vector>

#include "oracle.h"

static double
interpretation_max_alpha(int type)
{
	/* iofuncs/header.c:195-206: GREY16 (26), RGB16 (25) -> 65535; scRGB (28) -> 1; else 255 */
	return type == 26 || type == 25 ? 65535.0 : (type == 28 ? 1.0 : 255.0);
}

static double
format_max(int fmt)
{
	/* iofuncs/header.c:440-473 */
	switch (fmt) {
	case ORC_FORMAT_UCHAR: return UCHAR_MAX;
	case ORC_FORMAT_CHAR: return SCHAR_MAX;
	case ORC_FORMAT_USHORT: return USHRT_MAX;
	case ORC_FORMAT_SHORT: return SHRT_MAX;
	case ORC_FORMAT_UINT: return UINT_MAX;
	case ORC_FORMAT_INT: return INT_MAX;
	default: return 3.40282346638528859812e+38;
	}
}

#define CLIPD(A, V, B) ((V) < (A) ? (double) (A) : ((V) > (B) ? (double) (B) : (double) (V)))

/* CAST_FLOAT_INT, cast.c:231-238 with :123-131: clip in double, C truncation */
template <typename T> static T cast_from_double(double v);
template <> uint8_t cast_from_double<uint8_t>(double v) { return (uint8_t) CLIPD(0, v, UCHAR_MAX); }
template <> int8_t cast_from_double<int8_t>(double v) { return (int8_t) CLIPD(SCHAR_MIN, v, SCHAR_MAX); }
template <> uint16_t cast_from_double<uint16_t>(double v) { return (uint16_t) CLIPD(0, v, USHRT_MAX); }
template <> int16_t cast_from_double<int16_t>(double v) { return (int16_t) CLIPD(SHRT_MIN, v, SHRT_MAX); }
template <> uint32_t cast_from_double<uint32_t>(double v) { return (uint32_t) CLIPD(0, v, UINT_MAX); }
template <> int32_t cast_from_double<int32_t>(double v) { return (int32_t) CLIPD(INT_MIN, v, INT_MAX); }
template <> float cast_from_double<float>(double v) { return (float) v; }

/* integer image through double: cast up (exact), VIPS_FLATTEN[_BLACK]_FLOAT(double), cast back */
template <typename T>
static void
flatten_wide(const T *p, size_t n, int bands, const double *bg, int nbg, bool black, double max_alpha, T *q)
{
	std::vector<double> ink(bands - 1);
	for (int b = 0; b < bands - 1; b++)
		ink[b] = (double) (float) bg[nbg == 1 ? 0 : b];
	for (size_t i = 0; i < n; i++, p += bands, q += bands - 1) {
		const double alpha = (double) p[bands - 1];
		const double nalpha = max_alpha - alpha;
		for (int b = 0; b < bands - 1; b++) {
			const double v = black ? ((double) p[b] * alpha) / max_alpha
								   : ((double) p[b] * alpha + (double) ink[b] * nalpha) / max_alpha;
			q[b] = cast_from_double<T>(v);
		}
	}
}

/* VIPS_FLATTEN_FLOAT / VIPS_FLATTEN_BLACK_FLOAT(TYPE), flatten.c:108-166, for TYPE = ushort (max_alpha 65535) and float */
template <typename T>
static void
flatten_float_loops(const T *p, size_t n, int bands, const double *bg, int nbg, bool black, double max_alpha, T *q)
{
	std::vector<T> ink(bands - 1);
	for (int b = 0; b < bands - 1; b++)
		ink[b] = cast_from_double<T>((double) (float) bg[nbg == 1 ? 0 : b]);
	for (size_t i = 0; i < n; i++, p += bands, q += bands - 1) {
		const T alpha = p[bands - 1];
		const T nalpha = (T) (max_alpha - alpha);
		for (int b = 0; b < bands - 1; b++)
			if (black)
				q[b] = (T) (((double) p[b] * alpha) / max_alpha);
			else
				q[b] = (T) (((double) p[b] * alpha + (double) ink[b] * nalpha) / max_alpha);
	}
}

/* vips_flatten_black_gen_uchar / vips_flatten_gen_uchar, flatten.c:170-225, 293-354 */
static void
flatten_uchar(const uint8_t *p, size_t n, int bands, const double *bg, int nbg, bool black, double max_alpha, uint8_t *q)
{
	float alpha_lut[256], nalpha_lut[256];
	std::vector<uint8_t> ink(bands - 1);
	for (int i = 0; i < 256; i++) {
		alpha_lut[i] = (float) ((double) i / max_alpha);
		nalpha_lut[i] = (float) ((max_alpha - (double) i) / max_alpha);
	}
	for (int b = 0; b < bands - 1; b++)
		ink[b] = cast_from_double<uint8_t>((double) (float) bg[nbg == 1 ? 0 : b]);
	for (size_t i = 0; i < n; i++, p += bands, q += bands - 1) {
		const float fa = alpha_lut[p[bands - 1]], fn = nalpha_lut[p[bands - 1]];
		for (int b = 0; b < bands - 1; b++)
			if (black)
				q[b] = p[b] * fa;
			else
				q[b] = p[b] * fa + ink[b] * fn;
	}
}

/* 0 ok; -1 bad arguments (what the reference rejects); -2 declined (see the header) */
extern "C" int
orc_flatten(const void *in, int w, int h, int bands, int fmt, int type, const double *bg, int nbg, double max_alpha, void *out)
{
	const size_t n = (size_t) w * h;
	if (bands == 1) {
		memcpy(out, in, n * (fmt == ORC_FORMAT_UCHAR || fmt == ORC_FORMAT_CHAR ? 1 : (fmt == ORC_FORMAT_USHORT || fmt == ORC_FORMAT_SHORT ? 2 : 4)));
		return 0;
	}
	if (max_alpha <= 0)
		max_alpha = interpretation_max_alpha(type);
	bool black = true;
	for (int i = 0; i < nbg; i++)
		if (bg[i] != 0.0)
			black = false;
	if (!black && nbg != 1 && nbg != bands - 1)
		return -1;
	const bool isint = fmt >= ORC_FORMAT_UCHAR && fmt <= ORC_FORMAT_INT;
	if (isint && max_alpha < format_max(fmt)) {
		switch (fmt) {
		case ORC_FORMAT_UCHAR: flatten_wide((const uint8_t *) in, n, bands, bg, nbg, black, max_alpha, (uint8_t *) out); break;
		case ORC_FORMAT_CHAR: flatten_wide((const int8_t *) in, n, bands, bg, nbg, black, max_alpha, (int8_t *) out); break;
		case ORC_FORMAT_USHORT: flatten_wide((const uint16_t *) in, n, bands, bg, nbg, black, max_alpha, (uint16_t *) out); break;
		case ORC_FORMAT_SHORT: flatten_wide((const int16_t *) in, n, bands, bg, nbg, black, max_alpha, (int16_t *) out); break;
		case ORC_FORMAT_UINT: flatten_wide((const uint32_t *) in, n, bands, bg, nbg, black, max_alpha, (uint32_t *) out); break;
		default: flatten_wide((const int32_t *) in, n, bands, bg, nbg, black, max_alpha, (int32_t *) out); break;
		}
		return 0;
	}
	if (fmt == ORC_FORMAT_UCHAR)
		flatten_uchar((const uint8_t *) in, n, bands, bg, nbg, black, max_alpha, (uint8_t *) out);
	else if (fmt == ORC_FORMAT_USHORT && max_alpha == 65535.0)
		flatten_float_loops((const uint16_t *) in, n, bands, bg, nbg, black, max_alpha, (uint16_t *) out);
	else if (fmt == ORC_FORMAT_FLOAT)
		flatten_float_loops((const float *) in, n, bands, bg, nbg, black, max_alpha, (float *) out);
	else
		return -2;
	return 0;
}
