/* resample_oracle.cpp -- CPU restatement of the reference's resample hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  Build with -O2 -ffp-contract=off so
 * the float paths evaluate exactly as a baseline x86-64 libvips build would.
 *
 * Follows (reference = libvips 8.19, paths under libvips/):
 *   resample/templates.h:346-578   filters, calculate_coefficients, reduce_sum, LongT, *_fixed_round
 *   resample/reduceh.cpp:112-142   vips_reduce_get_points
 *   resample/reducev.cpp:420-619   reducev_block/line, vips_reducev_gen
 *   resample/reducev.cpp:859-981   vips_reducev_build (geometry, gap, masks, embed)
 *   resample/reduceh.cpp:145-336   reduceh_*_tab, vips_reduceh_gen
 *   resample/reduceh.cpp:396-521   vips_reduceh_build
 *   resample/shrinkv.c:158-388,474-565   ADD / *AVG macros, vips_shrinkv_gen, build
 *   resample/shrinkh.c:78-286,357-420    *SHRINK macros, vips_shrinkh_gen, build
 *   conversion/premultiply.c:86-260, conversion/unpremultiply.c:85-324
 *   conversion/embed.c:300-336     EXTEND_COPY == clamp addressing
 *   resample/resize.c:135-231, resample/thumbnail.c:413-466,848-902
 *   iofuncs/thread.c:288-325       tile geometry from demand hints
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"

#define TRANSFORM_SHIFT 6
#define TRANSFORM_SCALE (1 << TRANSFORM_SHIFT)
#define INTERPOLATE_SHIFT 12
#define INTERPOLATE_SCALE (1 << INTERPOLATE_SHIFT)
#define ORC_PI 3.14159265358979323846 /* VIPS_PI, include/vips/util.h */
#define ROUND_UINT(R) ((int) ((R) + 0.5))
#define MAX_POINT 2000

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

extern "C" size_t
orc_sizeof_format(int fmt)
{
	switch (fmt) {
	case ORC_FORMAT_UCHAR:
	case ORC_FORMAT_CHAR:
		return 1;
	case ORC_FORMAT_USHORT:
	case ORC_FORMAT_SHORT:
		return 2;
	case ORC_FORMAT_UINT:
	case ORC_FORMAT_INT:
	case ORC_FORMAT_FLOAT:
		return 4;
	case ORC_FORMAT_DOUBLE:
		return 8;
	}
	return 0;
}

/* ---------------------------------------------------------------- masks */

/* templates.h:346-354 */
static double
sinc_filter(double x)
{
	if (x == 0.0)
		return 1.0;
	x = x * ORC_PI;
	return sin(x) / x;
}

/* templates.h:321-344 */
static double
cubic_filter(double x, double B, double C)
{
	const double ax = fabs(x);
	const double ax2 = ax * ax;
	const double ax3 = ax2 * ax;

	if (ax <= 1)
		return ((12 - 9 * B - 6 * C) * ax3 + (-18 + 12 * B + 6 * C) * ax2 + (6 - 2 * B)) / 6;
	if (ax <= 2)
		return ((-B - 6 * C) * ax3 + (6 * B + 30 * C) * ax2 + (-12 * B - 48 * C) * ax + (8 * B + 24 * C)) / 6;
	return 0.0;
}

/* templates.h:358-448 */
static double
kernel_filter(int kernel, double x)
{
	switch (kernel) {
	case ORC_KERNEL_LINEAR:
		x = fabs(x);
		return x < 1.0 ? 1.0 - x : 0.0;
	case ORC_KERNEL_CUBIC:
		return cubic_filter(x, 0.0, 0.5);
	case ORC_KERNEL_MITCHELL:
		return cubic_filter(x, 1.0 / 3.0, 1.0 / 3.0);
	case ORC_KERNEL_LANCZOS2:
		if (x >= -2 && x <= 2)
			return sinc_filter(x) * sinc_filter(x / 2);
		return 0.0;
	case ORC_KERNEL_LANCZOS3:
		if (x >= -3 && x <= 3)
			return sinc_filter(x) * sinc_filter(x / 3);
		return 0.0;
	case ORC_KERNEL_MKS2013:
		x = fabs(x);
		if (x >= 2.5)
			return 0.0;
		if (x >= 1.5)
			return (x - 5.0 / 2.0) * (x - 5.0 / 2.0) / -8.0;
		if (x >= 0.5)
			return (4.0 * x * x - 11.0 * x + 7.0) / 4.0;
		return 17.0 / 16.0 - 7.0 * x * x / 4.0;
	case ORC_KERNEL_MKS2021:
		x = fabs(x);
		if (x >= 4.5)
			return 0.0;
		if (x >= 3.5)
			return (4.0 * x * x - 36.0 * x + 81.0) / -1152.0;
		if (x >= 2.5)
			return (4.0 * x * x - 27.0 * x + 45.0) / 144.0;
		if (x >= 1.5)
			return (24.0 * x * x - 113.0 * x + 130.0) / -144.0;
		if (x >= 0.5)
			return (140.0 * x * x - 379.0 * x + 239.0) / 144.0;
		return 577.0 / 576.0 - 239.0 * x * x / 144.0;
	}
	return 0.0;
}

/* templates.h:457-526, T = double or long double */
template <typename T>
static void
make_mask(T *c, int kernel, int n_points, double shrink, double x)
{
	if (kernel == ORC_KERNEL_NEAREST) {
		c[0] = 1.0;
		return;
	}

	const double half = x + n_points / 2.0 - 1;
	const double scale = 1.0 / shrink;
	T sum = 0.0;

	for (int i = 0; i < n_points; i++) {
		const double xp = (i - half) * scale;
		double l = kernel_filter(kernel, xp);

		c[i] = l;
		sum += l;
	}
	for (int i = 0; i < n_points; i++)
		c[i] /= sum;
}

extern "C" int
orc_reduce_get_points(int kernel, double shrink)
{
	/* reduceh.cpp:112-142 */
	switch (kernel) {
	case ORC_KERNEL_NEAREST:
		return 1;
	case ORC_KERNEL_LINEAR:
		return 2 * rint(shrink) + 1;
	case ORC_KERNEL_CUBIC:
	case ORC_KERNEL_MITCHELL:
	case ORC_KERNEL_LANCZOS2:
		return 2 * rint(2 * shrink) + 1;
	case ORC_KERNEL_LANCZOS3:
	case ORC_KERNEL_MKS2013:
		return 2 * rint(3 * shrink) + 1;
	case ORC_KERNEL_MKS2021:
		return 2 * rint(5 * shrink) + 1;
	}
	return 0;
}

extern "C" void
orc_reduce_make_mask(double *c, int kernel, int n_point, double shrink, double x)
{
	make_mask<double>(c, kernel, n_point, shrink, x);
}

extern "C" void
orc_reduce_tables(int kernel, int n_point, double residual, double *matrixf, short *matrixs)
{
	/* reducev.cpp:945-958 */
	for (int y = 0; y < TRANSFORM_SCALE + 1; y++) {
		double *f = matrixf + (size_t) y * n_point;
		short *s = matrixs + (size_t) y * n_point;

		make_mask<double>(f, kernel, n_point, residual, (float) y / TRANSFORM_SCALE);
		for (int i = 0; i < n_point; i++)
			s[i] = (short) (f[i] * INTERPOLATE_SCALE);
	}
}

extern "C" int
orc_reduce_geometry(int in_size, double shrink, int kernel, double gap, OrcReduceGeom *g)
{
	/* reducev.cpp:877-941 == reduceh.cpp:416-481 */
	if (shrink < 1.0)
		return -1;

	g->in_size = in_size;
	g->out_size = ROUND_UINT((double) in_size / shrink);
	double extra_pixels = g->out_size * shrink - in_size;
	g->residual = shrink;
	g->int_shrink = 1;
	g->shrunk_size = in_size;

	if (gap > 0.0 && kernel != ORC_KERNEL_NEAREST) {
		if (gap < 1.0)
			return -1;
		int int_shrink = std::max(1.0, floor((double) in_size / g->out_size / gap));
		if (int_shrink > 1) {
			g->int_shrink = int_shrink;
			g->shrunk_size = orc_shrink_size(in_size, int_shrink, 1);
			extra_pixels /= int_shrink;
			g->residual /= int_shrink;
		}
	}

	if (g->out_size <= 0)
		return -1;

	if (g->residual == 1.0) {
		g->n_point = 0;
		g->offset = 0;
		/* the reference just copies the (box-shrunk) image: its size wins */
		g->out_size = g->shrunk_size;
		return 0;
	}

	g->n_point = orc_reduce_get_points(kernel, g->residual);
	if (g->n_point > MAX_POINT)
		return -1;
	g->offset = (1 + extra_pixels) / 2.0 - 1;

	return 0;
}

/* ------------------------------------------------- fixed-point helpers */

/* templates.h:150-157, 203-210 */
template <typename IT>
static inline IT unsigned_fixed_round(IT v)
{
	const int round_by = INTERPOLATE_SCALE >> 1;
	return (v + round_by) >> INTERPOLATE_SHIFT;
}
template <typename IT>
static inline IT signed_fixed_round(IT v)
{
	const int sign_of_v = 2 * (v >= 0) - 1;
	const int round_by = sign_of_v * (INTERPOLATE_SCALE >> 1);
	return (v + round_by) >> INTERPOLATE_SHIFT;
}

template <typename T> struct LongT { typedef int32_t type; };
template <> struct LongT<int32_t> { typedef int64_t type; };
template <> struct LongT<uint32_t> { typedef int64_t type; };
template <> struct LongT<float> { typedef double type; };
template <> struct LongT<double> { typedef long double type; };

template <typename T> struct Lim;
template <> struct Lim<uint8_t> { static constexpr int64_t lo = 0, hi = UCHAR_MAX; static constexpr bool sgn = false; };
template <> struct Lim<int8_t> { static constexpr int64_t lo = SCHAR_MIN, hi = SCHAR_MAX; static constexpr bool sgn = true; };
template <> struct Lim<uint16_t> { static constexpr int64_t lo = 0, hi = USHRT_MAX; static constexpr bool sgn = false; };
template <> struct Lim<int16_t> { static constexpr int64_t lo = SHRT_MIN, hi = SHRT_MAX; static constexpr bool sgn = true; };
template <> struct Lim<uint32_t> { static constexpr int64_t lo = 0, hi = UINT_MAX; static constexpr bool sgn = false; };
template <> struct Lim<int32_t> { static constexpr int64_t lo = INT_MIN, hi = INT_MAX; static constexpr bool sgn = true; };

/* finalize of reduce{v,h}_{unsigned,signed}_int_tab */
template <typename T>
static inline T
int_finalize(typename LongT<T>::type sum)
{
	typedef typename LongT<T>::type IT;
	IT v = Lim<T>::sgn ? signed_fixed_round<IT>(sum) : unsigned_fixed_round<IT>(sum);
	int64_t w = v;
	w = std::max<int64_t>(Lim<T>::lo, std::min<int64_t>(w, Lim<T>::hi));
	return (T) w;
}

/* ------------------------------------------------------- reducev pass */

struct Tables {
	std::vector<double> f;
	std::vector<short> s;
	int n;
	const double *mf(int t) const { return f.data() + (size_t) t * n; }
	const short *ms(int t) const { return s.data() + (size_t) t * n; }
};

static void
build_tables(Tables &t, int kernel, int n_point, double residual)
{
	t.n = n_point;
	t.f.resize((size_t) (TRANSFORM_SCALE + 1) * n_point);
	t.s.resize((size_t) (TRANSFORM_SCALE + 1) * n_point);
	orc_reduce_tables(kernel, n_point, residual, t.f.data(), t.s.data());
}

/* Integer formats: reducev.cpp:461-484 via reducev_block (tap loop outer). */
template <typename T>
static void
reducev_int(const T *in, int w, int h, int bands, int out_h, double rs, double voffset, int n, const Tables &tab,
	int rect_h, T *out)
{
	typedef typename LongT<T>::type IT;
	const int ne = w * bands;
	const int emb = (int) ceil(n / 2.0) - 1; /* reducev.cpp:976 */
	std::vector<IT> sum(ne);

	if (rect_h <= 0)
		rect_h = out_h;
	for (int top = 0; top < out_h; top += rect_h) {
		const int height = std::min(rect_h, out_h - top);
		double Y = (top + 0.5) * rs - 0.5 - voffset;

		for (int y = 0; y < height; y++) {
			const int py = (int) Y;
			const int sy = Y * TRANSFORM_SCALE * 2;
			const int siy = sy & (TRANSFORM_SCALE * 2 - 1);
			const int ty = (siy + 1) >> 1;
			const short *cys = tab.ms(ty);
			T *q = out + (size_t) (top + y) * ne;

			std::fill(sum.begin(), sum.end(), (IT) 0);
			for (int i = 0; i < n; i++) {
				const IT c = cys[i];
				const T *p = in + (size_t) clampi(py + i - emb, 0, h - 1) * ne;
				for (int k = 0; k < ne; k++)
					sum[k] += c * p[k];
			}
			for (int k = 0; k < ne; k++)
				q[k] = int_finalize<T>(sum[k]);

			Y += rs;
		}
	}
}

/* float: reducev.cpp:487-496; double: reducev_notab :500-514 (long double mask per row) */
template <typename T>
static void
reducev_fp(const T *in, int w, int h, int bands, int out_h, double rs, double voffset, int n, int kernel,
	const Tables &tab, int rect_h, T *out)
{
	typedef typename LongT<T>::type IT;
	const int ne = w * bands;
	const int emb = (int) ceil(n / 2.0) - 1;
	std::vector<IT> sum(ne);
	std::vector<IT> cy(n);

	if (rect_h <= 0)
		rect_h = out_h;
	for (int top = 0; top < out_h; top += rect_h) {
		const int height = std::min(rect_h, out_h - top);
		double Y = (top + 0.5) * rs - 0.5 - voffset;

		for (int y = 0; y < height; y++) {
			const int py = (int) Y;
			const int sy = Y * TRANSFORM_SCALE * 2;
			const int siy = sy & (TRANSFORM_SCALE * 2 - 1);
			const int ty = (siy + 1) >> 1;
			T *q = out + (size_t) (top + y) * ne;

			if (sizeof(T) == sizeof(double))
				make_mask<IT>(cy.data(), kernel, n, rs, Y - py);
			else
				for (int i = 0; i < n; i++)
					cy[i] = tab.mf(ty)[i];

			std::fill(sum.begin(), sum.end(), (IT) 0);
			for (int i = 0; i < n; i++) {
				const IT c = cy[i];
				const T *p = in + (size_t) clampi(py + i - emb, 0, h - 1) * ne;
				for (int k = 0; k < ne; k++)
					sum[k] += c * p[k];
			}
			for (int k = 0; k < ne; k++)
				q[k] = sum[k];

			Y += rs;
		}
	}
}

extern "C" int
orc_reducev_pass(const void *in, int w, int h, int bands, int fmt, int out_h, double residual, double voffset,
	int n_point, int kernel, int rect_h, void *out)
{
	Tables tab;
	build_tables(tab, kernel, n_point, residual);

#define RV_INT(T) reducev_int<T>((const T *) in, w, h, bands, out_h, residual, voffset, n_point, tab, rect_h, (T *) out)
#define RV_FP(T) reducev_fp<T>((const T *) in, w, h, bands, out_h, residual, voffset, n_point, kernel, tab, rect_h, (T *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: RV_INT(uint8_t); break;
	case ORC_FORMAT_CHAR: RV_INT(int8_t); break;
	case ORC_FORMAT_USHORT: RV_INT(uint16_t); break;
	case ORC_FORMAT_SHORT: RV_INT(int16_t); break;
	case ORC_FORMAT_UINT: RV_INT(uint32_t); break;
	case ORC_FORMAT_INT: RV_INT(int32_t); break;
	case ORC_FORMAT_FLOAT: RV_FP(float); break;
	case ORC_FORMAT_DOUBLE: RV_FP(double); break;
	default: return -1;
	}
	return 0;
}

/* ------------------------------------------------------- reduceh pass */

template <typename T, bool FP>
static void
reduceh_any(const T *in, int w, int h, int bands, int out_w, double rs, double hoffset, int n, int kernel,
	const Tables &tab, int rect_w, T *out)
{
	typedef typename LongT<T>::type IT;
	const int emb = (int) ceil(n / 2.0) - 1; /* reduceh.cpp:516 */
	std::vector<IT> cx(n);

	if (rect_w <= 0)
		rect_w = out_w;
	for (int y = 0; y < h; y++) {
		const T *p0 = in + (size_t) y * w * bands;
		T *q = out + (size_t) y * out_w * bands;

		for (int left = 0; left < out_w; left += rect_w) {
			const int width = std::min(rect_w, out_w - left);
			double X = (left + 0.5) * rs - 0.5 - hoffset;

			for (int x = 0; x < width; x++) {
				const int ix = (int) X;
				const int sx = X * TRANSFORM_SCALE * 2;
				const int six = sx & (TRANSFORM_SCALE * 2 - 1);
				const int tx = (six + 1) >> 1;

				if (FP) {
					if (sizeof(T) == sizeof(double))
						make_mask<IT>(cx.data(), kernel, n, rs, X - ix);
					else
						for (int i = 0; i < n; i++)
							cx[i] = tab.mf(tx)[i];
				}
				else
					for (int i = 0; i < n; i++)
						cx[i] = tab.ms(tx)[i];

				for (int z = 0; z < bands; z++) {
					IT sum = 0; /* reduce_sum, templates.h:565-578 */
					for (int i = 0; i < n; i++) {
						const int sxp = clampi(ix + i - emb, 0, w - 1);
						sum += (IT) cx[i] * p0[(size_t) sxp * bands + z];
					}
					if constexpr (FP)
						q[(size_t) (left + x) * bands + z] = sum;
					else
						q[(size_t) (left + x) * bands + z] = int_finalize<T>(sum);
				}

				X += rs;
			}
		}
	}
}

extern "C" int
orc_reduceh_pass(const void *in, int w, int h, int bands, int fmt, int out_w, double residual, double hoffset,
	int n_point, int kernel, int rect_w, void *out)
{
	Tables tab;
	build_tables(tab, kernel, n_point, residual);

#define RH(T, FP) reduceh_any<T, FP>((const T *) in, w, h, bands, out_w, residual, hoffset, n_point, kernel, tab, rect_w, (T *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: RH(uint8_t, false); break;
	case ORC_FORMAT_CHAR: RH(int8_t, false); break;
	case ORC_FORMAT_USHORT: RH(uint16_t, false); break;
	case ORC_FORMAT_SHORT: RH(int16_t, false); break;
	case ORC_FORMAT_UINT: RH(uint32_t, false); break;
	case ORC_FORMAT_INT: RH(int32_t, false); break;
	case ORC_FORMAT_FLOAT: RH(float, true); break;
	case ORC_FORMAT_DOUBLE: RH(double, true); break;
	default: return -1;
	}
	return 0;
}

/* ------------------------------------------------------------ shrinkv */

extern "C" int
orc_shrink_size(int in_size, int shrink, int ceil_mode)
{
	/* shrinkv.c:561-563, shrinkh.c:412-414 */
	return ceil_mode ? (int) ceil((double) in_size / shrink) : ROUND_UINT((double) in_size / shrink);
}

template <typename T, typename ACC>
static void
shrinkv_t(const T *in, int w, int h, int bands, int vshrink, int out_h, int fmt, T *out)
{
	const int sz = w * bands;
	std::vector<ACC> sum(sz);
	const int amend = vshrink / 2;

	for (int y = 0; y < out_h; y++) {
		std::fill(sum.begin(), sum.end(), (ACC) 0);
		for (int k = 0; k < vshrink; k++) {
			/* embed to ROUND_UP(h, vshrink) rows with EXTEND_COPY, shrinkv.c:501 */
			const T *p = in + (size_t) clampi(y * vshrink + k, 0, h - 1) * sz;
			for (int x = 0; x < sz; x++)
				sum[x] += p[x];
		}
		T *q = out + (size_t) y * sz;
		switch (fmt) {
		case ORC_FORMAT_UCHAR: {
			/* UCHAR_AVG, shrinkv.c:218-227: int + int, times unsigned, >> 24, byte store */
			unsigned int multiplier = (1LL << 32) / ((1 << 8) * vshrink);
			for (int x = 0; x < sz; x++)
				q[x] = (T) ((((int) sum[x] + amend) * multiplier) >> 24);
			break;
		}
		case ORC_FORMAT_USHORT: {
			/* USHORT_AVG, shrinkv.c:232-242 */
			uint64_t multiplier = ((1ULL << 32) + vshrink - 1) / vshrink;
			for (int x = 0; x < sz; x++)
				q[x] = (T) (((int64_t) ((int) sum[x] + amend) * multiplier) >> 32);
			break;
		}
		case ORC_FORMAT_FLOAT:
		case ORC_FORMAT_DOUBLE: {
			/* FAVG */
			const double inv_vshrink = 1.0 / vshrink;
			for (int x = 0; x < sz; x++)
				q[x] = (T) ((double) sum[x] * inv_vshrink);
			break;
		}
		default:
			/* IAVG: C truncating division */
			for (int x = 0; x < sz; x++)
				q[x] = (T) ((sum[x] + (ACC) amend) / (ACC) vshrink);
			break;
		}
	}
}

extern "C" int
orc_shrinkv(const void *in, int w, int h, int bands, int fmt, int vshrink, int ceil_mode, void *out)
{
	if (vshrink < 1)
		return -1;
	const int out_h = vshrink == 1 ? h : orc_shrink_size(h, vshrink, ceil_mode);
	if (vshrink == 1) {
		memcpy(out, in, (size_t) w * h * bands * orc_sizeof_format(fmt));
		return 0;
	}
#define SV(T, ACC) shrinkv_t<T, ACC>((const T *) in, w, h, bands, vshrink, out_h, fmt, (T *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: SV(uint8_t, int); break;
	case ORC_FORMAT_CHAR: SV(int8_t, int); break;
	case ORC_FORMAT_USHORT: SV(uint16_t, int); break;
	case ORC_FORMAT_SHORT: SV(int16_t, int); break;
	case ORC_FORMAT_UINT: SV(uint32_t, int64_t); break;
	case ORC_FORMAT_INT: SV(int32_t, int64_t); break;
	case ORC_FORMAT_FLOAT: SV(float, double); break;
	case ORC_FORMAT_DOUBLE: SV(double, double); break;
	default: return -1;
	}
	return 0;
}

/* ------------------------------------------------------------ shrinkh */

template <typename T, typename ACC>
static void
shrinkh_t(const T *in, int w, int h, int bands, int hshrink, int out_w, int fmt, T *out)
{
	const int amend = hshrink / 2;
	const unsigned int multiplier = (1LL << 32) / ((1 << 8) * hshrink);
	const uint64_t ushort_multiplier = ((1ULL << 32) + hshrink - 1) / hshrink;
	const double inv_hshrink = 1.0 / hshrink;

	for (int y = 0; y < h; y++) {
		const T *p = in + (size_t) y * w * bands;
		T *q = out + (size_t) y * out_w * bands;

		for (int x = 0; x < out_w; x++)
			for (int b = 0; b < bands; b++) {
				/* embed to w + hshrink columns with EXTEND_COPY, shrinkh.c:383 */
				switch (fmt) {
				case ORC_FORMAT_UCHAR: {
					int sum = amend;
					for (int k = 0; k < hshrink; k++)
						sum += p[(size_t) clampi(x * hshrink + k, 0, w - 1) * bands + b];
					q[(size_t) x * bands + b] = (T) ((sum * multiplier) >> 24);
					break;
				}
				case ORC_FORMAT_USHORT: {
					int sum = amend;
					for (int k = 0; k < hshrink; k++)
						sum += p[(size_t) clampi(x * hshrink + k, 0, w - 1) * bands + b];
					q[(size_t) x * bands + b] = (T) (((int64_t) sum * ushort_multiplier) >> 32);
					break;
				}
				case ORC_FORMAT_FLOAT:
				case ORC_FORMAT_DOUBLE: {
					double sum = 0.0;
					for (int k = 0; k < hshrink; k++)
						sum += p[(size_t) clampi(x * hshrink + k, 0, w - 1) * bands + b];
					q[(size_t) x * bands + b] = (T) (sum * inv_hshrink);
					break;
				}
				default: {
					ACC sum = amend;
					for (int k = 0; k < hshrink; k++)
						sum += p[(size_t) clampi(x * hshrink + k, 0, w - 1) * bands + b];
					q[(size_t) x * bands + b] = (T) (sum / (ACC) hshrink);
					break;
				}
				}
			}
	}
}

extern "C" int
orc_shrinkh(const void *in, int w, int h, int bands, int fmt, int hshrink, int ceil_mode, void *out)
{
	if (hshrink < 1)
		return -1;
	if (hshrink == 1) {
		memcpy(out, in, (size_t) w * h * bands * orc_sizeof_format(fmt));
		return 0;
	}
	const int out_w = orc_shrink_size(w, hshrink, ceil_mode);
#define SH(T, ACC) shrinkh_t<T, ACC>((const T *) in, w, h, bands, hshrink, out_w, fmt, (T *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: SH(uint8_t, int); break;
	case ORC_FORMAT_CHAR: SH(int8_t, int); break;
	case ORC_FORMAT_USHORT: SH(uint16_t, int); break;
	case ORC_FORMAT_SHORT: SH(int16_t, int); break;
	case ORC_FORMAT_UINT: SH(uint32_t, int64_t); break;
	case ORC_FORMAT_INT: SH(int32_t, int64_t); break;
	case ORC_FORMAT_FLOAT: SH(float, double); break;
	case ORC_FORMAT_DOUBLE: SH(double, double); break;
	default: return -1;
	}
	return 0;
}

/* ------------------------------------------- vips_reducev / vips_reduceh */

extern "C" int
orc_reducev(const void *in, int w, int h, int bands, int fmt, double vshrink, int kernel, double gap, int rect_h,
	void *out)
{
	OrcReduceGeom g;
	if (orc_reduce_geometry(h, vshrink, kernel, gap, &g))
		return -1;
	const size_t es = orc_sizeof_format(fmt);
	std::vector<uint8_t> tmp;
	const void *src = in;
	int sh = h;
	if (g.int_shrink > 1) {
		tmp.resize((size_t) w * g.shrunk_size * bands * es);
		if (orc_shrinkv(in, w, h, bands, fmt, g.int_shrink, 1, tmp.data()))
			return -1;
		src = tmp.data();
		sh = g.shrunk_size;
	}
	if (g.n_point == 0) {
		memcpy(out, src, (size_t) w * sh * bands * es);
		return 0;
	}
	return orc_reducev_pass(src, w, sh, bands, fmt, g.out_size, g.residual, g.offset, g.n_point, kernel, rect_h, out);
}

extern "C" int
orc_reduceh(const void *in, int w, int h, int bands, int fmt, double hshrink, int kernel, double gap, int rect_w,
	void *out)
{
	OrcReduceGeom g;
	if (orc_reduce_geometry(w, hshrink, kernel, gap, &g))
		return -1;
	const size_t es = orc_sizeof_format(fmt);
	std::vector<uint8_t> tmp;
	const void *src = in;
	int sw = w;
	if (g.int_shrink > 1) {
		tmp.resize((size_t) g.shrunk_size * h * bands * es);
		if (orc_shrinkh(in, w, h, bands, fmt, g.int_shrink, 1, tmp.data()))
			return -1;
		src = tmp.data();
		sw = g.shrunk_size;
	}
	if (g.n_point == 0) {
		memcpy(out, src, (size_t) sw * h * bands * es);
		return 0;
	}
	return orc_reduceh_pass(src, sw, h, bands, fmt, g.out_size, g.residual, g.offset, g.n_point, kernel, rect_w, out);
}

/* ------------------------------------------- premultiply / unpremultiply */

template <typename IN, typename OUT>
static void
premultiply_t(const IN *p, size_t npix, int bands, double max_alpha, OUT *q)
{
	/* PRE_MANY / PRE_RGBA, premultiply.c:86-122 */
	for (size_t x = 0; x < npix; x++) {
		IN alpha = p[bands - 1];
		IN clip_alpha = std::max<double>(0, std::min<double>(max_alpha, alpha));
		OUT nalpha = (OUT) clip_alpha / max_alpha;
		int i;
		for (i = 0; i < bands - 1; i++)
			q[i] = p[i] * nalpha;
		q[i] = alpha;
		p += bands;
		q += bands;
	}
}

extern "C" int
orc_premultiply(const void *in, int w, int h, int bands, int fmt, double max_alpha, int uchar_mode, void *out)
{
	const size_t npix = (size_t) w * h;

	if (bands == 1) {
		memcpy(out, in, npix * orc_sizeof_format(fmt));
		return 0;
	}
	if (uchar_mode && fmt == ORC_FORMAT_UCHAR) {
		/* premultiply.c:152-166, LUT :253-259 */
		int scale[256];
		for (int i = 0; i < 256; i++) {
			double clip = std::max<double>(0, std::min<double>(max_alpha, i));
			scale[i] = 256 * clip / max_alpha;
		}
		const uint8_t *p = (const uint8_t *) in;
		uint8_t *q = (uint8_t *) out;
		for (size_t x = 0; x < npix; x++) {
			uint8_t alpha = p[bands - 1];
			int s = scale[alpha];
			int i;
			for (i = 0; i < bands - 1; i++)
				q[i] = (p[i] * s + 128) >> 8;
			q[i] = alpha;
			p += bands;
			q += bands;
		}
		return 0;
	}
#define PM(IN, OUT) premultiply_t<IN, OUT>((const IN *) in, npix, bands, max_alpha, (OUT *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: PM(uint8_t, float); break;
	case ORC_FORMAT_CHAR: PM(int8_t, float); break;
	case ORC_FORMAT_USHORT: PM(uint16_t, float); break;
	case ORC_FORMAT_SHORT: PM(int16_t, float); break;
	case ORC_FORMAT_UINT: PM(uint32_t, float); break;
	case ORC_FORMAT_INT: PM(int32_t, float); break;
	case ORC_FORMAT_FLOAT: PM(float, float); break;
	case ORC_FORMAT_DOUBLE: PM(double, double); break;
	default: return -1;
	}
	return 0;
}

template <typename IN, typename OUT, bool FP>
static void
unpremultiply_t(const IN *p, size_t npix, int bands, double max_alpha, OUT *q)
{
	/* UNPRE_* / FUNPRE_*, unpremultiply.c:85-183; alpha_band = bands - 1 */
	const int alpha_band = bands - 1;
	for (size_t x = 0; x < npix; x++) {
		IN alpha = p[alpha_band];
		OUT factor;
		if (FP)
			factor = fabs((double) alpha) < 0.01 ? 0 : max_alpha / alpha;
		else
			factor = alpha == 0 ? 0 : max_alpha / alpha;
		for (int i = 0; i < alpha_band; i++)
			q[i] = factor * p[i];
		q[alpha_band] = std::max<double>(0, std::min<double>(max_alpha, alpha));
		p += bands;
		q += bands;
	}
}

extern "C" int
orc_unpremultiply(const void *in, int w, int h, int bands, int fmt, double max_alpha, int uchar_mode, void *out)
{
	const size_t npix = (size_t) w * h;

	if (bands == 1) {
		memcpy(out, in, npix * orc_sizeof_format(fmt));
		return 0;
	}
	if (uchar_mode && fmt == ORC_FORMAT_UCHAR) {
		/* unpremultiply.c:209-222, LUT :313-324; byte store is unclipped */
		int scale[256];
		for (int i = 0; i < 256; i++) {
			double clip = std::max<double>(0, std::min<double>(max_alpha, i));
			scale[i] = clip == 0 ? 0 : 256 * max_alpha / clip;
		}
		const uint8_t *p = (const uint8_t *) in;
		uint8_t *q = (uint8_t *) out;
		for (size_t x = 0; x < npix; x++) {
			uint8_t alpha = p[bands - 1];
			int s = scale[alpha];
			int i;
			for (i = 0; i < bands - 1; i++)
				q[i] = (p[i] * s + 128) >> 8;
			q[i] = alpha;
			p += bands;
			q += bands;
		}
		return 0;
	}
#define UPM(IN, OUT, FP) unpremultiply_t<IN, OUT, FP>((const IN *) in, npix, bands, max_alpha, (OUT *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: UPM(uint8_t, float, false); break;
	case ORC_FORMAT_CHAR: UPM(int8_t, float, false); break;
	case ORC_FORMAT_USHORT: UPM(uint16_t, float, false); break;
	case ORC_FORMAT_SHORT: UPM(int16_t, float, false); break;
	case ORC_FORMAT_UINT: UPM(uint32_t, float, false); break;
	case ORC_FORMAT_INT: UPM(int32_t, float, false); break;
	case ORC_FORMAT_FLOAT: UPM(float, float, true); break;
	case ORC_FORMAT_DOUBLE: UPM(double, double, true); break;
	default: return -1;
	}
	return 0;
}

/* ---------------------------------------------------------------- resize */

/* The shrinks vips_resize hands to reducev / reduceh: resize.c:150-231. */
static void
resize_shrinks(int w, int h, double hscale, double vscale, double *hshrink, double *vshrink)
{
	hscale = std::max(hscale, 1.0 / w);
	vscale = std::max(vscale, 1.0 / h);
	*vshrink = vscale < 1.0 ? 1.0 / vscale : 1.0;
	*hshrink = hscale < 1.0 ? 1.0 / hscale : 1.0;
}

/* vips_resize with VIPS_KERNEL_NEAREST first drops whole pixels (resize.c:167-205): the integer part of the shrink, over
 * `gap`, goes to vips_subsample -- out(x, y) = in(x * xfac, y * yfac), size in / fac rounded DOWN (subsample.c:83-190,
 * 218-222) -- and the scales are multiplied up for the residual reducev / reduceh.
 */
static void
nearest_subsample(int w, int h, double hscale, double vscale, double gap, int *xfac, int *yfac)
{
	int int_hshrink, int_vshrink;
	if (gap < 1.0) {
		int_hshrink = (int) floor(1.0 / hscale);
		int_vshrink = (int) floor(1.0 / vscale);
	}
	else {
		const int target_width = (int) (w * hscale + 0.5);	 /* VIPS_ROUND_UINT */
		const int target_height = (int) (h * vscale + 0.5);
		int_hshrink = (int) floor((double) w / target_width / gap);
		int_vshrink = (int) floor((double) h / target_height / gap);
	}
	*xfac = std::max(1, int_hshrink);
	*yfac = std::max(1, int_vshrink);
}

extern "C" int orc_affine_size(int w, int h, double a, double b, double c, double d, int *ow, int *oh);
extern "C" int orc_affine(const void *in, int w, int h, int bands, int fmt, double a, double b, double c, double d,
	int interp, double idx, double idy, double odx, double ody, int tile_w, int tile_h, void *out);
extern "C" int orc_resize_affine_args(double hscale, double vscale, int kernel, double *a, double *d, double *idx,
	double *idy, int *interp);

/* upsizing (resize.c:233-307): pure enlargements only; a mixed up/down resize
 * chains reduce and affine with rect origins this whole-image oracle does not model
 */
static int
resize_is_upsize(int w, int h, double hscale, double vscale, int *mixed)
{
	hscale = std::max(hscale, 1.0 / w);
	vscale = std::max(vscale, 1.0 / h);
	*mixed = (hscale > 1.0 && vscale < 1.0) || (hscale < 1.0 && vscale > 1.0);
	return hscale > 1.0 || vscale > 1.0;
}

extern "C" int
orc_resize_size(int w, int h, double hscale, double vscale, int kernel, double gap, int *ow, int *oh)
{
	double hs, vs;
	OrcReduceGeom g;
	int mixed;
	if (kernel == ORC_KERNEL_NEAREST && hscale > 0 && vscale > 0 && (int) (w * hscale + 0.5) > 0 && (int) (h * vscale + 0.5) > 0) {
		int xfac, yfac;
		nearest_subsample(w, h, hscale, vscale, gap, &xfac, &yfac);
		if (xfac > 1 || yfac > 1) {
			w /= xfac;
			h /= yfac;
			if (w <= 0 || h <= 0)
				return -1; /* "image has shrunk to nothing", subsample.c:223-229 */
			hscale *= xfac;
			vscale *= yfac;
		}
	}
	if (resize_is_upsize(w, h, hscale, vscale, &mixed)) {
		double a, d, idx, idy;
		int interp;
		if (mixed)
			return -1;
		orc_resize_affine_args(std::max(hscale, 1.0 / w), std::max(vscale, 1.0 / h), kernel, &a, &d, &idx, &idy, &interp);
		return orc_affine_size(w, h, a, 0, 0, d, ow, oh);
	}
	resize_shrinks(w, h, hscale, vscale, &hs, &vs);
	*oh = h;
	*ow = w;
	if (vs > 1.0) {
		if (orc_reduce_geometry(h, vs, kernel, gap, &g))
			return -1;
		*oh = g.out_size;
	}
	if (hs > 1.0) {
		if (orc_reduce_geometry(w, hs, kernel, gap, &g))
			return -1;
		*ow = g.out_size;
	}
	return (*ow > 0 && *oh > 0) ? 0 : -1;
}

/* Sink tile geometry the reference would use for this chain: the demand hint
 * is the minimum over the pipeline (generate.c:275-293), shrinkv asks
 * SMALLTILE (shrinkv.c:553), reducev/reduceh FATSTRIP, the rest THINSTRIP;
 * vips_get_tile_size thread.c:288-325.
 */
static void
chain_tiles(int out_w, bool has_shrinkv, bool has_reduce, int *tile_w, int *tile_h)
{
	if (has_shrinkv) {
		*tile_w = 128;
		*tile_h = 128;
	}
	else if (has_reduce) {
		*tile_w = out_w;
		*tile_h = 16;
	}
	else {
		*tile_w = out_w;
		*tile_h = out_w > 10000 ? 1 : 16;
	}
}

/* vips_resize after the NEAREST pre-shrink (orc_resize below): residual reduce, or the affine for enlargements */
static int
resize_residual(const void *in, int w, int h, int bands, int fmt, double hscale, double vscale, int kernel, double gap,
	int tile_w, int tile_h, void *out)
{
	double hs, vs;
	int mixed;
	if (resize_is_upsize(w, h, hscale, vscale, &mixed)) {
		double a, d, idx, idy;
		int interp, ow, oh;
		if (mixed)
			return -1;
		const double ch = std::max(hscale, 1.0 / w), cv = std::max(vscale, 1.0 / h);
		if (kernel == ORC_KERNEL_NEAREST && ch == floor(ch) && cv == floor(cv)) {
			/* vips_zoom (resize.c:263-271; conversion/zoom.c:95-227 paints each input pixel as an
			 * xfac x yfac block): out(x, y) = in(x / xfac, y / yfac), any format
			 */
			const int xf = (int) floor(ch), yf = (int) floor(cv);
			const size_t ps = orc_sizeof_format(fmt) * bands;
			for (int y = 0; y < h * yf; y++)
				for (int x = 0; x < w * xf; x++)
					memcpy((char *) out + ((size_t) y * w * xf + x) * ps, (const char *) in + ((size_t) (y / yf) * w + x / xf) * ps, ps);
			return 0;
		}
		orc_resize_affine_args(std::max(hscale, 1.0 / w), std::max(vscale, 1.0 / h), kernel, &a, &d, &idx, &idy, &interp);
		if (orc_affine_size(w, h, a, 0, 0, d, &ow, &oh))
			return -1;
		/* a scale-only affine asks FATSTRIP (affine.c:571-575): full-width x 16-row sink tiles */
		if (tile_w <= 0 || tile_h <= 0) {
			tile_w = ow;
			tile_h = 16;
		}
		return orc_affine(in, w, h, bands, fmt, a, 0, 0, d, interp, idx, idy, 0, 0, tile_w, tile_h, out);
	}
	resize_shrinks(w, h, hscale, vscale, &hs, &vs);
	const size_t es = orc_sizeof_format(fmt);

	OrcReduceGeom gv, gh;
	if (vs > 1.0) {
		if (orc_reduce_geometry(h, vs, kernel, gap, &gv))
			return -1;
	}
	else {
		gv.int_shrink = 1;
		gv.n_point = 0;
		gv.out_size = h;
	}
	if (hs > 1.0) {
		if (orc_reduce_geometry(w, hs, kernel, gap, &gh))
			return -1;
	}
	else {
		gh.int_shrink = 1;
		gh.n_point = 0;
		gh.out_size = w;
	}

	if (tile_w <= 0 || tile_h <= 0)
		chain_tiles(gh.out_size, gv.int_shrink > 1, vs > 1.0 || hs > 1.0, &tile_w, &tile_h);

	/* reducev sees the rows of the sink tile, cut into fatstrip-height (16)
	 * chunks when a shrinkh sits downstream of it (shrinkh.c:247-272).
	 */
	int rect_h = tile_h;
	if (gh.int_shrink > 1)
		rect_h = std::min(rect_h, 16);
	const int rect_w = tile_w;

	std::vector<uint8_t> mid;
	const void *src = in;
	if (vs > 1.0) {
		mid.resize((size_t) w * gv.out_size * bands * es);
		if (orc_reducev(in, w, h, bands, fmt, vs, kernel, gap, rect_h, mid.data()))
			return -1;
		src = mid.data();
	}
	if (hs > 1.0)
		return orc_reduceh(src, w, gv.out_size, bands, fmt, hs, kernel, gap, rect_w, out);
	memcpy(out, src, (size_t) w * gv.out_size * bands * es);
	return 0;
}

extern "C" int
orc_resize(const void *in, int w, int h, int bands, int fmt, double hscale, double vscale, int kernel, double gap,
	int tile_w, int tile_h, void *out)
{
	if (kernel == ORC_KERNEL_NEAREST && hscale > 0 && vscale > 0 && (int) (w * hscale + 0.5) > 0 && (int) (h * vscale + 0.5) > 0) {
		int xfac, yfac;
		nearest_subsample(w, h, hscale, vscale, gap, &xfac, &yfac);
		if (xfac > 1 || yfac > 1) {
			const int sw = w / xfac, sh = h / yfac;
			if (sw <= 0 || sh <= 0)
				return -1; /* "image has shrunk to nothing", subsample.c:223-229 */
			const size_t ps = orc_sizeof_format(fmt) * bands;
			std::vector<uint8_t> sub((size_t) sw * sh * ps);
			for (int y = 0; y < sh; y++)
				for (int x = 0; x < sw; x++)
					memcpy(&sub[((size_t) y * sw + x) * ps], (const char *) in + ((size_t) y * yfac * w + (size_t) x * xfac) * ps, ps);
			return resize_residual(sub.data(), sw, sh, bands, fmt, hscale * xfac, vscale * yfac, kernel, gap, tile_w, tile_h, out);
		}
	}
	return resize_residual(in, w, h, bands, fmt, hscale, vscale, kernel, gap, tile_w, tile_h, out);
}

/* ------------------------------------------------------------- thumbnail */

extern "C" int
orc_thumbnail_size(int w, int h, int target_w, int target_h, int size_mode, double *hshrink, double *vshrink,
	int *ow, int *oh)
{
	/* vips_thumbnail_calculate_shrink, thumbnail.c:413-466 (crop NONE, no rotate) */
	double hs = (double) w / target_w;
	double vs = (double) h / target_h;
	bool horizontal = !(hs < vs);

	if (size_mode != 3 /*FORCE*/) {
		if (horizontal)
			vs = hs;
		else
			hs = vs;
	}
	if (size_mode == 1 /*UP*/) {
		hs = std::min(1.0, hs);
		vs = std::min(1.0, vs);
	}
	else if (size_mode == 2 /*DOWN*/) {
		hs = std::max(1.0, hs);
		vs = std::max(1.0, vs);
	}
	hs = std::min(hs, (double) w);
	vs = std::min(vs, (double) h);
	*hshrink = hs;
	*vshrink = vs;
	return orc_resize_size(w, h, 1.0 / hs, 1.0 / vs, ORC_KERNEL_LANCZOS3, 2.0, ow, oh);
}

extern "C" int
orc_thumbnail_image(const void *in, int w, int h, int bands, int target_w, int target_h, int size_mode,
	int has_alpha, int tile_w, int tile_h, void *out)
{
	double hs, vs;
	int ow, oh;
	if (orc_thumbnail_size(w, h, target_w, target_h, size_mode, &hs, &vs, &ow, &oh))
		return -1;
	/* enlarging thumbnails go through vips_resize's affine half (orc_resize handles it) */

	const bool premul = has_alpha && hs != 1.0 && vs != 1.0; /* thumbnail.c:848-861 */
	std::vector<uint8_t> pre, res;
	const void *src = in;
	if (premul) {
		pre.resize((size_t) w * h * bands);
		orc_premultiply(in, w, h, bands, ORC_FORMAT_UCHAR, 255.0, 1, pre.data());
		src = pre.data();
	}
	if (!premul)
		return orc_resize(src, w, h, bands, ORC_FORMAT_UCHAR, 1.0 / hs, 1.0 / vs, ORC_KERNEL_LANCZOS3, 2.0,
			tile_w, tile_h, out);
	res.resize((size_t) ow * oh * bands);
	if (orc_resize(src, w, h, bands, ORC_FORMAT_UCHAR, 1.0 / hs, 1.0 / vs, ORC_KERNEL_LANCZOS3, 2.0, tile_w,
			tile_h, res.data()))
		return -1;
	return orc_unpremultiply(res.data(), ow, oh, bands, ORC_FORMAT_UCHAR, 255.0, 1, out);
}

/* A batch of same-shaped frames, one frame per worker thread (the reference's
 * inter-image parallelism: many caller threads each running a pipeline,
 * doc/using-threads.md:61-95).  Used by bench.py's CPU baseline only.
 */
#include <atomic>
#include <thread>

extern "C" int
orc_thumbnail_image_batch(const void *in, int n_frames, int w, int h, int bands, int target_w, int target_h,
	int size_mode, int has_alpha, void *out, int ow, int oh, int n_threads)
{
	std::atomic<int> next{0}, rc{0};
	auto work = [&]() {
		for (;;) {
			const int i = next.fetch_add(1);
			if (i >= n_frames)
				break;
			const uint8_t *fi = (const uint8_t *) in + (size_t) i * w * h * bands;
			uint8_t *fo = (uint8_t *) out + (size_t) i * ow * oh * bands;
			if (orc_thumbnail_image(fi, w, h, bands, target_w, target_h, size_mode, has_alpha, 0, 0, fo))
				rc = -1;
		}
	};
	if (n_threads < 1)
		n_threads = 1;
	std::vector<std::thread> pool;
	for (int t = 1; t < n_threads; t++)
		pool.emplace_back(work);
	work();
	for (auto &t : pool)
		t.join();
	return rc;
}

/* vips_thumbnail_image(..., linear = TRUE) for an 8-bit sRGB image without ICC
 * profile: thumbnail.c:766-806 (to scRGB), :848-902 (float premultiply / resize /
 * unpremultiply + cast), :971-987 (back to sRGB).  bands >= 3.
 */
extern "C" int orc_colourspace(const void *in, int w, int h, int bands, int fmt, int from, int to, void *out);

extern "C" int
orc_thumbnail_image_linear(const void *in, int w, int h, int bands, int target_w, int target_h, int size_mode,
	int has_alpha, int tile_w, int tile_h, void *out)
{
	double hs, vs;
	int ow, oh;
	if (bands < 3)
		return -1; /* GREY16 route: not restated */
	if (orc_thumbnail_size(w, h, target_w, target_h, size_mode, &hs, &vs, &ow, &oh))
		return -1;
	if (hs < 1.0 || vs < 1.0)
		return -1;

	const size_t n_in = (size_t) w * h * bands, n_out = (size_t) ow * oh * bands;
	std::vector<float> lin(n_in), pre, res(n_out), unpre;
	if (orc_colourspace(in, w, h, bands, ORC_FORMAT_UCHAR, 22 /*sRGB*/, 28 /*scRGB*/, lin.data()))
		return -1;
	const bool premul = has_alpha && hs != 1.0 && vs != 1.0;
	const float *src = lin.data();
	if (premul) {
		pre.resize(n_in);
		/* max_alpha = vips_interpretation_max_alpha(scRGB) = 1.0 */
		if (orc_premultiply(lin.data(), w, h, bands, ORC_FORMAT_FLOAT, 1.0, 0, pre.data()))
			return -1;
		src = pre.data();
	}
	if (orc_resize(src, w, h, bands, ORC_FORMAT_FLOAT, 1.0 / hs, 1.0 / vs, ORC_KERNEL_LANCZOS3, 2.0, tile_w, tile_h,
			res.data()))
		return -1;
	const float *last = res.data();
	if (premul) {
		unpre.resize(n_out);
		if (orc_unpremultiply(res.data(), ow, oh, bands, ORC_FORMAT_FLOAT, 1.0, 0, unpre.data()))
			return -1;
		last = unpre.data(); /* vips_cast(float -> float) is the identity */
	}
	return orc_colourspace(last, ow, oh, bands, ORC_FORMAT_FLOAT, 28, 22, out);
}
