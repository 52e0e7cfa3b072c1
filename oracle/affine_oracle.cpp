/* affine_oracle.cpp -- CPU restatement of vips_affine + the nearest / bilinear /
 * bicubic interpolators, and the upsizing half of vips_resize.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  -O2 -ffp-contract=off.
 *
 * Follows (reference = libvips 8.19, libvips/resample/):
 *   affine.c:227-410        vips_affine_gen (sequential ix += ddx per rect row)
 *   affine.c:412-605        vips_affine_build (transform, oarea, embed window + 1, idx -= 1)
 *   transform.c:48-70,158-240   inverse, forward/invert point, rect, set_area
 *   interpolate.c:334-349   nearest;  :433-480, :527-554   bilinear
 *   bicubic.cpp:106-405,487-645   bicubic tab functions + tables
 *   templates.h:150-300     bicubic_unsigned_int / signed / float, catmull coefficients
 *   resize.c:116-132,233-307      interpolator choice and the affine calls of vips_resize
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <type_traits>
#include <vector>

#include "oracle.h"

#define TRANSFORM_SCALE 64
#define INTERPOLATE_SHIFT 12
#define INTERPOLATE_SCALE (1 << INTERPOLATE_SHIFT)
#define ROUND_INT(R) ((int) ((R) > 0 ? ((R) + 0.5) : ((R) -0.5)))
#define ROUND_UINT(R) ((int) ((R) + 0.5))

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

enum { INTERP_NEAREST = 0, INTERP_BILINEAR = 1, INTERP_BICUBIC = 2 };

struct Trn {
	double a, b, c, d, ia, ib, ic, id, idx, idy, odx, ody;
	int il, it, iw, ih; /* iarea */
	int ol, ot, ow, oh; /* oarea */
};

static int
trn_inverse(Trn *t)
{
	const double det = t->a * t->d - t->b * t->c;
	if (fabs(det) < 2.0 * 2.2250738585072014e-308)
		return -1;
	const double tmp = 1.0 / det;
	t->ia = tmp * t->d;
	t->ib = -tmp * t->b;
	t->ic = -tmp * t->c;
	t->id = tmp * t->a;
	return 0;
}

static void
forward_point(const Trn *t, double x, double y, double *ox, double *oy)
{
	x += t->idx;
	y += t->idy;
	*ox = t->a * x + t->b * y + t->odx;
	*oy = t->c * x + t->d * y + t->ody;
}

static void __attribute__((unused))
invert_point(const Trn *t, double x, double y, double *ox, double *oy)
{
	x -= t->odx;
	y -= t->ody;
	*ox = t->ia * x + t->ib * y - t->idx;
	*oy = t->ic * x + t->id * y - t->idy;
}

template <typename F>
static void
transform_rect(const Trn *t, F fn, int l, int tp, int w, int h, int *ol, int *ot, int *ow, int *oh)
{
	double x1, y1, x2, y2, x3, y3, x4, y4;
	fn(t, l, tp, &x1, &y1);
	fn(t, l, tp + h, &x3, &y3);
	fn(t, l + w, tp, &x2, &y2);
	fn(t, l + w, tp + h, &x4, &y4);
	const double left = std::min(x1, std::min(x2, std::min(x3, x4)));
	const double right = std::max(x1, std::max(x2, std::max(x3, x4)));
	const double top = std::min(y1, std::min(y2, std::min(y3, y4)));
	const double bottom = std::max(y1, std::max(y2, std::max(y3, y4)));
	*ol = ROUND_INT(left);
	*ot = ROUND_INT(top);
	*ow = ROUND_INT(right - left);
	*oh = ROUND_INT(bottom - top);
}

/* bicubic tables, bicubic.cpp:636-644 + calculate_coefficients_catmull templates.h:281-305 */
static double bicubic_f[TRANSFORM_SCALE + 1][4];
static int bicubic_i[TRANSFORM_SCALE + 1][4];
static bool bicubic_ready = false;

static void
catmull(double c[4], const double x)
{
	const double cr1 = 1. - x;
	const double cr2 = -.5 * x;
	const double cr3 = cr1 * cr2;
	const double cone = cr1 * cr3;
	const double cfou = x * cr3;
	const double cr4 = cfou - cone;
	const double ctwo = cr1 - cone + cr4;
	const double cthr = x - cfou - cr4;
	c[0] = cone;
	c[3] = cfou;
	c[1] = ctwo;
	c[2] = cthr;
}

static void
bicubic_tables()
{
	if (bicubic_ready)
		return;
	for (int x = 0; x < TRANSFORM_SCALE + 1; x++) {
		catmull(bicubic_f[x], (float) x / TRANSFORM_SCALE);
		for (int i = 0; i < 4; i++)
			bicubic_i[x][i] = bicubic_f[x][i] * INTERPOLATE_SCALE;
	}
	bicubic_ready = true;
}

extern "C" void
orc_bicubic_tables(double *f, int *i)
{
	bicubic_tables();
	memcpy(f, bicubic_f, sizeof(bicubic_f));
	memcpy(i, bicubic_i, sizeof(bicubic_i));
}

static inline int ufr(int v) { return (v + (INTERPOLATE_SCALE >> 1)) >> INTERPOLATE_SHIFT; }
static inline int
sfr(int v)
{
	const int sign_of_v = 2 * (v >= 0) - 1;
	const int round_by = sign_of_v * (INTERPOLATE_SCALE >> 1);
	return (v + round_by) >> INTERPOLATE_SHIFT;
}

/* the embedded image: (X, Y) -> source pixel with EXTEND_COPY */
template <typename T>
struct Src {
	const T *p;
	int w, h, bands, pad;
	inline T at(int X, int Y, int z) const
	{
		return p[((size_t) clampi(Y - pad, 0, h - 1) * w + clampi(X - pad, 0, w - 1)) * bands + z];
	}
};

template <typename T>
static void
interp_pixel(const Src<T> &s, int fmt, int interp, double x, double y, T *q)
{
	const int bands = s.bands;
	if (interp == INTERP_NEAREST) {
		const int xi = (int) x, yi = (int) y;
		for (int z = 0; z < bands; z++)
			q[z] = s.at(xi, yi, z);
		return;
	}
	if (interp == INTERP_BILINEAR) {
		const int ix = (int) x, iy = (int) y;
		if constexpr (std::is_integral<T>::value && sizeof(T) <= 2) {
			/* BILINEAR_INT, interpolate.c:433-457 */
			const int X = (x - ix) * INTERPOLATE_SCALE;
			const int Y = (y - iy) * INTERPOLATE_SCALE;
			const int Yd = INTERPOLATE_SCALE - Y;
			const int c4 = (Y * X) >> INTERPOLATE_SHIFT;
			const int c2 = (Yd * X) >> INTERPOLATE_SHIFT;
			const int c3 = Y - c4;
			const int c1 = Yd - c2;
			for (int z = 0; z < bands; z++)
				q[z] = (c1 * s.at(ix, iy, z) + c2 * s.at(ix + 1, iy, z) + c3 * s.at(ix, iy + 1, z) +
						   c4 * s.at(ix + 1, iy + 1, z) + (1 << INTERPOLATE_SHIFT) / 2) >>
					INTERPOLATE_SHIFT;
		}
		else {
			/* BILINEAR_FLOAT, interpolate.c:463-482 */
			const double X = x - ix;
			const double Y = y - iy;
			const double Yd = 1.0f - Y;
			const double c4 = Y * X;
			const double c2 = Yd * X;
			const double c3 = Y - c4;
			const double c1 = Yd - c2;
			for (int z = 0; z < bands; z++)
				q[z] = c1 * s.at(ix, iy, z) + c2 * s.at(ix + 1, iy, z) + c3 * s.at(ix, iy + 1, z) +
					c4 * s.at(ix + 1, iy + 1, z);
		}
		return;
	}

	/* bicubic, bicubic.cpp:487-619 */
	const int sx = x * TRANSFORM_SCALE * 2;
	const int sy = y * TRANSFORM_SCALE * 2;
	const int six = sx & (TRANSFORM_SCALE * 2 - 1);
	const int siy = sy & (TRANSFORM_SCALE * 2 - 1);
	const int tx = (six + 1) >> 1;
	const int ty = (siy + 1) >> 1;
	const int ix = (int) x, iy = (int) y;
	const int *cxi = bicubic_i[tx], *cyi = bicubic_i[ty];
	const double *cxf = bicubic_f[tx], *cyf = bicubic_f[ty];

	for (int z = 0; z < bands; z++) {
		T v[4][4];
		for (int j = 0; j < 4; j++)
			for (int i = 0; i < 4; i++)
				v[j][i] = s.at(ix - 1 + i, iy - 1 + j, z);
		if constexpr (sizeof(T) == 1) {
			int r[4];
			for (int j = 0; j < 4; j++) {
				const int sum = cxi[0] * v[j][0] + cxi[1] * v[j][1] + cxi[2] * v[j][2] + cxi[3] * v[j][3];
				r[j] = fmt == ORC_FORMAT_UCHAR ? ufr(sum) : sfr(sum);
			}
			const int sum = cyi[0] * r[0] + cyi[1] * r[1] + cyi[2] * r[2] + cyi[3] * r[3];
			int bicubic = fmt == ORC_FORMAT_UCHAR ? ufr(sum) : sfr(sum);
			bicubic = fmt == ORC_FORMAT_UCHAR ? clampi(bicubic, 0, UCHAR_MAX) : clampi(bicubic, SCHAR_MIN, SCHAR_MAX);
			q[z] = bicubic;
		}
		else if (fmt == ORC_FORMAT_FLOAT) {
			/* bicubic_float<float>: each cubic_float<T> returns T */
			double r[4];
			for (int j = 0; j < 4; j++)
				r[j] = (float) (cxf[0] * v[j][0] + cxf[1] * v[j][1] + cxf[2] * v[j][2] + cxf[3] * v[j][3]);
			q[z] = (float) (cyf[0] * (float) r[0] + cyf[1] * (float) r[1] + cyf[2] * (float) r[2] + cyf[3] * (float) r[3]);
		}
		else {
			/* bicubic_{un,}signed_int32_tab: bicubic_float<double>, clip, truncate */
			double r[4];
			for (int j = 0; j < 4; j++)
				r[j] = cxf[0] * (double) v[j][0] + cxf[1] * (double) v[j][1] + cxf[2] * (double) v[j][2] +
					cxf[3] * (double) v[j][3];
			double bicubic = cyf[0] * r[0] + cyf[1] * r[1] + cyf[2] * r[2] + cyf[3] * r[3];
			double lo, hi;
			switch (fmt) {
			case ORC_FORMAT_USHORT: lo = 0; hi = USHRT_MAX; break;
			case ORC_FORMAT_SHORT: lo = SHRT_MIN; hi = SHRT_MAX; break;
			case ORC_FORMAT_UINT: lo = 0; hi = INT_MAX; break;
			default: lo = INT_MIN; hi = INT_MAX; break;
			}
			/* VIPS_CLIP(lo, bicubic, hi) */
			const double m = hi < bicubic ? hi : bicubic;
			bicubic = lo > m ? lo : m;
			q[z] = (T) bicubic;
		}
	}
}

template <typename T>
static void
affine_t(const T *in, int w, int h, int bands, int fmt, const Trn &t, int interp, int tile_w, int tile_h, T *out)
{
	const int window_size = interp == INTERP_BICUBIC ? 4 : (interp == INTERP_BILINEAR ? 2 : 1);
	const int window_offset = std::max(0, window_size / 2 - 1);
	Src<T> s{in, w, h, bands, window_offset + 1};
	const int OW = t.ow, OH = t.oh;
	if (tile_w <= 0)
		tile_w = OW;
	if (tile_h <= 0)
		tile_h = OH;
	bicubic_tables();

	const int ile = t.il + window_offset;
	const int ito = t.it + window_offset;
	const int iri = ile + t.iw;
	const int ibo = ito + t.ih;
	const double ddx = t.ia;
	const double ddy = t.ic;

	for (int top = 0; top < OH; top += tile_h)
		for (int le = 0; le < OW; le += tile_w) {
			const int ri = std::min(le + tile_w, OW);
			const int bo = std::min(top + tile_h, OH);
			for (int y = top; y < bo; y++) {
				const double ox = le + t.ol - t.odx;
				const double oy = y + t.ot - t.ody;
				double ix = t.ia * ox + t.ib * oy;
				double iy = t.ic * ox + t.id * oy;
				ix -= t.idx;
				iy -= t.idy;
				ix += window_offset;
				iy += window_offset;
				T *q = out + ((size_t) y * OW + le) * bands;
				for (int x = le; x < ri; x++) {
					const int fx = floor(ix);
					const int fy = floor(iy);
					if (fx >= ile && fx <= iri && fy >= ito && fy <= ibo)
						interp_pixel<T>(s, fmt, interp, ix, iy, q);
					else
						for (int z = 0; z < bands; z++)
							q[z] = 0; /* background 0 */
					ix += ddx;
					iy += ddy;
					q += bands;
				}
			}
		}
}

static int
affine_setup(Trn *t, int w, int h, double a, double b, double c, double d, double idx, double idy, double odx, double ody)
{
	memset(t, 0, sizeof(*t));
	t->il = 0;
	t->it = 0;
	t->iw = w;
	t->ih = h;
	t->a = a;
	t->b = b;
	t->c = c;
	t->d = d;
	if (trn_inverse(t))
		return -1;
	/* vips__transform_set_area with idx = idy = odx = ody = 0, affine.c:466-481 */
	transform_rect(t, forward_point, 0, 0, w, h, &t->ol, &t->ot, &t->ow, &t->oh);
	t->odx = odx;
	t->ody = ody;
	t->idx = idx;
	t->idy = idy;
	/* the one-pixel border of the embed, affine.c:533-534 */
	t->idx -= 1;
	t->idy -= 1;
	return 0;
}

extern "C" int
orc_affine_size(int w, int h, double a, double b, double c, double d, int *ow, int *oh)
{
	Trn t;
	if (affine_setup(&t, w, h, a, b, c, d, 0, 0, 0, 0))
		return -1;
	*ow = t.ow;
	*oh = t.oh;
	return 0;
}

/* vips_affine(in, a, b, c, d, interpolate, idx, idy, odx, ody, extend = COPY, premultiplied = TRUE) */
extern "C" int
orc_affine(const void *in, int w, int h, int bands, int fmt, double a, double b, double c, double d, int interp,
	double idx, double idy, double odx, double ody, int tile_w, int tile_h, void *out)
{
	Trn t;
	if (affine_setup(&t, w, h, a, b, c, d, idx, idy, odx, ody))
		return -1;
	if (t.ow <= 0 || t.oh <= 0)
		return -1;
#define AF(T) affine_t<T>((const T *) in, w, h, bands, fmt, t, interp, tile_w, tile_h, (T *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: AF(uint8_t); break;
	case ORC_FORMAT_CHAR: AF(int8_t); break;
	case ORC_FORMAT_USHORT: AF(uint16_t); break;
	case ORC_FORMAT_SHORT: AF(int16_t); break;
	case ORC_FORMAT_UINT: AF(uint32_t); break;
	case ORC_FORMAT_INT: AF(int32_t); break;
	case ORC_FORMAT_FLOAT: AF(float); break;
	default: return -1;
	}
	return 0;
}

/* The upsizing tail of vips_resize, resize.c:233-307.  scales are the values
 * vips_resize holds at that point (hscale / vscale after the clamp to 1 / size);
 * returns the affine arguments.
 */
extern "C" int
orc_resize_affine_args(double hscale, double vscale, int kernel, double *a, double *d, double *idx, double *idy,
	int *interp)
{
	*interp = kernel == ORC_KERNEL_NEAREST ? INTERP_NEAREST : (kernel == ORC_KERNEL_LINEAR ? INTERP_BILINEAR : INTERP_BICUBIC);
	*idx = kernel == ORC_KERNEL_NEAREST ? 0.0 : 0.5 * (1.0 - 1.0 / hscale);
	*idy = kernel == ORC_KERNEL_NEAREST ? 0.0 : 0.5 * (1.0 - 1.0 / vscale);
	if (hscale > 1.0 && vscale > 1.0) {
		*a = hscale;
		*d = vscale;
	}
	else if (hscale > 1.0) {
		*a = hscale;
		*d = 1.0;
	}
	else {
		*a = 1.0;
		*d = vscale;
	}
	return 0;
}
