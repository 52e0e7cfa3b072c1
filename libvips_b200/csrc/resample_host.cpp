/* resample_host.cpp -- host-side setup for the resample kernels: the part of
 * vips_reducev_build / vips_reduceh_build that runs once per operation on the
 * CPU in the reference too (geometry, mask tables) plus the sampling tables
 * the kernels index instead of stepping a double per scanline.
 *
 * All tables are computed HERE with the host libm (sin) and uploaded: the
 * reference's masks depend on the host's sin() (resample/templates.h:346-354)
 * so they must never be recomputed with device intrinsics.
 *
 * reference: resample/reducev.cpp:877-965, resample/reduceh.cpp:112-142,416-505,
 *            resample/templates.h:300-526.
 */
#include <algorithm>
#include <cmath>

#include "vb200_internal.h"

namespace vb200 {

namespace {

const double kPi = 3.14159265358979323846; /* VIPS_PI */

double
sinc(double x)
{
	if (x == 0.0)
		return 1.0;
	x = x * kPi;
	return sin(x) / x;
}

/* Mitchell-Netravali family, templates.h:321-344 */
double
cubic_bc(double x, double B, double C)
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

double
lanczos(double x, int a)
{
	if (x >= -a && x <= a)
		return sinc(x) * sinc(x / a);
	return 0.0;
}

/* Magic Kernel Sharp 2013 / 2021, templates.h:408-448 */
double
mks2013(double x)
{
	x = fabs(x);
	if (x >= 2.5)
		return 0.0;
	if (x >= 1.5)
		return (x - 5.0 / 2.0) * (x - 5.0 / 2.0) / -8.0;
	if (x >= 0.5)
		return (4.0 * x * x - 11.0 * x + 7.0) / 4.0;
	return 17.0 / 16.0 - 7.0 * x * x / 4.0;
}

double
mks2021(double x)
{
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

double
kernel_value(int kernel, double x)
{
	switch (kernel) {
	case VB200_KERNEL_LINEAR: {
		const double ax = fabs(x);
		return ax < 1.0 ? 1.0 - ax : 0.0;
	}
	case VB200_KERNEL_CUBIC:
		return cubic_bc(x, 0.0, 0.5);
	case VB200_KERNEL_MITCHELL:
		return cubic_bc(x, 1.0 / 3.0, 1.0 / 3.0);
	case VB200_KERNEL_LANCZOS2:
		return lanczos(x, 2);
	case VB200_KERNEL_LANCZOS3:
		return lanczos(x, 3);
	case VB200_KERNEL_MKS2013:
		return mks2013(x);
	case VB200_KERNEL_MKS2021:
		return mks2021(x);
	}
	return 0.0;
}

} // namespace

int
reduce_get_points(int kernel, double shrink)
{
	switch (kernel) {
	case VB200_KERNEL_NEAREST:
		return 1;
	case VB200_KERNEL_LINEAR:
		return 2 * rint(shrink) + 1;
	case VB200_KERNEL_CUBIC:
	case VB200_KERNEL_MITCHELL:
	case VB200_KERNEL_LANCZOS2:
		return 2 * rint(2 * shrink) + 1;
	case VB200_KERNEL_LANCZOS3:
	case VB200_KERNEL_MKS2013:
		return 2 * rint(3 * shrink) + 1;
	case VB200_KERNEL_MKS2021:
		return 2 * rint(5 * shrink) + 1;
	}
	return 0;
}

void
reduce_make_mask(double *c, int kernel, int n_points, double shrink, double x)
{
	if (kernel == VB200_KERNEL_NEAREST) {
		c[0] = 1.0;
		return;
	}

	/* calculate_coefficients<double>, templates.h:457-480: sample the filter
	 * at the tap centres, normalise to unit DC gain.
	 */
	const double half = x + n_points / 2.0 - 1;
	const double scale = 1.0 / shrink;
	double sum = 0.0;
	for (int i = 0; i < n_points; i++) {
		const double v = kernel_value(kernel, (i - half) * scale);
		c[i] = v;
		sum += v;
	}
	for (int i = 0; i < n_points; i++)
		c[i] /= sum;
}

int
shrink_size(int in_size, int shrink, int ceil_mode)
{
	const double q = (double) in_size / shrink;
	return ceil_mode ? (int) ceil(q) : VB200_ROUND_UINT(q);
}

int
reduce_geometry(const char *domain, int in_size, double shrink, int kernel, double gap, ReduceGeom *g)
{
	if (shrink < 1.0) {
		error(domain, "reduce factor should be >= 1.0");
		return -1;
	}

	g->in_size = in_size;
	g->out_size = VB200_ROUND_UINT((double) in_size / shrink);
	g->int_shrink = 1;
	g->shrunk_size = in_size;
	g->residual = shrink;
	g->n_point = 0;
	g->offset = 0.0;

	/* pixels invented (+) or discarded (-) in the input */
	double extra = g->out_size * shrink - in_size;

	if (gap > 0.0 && kernel != VB200_KERNEL_NEAREST) {
		if (gap < 1.0) {
			error(domain, "reduce gap should be >= 1.0");
			return -1;
		}
		const int box = std::max(1.0, floor((double) in_size / g->out_size / gap));
		if (box > 1) {
			g->int_shrink = box;
			g->shrunk_size = shrink_size(in_size, box, 1);
			extra /= box;
			g->residual /= box;
		}
	}

	if (g->residual == 1.0) {
		g->out_size = g->shrunk_size;
		return 0;
	}

	g->n_point = reduce_get_points(kernel, g->residual);
	if (g->n_point > VB200_MAX_POINT) {
		error(domain, "reduce factor too large");
		return -1;
	}
	if (g->out_size <= 0) {
		error(domain, "image has shrunk to nothing");
		return -1;
	}
	g->offset = (1 + extra) / 2.0 - 1;
	return 0;
}

void
build_axis_table(AxisTable &t, int out_size, double residual, double offset, int n_point, int kernel,
	int rect_size, int rect_origin, int count)
{
	t.n_point = n_point;
	t.embed = (int) ceil(n_point / 2.0) - 1;

	t.mf.resize((size_t) (VB200_TRANSFORM_SCALE + 1) * n_point);
	t.ms.resize(t.mf.size());
	for (int p = 0; p <= VB200_TRANSFORM_SCALE; p++) {
		double *f = &t.mf[(size_t) p * n_point];
		reduce_make_mask(f, kernel, n_point, residual, (float) p / VB200_TRANSFORM_SCALE);
		for (int i = 0; i < n_point; i++)
			t.ms[(size_t) p * n_point + i] = (short) (f[i] * VB200_INTERPOLATE_SCALE);
	}

	/* One entry per output row/column.  Each rect restarts the coordinate at
	 * (origin + 0.5) * residual - 0.5 - offset and then ADDS residual per
	 * step, as the generate functions do (reducev.cpp:548-611,
	 * reduceh.cpp:254-326): not the same double as origin-free multiplication
	 * when residual is not exactly representable.
	 */
	if (count < 0)
		count = out_size;
	if (rect_size <= 0)
		rect_size = count;
	t.first.resize(count);
	t.phase.resize(count);
	for (int start = 0; start < count; start += rect_size) {
		const int len = std::min(rect_size, count - start);
		double pos = (rect_origin + start + 0.5) * residual - 0.5 - offset;
		for (int i = 0; i < len; i++) {
			const int whole = (int) pos;
			const int fixed = pos * VB200_TRANSFORM_SCALE * 2;
			const int frac = fixed & (VB200_TRANSFORM_SCALE * 2 - 1);
			t.first[start + i] = whole;
			t.phase[start + i] = (frac + 1) >> 1;
			pos += residual;
		}
	}
}

bool
build_mma_tables(const AxisTable &t, int out_size, int rows, std::vector<int> &vchunk, std::vector<unsigned> &bfrag)
{
	const int K = rows, QUADS = 8; /* rows: output rows per chunk, <= 8 (the N of the MMA; columns past it are zero) */
	if (K < 1 || K > 8)
		return false;
	const int chunks = (out_size + K - 1) / K;
	const int np = t.n_point;
	vchunk.assign((size_t) chunks * 2, 0);
	bfrag.assign((size_t) chunks * 32 * 4, 0);
	for (int c = 0; c < chunks; c++) {
		/* the rows this chunk's live taps touch (all-zero end taps do not count) */
		int lo = INT_MAX, hi = INT_MIN;
		for (int y = c * K; y < std::min(c * K + K, out_size); y++) {
			const short *m = &t.ms[(size_t) t.phase[y] * np];
			int a = 0, b = np - 1;
			while (a < b && m[a] == 0)
				a++;
			while (b > a && m[b] == 0)
				b--;
			lo = std::min(lo, t.first[y] + a);
			hi = std::max(hi, t.first[y] + b);
		}
		if (lo < 0)
			return false;
		int q0 = lo >> 2;
		const int q1 = hi >> 2;
		if (c > 0) {
			/* quads are produced in order, without gaps */
			q0 = std::min(q0, vchunk[(c - 1) * 2 + 1] + 1);
			if (q1 < vchunk[(c - 1) * 2 + 1])
				return false;
		}
		if (q1 - q0 + 1 > QUADS)
			return false;
		vchunk[c * 2] = q0;
		vchunk[c * 2 + 1] = q1;
		for (int lane = 0; lane < 32; lane++) {
			const int tig = lane & 3, g = lane >> 2;
			const int y = g < K ? c * K + g : out_size; /* fragment columns past the chunk's rows stay zero */
			unsigned w[4] = {0, 0, 0, 0}; /* hi b0, hi b1, lo b0, lo b1 */
			for (int half = 0; half < 2; half++) {
				const int slot = tig + 4 * half;
				int q = -1;
				for (int qq = q0; qq <= q1; qq++)
					if ((qq & (QUADS - 1)) == slot)
						q = qq;
				for (int i = 0; i < 4; i++) {
					int coef = 0;
					if (q >= 0 && y < out_size) {
						const int tap = 4 * q + i - t.first[y];
						if (tap >= 0 && tap < np)
							coef = t.ms[(size_t) t.phase[y] * np + tap];
					}
					w[half] |= (unsigned) ((coef >> 8) & 0xff) << (8 * i);
					w[2 + half] |= (unsigned) (coef & 0xff) << (8 * i);
				}
			}
			for (int k = 0; k < 4; k++)
				bfrag[((size_t) c * 32 + lane) * 4 + k] = w[k];
		}
	}
	return true;
}

int
pick_mma_rows(const AxisTable &t, int out_size, std::vector<int> &vchunk, std::vector<unsigned> &bfrag)
{
	/* most output rows per chunk first: fewer MMAs per row */
	for (int rows = 8; rows >= 4; rows--)
		if (build_mma_tables(t, out_size, rows, vchunk, bfrag))
			return rows;
	return 0;
}

} // namespace vb200

/* Test hook (tests/test_mma_tables.py, CPU): the reducev geometry, sampling table and tensor-pipe
 * tables of a vertical thumbnail shrink, exactly as the plan builds them.  Arrays are caller-sized:
 * first/phase [*out_size], vchunk [2 * chunks], bfrag [128 * chunks] with chunks = ceil(out_size / rows_per_chunk),
 * mask65 [65 * *n_point] shorts.
 * Returns 0, 1 when the window does not fit the quad ring, -1 on bad arguments.
 */
extern "C" int
vb200_debug_mma_tables(int in_size, double shrink, int rect_size, int *int_shrink, int *shrunk_size, int *out_size,
	int *n_point, int *embed, int *first, int *phase, short *mask65, int *vchunk, unsigned *bfrag, int cap_rows,
	int *rows_per_chunk)
{
	using namespace vb200;
	ReduceGeom g;
	if (reduce_geometry("debug", in_size, shrink, VB200_KERNEL_LANCZOS3, 2.0, &g))
		return -1;
	if (g.out_size > cap_rows)
		return -1;
	AxisTable t;
	build_axis_table(t, g.out_size, g.residual, g.offset, g.n_point, VB200_KERNEL_LANCZOS3, rect_size);
	*int_shrink = g.int_shrink;
	*shrunk_size = g.shrunk_size;
	*out_size = g.out_size;
	*n_point = g.n_point;
	*embed = t.embed;
	for (int i = 0; i < g.out_size; i++) {
		first[i] = t.first[i];
		phase[i] = t.phase[i];
	}
	memcpy(mask65, t.ms.data(), t.ms.size() * sizeof(short));
	std::vector<int> vc;
	std::vector<unsigned> bf;
	const int rows = pick_mma_rows(t, g.out_size, vc, bf);
	if (!rows)
		return 1;
	*rows_per_chunk = rows;
	memcpy(vchunk, vc.data(), vc.size() * sizeof(int));
	memcpy(bfrag, bf.data(), bf.size() * sizeof(unsigned));
	return 0;
}
