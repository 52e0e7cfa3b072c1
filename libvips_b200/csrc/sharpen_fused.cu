/* sharpen_fused.cu -- vips_sharpen on 8-bit sRGB frames as ONE kernel per batch.
 *
 * The reference builds vips_sharpen as a graph (convolution/sharpen.c:171-303):
 *     vips_colourspace(sRGB -> LABS)                      colourspace.c:366  (4 colour ops, float images between)
 *     extract L | extract the rest                         sharpen.c:262-264
 *     vips_convsep(L, integer gaussmat(sigma, 0.1))        two vips_convi passes over a short image, convi.c:698-717
 *     vips_sharpen_generate(L, blurred L)                  sharpen.c:116-168: L + lut[(L & 0x7fff) - (blur & 0x7fff) + 32768]
 *     bandjoin, vips_colourspace(LABS -> sRGB)             colourspace.c:317
 * Here one CTA owns a 32 x 32 output tile of one frame: it converts the tile plus a halo of the mask radius
 * to LabS (only L for the halo), blurs L horizontally then vertically in shared memory (the reference's
 * per-pass rounding and clip to short), applies the LUT, and converts (L', a, b) back to sRGB -- no
 * intermediate ever reaches HBM.  Every colour step is the device function the route kernels use
 * (colour_steps.cuh), so the pixels equal the unfused chain's bit for bit; a fourth band rides along through
 * both routes exactly as vips_colour_build carries it (colour.c:252-291).
 *
 * Edges: each vips_conv embeds its input with VIPS_EXTEND_COPY (convi.c:1194-1200); clamping the halo's
 * coordinates at load time is the same thing for both passes (a clamped row's horizontal blur is the
 * blurred edge row the second embed would replicate).
 *
 * Algorithmic bytes per frame: w * h * bands in + the same out.
 */
#include <algorithm>
#include <cmath>
#include <cstring>
#include <map>
#include <mutex>
#include <tuple>
#include <vector>

#include "colour_steps.cuh"
#include "vb200_internal.h"

namespace vb200 {

void host_gaussmat(double sigma, double min_ampl, bool separable, bool integer_precision, std::vector<double> &coeff, int *width,
	int *height, double *scale);

namespace {

constexpr int kTile = 32;
constexpr int kMaxTaps = 15;

struct SharpenParams {
	StepInfo fwd[6], bwd[6]; /* the alpha handling of sRGB -> LABS and LABS -> sRGB */
	int n_fwd, n_bwd;
	ColourTables t;
	const int *lut; /* [65536] */
	int w, h;
	size_t in_bpl, out_bpl, in_frame_stride, out_frame_stride;
	int n;				/* taps, odd */
	int coef[kMaxTaps]; /* rint(mask), vips__image_intize */
	int scale, rounding;
	/* magic = ceil(2^shift / scale), shift = 31 + bit length of scale: (n * magic) >> shift == n / scale
	 * for every 0 <= n < 2^31 (the error term magic * scale - 2^shift is < scale < 2^(shift - 31))
	 */
	unsigned long long magic;
	int shift;
};

__device__ __forceinline__ int
clip_short(int v)
{
	return max(-32768, min(v, 32767));
}

/* (sum + rounding) / scale with C's truncation toward zero (convi.c:711), without the integer divide
 * (a reciprocal on the conversion unit plus fix-up): one 32 x 32 -> 64 multiply by the host's magic number
 */
__device__ __forceinline__ int
div_scale(int n, const SharpenParams &P)
{
	const unsigned q = (unsigned) (((unsigned long long) (unsigned) abs(n) * P.magic) >> P.shift);
	return n < 0 ? -(int) q : (int) q;
}

/* L of one sRGB pixel in LabS, and optionally a / b: sRGB2scRGB, scRGB2XYZ, XYZ2Lab, Lab2LabS */
template <bool WANT_AB>
__device__ __forceinline__ void
srgb_to_labs(const float *s_v2Y, const float2 *__restrict__ cbrt2, int r8, int g8, int b8, int &L, int &A, int &B)
{
	float a = s_v2Y[r8], b = s_v2Y[g8], c = s_v2Y[b8];
	step_scRGB2XYZ(a, b, c); /* (the halo needs Y only: the compiler drops X and Z there) */
	const float nY = (float) DIVC((double) __fmul_rn(100000.0f, b), 100.0);
	const float cby = cbrt_lookup2(cbrt2, nY);
	const float fL = __fsub_rn(__fmul_rn(116.0F, cby), 16.0F);
	/* (short) VIPS_CLIP(0, L * 327.67, 32767), Lab2LabS.c:58-74: truncation commutes with the clip, and the
	 * saturating double -> int conversion plus an integer clamp is 3 instructions instead of 9
	 */
	L = max(0, min(__double2int_rz(__dmul_rn((double) fL, 32767.0 / 100.0)), 32767));
	if (WANT_AB) {
		const float nX = (float) DIVC((double) __fmul_rn(100000.0f, a), 95.0470);
		const float nZ = (float) DIVC((double) __fmul_rn(100000.0f, c), 108.8827);
		const float cbx = cbrt_lookup2(cbrt2, nX);
		const float cbz = cbrt_lookup2(cbrt2, nZ);
		const float fa = __fmul_rn(500.0F, __fsub_rn(cbx, cby));
		const float fb = __fmul_rn(200.0F, __fsub_rn(cby, cbz));
		A = max(-32768, min(__double2int_rz(__dmul_rn((double) fa, 32768.0 / 128.0)), 32767));
		B = max(-32768, min(__double2int_rz(__dmul_rn((double) fb, 32768.0 / 128.0)), 32767));
	}
}

template <int BANDS>
__global__ void __launch_bounds__(256)
sharpen_fused_kernel(const __grid_constant__ SharpenParams P, const uint8_t *__restrict__ in, uint8_t *__restrict__ out, int frame0)
{
	extern __shared__ __align__(16) unsigned char smem_raw[];
	const int r = P.n >> 1;
	const int HW = kTile + 2 * r, HH = kTile + 2 * r;
	short *sL = (short *) smem_raw;						   /* [HH][HW] L of tile + halo */
	short *sH = sL + ((HH * HW + 1) & ~1);				   /* [HH][kTile] after the horizontal pass */
	short2 *sAB = (short2 *) (sH + HH * kTile);			   /* [kTile][kTile] */
	float *s_v2Y = (float *) (sAB + kTile * kTile);		   /* [256] */
	float *s_Y2v = s_v2Y + 256;							   /* [257], integers held as floats */

	const int t = threadIdx.x;
	const int lx = t & 31, ly = t >> 5; /* 32 x 8 threads over the tile */
	/* i / HW for i < 64 * 64 without the integer divide: exact while i * (HW - 1) < 2^20 */
	const unsigned hw_magic = ((1u << 20) + HW - 1) / HW;
	for (int i = t; i < 257; i += 256) {
		if (i < 256)
			s_v2Y[i] = P.t.v2Y_8[i];
		s_Y2v[i] = (float) P.t.Y2v_8[i];
	}
	__syncthreads();

	const int x0 = blockIdx.x * kTile, y0 = blockIdx.y * kTile;
	const int frame = frame0 + blockIdx.z;
	const uint8_t *fin = in + (size_t) frame * P.in_frame_stride;
	uint8_t *fout = out + (size_t) frame * P.out_frame_stride;

	/* ---- 1: tile + halo to LabS (a / b only inside the tile) */
	for (int i = t; i < HH * HW; i += 256) { /* linear index: every lane busy (a 32-lane row walk leaves 30 idle on the halo columns) */
		const int hy = (int) (((unsigned) i * hw_magic) >> 20), hx = i - hy * HW;
		const int gy = max(0, min(y0 + hy - r, P.h - 1)), gx = max(0, min(x0 + hx - r, P.w - 1));
		const uint8_t *p = fin + (size_t) gy * P.in_bpl + (size_t) gx * BANDS;
		int r8, g8, b8;
		if (BANDS == 4) {
			const unsigned px = *(const unsigned *) p;
			r8 = px & 255;
			g8 = (px >> 8) & 255;
			b8 = (px >> 16) & 255;
		}
		else {
			r8 = p[0];
			g8 = p[1];
			b8 = p[2];
		}
		const int ty = hy - r, tx = hx - r;
		int L, A = 0, B = 0;
		if (ty >= 0 && ty < kTile && tx >= 0 && tx < kTile) {
			srgb_to_labs<true>(s_v2Y, P.t.cbrt2, r8, g8, b8, L, A, B);
			sAB[ty * kTile + tx] = make_short2((short) A, (short) B);
		}
		else
			srgb_to_labs<false>(s_v2Y, P.t.cbrt2, r8, g8, b8, L, A, B);
		sL[i] = (short) L;
	}
	__syncthreads();

	/* ---- 2: the n x 1 pass (convi.c:698-717 on shorts: int sum, (sum + scale / 2) / scale truncating, clip) */
	for (int hy = ly; hy < HH; hy += 8) {
		const short *row = sL + hy * HW + lx;
		int sum = 0;
		for (int k = 0; k < P.n; k++)
			sum += P.coef[k] * (int) row[k];
		sH[hy * kTile + lx] = (short) clip_short(div_scale(sum + P.rounding, P));
	}
	__syncthreads();

	/* ---- 3: the 1 x n pass, the LUT, and back to sRGB */
	for (int ty = ly; ty < kTile; ty += 8) {
		const int tx = lx, i = ty * kTile + lx;
		const int gx = x0 + tx, gy = y0 + ty;
		if (gx >= P.w || gy >= P.h)
			continue;
		int sum = 0;
		for (int k = 0; k < P.n; k++)
			sum += P.coef[k] * (int) sH[(ty + k) * kTile + tx];
		const int v2 = clip_short(div_scale(sum + P.rounding, P));
		const int v1 = sL[(ty + r) * HW + tx + r];
		/* sharpen.c:139-160 */
		const int diff = (v1 & 0x7fff) - (v2 & 0x7fff);
		int o = v1 + __ldg(P.lut + diff + 32768);
		o = max(0, min(o, 32767));
		const short2 ab = sAB[i];
		/* LabS2Lab, Lab2XYZ, XYZ2scRGB, scRGB2sRGB */
		float a = (float) DIVC((double) o, 32767.0 / 100.0);
		float b = (float) DIVC((double) ab.x, 32768.0 / 128.0);
		float c = (float) DIVC((double) ab.y, 32768.0 / 128.0);
		step_Lab2XYZ(a, b, c);
		step_XYZ2scRGB(a, b, c);
		unsigned R = 0, G = 0, Bb = 0;
		if (!(isnan(a) || isnan(b) || isnan(c))) {
			R = (unsigned) scRGB2sRGB_channel_f(s_Y2v, 255.0f, a) & 255u;
			G = (unsigned) scRGB2sRGB_channel_f(s_Y2v, 255.0f, b) & 255u;
			Bb = (unsigned) scRGB2sRGB_channel_f(s_Y2v, 255.0f, c) & 255u;
		}
		uint8_t *q = fout + (size_t) gy * P.out_bpl + (size_t) gx * BANDS;
		if (BANDS == 4) {
			/* the extra band: through both routes as vips_colour_build carries it */
			const uint8_t *p = fin + (size_t) gy * P.in_bpl + (size_t) gx * 4;
			double al = carry_extra_band((double) p[3], P.fwd, P.n_fwd);
			al = carry_extra_band(al, P.bwd, P.n_bwd);
			*(unsigned *) q = R | (G << 8) | (Bb << 16) | ((unsigned) (uint8_t) al << 24);
		}
		else {
			q[0] = (uint8_t) R;
			q[1] = (uint8_t) G;
			q[2] = (uint8_t) Bb;
		}
	}
}

/* the LUT of vips_sharpen_build (sharpen.c:227-257), cached per device and parameter set */
std::mutex g_lut_lock;
std::map<std::tuple<int, double, double, double, double, double>, int *> g_luts;

int
get_lut(const char *domain, double x1, double y2, double y3, double m1, double m2, const int **out)
{
	int dev = 0;
	VB200_CUDA(domain, cudaGetDevice(&dev));
	const auto key = std::make_tuple(dev, x1, y2, y3, m1, m2);
	std::lock_guard<std::mutex> lock(g_lut_lock);
	auto it = g_luts.find(key);
	if (it != g_luts.end()) {
		*out = it->second;
		return 0;
	}
	std::vector<int> lut(65536);
	for (int i = 0; i < 65536; i++) {
		const double v = (i - 32767) / 327.67;
		double y;
		if (v < -x1)
			y = (v + x1) * m2 + -x1 * m1;
		else if (v < x1)
			y = v * m1;
		else
			y = (v - x1) * m2 + x1 * m1;
		if (y < -y3)
			y = -y3;
		if (y > y2)
			y = y2;
		lut[i] = rint(y * 327.67);
	}
	if (g_luts.size() >= 64) {
		/* a caller sweeping parameters: do not grow without bound (no kernel can be using an entry:
		 * the upload below and the frees are ordered by the synchronising cudaFree)
		 */
		for (auto &e : g_luts)
			cudaFree(e.second);
		g_luts.clear();
	}
	int *d = nullptr;
	VB200_CUDA(domain, cudaMalloc(&d, lut.size() * sizeof(int)));
	VB200_CUDA(domain, cudaMemcpy(d, lut.data(), lut.size() * sizeof(int), cudaMemcpyHostToDevice));
	g_luts[key] = d;
	*out = d;
	return 0;
}

} // namespace

/* 0 = done, -1 = error, 1 = not eligible (the caller runs the unfused chain).  Frames are packed uchar sRGB,
 * 3 or 4 bands; in and out must not overlap (tiles read their neighbours' halo).
 */
int
dev_sharpen_fused(const char *domain, const void *in, size_t in_bpl, size_t in_frame_stride, void *out, size_t out_bpl,
	size_t out_frame_stride, int n_frames, int w, int h, int bands, double sigma, double x1, double y2, double y3, double m1,
	double m2, cudaStream_t s)
{
	if ((bands != 3 && bands != 4) || getenv("VB200_NO_SHARPEN_FUSED") != nullptr)
		return 1;
	if (bands == 4 && ((in_bpl | out_bpl | in_frame_stride | out_frame_stride | (uintptr_t) in | (uintptr_t) out) & 3) != 0)
		return 1;
	std::vector<double> m;
	int mw, mh;
	double scale;
	host_gaussmat(sigma, 0.1, true, true, m, &mw, &mh, &scale); /* sharpen.c:205-210 */
	if (mw > kMaxTaps || mw < 1)
		return 1;

	RouteParams fwd, bwd;
	if (colour_route_params(domain, VB200_INTERPRETATION_sRGB, VB200_INTERPRETATION_LABS, &fwd) ||
		colour_route_params(domain, VB200_INTERPRETATION_LABS, VB200_INTERPRETATION_sRGB, &bwd))
		return -1;
	SharpenParams P;
	memset(&P, 0, sizeof(P));
	memcpy(P.fwd, fwd.steps, sizeof(P.fwd));
	memcpy(P.bwd, bwd.steps, sizeof(P.bwd));
	P.n_fwd = fwd.n_steps;
	P.n_bwd = bwd.n_steps;
	P.t = fwd.t;
	if (get_lut(domain, x1, y2, y3, m1, m2, &P.lut))
		return -1;
	P.w = w;
	P.h = h;
	P.in_bpl = in_bpl;
	P.out_bpl = out_bpl;
	P.in_frame_stride = in_frame_stride;
	P.out_frame_stride = out_frame_stride;
	P.n = mw;
	long abs_sum = 0;
	for (int i = 0; i < mw; i++) {
		P.coef[i] = (int) rint(m[i]); /* vips__image_intize, convi.c:859-923 */
		abs_sum += labs((long) P.coef[i]);
	}
	P.scale = (int) rint(scale); /* convi.c:760-763 */
	P.rounding = P.scale / 2;
	if (P.scale <= 0)
		return 1; /* a negative scale: leave it to the general path */
	int bits = 0;
	while ((P.scale >> bits) != 0)
		bits++;
	P.shift = 31 + bits;
	P.magic = (unsigned long long) ((((unsigned __int128) 1 << P.shift) + P.scale - 1) / P.scale);
	if (P.scale == 0 || abs_sum * 32768 >= (1L << 31))
		return 1; /* the reference divides by it / sums in int64: keep the general path */

	const int r = mw / 2, hw = kTile + 2 * r;
	const size_t smem = (size_t) ((hw * hw + 1) & ~1) * 2 + (size_t) hw * kTile * 2 + (size_t) kTile * kTile * 4 + 256 * 4 + 257 * 4;
	for (int f0 = 0; f0 < n_frames; f0 += 32768) {
		const dim3 grid((w + kTile - 1) / kTile, (h + kTile - 1) / kTile, std::min(32768, n_frames - f0));
		if (bands == 4)
			sharpen_fused_kernel<4><<<grid, 256, smem, s>>>(P, (const uint8_t *) in, (uint8_t *) out, f0);
		else
			sharpen_fused_kernel<3><<<grid, 256, smem, s>>>(P, (const uint8_t *) in, (uint8_t *) out, f0);
		cudaError_t e = cudaGetLastError();
		if (e != cudaSuccess)
			return cuda_fail(domain, e, "sharpen_fused_kernel launch");
		count_launch();
	}
	return 0;
}

} // namespace vb200

using namespace vb200;

/* A batch of same-shaped 8-bit sRGB frames on the device, one kernel: the second stage of the
 * thumbnail + sharpen stream (BASELINE config 5), also usable on its own.
 */
extern "C" int
vb200_sharpen_batch_device(const void *in, size_t in_frame_stride, void *out, size_t out_frame_stride, int n_frames, int width,
	int height, int bands, double sigma, double x1, double y2, double y3, double m1, double m2)
{
	const char *domain = "sharpen_batch_device";
	if (!in || !out || n_frames < 0 || width <= 0 || height <= 0) {
		error(domain, "bad argument");
		return -1;
	}
	if (ensure_init(domain))
		return -1;
	const size_t line = (size_t) width * bands;
	const char *a = (const char *) in, *o = (const char *) out;
	const size_t in_span = n_frames ? (size_t) (n_frames - 1) * in_frame_stride + line * height : 0;
	const size_t out_span = n_frames ? (size_t) (n_frames - 1) * out_frame_stride + line * height : 0;
	if (o < a + in_span && a < o + out_span) {
		error(domain, "in and out overlap");
		return -1;
	}
	const int rc = dev_sharpen_fused(domain, in, line, in_frame_stride, out, line, out_frame_stride, n_frames, width, height, bands,
		sigma, x1, y2, y3, m1, m2, current_stream());
	if (rc == 1) {
		error(domain, "only 3- or 4-band 8-bit sRGB frames with a mask of at most %d taps are on the batched path", kMaxTaps);
		return -1;
	}
	return rc;
}
