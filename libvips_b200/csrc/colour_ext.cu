/* colour_ext.cu -- the B_W, GREY16 and HSV rows of vips_colourspace on the device.
 *
 * reference: colour/colourspace.c:223-497.  Every row that STARTS in one of these spaces opens with
 *     vips_BW2sRGB / vips_GREY162RGB16 (:152-188: vips__colourspace_process_n over a bandjoin of the first band with
 *     itself, other bands re-attached; the format stays, the Type becomes sRGB / RGB16) or vips_HSV2sRGB (HSV2sRGB.c:54-108)
 * and goes on as the sRGB / RGB16 row does; every row that ENDS there is the row to scRGB followed by
 *     vips_scRGB2BW (scRGB2BW.c:58-105 over vips_col_scRGB2BW, LabQ2sRGB.c:385-429: Y = 0.2126 R + 0.7152 G + 0.0722 B in
 *     float, then the interpolated gamma table of scRGB -> sRGB; depth 8 -> B_W uchar, depth 16 -> GREY16 ushort)
 * or the row to sRGB followed by vips_sRGB2HSV (sRGB2HSV.c:50-126).  dev_colourspace_ext composes exactly those rows out
 * of the route kernels of colour.cu (untouched) and three leaf kernels here:
 *   replicate_kernel   one band -> three (+ extra bands copied)
 *   hsv_kernel         uchar sRGB <-> uchar HSV; the per-pixel arithmetic (float / double mix as the reference writes it,
 *                      explicit round-to-nearest operations, fmodf(x, 2) as the exact x - 2 floor(x / 2)) is
 *                      __host__ __device__: vb200_debug_hsv_host runs it on the CPU over all 2^24 inputs
 *   bw_kernel          sRGB uchar / RGB16 ushort / scRGB float -> one band: for integer sources sRGB2scRGB (the v2Y tables)
 *                      is fused in, so sRGB -> B_W -- the common call -- is one launch reading 3 bytes and writing 1 per pixel
 * Extra bands follow vips_colour_build per step (colour.c:252-291) through carry_extra_band, as in the route kernel.
 * The device path wants the format the source space implies (uchar B_W / HSV, ushort GREY16), like dev_colourspace.
 */
#include <cstring>

#include "colour_steps.cuh"

namespace vb200 {

namespace {

constexpr int BW = VB200_INTERPRETATION_B_W, GREY16 = VB200_INTERPRETATION_GREY16, HSV = VB200_INTERPRETATION_HSV;
constexpr int sRGB = VB200_INTERPRETATION_sRGB, RGB16 = VB200_INTERPRETATION_RGB16, scRGB = VB200_INTERPRETATION_scRGB;

#ifdef __CUDA_ARCH__
#define CX_FMUL(a, b) __fmul_rn((a), (b))
#define CX_FADD(a, b) __fadd_rn((a), (b))
#define CX_FSUB(a, b) __fsub_rn((a), (b))
#define CX_FDIV(a, b) __fdiv_rn((a), (b))
#define CX_DMUL(a, b) __dmul_rn((a), (b))
#define CX_DADD(a, b) __dadd_rn((a), (b))
#define CX_DDIV(a, b) __ddiv_rn((a), (b))
#define CX_D2F(a) __double2float_rn(a)
#else
#define CX_FMUL(a, b) ((float) (a) * (float) (b))
#define CX_FADD(a, b) ((float) (a) + (float) (b))
#define CX_FSUB(a, b) ((float) (a) - (float) (b))
#define CX_FDIV(a, b) ((float) (a) / (float) (b))
#define CX_DMUL(a, b) ((double) (a) * (double) (b))
#define CX_DADD(a, b) ((double) (a) + (double) (b))
#define CX_DDIV(a, b) ((double) (a) / (double) (b))
#define CX_D2F(a) ((float) (a))
#endif

/* vips_sRGB2HSV_line's body, sRGB2HSV.c:57-122 */
__host__ __device__ __forceinline__ void
srgb2hsv_px(const uint8_t *p, uint8_t *q)
{
	int c_max, c_min;
	float secondary_diff, wrap_around_hue;
	if (p[1] < p[2]) {
		if (p[2] < p[0]) {
			c_max = p[0];
			c_min = p[1];
			secondary_diff = (float) (p[1] - p[2]);
			wrap_around_hue = 255.0f;
		}
		else {
			c_max = p[2];
			c_min = p[1] < p[0] ? p[1] : p[0];
			secondary_diff = (float) (p[0] - p[1]);
			wrap_around_hue = 170.0f;
		}
	}
	else {
		if (p[1] < p[0]) {
			c_max = p[0];
			c_min = p[2];
			secondary_diff = (float) (p[1] - p[2]);
			wrap_around_hue = 0.0f;
		}
		else {
			c_max = p[1];
			c_min = p[2] < p[0] ? p[2] : p[0];
			secondary_diff = (float) (p[2] - p[0]);
			wrap_around_hue = 85.0f;
		}
	}
	if (c_max == 0) {
		q[0] = q[1] = q[2] = 0;
		return;
	}
	const int delta = c_max - c_min;
	q[2] = (uint8_t) c_max;
	if (delta == 0)
		q[0] = 0;
	else /* 42.5 * (float quotient) + hue: double arithmetic, stored by truncation; always within 0 .. 255 */
		q[0] = (uint8_t) (int) CX_DADD(CX_DMUL(42.5, (double) CX_FDIV(secondary_diff, (float) delta)), (double) wrap_around_hue);
	q[1] = (uint8_t) (int) CX_DDIV((double) delta * 255.0, (double) (float) c_max);
}

/* vips_HSV2sRGB_line's body, HSV2sRGB.c:62-104; SIXTH_OF_CHAR is the double 42.5 */
__host__ __device__ __forceinline__ void
hsv2srgb_px(const uint8_t *p, uint8_t *q)
{
	const float c = CX_D2F(CX_DDIV((double) ((int) p[2] * (int) p[1]), 255.0));
	const float h = CX_D2F(CX_DDIV((double) p[0], 42.5)); /* the argument of fmodf: 0 <= h <= 6 */
	const float h2 = CX_FMUL(h, 0.5f);
	const float fl = (float) (int) h2;				  /* floorf, h2 >= 0 */
	const float r = CX_FSUB(h, CX_FMUL(2.0f, fl)); /* fmodf(h, 2): every step exact */
	float t = CX_FSUB(r, 1.0f);
	t = t < 0 ? -t : t;
	const float x = CX_FMUL(c, CX_FSUB(1.0f, t));
	const float m = CX_FSUB((float) p[2], c);
	const float cm = CX_FADD(c, m), xm = CX_FADD(x, m), zm = CX_FADD(0.0f, m);
	float r0, r1, r2;
	if (p[0] < 42) {
		r0 = cm, r1 = xm, r2 = zm;
	}
	else if (p[0] < 85) {
		r0 = xm, r1 = cm, r2 = zm;
	}
	else if (p[0] < 127) {
		r0 = zm, r1 = cm, r2 = xm;
	}
	else if (p[0] < 170) {
		r0 = zm, r1 = xm, r2 = cm;
	}
	else if (p[0] < 212) {
		r0 = xm, r1 = zm, r2 = cm;
	}
	else {
		r0 = cm, r1 = zm, r2 = xm;
	}
	q[0] = (uint8_t) (int) r0;
	q[1] = (uint8_t) (int) r1;
	q[2] = (uint8_t) (int) r2;
}

struct ExtDev {
	int w, h, in_bands, out_bands, in_fmt, out_fmt, depth, n_steps;
	size_t in_bpl, out_bpl;
	ColourTables t;
	StepInfo steps[2]; /* what an extra band goes through (vips_colour_build per fused step) */
};

template <bool TO_HSV>
__global__ void __launch_bounds__(256)
hsv_kernel(const __grid_constant__ ExtDev P, const unsigned char *__restrict__ in, unsigned char *__restrict__ out)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= P.w)
		return;
	for (int y = blockIdx.y; y < P.h; y += gridDim.y) {
		const uint8_t *p = in + (size_t) y * P.in_bpl + (size_t) x * P.in_bands;
		uint8_t *q = out + (size_t) y * P.out_bpl + (size_t) x * P.out_bands;
		const uint8_t px[3] = {p[0], p[1], p[2]};
		uint8_t res[3];
		if (TO_HSV)
			srgb2hsv_px(px, res);
		else
			hsv2srgb_px(px, res);
		q[0] = res[0];
		q[1] = res[1];
		q[2] = res[2];
		for (int e = 3; e < P.in_bands; e++) /* 255 -> 255, uchar -> uchar: a copy */
			q[e] = p[e];
	}
}

/* T: element type (1 or 2 bytes) */
template <typename T>
__global__ void __launch_bounds__(256)
replicate_kernel(const __grid_constant__ ExtDev P, const unsigned char *__restrict__ in, unsigned char *__restrict__ out)
{
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= P.w)
		return;
	for (int y = blockIdx.y; y < P.h; y += gridDim.y) {
		const T *p = reinterpret_cast<const T *>(in + (size_t) y * P.in_bpl) + (size_t) x * P.in_bands;
		T *q = reinterpret_cast<T *>(out + (size_t) y * P.out_bpl) + (size_t) x * P.out_bands;
		const T v = p[0];
		q[0] = v;
		q[1] = v;
		q[2] = v;
		for (int e = 1; e < P.in_bands; e++)
			q[2 + e] = p[e];
	}
}

__global__ void __launch_bounds__(256)
bw_kernel(const __grid_constant__ ExtDev P, const unsigned char *__restrict__ in, unsigned char *__restrict__ out)
{
	__shared__ float s_v2Y_8[256];
	if (P.in_fmt == VB200_FORMAT_UCHAR) {
		for (int i = threadIdx.x; i < 256; i += blockDim.x)
			s_v2Y_8[i] = P.t.v2Y_8[i];
		__syncthreads();
	}
	const int x = blockIdx.x * blockDim.x + threadIdx.x;
	if (x >= P.w)
		return;
	for (int y = blockIdx.y; y < P.h; y += gridDim.y) {
		const unsigned char *pin = in + (size_t) y * P.in_bpl;
		unsigned char *pout = out + (size_t) y * P.out_bpl;
		const int ib = x * P.in_bands, ob = x * P.out_bands;
		float R, G, B;
		if (P.in_fmt == VB200_FORMAT_UCHAR) { /* vips_sRGB2scRGB_line, 8 bit: the table */
			R = s_v2Y_8[pin[ib]];
			G = s_v2Y_8[pin[ib + 1]];
			B = s_v2Y_8[pin[ib + 2]];
		}
		else if (P.in_fmt == VB200_FORMAT_USHORT) {
			const uint16_t *p16 = reinterpret_cast<const uint16_t *>(pin);
			R = __ldg(P.t.v2Y_16 + p16[ib]);
			G = __ldg(P.t.v2Y_16 + p16[ib + 1]);
			B = __ldg(P.t.v2Y_16 + p16[ib + 2]);
		}
		else {
			const float *pf = reinterpret_cast<const float *>(pin);
			R = pf[ib];
			G = pf[ib + 1];
			B = pf[ib + 2];
		}
		/* LabQ2sRGB.c:400, left to right, no contraction */
		const float Y = __fadd_rn(__fadd_rn(__fmul_rn(0.2126f, R), __fmul_rn(0.7152f, G)), __fmul_rn(0.0722f, B));
		int g = 0;
		if (!isnan(Y))
			g = P.depth == 16 ? scRGB2sRGB_channel(P.t.Y2v_16, 65535, Y) : scRGB2sRGB_channel(P.t.Y2v_8, 255, Y);
		if (P.depth == 16)
			reinterpret_cast<uint16_t *>(pout)[ob] = (uint16_t) g;
		else
			pout[ob] = (uint8_t) g;
		for (int e = 3; e < P.in_bands; e++)
			store_elem(pout, P.out_fmt, ob + e - 2, carry_extra_band(load_elem(pin, P.in_fmt, ib + e), P.steps, P.n_steps));
	}
}

dim3
ext_grid(int w, int h)
{
	return dim3((w + 255) / 256, h < 4096 ? h : 4096);
}

int
launched(const char *domain, const char *what)
{
	cudaError_t e = cudaGetLastError();
	if (e != cudaSuccess)
		return cuda_fail(domain, e, what);
	count_launch();
	return 0;
}

void
ext_geometry(ExtDev *P, const DevImage &in, const DevImage &out)
{
	P->w = in.w;
	P->h = in.h;
	P->in_bands = in.bands;
	P->out_bands = out.bands;
	P->in_fmt = in.fmt;
	P->out_fmt = out.fmt;
	P->in_bpl = in.bpl;
	P->out_bpl = out.bpl;
}

int
want_format(const char *domain, const DevImage &in, int space, int fmt)
{
	if (in.fmt != fmt) {
		/* the reference would insert a vips_cast first (colour.c:421-428, 338-342) */
		error(domain, "source space %d wants band format %d, image has %d", space, fmt, in.fmt);
		return -1;
	}
	return 0;
}

/* vips_BW2sRGB / vips_GREY162RGB16 */
int
dev_replicate(const char *domain, const DevImage &in, DevImage *out, int type, cudaStream_t s)
{
	if (dev_image_new(domain, out, in.w, in.h, in.bands + 2, in.fmt, type, s))
		return -1;
	ExtDev P;
	memset(&P, 0, sizeof(P));
	ext_geometry(&P, in, *out);
	if (in.fmt == VB200_FORMAT_UCHAR)
		replicate_kernel<uint8_t><<<ext_grid(in.w, in.h), 256, 0, s>>>(P, (const unsigned char *) in.data, (unsigned char *) out->data);
	else
		replicate_kernel<uint16_t><<<ext_grid(in.w, in.h), 256, 0, s>>>(P, (const unsigned char *) in.data, (unsigned char *) out->data);
	return launched(domain, "replicate_kernel");
}

int
dev_hsv(const char *domain, const DevImage &in, DevImage *out, bool to_hsv, cudaStream_t s)
{
	if (in.bands < 3) {
		error(domain, "too few bands for operation");
		return -1;
	}
	if (dev_image_new(domain, out, in.w, in.h, in.bands, VB200_FORMAT_UCHAR, to_hsv ? HSV : sRGB, s))
		return -1;
	ExtDev P;
	memset(&P, 0, sizeof(P));
	ext_geometry(&P, in, *out);
	if (to_hsv)
		hsv_kernel<true><<<ext_grid(in.w, in.h), 256, 0, s>>>(P, (const unsigned char *) in.data, (unsigned char *) out->data);
	else
		hsv_kernel<false><<<ext_grid(in.w, in.h), 256, 0, s>>>(P, (const unsigned char *) in.data, (unsigned char *) out->data);
	return launched(domain, "hsv_kernel");
}

/* [sRGB2scRGB +] scRGB2BW: in is sRGB uchar, RGB16 ushort or scRGB float */
int
dev_bw(const char *domain, const DevImage &in, int in_space, DevImage *out, int depth, cudaStream_t s)
{
	if (in.bands < 3) {
		error(domain, "too few bands for operation");
		return -1;
	}
	ExtDev P;
	memset(&P, 0, sizeof(P));
	if (get_tables(domain, &P.t))
		return -1;
	const int out_fmt = depth == 16 ? VB200_FORMAT_USHORT : VB200_FORMAT_UCHAR;
	if (dev_image_new(domain, out, in.w, in.h, in.bands - 2, out_fmt, depth == 16 ? GREY16 : BW, s))
		return -1;
	ext_geometry(&P, in, *out);
	P.depth = depth;
	int n = 0;
	double before = interpretation_max_alpha(in_space);
	if (in_space != scRGB) {
		P.steps[n].step = S_sRGB2scRGB;
		P.steps[n].out_fmt = VB200_FORMAT_FLOAT;
		P.steps[n].rescale = before != 1.0;
		P.steps[n].alpha_a = (float) (1.0 / before);
		before = 1.0;
		n++;
	}
	const double after = depth == 16 ? 65535.0 : 255.0;
	P.steps[n].step = 0;
	P.steps[n].out_fmt = out_fmt;
	P.steps[n].rescale = before != after;
	P.steps[n].alpha_a = (float) (after / before);
	P.n_steps = n + 1;
	bw_kernel<<<ext_grid(in.w, in.h), 256, 0, s>>>(P, (const unsigned char *) in.data, (unsigned char *) out->data);
	return launched(domain, "bw_kernel");
}

bool
is_ext(int space)
{
	return space == BW || space == GREY16 || space == HSV;
}

} // namespace

bool
colour_ext_space(int space)
{
	return is_ext(space);
}

int
dev_colourspace_ext(const char *domain, const DevImage &in, DevImage *out, int space, int source_space, cudaStream_t s)
{
	const int src_fmt = source_space == GREY16 ? VB200_FORMAT_USHORT : VB200_FORMAT_UCHAR;
	if (source_space == space) {
		/* the identity rows are casts to the space's format (colourspace.c:413, 430, 447): on this path, a copy */
		if (want_format(domain, in, source_space, src_fmt) || dev_image_new(domain, out, in.w, in.h, in.bands, in.fmt, space, s))
			return -1;
		VB200_CUDA(domain, cudaMemcpy2DAsync(out->data, out->bpl, in.data, in.bpl, (size_t) in.w * in.bands * format_sizeof(in.fmt), in.h,
			cudaMemcpyDeviceToDevice, s));
		return 0;
	}
	DevImage t1, t2;
	const DevImage *cur = &in;
	int cur_space = source_space;
	int rc = 0;
	/* leave B_W / GREY16 / HSV */
	if (is_ext(source_space)) {
		rc = want_format(domain, in, source_space, src_fmt);
		const int hub = source_space == GREY16 ? RGB16 : sRGB;
		DevImage *dst = space == hub ? out : &t1; /* B_W -> sRGB, GREY16 -> RGB16, HSV -> sRGB: the whole row */
		if (!rc)
			rc = source_space == HSV ? dev_hsv(domain, in, dst, false, s) : dev_replicate(domain, in, dst, hub, s);
		if (rc || space == hub) {
			dev_image_release(&t1, s);
			return rc;
		}
		cur = &t1;
		cur_space = hub;
	}
	/* arrive */
	if (!is_ext(space))
		rc = dev_colourspace(domain, *cur, out, space, cur_space, s);
	else if (space == HSV) {
		if (cur_space != sRGB) {
			rc = dev_colourspace(domain, *cur, &t2, sRGB, cur_space, s);
			cur = &t2;
		}
		if (!rc)
			rc = dev_hsv(domain, *cur, out, true, s);
	}
	else {
		if (cur_space != sRGB && cur_space != RGB16 && cur_space != scRGB) {
			rc = dev_colourspace(domain, *cur, &t2, scRGB, cur_space, s);
			cur = &t2;
			cur_space = scRGB;
		}
		if (!rc && cur_space != scRGB)
			rc = want_format(domain, *cur, cur_space, cur_space == RGB16 ? VB200_FORMAT_USHORT : VB200_FORMAT_UCHAR);
		if (!rc && cur_space == scRGB)
			rc = want_format(domain, *cur, cur_space, VB200_FORMAT_FLOAT);
		if (!rc)
			rc = dev_bw(domain, *cur, cur_space, out, space == GREY16 ? 16 : 8, s);
	}
	dev_image_release(&t1, s);
	dev_image_release(&t2, s);
	return rc;
}

} // namespace vb200

using namespace vb200;

/* test hook, host only: hsv_kernel's per-pixel code on the CPU; n pixels of 3 bytes; to_hsv != 0: sRGB -> HSV */
extern "C" int
vb200_debug_hsv_host(const void *in, size_t n, int to_hsv, void *out)
{
	if (!in || !out) {
		error("colourspace", "null argument");
		return -1;
	}
	const uint8_t *p = (const uint8_t *) in;
	uint8_t *q = (uint8_t *) out;
	for (size_t i = 0; i < n; i++, p += 3, q += 3)
		if (to_hsv)
			srgb2hsv_px(p, q);
		else
			hsv2srgb_px(p, q);
	return 0;
}
