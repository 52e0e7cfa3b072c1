/* chain.cu -- a pump for UNFUSED operation graphs.
 *
 * The reference evaluates an arbitrary chain of operations by pulling tiles through every op's
 * generate() from a sink (iofuncs/sinkmemory.c:324, threadpool.c:625).  On the device the unit is the
 * whole image: a VB200Chain is a list of the operations this library implements; running it over a
 * batch of host images uploads image i on one of three streams, runs every op leaf by leaf on DEVICE
 * images of that stream (intermediates come from the stream-ordered pool and never visit the host) and
 * downloads the result, while images i + 1 and i + 2 are in their own upload / compute / download
 * phases on the other streams.  The same ops called one by one through vb200_resize(), vb200_conv() ...
 * on host images cost a synchronous host round trip each.
 *
 * Images of one batch may differ in size and format; every op keeps the exact arithmetic of its
 * stand-alone entry point (it IS the same dev_* function).
 */
#include <cstring>
#include <mutex>
#include <vector>

#include "vb200_internal.h"

using namespace vb200;

namespace {

enum ChainOp { C_RESIZE, C_REDUCE, C_COLOURSPACE, C_CONV, C_CONVSEP, C_GAUSSBLUR, C_SHARPEN, C_PREMULTIPLY, C_UNPREMULTIPLY, C_MORPH, C_RANK, C_FLATTEN };

struct ChainStep {
	ChainOp op = C_RESIZE;
	double d[6] = {0, 0, 0, 0, 0, 0};
	int i[3] = {0, 0, 0};
	std::vector<double> mask;
	int mw = 0, mh = 0;
};

constexpr int kChainStreams = 3;

} // namespace

struct VB200Chain {
	std::vector<ChainStep> steps;
	cudaStream_t streams[kChainStreams] = {nullptr, nullptr, nullptr};
	std::mutex lock;
};

extern "C" VB200Chain *
vb200_chain_new(void)
{
	if (ensure_init("chain"))
		return nullptr;
	return new VB200Chain();
}

extern "C" void
vb200_chain_free(VB200Chain *chain)
{
	if (!chain)
		return;
	for (auto &s : chain->streams)
		if (s) {
			cudaStreamSynchronize(s);
			cudaStreamDestroy(s);
		}
	delete chain;
}

static int
chain_push(VB200Chain *chain, ChainStep &&st)
{
	if (!chain) {
		error("chain", "null chain");
		return -1;
	}
	chain->steps.push_back(std::move(st));
	return 0;
}

extern "C" int
vb200_chain_add_resize(VB200Chain *chain, double scale, double vscale, int kernel, double gap)
{
	ChainStep st;
	st.op = C_RESIZE;
	st.d[0] = scale;
	st.d[1] = vscale > 0 ? vscale : scale;
	st.d[2] = gap < 0 ? 2.0 : gap;
	st.i[0] = kernel;
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_reduce(VB200Chain *chain, double hshrink, double vshrink, int kernel, double gap)
{
	ChainStep st;
	st.op = C_REDUCE;
	st.d[0] = hshrink;
	st.d[1] = vshrink;
	st.d[2] = gap;
	st.i[0] = kernel;
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_colourspace(VB200Chain *chain, int space)
{
	ChainStep st;
	st.op = C_COLOURSPACE;
	st.i[0] = space;
	return chain_push(chain, std::move(st));
}

static int
chain_add_mask(VB200Chain *chain, ChainOp op, const VB200Mask *mask, int precision)
{
	if (!mask || !mask->coeff || mask->width <= 0 || mask->height <= 0) {
		error("chain", "no mask");
		return -1;
	}
	if (op == C_CONVSEP && mask->width != 1 && mask->height != 1) {
		error("chain", "mask must be 1xn or nx1 elements");
		return -1;
	}
	ChainStep st;
	st.op = op;
	st.mask.assign(mask->coeff, mask->coeff + (size_t) mask->width * mask->height);
	st.mw = mask->width;
	st.mh = mask->height;
	st.d[0] = mask->scale;
	st.d[1] = mask->offset;
	st.i[0] = precision;
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_conv(VB200Chain *chain, const VB200Mask *mask, int precision)
{
	return chain_add_mask(chain, C_CONV, mask, precision);
}

extern "C" int
vb200_chain_add_convsep(VB200Chain *chain, const VB200Mask *mask, int precision)
{
	return chain_add_mask(chain, C_CONVSEP, mask, precision);
}

extern "C" int
vb200_chain_add_morph(VB200Chain *chain, const VB200Mask *mask, int morph)
{
	return chain_add_mask(chain, C_MORPH, mask, morph);
}

extern "C" int
vb200_chain_add_flatten(VB200Chain *chain, const double *background, int n, double max_alpha)
{
	ChainStep st;
	st.op = C_FLATTEN;
	if (background && n > 0)
		st.mask.assign(background, background + n);
	st.d[0] = max_alpha;
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_rank(VB200Chain *chain, int width, int height, int index)
{
	ChainStep st;
	st.op = C_RANK;
	st.i[0] = width;
	st.i[1] = height;
	st.i[2] = index;
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_gaussblur(VB200Chain *chain, double sigma, double min_ampl, int precision)
{
	ChainStep st;
	st.op = C_GAUSSBLUR;
	st.d[0] = sigma;
	st.d[1] = min_ampl <= 0 ? 0.2 : min_ampl;
	st.i[0] = precision;
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_sharpen(VB200Chain *chain, double sigma, double x1, double y2, double y3, double m1, double m2)
{
	ChainStep st;
	st.op = C_SHARPEN;
	const double v[6] = {sigma, x1, y2, y3, m1, m2};
	memcpy(st.d, v, sizeof(v));
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_premultiply(VB200Chain *chain, double max_alpha, int uchar_mode)
{
	ChainStep st;
	st.op = C_PREMULTIPLY;
	st.d[0] = max_alpha;
	st.i[0] = uchar_mode;
	return chain_push(chain, std::move(st));
}

extern "C" int
vb200_chain_add_unpremultiply(VB200Chain *chain, double max_alpha, int uchar_mode)
{
	ChainStep st;
	st.op = C_UNPREMULTIPLY;
	st.d[0] = max_alpha;
	st.i[0] = uchar_mode;
	return chain_push(chain, std::move(st));
}

static int
chain_step(const char *domain, const ChainStep &st, const DevImage &in, DevImage *out, cudaStream_t s)
{
	switch (st.op) {
	case C_RESIZE:
		return dev_resize(domain, in, out, st.d[0], st.d[1], st.i[0], st.d[2], s);
	case C_REDUCE:
		if (st.d[0] < 1.0 || st.d[1] < 1.0) {
			error(domain, "reduce factor should be >= 1.0");
			return -1;
		}
		return dev_reduce_chain(domain, in, out, st.d[0], st.d[1], st.i[0], st.d[2], s);
	case C_COLOURSPACE:
		return dev_colourspace(domain, in, out, st.i[0], in.type, s);
	case C_CONV:
		return dev_conv(domain, in, out, st.mask.data(), st.mw, st.mh, st.d[0], st.d[1], st.i[0], s, true);
	case C_CONVSEP:
		return dev_convsep(domain, in, out, st.mask.data(), st.mw, st.mh, st.d[0], st.d[1], st.i[0], s, true);
	case C_MORPH:
		return dev_morph(domain, in, out, st.mask.data(), st.mw, st.mh, st.i[0], s);
	case C_FLATTEN:
		return dev_flatten(domain, in, out, st.mask.empty() ? nullptr : st.mask.data(), (int) st.mask.size(), st.d[0], s);
	case C_RANK:
		return dev_rank(domain, in, out, st.i[0], st.i[1], st.i[2], s);
	case C_GAUSSBLUR:
		return dev_gaussblur(domain, in, out, st.d[0], st.d[1], st.i[0], s);
	case C_SHARPEN:
		return dev_sharpen(domain, in, out, st.d[0], st.d[1], st.d[2], st.d[3], st.d[4], st.d[5], s);
	case C_PREMULTIPLY:
		return dev_premultiply(domain, in, out, st.d[0], st.i[0], s);
	case C_UNPREMULTIPLY:
		return dev_unpremultiply(domain, in, out, st.d[0], st.i[0], s);
	}
	return -1;
}

/* in[i]: host images (pinned memory lets the three phases overlap; pageable memory is correct but its
 * copies are staged synchronously).  out[i]: data == NULL -> malloc'ed by the library (vb200_image_free),
 * else the caller's buffer (it must be large enough: run the chain once with NULL to learn the geometry).
 * Returns after the last result has landed.
 */
extern "C" int
vb200_chain_run_host(VB200Chain *chain, const VB200Image *in, VB200Image *out, int n_images)
{
	const char *domain = "chain_run_host";
	if (!chain || !in || !out || n_images < 0) {
		error(domain, "bad argument");
		return -1;
	}
	if (ensure_init(domain))
		return -1;
	std::lock_guard<std::mutex> lock(chain->lock);
	for (auto &s : chain->streams)
		if (!s)
			VB200_CUDA(domain, cudaStreamCreateWithFlags(&s, cudaStreamNonBlocking));

	int rc = 0;
	for (int i = 0; i < n_images && !rc; i++) {
		cudaStream_t s = chain->streams[i % kChainStreams];
		if (in[i].where != VB200_HOST) {
			error(domain, "image %d is not a host image", i);
			rc = -1;
			break;
		}
		DevImage cur;
		if (to_device(domain, &in[i], &cur, s)) { /* cudaMemcpy2DAsync on s */
			rc = -1;
			break;
		}
		for (const ChainStep &st : chain->steps) {
			DevImage next;
			if (chain_step(domain, st, cur, &next, s)) {
				dev_image_release(&cur, s);
				rc = -1;
				break;
			}
			if (next.data == cur.data) {
				/* a pass-through step (e.g. premultiply of a 1-band image) handed its input on */
				next.owned = cur.owned;
				cur.owned = false;
			}
			dev_image_release(&cur, s);
			cur = next;
		}
		if (rc)
			break;
		/* download: async on s; the buffer is released to the pool in stream order */
		const size_t line = (size_t) cur.w * cur.bands * format_sizeof(cur.fmt);
		VB200Image *o = &out[i];
		void *dst = o->data;
		size_t dst_bpl = (o->data && o->bpl) ? o->bpl : line;
		if (!dst) {
			dst = malloc(line * cur.h > 0 ? line * cur.h : 1);
			if (!dst) {
				error(domain, "out of memory");
				dev_image_release(&cur, s);
				rc = -1;
				break;
			}
		}
		o->Xsize = cur.w;
		o->Ysize = cur.h;
		o->Bands = cur.bands;
		o->BandFmt = cur.fmt;
		o->Type = cur.type;
		o->where = VB200_HOST;
		o->data = dst;
		o->bpl = dst_bpl;
		if (cudaMemcpy2DAsync(dst, dst_bpl, cur.data, cur.bpl, line, cur.h, cudaMemcpyDeviceToHost, s) != cudaSuccess)
			rc = cuda_fail(domain, cudaGetLastError(), "chain download");
		dev_image_release(&cur, s);
	}
	for (auto &s : chain->streams)
		if (s && cudaStreamSynchronize(s) != cudaSuccess && !rc)
			rc = cuda_fail(domain, cudaGetLastError(), "chain sync");
	return rc;
}
