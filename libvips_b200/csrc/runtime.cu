#include <algorithm>
/* runtime.cu -- init, error buffer, stream and memory plumbing of libvb200.so.
 *
 * Mirrors the reference's conventions: 0 / -1 returns with a text buffer
 * (libvips/iofuncs/error.c:120-350), an on/off switch like
 * vips_vector_isenabled() (iofuncs/vector.cpp:83-110, env VB200_DISABLE like
 * VIPS_NOVECTOR), and tile-geometry globals (iofuncs/thread.c:74-77).
 */
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <mutex>
#include <string>

#include "vb200_internal.h"

namespace vb200 {

static thread_local std::string g_error;
static thread_local cudaStream_t g_stream = nullptr;
static std::atomic<uint64_t> g_launches{0};
static std::atomic<int> g_device{-1};
static std::mutex g_init_lock;
/* reference: VIPS__TILE_WIDTH/HEIGHT 128, FATSTRIP 16, THINSTRIP 1,
 * include/vips/private.h:147-153
 */
static TileGeometry g_tiles = {128, 128, 16, 1};

void
error(const char *domain, const char *fmt, ...)
{
	char buf[1024];
	va_list ap;
	va_start(ap, fmt);
	vsnprintf(buf, sizeof(buf), fmt, ap);
	va_end(ap);
	g_error += domain;
	g_error += ": ";
	g_error += buf;
	g_error += "\n";
}

int
cuda_fail(const char *domain, cudaError_t e, const char *what)
{
	error(domain, "CUDA error %s (%s) in %s", cudaGetErrorName(e), cudaGetErrorString(e), what);
	return -1;
}

cudaStream_t
current_stream()
{
	return g_stream;
}

void
count_launch(int n)
{
	g_launches.fetch_add(n, std::memory_order_relaxed);
}

TileGeometry
tile_geometry()
{
	return g_tiles;
}

int
ensure_init(const char *domain)
{
	if (g_device.load() >= 0)
		return 0;
	return vb200_init(0) ? (error(domain, "libvb200 is not initialised and no CUDA device is usable"), -1) : 0;
}

int
dev_alloc(const char *domain, void **p, size_t bytes, cudaStream_t s)
{
	*p = nullptr;
	if (bytes == 0)
		bytes = 16;
	VB200_CUDA(domain, cudaMallocAsync(p, bytes, s));
	return 0;
}

void
dev_free(void *p, cudaStream_t s)
{
	if (p)
		cudaFreeAsync(p, s);
}

size_t
format_sizeof(int fmt)
{
	switch (fmt) {
	case VB200_FORMAT_UCHAR:
	case VB200_FORMAT_CHAR:
		return 1;
	case VB200_FORMAT_USHORT:
	case VB200_FORMAT_SHORT:
		return 2;
	case VB200_FORMAT_UINT:
	case VB200_FORMAT_INT:
	case VB200_FORMAT_FLOAT:
		return 4;
	case VB200_FORMAT_COMPLEX:
	case VB200_FORMAT_DOUBLE:
		return 8;
	case VB200_FORMAT_DPCOMPLEX:
		return 16;
	}
	return 0;
}

bool
format_is_supported(int fmt)
{
	switch (fmt) {
	case VB200_FORMAT_UCHAR:
	case VB200_FORMAT_CHAR:
	case VB200_FORMAT_USHORT:
	case VB200_FORMAT_SHORT:
	case VB200_FORMAT_UINT:
	case VB200_FORMAT_INT:
	case VB200_FORMAT_FLOAT:
		return true;
	}
	return false;
}

double
interpretation_max_alpha(int type)
{
	/* reference: vips_interpretation_max_alpha, iofuncs/header.c:195-206 */
	switch (type) {
	case VB200_INTERPRETATION_GREY16:
	case VB200_INTERPRETATION_RGB16:
		return 65535.0;
	case VB200_INTERPRETATION_scRGB:
		return 1.0;
	default:
		return 255.0;
	}
}

int
dev_image_new(const char *domain, DevImage *d, int w, int h, int bands, int fmt, int type, cudaStream_t s)
{
	d->w = w;
	d->h = h;
	d->bands = bands;
	d->fmt = fmt;
	d->type = type;
	const size_t line = (size_t) w * bands * format_sizeof(fmt);
	if (d->preset && d->data) {
		/* the caller's buffer (preset_output): sized by contract for this op's result */
		d->preset = false;
		d->owned = false;
		if (d->bpl < line)
			d->bpl = line;
		return 0;
	}
	d->bpl = line;
	d->owned = true;
	return dev_alloc(domain, &d->data, d->bpl * h, s);
}

void
preset_output(DevImage *dout, const VB200Image *in, const VB200Image *out)
{
	if (in->where != VB200_DEVICE || !out->data || !in->data)
		return;
	const size_t in_bytes = (in->bpl ? in->bpl : (size_t) in->Xsize * in->Bands * format_sizeof(in->BandFmt)) * in->Ysize;
	const char *a = (const char *) in->data, *o = (const char *) out->data;
	/* no aliasing: the result must lie wholly before or after the input.  Its extent is not known
	 * yet; no device op produces wider than 4-byte elements, and none changes the pixel count.
	 */
	const size_t out_line = std::max((size_t) out->bpl, (size_t) in->Xsize * in->Bands * 4);
	if (o >= a + in_bytes || o + out_line * in->Ysize <= a) {
		dout->data = out->data;
		dout->bpl = out->bpl;
		dout->preset = true;
	}
}

void
dev_image_release(DevImage *d, cudaStream_t s)
{
	if (d->owned && d->data)
		dev_free(d->data, s);
	d->data = nullptr;
	d->owned = false;
}

int
to_device(const char *domain, const VB200Image *in, DevImage *d, cudaStream_t s)
{
	if (!in || !in->data) {
		error(domain, "no input image");
		return -1;
	}
	if (in->Xsize <= 0 || in->Ysize <= 0 || in->Bands <= 0) {
		error(domain, "bad image dimensions %d x %d x %d", in->Xsize, in->Ysize, in->Bands);
		return -1;
	}
	const size_t line = (size_t) in->Xsize * in->Bands * format_sizeof(in->BandFmt);
	if (line == 0) {
		error(domain, "unknown band format %d", in->BandFmt);
		return -1;
	}
	const size_t bpl = in->bpl ? in->bpl : line;

	d->w = in->Xsize;
	d->h = in->Ysize;
	d->bands = in->Bands;
	d->fmt = in->BandFmt;
	d->type = in->Type;
	if (in->where == VB200_DEVICE) {
		d->data = in->data;
		d->bpl = bpl;
		d->owned = false;
		return 0;
	}
	d->bpl = line;
	d->owned = true;
	if (dev_alloc(domain, &d->data, line * in->Ysize, s))
		return -1;
	VB200_CUDA(domain,
		cudaMemcpy2DAsync(d->data, line, in->data, bpl, line, in->Ysize, cudaMemcpyHostToDevice, s));
	return 0;
}

int
deliver(const char *domain, DevImage *d, const VB200Image *like, VB200Image *out, cudaStream_t s)
{
	const size_t line = (size_t) d->w * d->bands * format_sizeof(d->fmt);
	const int where = like->where;
	void *dst = out->data;
	size_t dst_bpl = (out->data && out->bpl) ? out->bpl : line;

	out->Xsize = d->w;
	out->Ysize = d->h;
	out->Bands = d->bands;
	out->BandFmt = d->fmt;
	out->Type = d->type;
	out->where = where;

	if (where == VB200_DEVICE) {
		if (!dst && d->owned) {
			/* hand our buffer over */
			out->data = d->data;
			out->bpl = d->bpl;
			d->owned = false;
			return 0;
		}
		if (dst && dst == d->data) {
			/* the op wrote into the caller's buffer (preset_output) */
			out->bpl = d->bpl;
			return 0;
		}
		if (!dst) {
			if (dev_alloc(domain, &dst, line * d->h, s))
				return -1;
			out->data = dst;
		}
		out->bpl = dst_bpl;
		VB200_CUDA(domain,
			cudaMemcpy2DAsync(dst, dst_bpl, d->data, d->bpl, line, d->h, cudaMemcpyDeviceToDevice, s));
		dev_image_release(d, s);
		return 0;
	}

	if (!dst) {
		dst = malloc(line * d->h > 0 ? line * d->h : 1);
		if (!dst) {
			error(domain, "out of memory");
			return -1;
		}
		out->data = dst;
	}
	out->bpl = dst_bpl;
	VB200_CUDA(domain, cudaMemcpy2DAsync(dst, dst_bpl, d->data, d->bpl, line, d->h, cudaMemcpyDeviceToHost, s));
	VB200_CUDA(domain, cudaStreamSynchronize(s));
	dev_image_release(d, s);
	return 0;
}

} // namespace vb200

using namespace vb200;

extern "C" int
vb200_init(int device)
{
	std::lock_guard<std::mutex> lock(g_init_lock);
	int n = 0;
	cudaError_t e = cudaGetDeviceCount(&n);
	if (e != cudaSuccess || n <= 0) {
		error("vb200_init", "no CUDA device (%s)", cudaGetErrorString(e));
		return -1;
	}
	if (device < 0 || device >= n) {
		error("vb200_init", "device %d out of range (%d visible)", device, n);
		return -1;
	}
	VB200_CUDA("vb200_init", cudaSetDevice(device));
	cudaDeviceProp prop;
	VB200_CUDA("vb200_init", cudaGetDeviceProperties(&prop, device));
	if (prop.major < 10) {
		error("vb200_init", "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
			prop.minor);
		return -1;
	}
	/* keep freed scratch in the pool instead of returning it to the driver */
	cudaMemPool_t pool;
	if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
		uint64_t threshold = UINT64_MAX;
		cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
	}
	g_device.store(device);
	return 0;
}

extern "C" void
vb200_shutdown(void)
{
	std::lock_guard<std::mutex> lock(g_init_lock);
	if (g_device.load() >= 0)
		cudaDeviceSynchronize();
	g_device.store(-1);
}

extern "C" const char *
vb200_error_buffer(void)
{
	return g_error.c_str();
}

extern "C" void
vb200_error_clear(void)
{
	g_error.clear();
}

extern "C" int
vb200_isenabled(void)
{
	const char *off = getenv("VB200_DISABLE");
	if (off && *off && strcmp(off, "0") != 0)
		return 0;
	int n = 0;
	return cudaGetDeviceCount(&n) == cudaSuccess && n > 0;
}

extern "C" void
vb200_set_stream(void *cuda_stream)
{
	g_stream = (cudaStream_t) cuda_stream;
}

extern "C" void *
vb200_get_stream(void)
{
	return (void *) g_stream;
}

extern "C" void
vb200_set_tile_geometry(int tile_width, int tile_height, int fatstrip_height, int thinstrip_height)
{
	if (tile_width > 0)
		g_tiles.tile_width = tile_width;
	if (tile_height > 0)
		g_tiles.tile_height = tile_height;
	if (fatstrip_height > 0)
		g_tiles.fatstrip_height = fatstrip_height;
	if (thinstrip_height > 0)
		g_tiles.thinstrip_height = thinstrip_height;
}

extern "C" uint64_t
vb200_launch_count(void)
{
	return g_launches.load();
}

extern "C" void
vb200_image_free(VB200Image *image)
{
	if (!image || !image->data)
		return;
	if (image->where == VB200_DEVICE)
		cudaFree(image->data);
	else
		free(image->data);
	image->data = nullptr;
}

extern "C" size_t
vb200_format_sizeof(int band_format)
{
	return format_sizeof(band_format);
}

extern "C" void *
vb200_host_alloc(size_t bytes)
{
	void *p = nullptr;
	if (cudaHostAlloc(&p, bytes ? bytes : 1, cudaHostAllocDefault) != cudaSuccess)
		return nullptr;
	return p;
}

extern "C" void
vb200_host_free(void *p)
{
	if (p)
		cudaFreeHost(p);
}
