/* oracle.h -- C interface of the CPU oracle.
 *
 * TEST INFRASTRUCTURE ONLY.  This directory holds a plain-C++ CPU restatement
 * of the reference's (libvips 8.19) per-tile pixel arithmetic.  It is the
 * checker for the CUDA path and the "port" CPU baseline; nothing under
 * libvips_b200/ may include, link or call it.  Only tests/, bench.py's
 * cpu_baseline / --impl reference leg and __graft_entry__.smoke() use it.
 *
 * Parity pinning: every function here is cross-checked against the reference's
 * own sources compiled under a GLib-free shim (oracle/ref_shim -> oracle/_ref,
 * see oracle/Makefile) by tests/test_oracle_vs_ref.py, and against the
 * known-answer values the reference's test-suite holds for this path
 * (tests/test_oracle_known_answers.py).
 *
 * All images are host memory, row-major, band-interleaved, packed rows
 * (bpl = width * bands * sizeof(elem)), exactly as VipsRegion lays them out
 * (reference: libvips/iofuncs/region.c:583-587).
 */
#ifndef VIPS_B200_ORACLE_H
#define VIPS_B200_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Same numeric values as VipsBandFormat (include/vips/image.h:121-132). */
enum {
	ORC_FORMAT_UCHAR = 0,
	ORC_FORMAT_CHAR = 1,
	ORC_FORMAT_USHORT = 2,
	ORC_FORMAT_SHORT = 3,
	ORC_FORMAT_UINT = 4,
	ORC_FORMAT_INT = 5,
	ORC_FORMAT_FLOAT = 6,
	ORC_FORMAT_DOUBLE = 8
};

/* Same numeric values as VipsKernel (include/vips/resample.h:41-51). */
enum {
	ORC_KERNEL_NEAREST = 0,
	ORC_KERNEL_LINEAR = 1,
	ORC_KERNEL_CUBIC = 2,
	ORC_KERNEL_MITCHELL = 3,
	ORC_KERNEL_LANCZOS2 = 4,
	ORC_KERNEL_LANCZOS3 = 5,
	ORC_KERNEL_MKS2013 = 6,
	ORC_KERNEL_MKS2021 = 7
};

/* Geometry of one reduce axis: what vips_reducev_build / vips_reduceh_build
 * derive before any pixel is touched (reducev.cpp:887-941, reduceh.cpp:426-481).
 */
typedef struct {
	int in_size;	 /* axis length of the input */
	int out_size;	 /* ROUND_UINT(in / shrink) */
	int int_shrink;	 /* box pre-shrink (1 = none) */
	int shrunk_size; /* axis length after the box pre-shrink (ceil) */
	double residual; /* shrink left for the kernel pass */
	int n_point;	 /* taps of the kernel pass, 0 if residual == 1 */
	double offset;	 /* voffset / hoffset */
} OrcReduceGeom;

int orc_reduce_get_points(int kernel, double shrink);
void orc_reduce_make_mask(double *c, int kernel, int n_point, double shrink, double x);
/* 65 x n_point tables, matrixf (double) and matrixs (short, truncated x4096). */
void orc_reduce_tables(int kernel, int n_point, double residual, double *matrixf, short *matrixs);
int orc_reduce_geometry(int in_size, double shrink, int kernel, double gap, OrcReduceGeom *g);

int orc_shrink_size(int in_size, int shrink, int ceil_mode);
int orc_shrinkv(const void *in, int w, int h, int bands, int fmt, int vshrink, int ceil_mode, void *out);
int orc_shrinkh(const void *in, int w, int h, int bands, int fmt, int hshrink, int ceil_mode, void *out);

/* Kernel pass only (residual shrink, n_point taps) on an un-embedded image:
 * the EXTEND_COPY embed is applied by clamping.  rect_h / rect_w give the
 * height / width of the output rects the reference's generate() would be
 * called with (rect origins at multiples of it); 0 = one rect for the axis.
 */
int orc_reducev_pass(const void *in, int w, int h, int bands, int fmt, int out_h, double residual,
	double voffset, int n_point, int kernel, int rect_h, void *out);
int orc_reduceh_pass(const void *in, int w, int h, int bands, int fmt, int out_w, double residual,
	double hoffset, int n_point, int kernel, int rect_w, void *out);

/* vips_reducev / vips_reduceh: box pre-shrink (gap) + kernel pass. */
int orc_reducev(const void *in, int w, int h, int bands, int fmt, double vshrink, int kernel, double gap,
	int rect_h, void *out);
int orc_reduceh(const void *in, int w, int h, int bands, int fmt, double hshrink, int kernel, double gap,
	int rect_w, void *out);

/* premultiply.c / unpremultiply.c.  uchar_mode selects the 8.8 LUT fast path
 * (uchar in -> uchar out); otherwise out is float (double for double in).
 */
int orc_premultiply(const void *in, int w, int h, int bands, int fmt, double max_alpha, int uchar_mode, void *out);
int orc_unpremultiply(const void *in, int w, int h, int bands, int fmt, double max_alpha, int uchar_mode, void *out);

/* vips_resize downsizing part (resize.c:135-231): reducev then reduceh.
 * tile_w/tile_h = the sink tile geometry (0,0 = derive from the demand hints
 * the way vips_get_tile_size would).
 */
int orc_resize_size(int w, int h, double hscale, double vscale, int kernel, double gap, int *ow, int *oh);
size_t orc_sizeof_format(int fmt);
int orc_resize(const void *in, int w, int h, int bands, int fmt, double hscale, double vscale, int kernel,
	double gap, int tile_w, int tile_h, void *out);

/* vips_thumbnail_image for uchar images already in sRGB/B_W (thumbnail.c:678-902):
 * [premultiply uchar] -> resize -> [unpremultiply uchar].  size_mode as VipsSize.
 * has_alpha: treat the last band as alpha (bands 2 or 4).
 */
int orc_thumbnail_size(int w, int h, int target_w, int target_h, int size_mode, double *hshrink,
	double *vshrink, int *ow, int *oh);
int orc_thumbnail_image(const void *in, int w, int h, int bands, int target_w, int target_h, int size_mode,
	int has_alpha, int tile_w, int tile_h, void *out);
int orc_thumbnail_image_batch(const void *in, int n_frames, int w, int h, int bands, int target_w, int target_h,
	int size_mode, int has_alpha, void *out, int ow, int oh, int n_threads);
/* linear=TRUE variant (thumbnail.c:766-806, 971-987): sRGB->scRGB float, float
 * premultiply, float resize, float unpremultiply, scRGB->sRGB.
 */
int orc_thumbnail_image_linear(const void *in, int w, int h, int bands, int target_w, int target_h,
	int size_mode, int has_alpha, int tile_w, int tile_h, void *out);

#ifdef __cplusplus
}
#endif

#endif
