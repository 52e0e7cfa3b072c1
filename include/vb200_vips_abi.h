/* vb200_vips_abi.h -- layout mirrors of the two GObject-derived structs that libvips'
 * vector kernels receive by pointer, so that libvb200.so can export
 * vips_convi_uchar_hwy() with the reference's exact signature
 * (reference: convolution/pconvolution.h:74-76) without GLib headers.
 *
 * LP64 (x86-64 / aarch64 Linux) layouts of
 *   GObject      { GTypeInstance{GTypeClass *g_class}; guint ref_count; GData *qdata; }    24 bytes
 *   VipsObject   include/vips/object.h:441-474                                             80 bytes
 *   VipsImage    include/vips/image.h:189-260    (head: Xsize .. Type)
 *   VipsRegion   include/vips/region.h:96-130    (head: im, valid, type, data, bpl)
 * Only the fields VIPS_REGION_ADDR / VIPS_REGION_LSKIP / VIPS_IMAGE_SIZEOF_PEL read are named.
 * A host build checks them against the real headers with the static asserts at the end of
 * INTEGRATION.md's binding; here they are checked against the offsets those headers give.
 */
#ifndef VB200_VIPS_ABI_H
#define VB200_VIPS_ABI_H

#include <stddef.h>
#include <stdint.h>

#include "vb200.h"

#ifdef __cplusplus
extern "C" {
#endif

typedef struct {
	void *g_class;
	unsigned int ref_count;
	void *qdata;
} VB200GObjectHead;

typedef struct {
	VB200GObjectHead parent_instance;
	int constructed;
	int static_object;
	void *argument_table;
	char *nickname;
	char *description;
	int preclose;
	int close;
	int postclose;
	size_t local_memory;
} VB200VipsObjectHead;

typedef struct {
	VB200VipsObjectHead parent_instance;
	int Xsize;
	int Ysize;
	int Bands;
	int BandFmt;
	int Coding;
	int Type;
	/* ... the rest of VipsImage is never read here */
} VB200VipsImageHead;

typedef struct {
	VB200VipsObjectHead parent_object;
	VB200VipsImageHead *im;
	VB200Rect valid;
	int type;
	uint8_t *data;
	int bpl;
	/* ... seq, thread, window, buffer, invalid: never read here */
} VB200VipsRegionHead;

#if defined(__LP64__) || defined(_LP64)
#ifdef __cplusplus
#define VB200_ABI_ASSERT(c, m) static_assert(c, m)
#else
#define VB200_ABI_ASSERT(c, m) _Static_assert(c, m)
#endif
VB200_ABI_ASSERT(sizeof(VB200GObjectHead) == 24, "GObject is 24 bytes on LP64");
VB200_ABI_ASSERT(sizeof(VB200VipsObjectHead) == 80, "VipsObject is 80 bytes on LP64");
VB200_ABI_ASSERT(offsetof(VB200VipsImageHead, Bands) == 88, "VipsImage.Bands");
VB200_ABI_ASSERT(offsetof(VB200VipsRegionHead, im) == 80, "VipsRegion.im");
VB200_ABI_ASSERT(offsetof(VB200VipsRegionHead, valid) == 88, "VipsRegion.valid");
VB200_ABI_ASSERT(offsetof(VB200VipsRegionHead, data) == 112, "VipsRegion.data");
VB200_ABI_ASSERT(offsetof(VB200VipsRegionHead, bpl) == 120, "VipsRegion.bpl");
#endif

/* reference: convolution/pconvolution.h:74-76, convi_hwy.cpp:93-276 -- the Highway kernel of
 * vips_convi_uchar_vector_gen (convi.c:306-362).  Same name, same signature (VipsRegion * /
 * VipsRect * by layout): for every y in r and x in 0 .. ne - 1,
 *   q[x] = clip(0, (((1 << (exp - 1)) + sum_i p[x + offsets[i]] * mant[i]) >> exp) + offset, 255)
 * with p = VIPS_REGION_ADDR(ir, r->left, y), q = VIPS_REGION_ADDR(out_region, r->left, y).
 * Host pointers; the region is staged through the device.
 */
void vips_convi_uchar_hwy(VB200VipsRegionHead *out_region, VB200VipsRegionHead *ir, VB200Rect *r, int32_t ne, int32_t nnz,
	int32_t offset, const int32_t *offsets, const int16_t *mant, int32_t exp);

/* The same kernel on plain pointers, for hosts that would rather not pass GObject structs:
 * p0 / q0 = the addresses of (r->left, r->top) in the input / output region.  in_rows x in_line_bytes is the
 * extent of the input that may be read starting at p0 (what vips_region_prepare() made valid).
 */
int vb200_convi_uchar_vector(uint8_t *q0, int out_bpl, const uint8_t *p0, int in_bpl, int in_line_bytes, int in_rows, int ne,
	int rows, int nnz, int offset, const int32_t *offsets, const int16_t *mant, int exp);

#ifdef __cplusplus
}
#endif

#endif /* VB200_VIPS_ABI_H */
