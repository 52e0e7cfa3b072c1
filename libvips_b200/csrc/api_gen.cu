/* api_gen.cu -- generate()-shaped seams for the convolution and colour rows of SURVEY 8(b):
 * "fill out->valid from an input region", the shape of vips_convf_gen / vips_convi_gen
 * (convf.c:185-282, convi.c:752-852) and vips_colour_gen (colour.c:119-156).  They are
 * host-pointer compatibility entry points built on the whole-image operations (one
 * region = one small image); throughput comes from vb200_conv / vb200_colourspace on
 * device images, not from here.
 */
#include <cstring>

#include "vb200_internal.h"

using namespace vb200;

namespace {

int
sub_image(const char *domain, const VB200Region *in, const VB200Rect &need, VB200Image *sub)
{
	if (!in || !in->data) {
		error(domain, "null region");
		return -1;
	}
	if (need.left < in->valid.left || need.top < in->valid.top ||
		need.left + need.width > in->valid.left + in->valid.width ||
		need.top + need.height > in->valid.top + in->valid.height) {
		error(domain, "input region does not cover %d x %d at (%d, %d)", need.width, need.height, need.left, need.top);
		return -1;
	}
	const size_t ps = format_sizeof(in->im.BandFmt) * in->im.Bands;
	*sub = in->im;
	sub->Xsize = need.width;
	sub->Ysize = need.height;
	sub->where = VB200_HOST;
	sub->data = (char *) in->data + (size_t) (need.top - in->valid.top) * in->bpl + (size_t) (need.left - in->valid.left) * ps;
	sub->bpl = in->bpl;
	return 0;
}

int
copy_out(const char *domain, const VB200Region *out, const VB200Image *res, int x0, int y0)
{
	const size_t ps = format_sizeof(res->BandFmt) * res->Bands;
	if (x0 + out->valid.width > res->Xsize || y0 + out->valid.height > res->Ysize) {
		error(domain, "result smaller than the output rect");
		return -1;
	}
	for (int y = 0; y < out->valid.height; y++)
		memcpy((char *) out->data + (size_t) y * out->bpl, (const char *) res->data + (size_t) (y + y0) * res->bpl + (size_t) x0 * ps,
			ps * out->valid.width);
	return 0;
}

} // namespace

/* Regions are on the EMBEDDED image (vips_conv embeds by the mask size first, conv.c:95-104): output
 * pixel (x, y) reads embedded (x .. x + mw - 1, y .. y + mh - 1), so `in` must cover
 * {left, top, width + mw - 1, height + mh - 1} of out->valid (convf.c:206-210, convi.c:775-779).
 */
extern "C" int
vb200_conv_gen(const VB200Region *out, const VB200Region *in, const VB200Mask *mask, int precision)
{
	const char *domain = "conv_gen";
	if (!out || !out->data || !mask || !mask->coeff) {
		error(domain, "null argument");
		return -1;
	}
	const VB200Rect need = {out->valid.left, out->valid.top, out->valid.width + mask->width - 1,
		out->valid.height + mask->height - 1};
	VB200Image sub, res;
	if (sub_image(domain, in, need, &sub))
		return -1;
	memset(&res, 0, sizeof(res));
	if (vb200_conv(&sub, &res, mask, precision))
		return -1;
	/* conv(sub) clamps at sub's edges; the rect that never touched an edge is the answer */
	const int rc = copy_out(domain, out, &res, mask->width / 2, mask->height / 2);
	vb200_image_free(&res);
	return rc;
}

/* vips_colour_gen: in covers out->valid; in->im.Type is the source interpretation. */
extern "C" int
vb200_colour_gen(const VB200Region *out, const VB200Region *in, int space)
{
	const char *domain = "colour_gen";
	if (!out || !out->data) {
		error(domain, "null argument");
		return -1;
	}
	VB200Image sub, res;
	if (sub_image(domain, in, out->valid, &sub))
		return -1;
	memset(&res, 0, sizeof(res));
	if (vb200_colourspace(&sub, &res, space))
		return -1;
	const int rc = copy_out(domain, out, &res, 0, 0);
	vb200_image_free(&res);
	return rc;
}
