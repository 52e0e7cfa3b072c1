/* morph_oracle.cpp -- CPU restatement of vips_morph (binary erode / dilate), SURVEY 8f rank 4.
 * TEST INFRASTRUCTURE ONLY (see oracle.h).
 *
 * Follows morphology/morph.c:
 *   :829-935  vips_morph_build: embed by the mask (VIPS_EXTEND_COPY, origin M / 2), cast to uchar,
 *             mask elements must be 0, 128 (do not care) or 255 after vips__image_intize (rint)
 *   :657-739  vips_dilate_gen: result = 0;   result |= coeff ? p : ~p   over the non-128 elements
 *   :744-826  vips_erode_gen:  result = 255; result &= coeff ? p : ~p
 * applied to every ELEMENT (band-interleaved bytes; offsets are in pixels of the embedded image).
 */
#include <cmath>
#include <cstdint>
#include <vector>

#include "oracle.h"

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* op: 0 erode, 1 dilate (VipsOperationMorphology).  uchar in, uchar out.  -1: bad mask element */
extern "C" int
orc_morph(const uint8_t *in, int w, int h, int bands, const double *mask, int mw, int mh, int op, uint8_t *out)
{
	std::vector<int> dx, dy, co;
	for (int y = 0; y < mh; y++)
		for (int x = 0; x < mw; x++) {
			const double c = rint(mask[y * mw + x]);
			if (c != 0 && c != 128 && c != 255)
				return -1;
			if (c == 128)
				continue;
			dx.push_back(x - mw / 2);
			dy.push_back(y - mh / 2);
			co.push_back((int) c);
		}
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
			for (int b = 0; b < bands; b++) {
				int result = op ? 0 : 255;
				for (size_t i = 0; i < co.size(); i++) {
					const int sx = clampi(x + dx[i], 0, w - 1), sy = clampi(y + dy[i], 0, h - 1);
					const int p = in[((size_t) sy * w + sx) * bands + b];
					const int v = !co[i] ? ~p : p;
					if (op)
						result |= v;
					else
						result &= v;
				}
				out[((size_t) y * w + x) * bands + b] = (uint8_t) result;
			}
	return 0;
}

/* ------------------------------------------------------------------ vips_rank
 * Follows morphology/rank.c:
 *   :458-490  vips_rank_build: window no larger than the image, 0 <= index < width * height
 *   :507-512  the image is embedded at (width / 2, height / 2) with VIPS_EXTEND_COPY, so output (x, y) sees the
 *             window of input pixels (x - width / 2 + i, y - height / 2 + j), coordinates clamped
 *   :165-232  (uchar histogram), :236-323 (select), :327-381 (max / min): all four paths return the index-th
 *             smallest element of the window, per band; the restatement sorts the window.
 * fmt: VipsBandFormat (0 uchar .. 6 float).  -1: bad window / index / format.
 */
#include <algorithm>

template <typename T>
static void
rank_typed(const T *in, int w, int h, int bands, int rw, int rh, int index, T *out)
{
	std::vector<T> win((size_t) rw * rh);
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
			for (int b = 0; b < bands; b++) {
				size_t k = 0;
				for (int j = 0; j < rh; j++)
					for (int i = 0; i < rw; i++) {
						const int sx = clampi(x - rw / 2 + i, 0, w - 1), sy = clampi(y - rh / 2 + j, 0, h - 1);
						win[k++] = in[((size_t) sy * w + sx) * bands + b];
					}
				std::nth_element(win.begin(), win.begin() + index, win.end());
				out[((size_t) y * w + x) * bands + b] = win[index];
			}
}

extern "C" int
orc_rank(const void *in, int w, int h, int bands, int fmt, int rw, int rh, int index, void *out)
{
	if (rw < 1 || rh < 1 || rw > w || rh > h || index < 0 || index > rw * rh - 1)
		return -1;
	switch (fmt) {
	case 0: rank_typed((const uint8_t *) in, w, h, bands, rw, rh, index, (uint8_t *) out); break;
	case 1: rank_typed((const int8_t *) in, w, h, bands, rw, rh, index, (int8_t *) out); break;
	case 2: rank_typed((const uint16_t *) in, w, h, bands, rw, rh, index, (uint16_t *) out); break;
	case 3: rank_typed((const int16_t *) in, w, h, bands, rw, rh, index, (int16_t *) out); break;
	case 4: rank_typed((const uint32_t *) in, w, h, bands, rw, rh, index, (uint32_t *) out); break;
	case 5: rank_typed((const int32_t *) in, w, h, bands, rw, rh, index, (int32_t *) out); break;
	case 6: rank_typed((const float *) in, w, h, bands, rw, rh, index, (float *) out); break;
	default: return -1;
	}
	return 0;
}
