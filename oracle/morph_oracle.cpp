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
