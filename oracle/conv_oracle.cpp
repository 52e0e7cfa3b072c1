/* conv_oracle.cpp -- CPU restatement of the reference's convolution hot path.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle.h).  -O2 -ffp-contract=off.
 *
 * Follows (reference = libvips 8.19, libvips/):
 *   convolution/convf.c:163-367      CONV_FLOAT, vips_convf_gen, vips_convf_build
 *   convolution/convi.c:698-852      CONV_INT / CONV_FLOAT, vips_convi_gen (the C path)
 *   convolution/convi.c:859-923      vips__image_intize
 *   convolution/convi.c:931-1119     vips_convi_intize (8-bit mantissa + shared exponent, HAVE_HWY branch)
 *   convolution/convi_hwy.cpp:265-273  the scalar statement of the vector arithmetic
 *   convolution/convi.c:1123-1230    vips_convi_build (path selection, embed, geometry)
 *   convolution/conv.c:60-121, convsep.c:61-114, gaussblur.c:70-114
 *   create/gaussmat.c:93-170
 *   convolution/sharpen.c:116-303
 */
#include <algorithm>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <vector>

#include "oracle.h"

extern "C" int orc_colourspace(const void *in, int w, int h, int bands, int fmt, int from, int to, void *out);
extern "C" int orc_colourspace_format(int space);

static inline int clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

/* ---------------------------------------------------------------- gaussmat */

extern "C" int
orc_gaussmat_size(double sigma, double min_ampl)
{
	/* gaussmat.c:99-127 */
	const double sig2 = 2. * sigma * sigma;
	const double m = 8 * sigma;
	const int max_x = (int) std::max(0.0, std::min(5000.0, m));
	int x;
	for (x = 0; x < max_x; x++) {
		double v = exp(-((double) (x * x)) / sig2);
		if (v < min_ampl)
			break;
	}
	return 2 * std::max(x - 1, 0) + 1;
}

/* coeff: width x (separable ? 1 : width) doubles.  Returns the scale (sum). */
extern "C" double
orc_gaussmat(double sigma, double min_ampl, int separable, int integer_precision, double *coeff)
{
	const double sig2 = 2. * sigma * sigma;
	const int width = orc_gaussmat_size(sigma, min_ampl);
	const int height = separable ? 1 : width;
	double sum = 0.0;
	for (int y = 0; y < height; y++)
		for (int x = 0; x < width; x++) {
			int xo = x - width / 2;
			int yo = y - height / 2;
			double distance = xo * xo + yo * yo;
			double v = exp(-distance / sig2);
			if (integer_precision)
				v = rint(20 * v);
			coeff[y * width + x] = v;
			sum += v;
		}
	if (sum == 0)
		sum = 1;
	return sum;
}

/* -------------------------------------------------------------------- conv */

template <typename T>
static inline double
elem(const void *in, size_t idx)
{
	return ((const T *) in)[idx];
}

struct Sparse {
	std::vector<int> mx, my;
};

/* vips_convf: out float (double for double in), convf.c */
template <typename T, typename O>
static void
convf_t(const T *in, int w, int h, int bands, const double *mask, int mw, int mh, double scale, double offset, O *out)
{
	std::vector<double> coeff;
	std::vector<int> pos;
	for (int i = 0; i < mw * mh; i++) {
		const double c = mask[i] / scale; /* convf.c:307-309 */
		if (c) {
			coeff.push_back(c);
			pos.push_back(i);
		}
	}
	if (coeff.empty()) {
		coeff.push_back(0);
		pos.push_back(0);
	}
	const int nnz = (int) coeff.size();
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
			for (int b = 0; b < bands; b++) {
				double sum = offset;
				for (int i = 0; i < nnz; i++) {
					const int mxp = pos[i] % mw, myp = pos[i] / mw;
					const int sx = clampi(x + mxp - mw / 2, 0, w - 1);
					const int sy = clampi(y + myp - mh / 2, 0, h - 1);
					sum += coeff[i] * in[((size_t) sy * w + sx) * bands + b];
				}
				out[((size_t) y * w + x) * bands + b] = sum;
			}
}

/* vips__image_intize, convi.c:859-923 */
static void
image_intize(const double *mask, int n, double scale, double offset, std::vector<int> &icoeff, int *iscale,
	int *ioffset)
{
	double double_result = 0;
	for (int i = 0; i < n; i++)
		double_result += mask[i];
	double_result /= scale;
	std::vector<double> r(n);
	for (int i = 0; i < n; i++)
		r[i] = rint(mask[i]);
	double out_scale = rint(scale);
	if (out_scale == 0)
		out_scale = 1;
	const double out_offset = rint(offset);
	int int_result = 0;
	for (int i = 0; i < n; i++)
		int_result += r[i];
	int_result /= out_scale;
	out_scale = rint(out_scale + (int_result - double_result));
	if (out_scale == 0)
		out_scale = 1;
	icoeff.resize(n);
	for (int i = 0; i < n; i++)
		icoeff[i] = r[i];
	/* vips_convi_gen reads scale / offset from convolution->M, the ORIGINAL
	 * double matrix (convi.c:760-763) -- the adjusted out_scale computed above
	 * lives on the build()-local intized copy and never reaches the pixels.
	 */
	(void) out_scale;
	(void) out_offset;
	*iscale = rint(scale);
	*ioffset = rint(offset);
}

template <typename T>
static void
convi_int_t(const T *in, int w, int h, int bands, const std::vector<int> &icoeff, int mw, int mh, int scale, int offset,
	int64_t lo, int64_t hi, bool clip, T *out)
{
	std::vector<int> t, pos;
	for (int i = 0; i < mw * mh; i++)
		if (icoeff[i]) {
			t.push_back(icoeff[i]);
			pos.push_back(i);
		}
	if (t.empty()) {
		t.push_back(0);
		pos.push_back(0);
	}
	const int nnz = (int) t.size();
	const int rounding = scale / 2;
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
			for (int b = 0; b < bands; b++) {
				int64_t sum = 0;
				for (int i = 0; i < nnz; i++) {
					const int mxp = pos[i] % mw, myp = pos[i] / mw;
					const int sx = clampi(x + mxp - mw / 2, 0, w - 1);
					const int sy = clampi(y + myp - mh / 2, 0, h - 1);
					sum += (int64_t) t[i] * in[((size_t) sy * w + sx) * bands + b];
				}
				sum = ((sum + rounding) / scale) + offset;
				if (clip)
					sum = std::max(lo, std::min(hi, sum));
				out[((size_t) y * w + x) * bands + b] = (T) sum;
			}
}

/* float input through convi (CONV_FLOAT, convi.c:721-739): int coefficients */
template <typename T>
static void
convi_float_t(const T *in, int w, int h, int bands, const std::vector<int> &icoeff, int mw, int mh, int scale,
	int offset, T *out)
{
	std::vector<int> t, pos;
	for (int i = 0; i < mw * mh; i++)
		if (icoeff[i]) {
			t.push_back(icoeff[i]);
			pos.push_back(i);
		}
	if (t.empty()) {
		t.push_back(0);
		pos.push_back(0);
	}
	const int nnz = (int) t.size();
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
			for (int b = 0; b < bands; b++) {
				double sum = 0;
				for (int i = 0; i < nnz; i++) {
					const int mxp = pos[i] % mw, myp = pos[i] / mw;
					const int sx = clampi(x + mxp - mw / 2, 0, w - 1);
					const int sy = clampi(y + myp - mh / 2, 0, h - 1);
					sum += (double) t[i] * in[((size_t) sy * w + sx) * bands + b];
				}
				sum = (sum / scale) + offset;
				out[((size_t) y * w + x) * bands + b] = sum;
			}
}

/* vips_convi_intize (HAVE_HWY branch), convi.c:931-1119.  Returns 0 and fills
 * mant/pos/exp when the 8-bit-mantissa form is accurate enough, else -1.
 */
extern "C" int
orc_convi_intize8(const double *mask, int n_point, double scale, short *mant, int *pos, int *nnz_out, int *exp_out)
{
	std::vector<double> scaled(n_point);
	for (int i = 0; i < n_point; i++)
		scaled[i] = mask[i] / scale;
	double mx = scaled[0];
	for (int i = 1; i < n_point; i++)
		if (scaled[i] > mx)
			mx = scaled[i];
	const int shift = ceil(log2(mx) + 1);
	if (shift > 6 || shift < -24)
		return -1;
	if (ceil(log2(n_point)) > 10)
		return -1;
	const int exp = 7 - shift;
	int nnz = 0;
	std::vector<short> all(n_point);
	for (int i = 0; i < n_point; i++) {
		all[i] = rint(128 * scaled[i] * pow(2, -shift));
		if (all[i] < -128 || all[i] > 127)
			return -1;
		if (all[i]) {
			mant[nnz] = all[i];
			pos[nnz] = i;
			nnz += 1;
		}
	}
	if (nnz == 0) {
		mant[0] = 0;
		pos[0] = 0;
		nnz = 1;
	}
	double true_sum = 0.0;
	int int_sum = 0;
	for (int i = 0; i < nnz; i++) {
		true_sum += 128 * scaled[pos[i]];
		int_sum += 128 * mant[i];
	}
	const int true_value = std::max(0.0, std::min(255.0, true_sum));
	int int_value = (int_sum + (1 << (exp - 1))) >> exp;
	int_value = std::max(0, std::min(255, int_value));
	if (abs(true_value - int_value) > 2)
		return -1;
	*nnz_out = nnz;
	*exp_out = exp;
	return 0;
}

/* convi_hwy.cpp:265-273 */
static void
convi_vector_u8(const uint8_t *in, int w, int h, int bands, const short *mant, const int *pos, int nnz, int exp, int mw,
	int mh, int offset, uint8_t *out)
{
	for (int y = 0; y < h; y++)
		for (int x = 0; x < w; x++)
			for (int b = 0; b < bands; b++) {
				int32_t sum = 1 << (exp - 1);
				for (int i = 0; i < nnz; i++) {
					const int mxp = pos[i] % mw, myp = pos[i] / mw;
					const int sx = clampi(x + mxp - mw / 2, 0, w - 1);
					const int sy = clampi(y + myp - mh / 2, 0, h - 1);
					sum += in[((size_t) sy * w + sx) * bands + b] * mant[i];
				}
				out[((size_t) y * w + x) * bands + b] = std::max(0, std::min(255, (sum >> exp) + offset));
			}
}

extern "C" int
orc_conv_out_format(int fmt, int precision)
{
	if (precision == 1 /*FLOAT*/)
		return fmt == ORC_FORMAT_DOUBLE ? ORC_FORMAT_DOUBLE : ORC_FORMAT_FLOAT;
	return fmt;
}

/* vips_conv(in, mask, precision).  vector != 0 selects the Highway arithmetic
 * for uchar INTEGER convolutions when vips_convi_intize succeeds.
 */
extern "C" int
orc_conv(const void *in, int w, int h, int bands, int fmt, const double *mask, int mw, int mh, double scale,
	double offset, int precision, int vector, void *out)
{
	if (precision == 1) {
#define CF(T, O) convf_t<T, O>((const T *) in, w, h, bands, mask, mw, mh, scale, offset, (O *) out)
		switch (fmt) {
		case ORC_FORMAT_UCHAR: CF(uint8_t, float); break;
		case ORC_FORMAT_CHAR: CF(int8_t, float); break;
		case ORC_FORMAT_USHORT: CF(uint16_t, float); break;
		case ORC_FORMAT_SHORT: CF(int16_t, float); break;
		case ORC_FORMAT_UINT: CF(uint32_t, float); break;
		case ORC_FORMAT_INT: CF(int32_t, float); break;
		case ORC_FORMAT_FLOAT: CF(float, float); break;
		case ORC_FORMAT_DOUBLE: CF(double, double); break;
		default: return -1;
		}
		return 0;
	}
	if (precision != 0)
		return -1;

	if (vector && fmt == ORC_FORMAT_UCHAR) {
		std::vector<short> mant(mw * mh);
		std::vector<int> pos(mw * mh);
		int nnz, exp;
		if (!orc_convi_intize8(mask, mw * mh, scale, mant.data(), pos.data(), &nnz, &exp)) {
			convi_vector_u8((const uint8_t *) in, w, h, bands, mant.data(), pos.data(), nnz, exp, mw, mh, (int) rint(offset),
				(uint8_t *) out);
			return 0;
		}
	}

	std::vector<int> icoeff;
	int iscale, ioffset;
	image_intize(mask, mw * mh, scale, offset, icoeff, &iscale, &ioffset);
	if (iscale == 0)
		return -1; /* the reference would divide by zero */
#define CI(T, LO, HI, CLIP) convi_int_t<T>((const T *) in, w, h, bands, icoeff, mw, mh, iscale, ioffset, LO, HI, CLIP, (T *) out)
	switch (fmt) {
	case ORC_FORMAT_UCHAR: CI(uint8_t, 0, UCHAR_MAX, true); break;
	case ORC_FORMAT_CHAR: CI(int8_t, SCHAR_MIN, SCHAR_MAX, true); break;
	case ORC_FORMAT_USHORT: CI(uint16_t, 0, USHRT_MAX, true); break;
	case ORC_FORMAT_SHORT: CI(int16_t, SHRT_MIN, SHRT_MAX, true); break;
	case ORC_FORMAT_UINT: CI(uint32_t, 0, 0, false); break;
	case ORC_FORMAT_INT: CI(int32_t, 0, 0, false); break;
	case ORC_FORMAT_FLOAT: convi_float_t<float>((const float *) in, w, h, bands, icoeff, mw, mh, iscale, ioffset, (float *) out); break;
	case ORC_FORMAT_DOUBLE: convi_float_t<double>((const double *) in, w, h, bands, icoeff, mw, mh, iscale, ioffset, (double *) out); break;
	default: return -1;
	}
	return 0;
}

/* vips_convsep, convsep.c:61-114: conv(M) as given, then conv(rot90(M)) with offset 0, same scale.
 * vips_rot90 (conversion/rot.c:100-156): out(x, y) = in(y, Ysize - 1 - x), so an n x 1 row becomes
 * the 1 x n column in the same order and a 1 x n column becomes the n x 1 row reversed.
 */
extern "C" int
orc_convsep2(const void *in, int w, int h, int bands, int fmt, const double *mask, int mw, int mh, double scale,
	double offset, int precision, int vector, void *out)
{
	if (mw != 1 && mh != 1)
		return -1; /* vips_check_separable */
	const int n = mw * mh;
	std::vector<double> rot(mask, mask + n);
	if (mw == 1)
		std::reverse(rot.begin(), rot.end());
	const int mid_fmt = orc_conv_out_format(fmt, precision);
	std::vector<uint8_t> mid((size_t) w * h * bands * orc_sizeof_format(mid_fmt));
	if (orc_conv(in, w, h, bands, fmt, mask, mw, mh, scale, offset, precision, vector, mid.data()))
		return -1;
	return orc_conv(mid.data(), w, h, bands, mid_fmt, rot.data(), mh, mw, scale, 0.0, precision, vector, out);
}

/* the n x 1 (horizontal first) case under its old name */
extern "C" int
orc_convsep(const void *in, int w, int h, int bands, int fmt, const double *mask, int n, double scale, double offset,
	int precision, int vector, void *out)
{
	return orc_convsep2(in, w, h, bands, fmt, mask, n, 1, scale, offset, precision, vector, out);
}

/* vips_gaussblur, gaussblur.c:70-114 */
extern "C" int
orc_gaussblur(const void *in, int w, int h, int bands, int fmt, double sigma, double min_ampl, int precision,
	int vector, void *out)
{
	if (sigma < 0.2) {
		memcpy(out, in, (size_t) w * h * bands * orc_sizeof_format(fmt));
		return 0;
	}
	const int n = orc_gaussmat_size(sigma, min_ampl);
	std::vector<double> m(n);
	const double scale = orc_gaussmat(sigma, min_ampl, 1, precision != 1, m.data());
	return orc_convsep(in, w, h, bands, fmt, m.data(), n, scale, 0.0, precision, vector, out);
}

/* the LUT of vips_sharpen_build, sharpen.c:227-257: index = signed L difference + 32768 */
extern "C" void
orc_sharpen_lut(double x1, double y2, double y3, double m1, double m2, int *lut)
{
	for (int i = 0; i < 65536; i++) {
		double v = (i - 32767) / 327.67;
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
}

/* vips_sharpen on an image of interpretation `type` whose format is the
 * colourspace's native one (uchar sRGB, float Lab...).  sharpen.c:171-303.
 */
extern "C" int
orc_sharpen(const void *in, int w, int h, int bands, int fmt, int type, double sigma, double x1, double y2, double y3,
	double m1, double m2, void *out)
{
	const size_t n = (size_t) w * h;
	std::vector<int16_t> labs(n * bands);
	if (orc_colourspace(in, w, h, bands, fmt, type, 21 /*LABS*/, labs.data()))
		return -1;

	const int mn = orc_gaussmat_size(sigma, 0.1);
	std::vector<double> m(mn);
	const double scale = orc_gaussmat(sigma, 0.1, 1, 1, m.data());

	std::vector<int> lut(65536);
	orc_sharpen_lut(x1, y2, y3, m1, m2, lut.data());

	std::vector<int16_t> L(n), blur(n);
	for (size_t i = 0; i < n; i++)
		L[i] = labs[i * bands];
	/* short input: always the exact C path (convi.c:1152-1172 needs uchar for the vector path) */
	if (orc_convsep(L.data(), w, h, 1, ORC_FORMAT_SHORT, m.data(), mn, scale, 0.0, 0, 0, blur.data()))
		return -1;
	for (size_t i = 0; i < n; i++) {
		int v1 = L[i];
		int v2 = blur[i];
		int diff = ((v1 & 0x7fff) - (v2 & 0x7fff));
		int o = v1 + lut[diff + 32768];
		if (o < 0)
			o = 0;
		if (o > 32767)
			o = 32767;
		labs[i * bands] = o;
	}
	return orc_colourspace(labs.data(), w, h, bands, ORC_FORMAT_SHORT, 21, type, out);
}
