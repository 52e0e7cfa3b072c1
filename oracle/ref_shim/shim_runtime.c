/* shim_runtime.c -- a miniature, single-threaded, cache-free restatement of the
 * libvips demand-driven region engine, just enough to RUN the reference's own
 * build() and generate() functions (compiled from /root/reference in place)
 * on in-memory images.  TEST INFRASTRUCTURE ONLY.
 *
 * Semantics kept from the reference (they decide which rects generate() sees):
 *   - vips_image_generate() records (start, generate, stop, a, b)   iofuncs/generate.c:679
 *   - vips_region_prepare() clips the request to the image and calls the
 *     image's generate fn on the calling thread                      iofuncs/region.c:1646
 *   - demand hint = min over the pipeline                            iofuncs/generate.c:275-293
 *   - sink tile geometry from the hint                               iofuncs/thread.c:288-325
 *   - vips_embed(EXTEND_COPY) = edge replication                     conversion/embed.c:300-336
 * Not kept: threads, buffer caching, sequential mode, ref counting (leaks).
 */
#include <stdarg.h>
#include <float.h>
#include <limits.h>
#include <vips/vips.h>

/* --------------------------------------------------------------- type system */

typedef struct _ShimType {
	struct _ShimType *parent;
	size_t class_size, instance_size;
	void (*instance_init)(void *);
	/* the class struct follows */
} ShimType;

#define SHIM_CLASS(T) ((void *) ((ShimType *) (T) + 1))
#define SHIM_TYPE_OF_CLASS(K) ((ShimType *) (K) -1)

GType vips__shim_type_register(GType parent, size_t class_size, size_t instance_size, void (*class_init)(void *),
	void (*instance_init)(void *), gpointer *parent_class_out)
{
	ShimType *pt = parent ? SHIM_TYPE_OF_CLASS((void *) parent) : NULL;
	ShimType *t;
	if (pt && class_size < pt->class_size)
		class_size = pt->class_size;
	if (pt && instance_size < pt->instance_size)
		instance_size = pt->instance_size;
	t = (ShimType *) calloc(1, sizeof(ShimType) + class_size + 64);
	t->parent = pt;
	t->class_size = class_size;
	t->instance_size = instance_size;
	t->instance_init = instance_init;
	if (pt)
		memcpy(SHIM_CLASS(t), SHIM_CLASS(pt), pt->class_size);
	if (parent_class_out)
		*parent_class_out = pt ? SHIM_CLASS(pt) : NULL;
	if (class_init)
		class_init(SHIM_CLASS(t));
	return (GType) SHIM_CLASS(t);
}

static void shim_init_chain(ShimType *t, void *obj)
{
	if (!t)
		return;
	shim_init_chain(t->parent, obj);
	if (t->instance_init)
		t->instance_init(obj);
}

void *vips__shim_object_new(GType type)
{
	ShimType *t = SHIM_TYPE_OF_CLASS((void *) type);
	GObject *obj = (GObject *) calloc(1, t->instance_size + 64);
	obj->kind = 3;
	obj->klass = (void *) type;
	shim_init_chain(t, obj);
	return obj;
}

static int shim_build_ok(VipsObject *object) { return 0; }
static void shim_operation_class_init(void *klass)
{
	VipsObjectClass *oc = (VipsObjectClass *) klass;
	oc->build = shim_build_ok;
	oc->nickname = "operation";
	oc->description = "shim root";
}

GType vips__shim_operation_get_type(void)
{
	static GType type = 0;
	if (!type)
		type = vips__shim_type_register(0, sizeof(VipsOperationClass) + 256, sizeof(VipsOperation), shim_operation_class_init, NULL, NULL);
	return type;
}

/* parent classes whose own source files are not compiled: plain children of the root */
#define SHIM_PLAIN_TYPE(fn) \
	GType fn(void) \
	{ \
		static GType type = 0; \
		if (!type) \
			type = vips__shim_type_register(vips__shim_operation_get_type(), sizeof(VipsOperationClass) + 256, \
				sizeof(VipsOperation) + 64, NULL, NULL, NULL); \
		return type; \
	}
SHIM_PLAIN_TYPE(vips_resample_get_type)
SHIM_PLAIN_TYPE(vips_conversion_get_type)
SHIM_PLAIN_TYPE(vips_convolution_get_type)
SHIM_PLAIN_TYPE(vips_create_get_type)
SHIM_PLAIN_TYPE(vips_morphology_get_type)
SHIM_PLAIN_TYPE(vips_unary_get_type)

void *g_object_ref(void *p) { return p; }

VipsArrayDouble *vips_array_double_newv(int n, ...)
{
	VipsArea *area = (VipsArea *) calloc(1, sizeof(VipsArea));
	double *d = (double *) calloc(n > 0 ? n : 1, sizeof(double));
	va_list ap;
	int i;
	va_start(ap, n);
	for (i = 0; i < n; i++)
		d[i] = va_arg(ap, double);
	va_end(ap);
	area->data = d;
	area->n = n;
	return (VipsArrayDouble *) area;
}
void vips_area_unref(VipsArea *area) {}
double *vips_array_double_get(VipsArrayDouble *array, int *n)
{
	/* iofuncs/type.c:1115-1126 */
	if (n)
		*n = VIPS_AREA(array)->n;
	return (double *) VIPS_AREA(array)->data;
}
double vips_image_get_format_max(VipsBandFormat format)
{
	/* iofuncs/header.c:440-473 */
	static const double max[] = { UCHAR_MAX, SCHAR_MAX, USHRT_MAX, SHRT_MAX, UINT_MAX, INT_MAX, FLT_MAX, FLT_MAX, DBL_MAX, DBL_MAX };
	return format >= VIPS_FORMAT_UCHAR && format <= VIPS_FORMAT_DPCOMPLEX ? max[format] : -1;
}
void vips_object_set_static(VipsObject *object, gboolean static_object) {}

/* ------------------------------------------------------------------- rects */

void vips_rect_marginadjust(VipsRect *r, int n)
{
	r->left -= n;
	r->top -= n;
	r->width += 2 * n;
	r->height += 2 * n;
}

void vips_rect_intersectrect(const VipsRect *r1, const VipsRect *r2, VipsRect *out)
{
	int left = VIPS_MAX(r1->left, r2->left);
	int top = VIPS_MAX(r1->top, r2->top);
	int right = VIPS_MIN(VIPS_RECT_RIGHT(r1), VIPS_RECT_RIGHT(r2));
	int bottom = VIPS_MIN(VIPS_RECT_BOTTOM(r1), VIPS_RECT_BOTTOM(r2));
	int width = VIPS_MAX(0, right - left);
	int height = VIPS_MAX(0, bottom - top);
	out->left = left;
	out->top = top;
	out->width = width;
	out->height = height;
}

gboolean vips_rect_isempty(const VipsRect *r) { return r->width <= 0 || r->height <= 0; }

void vips_region_paint_pel(VipsRegion *reg, const VipsRect *r, const VipsPel *ink)
{
	const size_t ps = VIPS_IMAGE_SIZEOF_PEL(reg->im);
	int x, y;
	for (y = 0; y < r->height; y++)
		for (x = 0; x < r->width; x++)
			memcpy(VIPS_REGION_ADDR(reg, r->left + x, r->top + y), ink, ps);
}

VipsPel *vips__vector_to_ink(const char *domain, VipsImage *im, double *real, double *imag, int n)
{
	/* background 0 only */
	return (VipsPel *) calloc(1, VIPS_IMAGE_SIZEOF_PEL(im) + 16);
}

int vips_check_vector_length(const char *domain, int n, int len) { return n == len ? 0 : -1; }

int vips__tile_width = 128, vips__tile_height = 128, vips__fatstrip_height = 16, vips__thinstrip_height = 1;

static char shim_error_buf[4096];
void vips_error(const char *domain, const char *fmt, ...)
{
	va_list ap;
	size_t n = strlen(shim_error_buf);
	va_start(ap, fmt);
	n += snprintf(shim_error_buf + n, sizeof(shim_error_buf) - n, "%s: ", domain ? domain : "?");
	if (n < sizeof(shim_error_buf))
		vsnprintf(shim_error_buf + n, sizeof(shim_error_buf) - n, fmt, ap);
	va_end(ap);
}
const char *vips__shim_error(void) { return shim_error_buf; }

void vips_object_set_property(void) {}
void vips_object_get_property(void) {}
void *vips_malloc(VipsObject *object, size_t size) { return calloc(1, size ? size : 1); }

gboolean vips_band_format_iscomplex(VipsBandFormat f) { return f == VIPS_FORMAT_COMPLEX || f == VIPS_FORMAT_DPCOMPLEX; }
gboolean vips_band_format_isint(VipsBandFormat f) { return f >= VIPS_FORMAT_UCHAR && f <= VIPS_FORMAT_INT; }
gboolean vips_band_format_isfloat(VipsBandFormat f) { return f == VIPS_FORMAT_FLOAT || f == VIPS_FORMAT_DOUBLE; }
gboolean vips_band_format_isuint(VipsBandFormat f) { return f == VIPS_FORMAT_UCHAR || f == VIPS_FORMAT_USHORT || f == VIPS_FORMAT_UINT; }
gboolean vips_vector_isenabled(void) { return 0; }

double vips_interpretation_max_alpha(VipsInterpretation interpretation)
{
	/* iofuncs/header.c:195-206 */
	switch (interpretation) {
	case VIPS_INTERPRETATION_GREY16:
	case VIPS_INTERPRETATION_RGB16:
		return 65535.0;
	case VIPS_INTERPRETATION_scRGB:
		return 1.0;
	default:
		return 255.0;
	}
}

/* ------------------------------------------------------------------ images */

VipsImage *vips_image_new(void)
{
	VipsImage *im = (VipsImage *) calloc(1, sizeof(VipsImage));
	im->parent_instance.parent_instance.kind = 1;
	im->dhint = VIPS_DEMAND_STYLE_ANY;
	im->Coding = VIPS_CODING_NONE;
	im->Xres = im->Yres = 1.0;
	return im;
}

VipsImage *vips__shim_image_from_memory(const void *data, int w, int h, int bands, VipsBandFormat fmt,
	VipsInterpretation type)
{
	VipsImage *im = vips_image_new();
	im->Xsize = w;
	im->Ysize = h;
	im->Bands = bands;
	im->BandFmt = fmt;
	im->Type = type;
	im->data = (VipsPel *) data;
	return im;
}

int vips_image_pipelinev(VipsImage *image, VipsDemandStyle hint, ...)
{
	va_list ap;
	VipsImage *in;
	int first = 1;
	VipsDemandStyle set_hint = hint;
	va_start(ap, hint);
	while ((in = va_arg(ap, VipsImage *))) {
		if (first) {
			/* vips__image_copy_fields_array: header from the first input */
			image->Xsize = in->Xsize;
			image->Ysize = in->Ysize;
			image->Bands = in->Bands;
			image->BandFmt = in->BandFmt;
			image->Coding = in->Coding;
			image->Type = in->Type;
			image->Xres = in->Xres;
			image->Yres = in->Yres;
			/* ... and its metadata: the only items this shim models are a matrix's scale / offset */
			image->mat_scale = in->mat_scale;
			image->mat_offset = in->mat_offset;
			image->mat_meta_set = in->mat_meta_set;
			first = 0;
		}
		if ((int) in->dhint < (int) set_hint)
			set_hint = in->dhint;
	}
	va_end(ap);
	image->dhint = set_hint;
	return 0;
}

int vips_image_generate(VipsImage *image, VipsStartFn start_fn, VipsGenerateFn generate_fn, VipsStopFn stop_fn,
	void *a, void *b)
{
	image->start_fn = start_fn;
	image->generate_fn = generate_fn;
	image->stop_fn = stop_fn;
	image->client1 = a;
	image->client2 = b;
	return 0;
}

int vips_image_decode(VipsImage *in, VipsImage **out) { *out = in; return 0; }
gboolean vips_image_is_sequential(VipsImage *image) { return 0; }
void vips_reorder_margin_hint(VipsImage *image, int margin) {}
int vips_check_noncomplex(const char *domain, VipsImage *im) { return vips_band_format_iscomplex(im->BandFmt) ? -1 : 0; }
int vips_check_uncoded(const char *domain, VipsImage *im) { return 0; }
int vips_check_coding_known(const char *domain, VipsImage *im) { return 0; }
VipsImage **vips_object_local_array(VipsObject *parent, int n) { return (VipsImage **) calloc(n + 1, sizeof(VipsImage *)); }
void vips_object_local(void *parent, void *child) {}
gboolean vips_image_hasalpha(VipsImage *image) { return image->Bands == 2 || image->Bands == 4 || image->Bands > 4; }

gboolean vips_object_argument_isset(VipsObject *object, const char *name)
{
	const char **p;
	for (p = object->set_args; p && *p; p++)
		if (strcmp(*p, name) == 0)
			return TRUE;
	return FALSE;
}

/* --------------------------------------------------------- matrix images */

double vips_image_get_scale(const VipsImage *image) { return image->mat_meta_set & 1 ? image->mat_scale : 1.0; }
double vips_image_get_offset(const VipsImage *image) { return image->mat_meta_set & 2 ? image->mat_offset : 0.0; }

void vips_image_set_double(VipsImage *image, const char *name, double d)
{
	if (strcmp(name, "scale") == 0) {
		image->mat_scale = d;
		image->mat_meta_set |= 1;
	}
	else if (strcmp(name, "offset") == 0) {
		image->mat_offset = d;
		image->mat_meta_set |= 2;
	}
}

void vips_image_init_fields(VipsImage *image, int xsize, int ysize, int bands, VipsBandFormat format, VipsCoding coding,
	VipsInterpretation interpretation, double xres, double yres)
{
	image->Xsize = xsize;
	image->Ysize = ysize;
	image->Bands = bands;
	image->BandFmt = format;
	image->Coding = coding;
	image->Type = interpretation;
	image->Xres = xres;
	image->Yres = yres;
}

int vips_image_write_prepare(VipsImage *image)
{
	image->data = (VipsPel *) calloc(1, VIPS_IMAGE_SIZEOF_LINE(image) * image->Ysize + 16);
	return image->data ? 0 : -1;
}

VipsImage *vips_image_new_matrix(int width, int height)
{
	VipsImage *im = vips_image_new();
	vips_image_init_fields(im, width, height, 1, VIPS_FORMAT_DOUBLE, VIPS_CODING_NONE, VIPS_INTERPRETATION_MULTIBAND, 1.0, 1.0);
	vips_image_write_prepare(im);
	return im;
}

/* vips_check_matrix, iofuncs/error.c: a 1-band double memory copy carrying scale/offset */
int vips_check_matrix(const char *domain, VipsImage *im, VipsImage **out)
{
	VipsImage *t;
	int i, n = im->Xsize * im->Ysize;
	if (im->Bands != 1 || im->BandFmt != VIPS_FORMAT_DOUBLE || !im->data) {
		vips_error(domain, "shim: matrix images must be 1-band double in memory");
		return -1;
	}
	t = vips_image_new_matrix(im->Xsize, im->Ysize);
	for (i = 0; i < n; i++)
		((double *) t->data)[i] = ((double *) im->data)[i];
	vips_image_set_double(t, "scale", vips_image_get_scale(im));
	vips_image_set_double(t, "offset", vips_image_get_offset(im));
	*out = t;
	return 0;
}

/* ----------------------------------------------------------------- regions */

VipsRegion *vips_region_new(VipsImage *image)
{
	VipsRegion *reg = (VipsRegion *) calloc(1, sizeof(VipsRegion));
	reg->parent_object.parent_instance.kind = 2;
	reg->im = image;
	return reg;
}

void g_object_unref(void *p)
{
	GObject *obj = (GObject *) p;
	if (obj && obj->kind == 2) {
		VipsRegion *reg = (VipsRegion *) p;
		if (reg->seq && reg->im && reg->im->stop_fn)
			reg->im->stop_fn(reg->seq, reg->im->client1, reg->im->client2);
		free(reg->buffer);
		free(reg);
	}
}

int vips_region_prepare(VipsRegion *reg, const VipsRect *r)
{
	VipsImage *im = reg->im;
	const size_t ps = VIPS_IMAGE_SIZEOF_PEL(im);
	VipsRect c;
	int right, bottom;

	/* clip against the image, region.c:1663-1670 */
	c.left = VIPS_MAX(r->left, 0);
	c.top = VIPS_MAX(r->top, 0);
	right = VIPS_MIN(r->left + r->width, im->Xsize);
	bottom = VIPS_MIN(r->top + r->height, im->Ysize);
	c.width = right - c.left;
	c.height = bottom - c.top;
	if (c.width <= 0 || c.height <= 0) {
		vips_error("vips_region_prepare", "valid clipped to nothing");
		return -1;
	}

	if (im->data) {
		reg->valid = c;
		reg->bpl = (int) VIPS_IMAGE_SIZEOF_LINE(im);
		reg->data = im->data + (size_t) c.top * reg->bpl + (size_t) c.left * ps;
		return 0;
	}
	if (!im->generate_fn) {
		vips_error("vips_region_prepare", "image has no pixels");
		return -1;
	}
	if (!reg->seq && im->start_fn) {
		reg->seq = im->start_fn(im, im->client1, im->client2);
		if (!reg->seq)
			return -1;
	}
	{
		const size_t need = (size_t) c.width * c.height * ps;
		gboolean stop = FALSE;
		if (need > reg->buffer_size) {
			free(reg->buffer);
			reg->buffer = (VipsPel *) malloc(need);
			reg->buffer_size = need;
		}
		reg->valid = c;
		reg->bpl = (int) (c.width * ps);
		reg->data = reg->buffer;
		return im->generate_fn(reg, reg->seq, im->client1, im->client2, &stop);
	}
}

void *vips_start_one(VipsImage *out, void *a, void *b) { return vips_region_new((VipsImage *) a); }
int vips_stop_one(void *seq, void *a, void *b) { g_object_unref(seq); return 0; }

void *vips_start_many(VipsImage *out, void *a, void *b)
{
	VipsImage **in = (VipsImage **) a;
	int n, i;
	VipsRegion **ar;
	for (n = 0; in[n]; n++)
		;
	ar = (VipsRegion **) calloc(n + 1, sizeof(VipsRegion *));
	for (i = 0; i < n; i++)
		ar[i] = vips_region_new(in[i]);
	return ar;
}

int vips_stop_many(void *seq, void *a, void *b)
{
	VipsRegion **ar = (VipsRegion **) seq;
	int i;
	if (ar) {
		for (i = 0; ar[i]; i++)
			g_object_unref(ar[i]);
		free(ar);
	}
	return 0;
}

/* --------------------------------------------------------- copy (image_write) */

static int shim_copy_gen(VipsRegion *out_region, void *seq, void *a, void *b, gboolean *stop)
{
	VipsRegion *ir = (VipsRegion *) seq;
	VipsRect *r = &out_region->valid;
	const size_t line = VIPS_REGION_SIZEOF_LINE(out_region);
	int y;
	if (vips_region_prepare(ir, r))
		return -1;
	for (y = 0; y < r->height; y++)
		memcpy(VIPS_REGION_ADDR(out_region, r->left, r->top + y), VIPS_REGION_ADDR(ir, r->left, r->top + y), line);
	return 0;
}

int vips_image_write(VipsImage *image, VipsImage *out)
{
	/* iofuncs/image.c:2646: a THINSTRIP copy of image into out */
	if (vips_image_pipelinev(out, VIPS_DEMAND_STYLE_THINSTRIP, image, NULL))
		return -1;
	return vips_image_generate(out, vips_start_one, shim_copy_gen, vips_stop_one, image, NULL);
}

/* ------------------------------------------------------- embed, EXTEND_COPY */

typedef struct { VipsImage *in; int x, y; } ShimEmbed;

static int shim_embed_gen(VipsRegion *out_region, void *seq, void *a, void *b, gboolean *stop)
{
	VipsRegion *ir = (VipsRegion *) seq;
	ShimEmbed *embed = (ShimEmbed *) b;
	VipsImage *in = embed->in;
	VipsRect *r = &out_region->valid;
	const size_t ps = VIPS_IMAGE_SIZEOF_PEL(in);
	VipsRect need;
	int x, y;
	int x0 = VIPS_CLIP(0, r->left - embed->x, in->Xsize - 1);
	int x1 = VIPS_CLIP(0, r->left + r->width - 1 - embed->x, in->Xsize - 1);
	int y0 = VIPS_CLIP(0, r->top - embed->y, in->Ysize - 1);
	int y1 = VIPS_CLIP(0, r->top + r->height - 1 - embed->y, in->Ysize - 1);

	/* the part of the input under the request (rows/columns preserved), or
	 * the nearest edge line when the request is wholly in the border
	 */
	need.left = x0;
	need.top = y0;
	need.width = x1 - x0 + 1;
	need.height = y1 - y0 + 1;
	if (vips_region_prepare(ir, &need))
		return -1;
	for (y = 0; y < r->height; y++) {
		const int sy = VIPS_CLIP(0, r->top + y - embed->y, in->Ysize - 1);
		VipsPel *q = VIPS_REGION_ADDR(out_region, r->left, r->top + y);
		for (x = 0; x < r->width; x++) {
			const int sx = VIPS_CLIP(0, r->left + x - embed->x, in->Xsize - 1);
			memcpy(q + x * ps, VIPS_REGION_ADDR(ir, sx, sy), ps);
		}
	}
	return 0;
}

int vips_embed(VipsImage *in, VipsImage **out, int x, int y, int width, int height, ...)
{
	/* only "extend", VIPS_EXTEND_COPY is used on the hot path */
	ShimEmbed *embed = (ShimEmbed *) calloc(1, sizeof(ShimEmbed));
	VipsImage *o = vips_image_new();
	embed->in = in;
	embed->x = x;
	embed->y = y;
	vips_image_pipelinev(o, VIPS_DEMAND_STYLE_ANY, in, NULL); /* conversion/embed.c:440 */
	o->Xsize = width;
	o->Ysize = height;
	vips_image_generate(o, vips_start_one, shim_embed_gen, vips_stop_one, in, embed);
	*out = o;
	return 0;
}

/* -------------------------------------------------------------------- sink */

void vips__shim_tile_size(VipsImage *im, int *tile_w, int *tile_h)
{
	/* vips_get_tile_size, iofuncs/thread.c:288-325 */
	switch (im->dhint) {
	case VIPS_DEMAND_STYLE_SMALLTILE:
		*tile_w = vips__tile_width;
		*tile_h = vips__tile_height;
		break;
	case VIPS_DEMAND_STYLE_THINSTRIP:
		*tile_w = im->Xsize;
		*tile_h = im->Xsize > 10000 ? vips__thinstrip_height : vips__fatstrip_height;
		break;
	default:
		*tile_w = im->Xsize;
		*tile_h = vips__fatstrip_height;
		break;
	}
}

int vips__shim_write_to_memory(VipsImage *im, void *out, int tile_w, int tile_h)
{
	/* vips_sink_memory, iofuncs/sinkmemory.c:171-274: tiles left-to-right, top-to-bottom */
	const size_t ps = VIPS_IMAGE_SIZEOF_PEL(im);
	const size_t line = VIPS_IMAGE_SIZEOF_LINE(im);
	VipsRegion *reg = vips_region_new(im);
	int x, y, yy, rc = 0;
	if (tile_w <= 0 || tile_h <= 0)
		vips__shim_tile_size(im, &tile_w, &tile_h);
	for (y = 0; y < im->Ysize && !rc; y += tile_h)
		for (x = 0; x < im->Xsize && !rc; x += tile_w) {
			VipsRect r;
			r.left = x;
			r.top = y;
			r.width = VIPS_MIN(tile_w, im->Xsize - x);
			r.height = VIPS_MIN(tile_h, im->Ysize - y);
			if (vips_region_prepare(reg, &r)) {
				rc = -1;
				break;
			}
			for (yy = 0; yy < r.height; yy++)
				memcpy((VipsPel *) out + (size_t) (y + yy) * line + (size_t) x * ps,
					VIPS_REGION_ADDR(reg, x, y + yy), (size_t) r.width * ps);
		}
	g_object_unref(reg);
	return rc;
}
