/* vips/vips.h -- GLib/GObject-FREE stand-in for the libvips headers.
 *
 * TEST INFRASTRUCTURE ONLY.  It exists so that the reference's OWN source files
 * (read in place from /root/reference, never copied) can be compiled into
 * oracle/_ref/libvipsref.so and their static generate functions / scanline
 * kernels called on plain memory buffers.  It declares just the types, enums
 * (same numeric values as the real headers) and macros those files touch; the
 * GObject class machinery (G_DEFINE_TYPE, VIPS_ARG_*, class_init bodies) is
 * compiled but never executed.  Nothing here is product code.
 */
#ifndef SHIM_VIPS_H
#define SHIM_VIPS_H

#include <limits.h>
#include <stdarg.h>
#include <math.h>
#include <stddef.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#ifdef __cplusplus
#define restrict __restrict
extern "C" {
#endif

/* ---- glib scalar types */
typedef int gboolean;
typedef void *gpointer;
typedef const void *gconstpointer;
typedef char gchar;
typedef int gint;
typedef unsigned int guint;
typedef unsigned char guchar;
typedef unsigned char guint8;
typedef signed char gint8;
typedef long long gint64;
typedef unsigned long long guint64;
typedef int gint32;
typedef unsigned int guint32;
typedef short gint16;
typedef unsigned short guint16;
typedef size_t gsize;
typedef unsigned long GType;
typedef double gdouble;
typedef struct _GThread GThread;
typedef struct _GMutex { int dummy; } GMutex;
typedef struct _GSList GSList;
typedef struct _GParamSpec GParamSpec;
typedef struct _GValue GValue;
#ifndef TRUE
#define TRUE 1
#define FALSE 0
#endif
#define G_BEGIN_DECLS
#define G_END_DECLS
#define G_STMT_START do
#define G_STMT_END while (0)
#define G_GNUC_UNUSED __attribute__((unused))
#define G_MAXINT INT_MAX
#define G_MAXDOUBLE 1.7976931348623157e308
#define G_STRUCT_OFFSET(T, M) offsetof(T, M)
#define G_N_ELEMENTS(A) (sizeof(A) / sizeof((A)[0]))
#define g_assert(X) ((void) 0)
#define g_assert_not_reached() ((void) 0)
#define g_info(...) ((void) 0)
#define g_warning(...) ((void) 0)
#define g_free free
#define GINT_TO_POINTER(I) ((gpointer) (intptr_t) (I))
#define GPOINTER_TO_INT(P) ((int) (intptr_t) (P))

/* ---- vips basics */
typedef unsigned char VipsPel;
#define VIPS_PI (3.14159265358979323846)
#define VIPS_API
#define VIPS_TARGET_CLONES(T)
#define VIPS_DEPRECATED_MACRO_FOR(F)
#define VIPS_MAX(A, B) ((A) > (B) ? (A) : (B))
#define VIPS_MIN(A, B) ((A) < (B) ? (A) : (B))
#define VIPS_CLIP(A, V, B) VIPS_MAX((A), VIPS_MIN((B), (V)))
#define VIPS_FCLIP(A, V, B) fmax((A), fmin((B), (V)))
#define VIPS_NUMBER(R) ((int) (sizeof(R) / sizeof(R[0])))
#define VIPS_ABS(V) (((V) >= 0) ? (V) : -(V))
#define VIPS_FABS(V) fabs(V)
#define VIPS_RINT(V) rint(V)
#define VIPS_FLOOR(V) floor(V)
#define VIPS_CEIL(V) ceil(V)
#define VIPS_RAD(R) (((R) / 360.0) * 2.0 * VIPS_PI) /* include/vips/util.h:50-51 */
#define VIPS_DEG(A) (((A) / (2.0 * VIPS_PI)) * 360.0)
#define VIPS_ROUND_UINT(R) ((int) ((R) + 0.5))
#define VIPS_ROUND_DOWN(N, P) ((N) - ((N) % (P)))
#define VIPS_ROUND_UP(N, P) (VIPS_ROUND_DOWN((N) + (P) -1, (P)))
#define VIPS_ISNAN(V) isnan(V)
#define VIPS_ISINF(V) isinf(V)
#define VIPS_UNROLL(N, OPER) \
	G_STMT_START { int duff_count = (N); while (duff_count-- > 0) { OPER; } } G_STMT_END

/* include/vips/interpolate.h:109-118 */
#define VIPS_TRANSFORM_SHIFT (6)
#define VIPS_TRANSFORM_SCALE (1 << VIPS_TRANSFORM_SHIFT)
#define VIPS_INTERPOLATE_SHIFT (12)
#define VIPS_INTERPOLATE_SCALE (1 << VIPS_INTERPOLATE_SHIFT)

/* include/vips/colour.h illuminants */
#define VIPS_D65_X0 (95.0470)
#define VIPS_D65_Y0 (100.0)
#define VIPS_D65_Z0 (108.8827)
#define VIPS_D50_X0 (96.4250)
#define VIPS_D50_Y0 (100.0)
#define VIPS_D50_Z0 (82.4680)

typedef enum {
	VIPS_FORMAT_NOTSET = -1, VIPS_FORMAT_UCHAR = 0, VIPS_FORMAT_CHAR = 1, VIPS_FORMAT_USHORT = 2,
	VIPS_FORMAT_SHORT = 3, VIPS_FORMAT_UINT = 4, VIPS_FORMAT_INT = 5, VIPS_FORMAT_FLOAT = 6,
	VIPS_FORMAT_COMPLEX = 7, VIPS_FORMAT_DOUBLE = 8, VIPS_FORMAT_DPCOMPLEX = 9, VIPS_FORMAT_LAST = 10
} VipsBandFormat;

typedef enum {
	VIPS_INTERPRETATION_ERROR = -1, VIPS_INTERPRETATION_MULTIBAND = 0, VIPS_INTERPRETATION_B_W = 1,
	VIPS_INTERPRETATION_HISTOGRAM = 10, VIPS_INTERPRETATION_XYZ = 12, VIPS_INTERPRETATION_LAB = 13,
	VIPS_INTERPRETATION_CMYK = 15, VIPS_INTERPRETATION_LABQ = 16, VIPS_INTERPRETATION_RGB = 17,
	VIPS_INTERPRETATION_CMC = 18, VIPS_INTERPRETATION_LCH = 19, VIPS_INTERPRETATION_LABS = 21,
	VIPS_INTERPRETATION_sRGB = 22, VIPS_INTERPRETATION_YXY = 23, VIPS_INTERPRETATION_FOURIER = 24,
	VIPS_INTERPRETATION_RGB16 = 25, VIPS_INTERPRETATION_GREY16 = 26, VIPS_INTERPRETATION_MATRIX = 27,
	VIPS_INTERPRETATION_scRGB = 28, VIPS_INTERPRETATION_HSV = 29, VIPS_INTERPRETATION_OKLAB = 30,
	VIPS_INTERPRETATION_OKLCH = 31, VIPS_INTERPRETATION_LAST = 32
} VipsInterpretation;

typedef enum { VIPS_CODING_ERROR = -1, VIPS_CODING_NONE = 0, VIPS_CODING_LABQ = 2, VIPS_CODING_RAD = 6 } VipsCoding;
typedef enum {
	VIPS_DEMAND_STYLE_ERROR = -1, VIPS_DEMAND_STYLE_SMALLTILE, VIPS_DEMAND_STYLE_FATSTRIP,
	VIPS_DEMAND_STYLE_THINSTRIP, VIPS_DEMAND_STYLE_ANY
} VipsDemandStyle;
typedef enum {
	VIPS_KERNEL_NEAREST, VIPS_KERNEL_LINEAR, VIPS_KERNEL_CUBIC, VIPS_KERNEL_MITCHELL, VIPS_KERNEL_LANCZOS2,
	VIPS_KERNEL_LANCZOS3, VIPS_KERNEL_MKS2013, VIPS_KERNEL_MKS2021, VIPS_KERNEL_LAST
} VipsKernel;
typedef enum { VIPS_EXTEND_BLACK, VIPS_EXTEND_COPY, VIPS_EXTEND_REPEAT, VIPS_EXTEND_MIRROR, VIPS_EXTEND_WHITE,
	VIPS_EXTEND_BACKGROUND } VipsExtend;
typedef enum { VIPS_PRECISION_INTEGER, VIPS_PRECISION_FLOAT, VIPS_PRECISION_APPROXIMATE } VipsPrecision;
typedef enum { VIPS_OPERATION_NONE = 0, VIPS_OPERATION_SEQUENTIAL = 1, VIPS_OPERATION_NOCACHE = 4,
	VIPS_OPERATION_DEPRECATED = 8 } VipsOperationFlags;
typedef enum { VIPS_ARGUMENT_NONE = 0, VIPS_ARGUMENT_REQUIRED = 1, VIPS_ARGUMENT_INPUT = 16, VIPS_ARGUMENT_OUTPUT = 32,
	VIPS_ARGUMENT_DEPRECATED = 64 } VipsArgumentFlags;
#define VIPS_ARGUMENT_REQUIRED_INPUT (VIPS_ARGUMENT_INPUT | VIPS_ARGUMENT_REQUIRED)
#define VIPS_ARGUMENT_OPTIONAL_INPUT (VIPS_ARGUMENT_INPUT)
#define VIPS_ARGUMENT_REQUIRED_OUTPUT (VIPS_ARGUMENT_OUTPUT | VIPS_ARGUMENT_REQUIRED)
#define VIPS_ARGUMENT_OPTIONAL_OUTPUT (VIPS_ARGUMENT_OUTPUT)

typedef struct _VipsRect { int left, top, width, height; } VipsRect;
#define VIPS_RECT_RIGHT(R) ((R)->left + (R)->width)
#define VIPS_RECT_BOTTOM(R) ((R)->top + (R)->height)

/* ---- object model: plain structs, never instantiated through a type system */
typedef struct _GObject {
	int kind;    /* shim: 1 image, 2 region, 3 anything made by vips__shim_object_new() */
	void *klass; /* shim: the class struct (see vips__shim_type_register) */
} GObject;
typedef struct _GObjectClass {
	void (*set_property)(void);
	void (*get_property)(void);
	void (*dispose)(GObject *);
	void (*finalize)(GObject *);
} GObjectClass;
typedef struct _VipsObject {
	GObject parent_instance;
	gboolean constructed;
	const char **set_args; /* shim: NULL-terminated names vips_object_argument_isset() answers TRUE for */
} VipsObject;
typedef struct _VipsObjectClass {
	GObjectClass parent_class;
	int (*build)(VipsObject *object);
	const char *nickname;
	const char *description;
} VipsObjectClass;
typedef struct _VipsOperation { VipsObject parent_instance; } VipsOperation;
typedef struct _VipsOperationClass {
	VipsObjectClass parent_class;
	VipsOperationFlags flags;
} VipsOperationClass;

struct _VipsRegion;
typedef struct _VipsImage {
	VipsObject parent_instance;
	int Xsize, Ysize, Bands;
	VipsBandFormat BandFmt;
	VipsCoding Coding;
	VipsInterpretation Type;
	double Xres, Yres;
	int Xoffset, Yoffset;
	VipsDemandStyle dhint;
	/* shim: either the whole image in memory ... */
	VipsPel *data;
	/* ... or a lazily evaluated op, exactly as vips_image_generate() records it
	 * (iofuncs/generate.c:679-788)
	 */
	void *(*start_fn)(struct _VipsImage *out, void *a, void *b);
	int (*generate_fn)(struct _VipsRegion *out, void *seq, void *a, void *b, gboolean *stop);
	int (*stop_fn)(void *seq, void *a, void *b);
	void *client1;
	void *client2;
	/* shim: the "scale" / "offset" metadata of matrix images */
	double mat_scale, mat_offset;
	int mat_meta_set;
} VipsImage;

typedef struct _VipsRegion {
	VipsObject parent_object;
	VipsImage *im;
	VipsRect valid;
	VipsPel *data;
	int bpl;
	void *seq;
	/* shim: the buffer this region owns when its image is generated */
	VipsPel *buffer;
	size_t buffer_size;
} VipsRegion;

#define VIPS_FORMAT_SIZEOF_UNSAFE(F) vips__shim_sizeof(F)
static inline size_t vips__shim_sizeof(int f)
{
	static const size_t s[10] = { 1, 1, 2, 2, 4, 4, 4, 8, 8, 16 };
	return f >= 0 && f < 10 ? s[f] : 0;
}
#define VIPS_IMAGE_SIZEOF_ELEMENT(I) (vips__shim_sizeof((I)->BandFmt))
#define VIPS_IMAGE_SIZEOF_PEL(I) (VIPS_IMAGE_SIZEOF_ELEMENT(I) * (I)->Bands)
#define VIPS_IMAGE_SIZEOF_LINE(I) (VIPS_IMAGE_SIZEOF_PEL(I) * (I)->Xsize)
#define VIPS_IMAGE_N_ELEMENTS(I) ((I)->Bands * (I)->Xsize)
#define VIPS_IMAGE_ADDR(I, X, Y) ((I)->data + (size_t) (Y) * VIPS_IMAGE_SIZEOF_LINE(I) + (size_t) (X) * VIPS_IMAGE_SIZEOF_PEL(I))
/* include/vips/region.h:198-233 */
#define VIPS_REGION_LSKIP(R) ((size_t) ((R)->bpl))
#define VIPS_REGION_N_ELEMENTS(R) ((size_t) ((R)->valid.width * (R)->im->Bands))
#define VIPS_REGION_SIZEOF_LINE(R) ((size_t) ((R)->valid.width * VIPS_IMAGE_SIZEOF_PEL((R)->im)))
#define VIPS_REGION_ADDR(R, X, Y) \
	((R)->data + ((Y) - (R)->valid.top) * VIPS_REGION_LSKIP(R) + ((X) - (R)->valid.left) * VIPS_IMAGE_SIZEOF_PEL((R)->im))
#define VIPS_REGION_ADDR_TOPLEFT(R) ((R)->data)
#define VIPS_COUNT_PIXELS(R, N)
#define VIPS_GATE_START(NAME)
#define VIPS_GATE_STOP(NAME)
#define VIPS_DEBUG_MSG(...) ((void) 0)

typedef int (*VipsGenerateFn)(VipsRegion *out, void *seq, void *a, void *b, gboolean *stop);
typedef void *(*VipsStartFn)(VipsImage *out, void *a, void *b);
typedef int (*VipsStopFn)(void *seq, void *a, void *b);

/* ---- class plumbing: compiled, class_init never runs; build() does */
/* A miniature type system: one heap class struct per type, made on first use by
 * copying the parent's class struct and running the type's own class_init --
 * exactly the part of GObject the reference's build()/dispatch code relies on.
 */
GType vips__shim_type_register(GType parent, size_t class_size, size_t instance_size, void (*class_init)(void *),
	void (*instance_init)(void *), gpointer *parent_class_out);
void *vips__shim_object_new(GType type);
GType vips__shim_operation_get_type(void);
#define G_DEFINE_TYPE(TN, t_n, T_P) \
	static void t_n##_class_init(TN##Class *klass); \
	static void t_n##_init(TN *self); \
	static gpointer t_n##_parent_class = 0; \
	GType t_n##_get_type(void) \
	{ \
		static GType type = 0; \
		if (!type) \
			type = vips__shim_type_register((GType) (T_P), sizeof(TN##Class), sizeof(TN), \
				(void (*)(void *)) t_n##_class_init, (void (*)(void *)) t_n##_init, &t_n##_parent_class); \
		return type; \
	}
#define G_DEFINE_ABSTRACT_TYPE(TN, t_n, T_P) G_DEFINE_TYPE(TN, t_n, T_P)
#define G_TYPE_CHECK_INSTANCE_CAST(O, T, C) ((C *) (O))
#define G_TYPE_CHECK_CLASS_CAST(K, T, C) ((C *) (K))
#define G_TYPE_CHECK_INSTANCE_TYPE(O, T) (1)
#define G_TYPE_CHECK_CLASS_TYPE(K, T) (1)
#define G_TYPE_INSTANCE_GET_CLASS(O, T, C) ((C *) ((GObject *) (O))->klass)
#define G_OBJECT(O) ((GObject *) (O))
#define G_OBJECT_CLASS(K) ((GObjectClass *) (K))
#define VIPS_OBJECT(O) ((VipsObject *) (O))
#define VIPS_OBJECT_CLASS(K) ((VipsObjectClass *) (K))
#define VIPS_OBJECT_GET_CLASS(O) ((VipsObjectClass *) ((GObject *) (O))->klass)
#define VIPS_OPERATION(O) ((VipsOperation *) (O))
#define VIPS_OPERATION_CLASS(K) ((VipsOperationClass *) (K))
#define VIPS_IMAGE(O) ((VipsImage *) (O))
#define VIPS_TYPE_OPERATION (vips__shim_operation_get_type())
#define VIPS_TYPE_OBJECT (vips__shim_operation_get_type())
#define VIPS_TYPE_IMAGE 0
#define VIPS_TYPE_KERNEL 0
#define VIPS_TYPE_PRECISION 0
#define VIPS_TYPE_INTERPRETATION 0
#define VIPS_TYPE_EXTEND 0
#define VIPS_TYPE_ARRAY_DOUBLE 0
#define VIPS_TYPE_INTERPOLATE 0
#define VIPS_TYPE_PCS 0
#define VIPS_TYPE_INTENT 0
#define VIPS_ARG_IMAGE(...)
#define VIPS_ARG_INT(...)
#define VIPS_ARG_DOUBLE(...)
#define VIPS_ARG_BOOL(...)
#define VIPS_ARG_ENUM(...)
#define VIPS_ARG_BOXED(...)
#define VIPS_ARG_OBJECT(...)
#define VIPS_ARG_STRING(...)
#define VIPS_ARG_INTERPOLATE(...)
void vips_object_set_property(void);
void vips_object_get_property(void);

/* ---- memory helpers */
#define VIPS_FREEF(F, S) G_STMT_START { if (S) { (void) F((S)); (S) = 0; } } G_STMT_END
#define VIPS_FREE(S) G_STMT_START { if (S) { free((void *) (S)); (S) = 0; } } G_STMT_END
#define VIPS_UNREF(X) VIPS_FREEF(g_object_unref, X)
#define VIPS_MALLOC(OBJ, S) (vips_malloc(VIPS_OBJECT(OBJ), S))
#define VIPS_NEW(OBJ, T) ((T *) VIPS_MALLOC(OBJ, sizeof(T)))
#define VIPS_ARRAY(OBJ, N, T) ((T *) VIPS_MALLOC(OBJ, (N) * sizeof(T)))
void *vips_malloc(VipsObject *object, size_t size);
void g_object_unref(void *p);

/* ---- the functions the hot-path files call.  Only the few the generate
 * functions reach have bodies (shim_runtime.c); the graph-building ones are
 * declared so the build() functions compile and are stubbed to abort().
 */
int vips_region_prepare(VipsRegion *reg, const VipsRect *r);
VipsRegion *vips_region_new(VipsImage *image);
void vips_error(const char *domain, const char *fmt, ...);
gboolean vips_band_format_iscomplex(VipsBandFormat format);
gboolean vips_band_format_isint(VipsBandFormat format);
gboolean vips_band_format_isfloat(VipsBandFormat format);
gboolean vips_band_format_isuint(VipsBandFormat format);
gboolean vips_vector_isenabled(void);
double vips_interpretation_max_alpha(VipsInterpretation interpretation);
void *vips_start_one(VipsImage *out, void *a, void *b);
int vips_stop_one(void *seq, void *a, void *b);
void *vips_start_many(VipsImage *out, void *a, void *b);
int vips_stop_many(void *seq, void *a, void *b);
extern int vips__fatstrip_height;
extern int vips__tile_width, vips__tile_height, vips__thinstrip_height;

VipsImage **vips_object_local_array(VipsObject *parent, int n);
void vips_object_local(void *parent, void *child);
gboolean vips_object_argument_isset(VipsObject *object, const char *name);
VipsImage *vips_image_new(void);
int vips_image_pipelinev(VipsImage *image, VipsDemandStyle hint, ...);
int vips_image_generate(VipsImage *image, VipsStartFn start_fn, VipsGenerateFn generate_fn, VipsStopFn stop_fn,
	void *a, void *b);
int vips_image_write(VipsImage *image, VipsImage *out);
int vips_image_decode(VipsImage *in, VipsImage **out);
gboolean vips_image_is_sequential(VipsImage *image);
gboolean vips_image_hasalpha(VipsImage *image);
void vips_reorder_margin_hint(VipsImage *image, int margin);
int vips_check_noncomplex(const char *domain, VipsImage *im);
int vips_check_uncoded(const char *domain, VipsImage *im);
int vips_check_coding_known(const char *domain, VipsImage *im);
int vips_check_matrix(const char *domain, VipsImage *im, VipsImage **out);
int vips_check_separable(const char *domain, VipsImage *im);
int vips_check_bands_atleast(const char *domain, VipsImage *im, int bands);
int vips_embed(VipsImage *in, VipsImage **out, int x, int y, int width, int height, ...);
int vips_shrinkv(VipsImage *in, VipsImage **out, int vshrink, ...);
int vips_shrinkh(VipsImage *in, VipsImage **out, int hshrink, ...);
int vips_sequential(VipsImage *in, VipsImage **out, ...);
int vips_linecache(VipsImage *in, VipsImage **out, ...);
int vips_tilecache(VipsImage *in, VipsImage **out, ...);
int vips_cast(VipsImage *in, VipsImage **out, VipsBandFormat format, ...);
int vips_copy(VipsImage *in, VipsImage **out, ...);
int vips_call_split(const char *operation_name, va_list optional, ...);
void vips_image_set_double(VipsImage *image, const char *name, double d);
void vips_image_init_fields(VipsImage *image, int xsize, int ysize, int bands, VipsBandFormat format, VipsCoding coding,
	VipsInterpretation interpretation, double xres, double yres);
int vips_image_write_prepare(VipsImage *image);
VipsImage *vips_image_new_matrix(int width, int height);
void g_object_set(void *object, const char *first, ...);
double vips_image_get_offset(const VipsImage *image);
double vips_image_get_scale(const VipsImage *image);
#define VIPS_MATRIX(I, X, Y) ((double *) VIPS_IMAGE_ADDR(I, X, Y))


/* ---- bits the colour files touch */
typedef struct { int done; } GOnce;
#define G_ONCE_INIT { 0 }
#define VIPS_ONCE(ONCE, FN, CLIENT) G_STMT_START { if (!(ONCE)->done) { (void) FN(CLIENT); (ONCE)->done = 1; } } G_STMT_END
typedef struct _VipsArea { void *data; size_t length; int n; } VipsArea;
typedef VipsArea VipsArrayDouble;
typedef enum { VIPS_INTENT_PERCEPTUAL = 0, VIPS_INTENT_RELATIVE, VIPS_INTENT_SATURATION, VIPS_INTENT_ABSOLUTE, VIPS_INTENT_AUTO = 32 } VipsIntent;
typedef enum { VIPS_PCS_LAB, VIPS_PCS_XYZ } VipsPCS;
/* include/vips/colour.h:124-138 */
typedef enum {
	VIPS_CICP_COLOUR_PRIMARIES_BT709 = 1, VIPS_CICP_COLOUR_PRIMARIES_UNSPECIFIED = 2,
	VIPS_CICP_COLOUR_PRIMARIES_BT470M = 4, VIPS_CICP_COLOUR_PRIMARIES_BT470BG = 5, VIPS_CICP_COLOUR_PRIMARIES_BT601 = 6,
	VIPS_CICP_COLOUR_PRIMARIES_SMPTE240 = 7, VIPS_CICP_COLOUR_PRIMARIES_GENERIC_FILM = 8,
	VIPS_CICP_COLOUR_PRIMARIES_BT2020 = 9, VIPS_CICP_COLOUR_PRIMARIES_SMPTE428 = 10,
	VIPS_CICP_COLOUR_PRIMARIES_SMPTE431 = 11, VIPS_CICP_COLOUR_PRIMARIES_SMPTE432 = 12,
	VIPS_CICP_COLOUR_PRIMARIES_EBU3213 = 22
} VipsCICPColourPrimaries;
int vips_check_vector_length(const char *domain, int n, int len);
int vips_check_coding(const char *domain, VipsImage *im, VipsCoding coding);
/* include/vips/colour.h: the sRGB <-> linear tables and helpers LabQ2sRGB.c defines */
extern float vips_v2Y_8[256];
extern float vips_v2Y_16[65536];
void vips_col_make_tables_RGB_8(void);
void vips_col_make_tables_RGB_16(void);
int vips_col_sRGB2scRGB_8(int r, int g, int b, float *R, float *G, float *B);
int vips_col_sRGB2scRGB_16(int r, int g, int b, float *R, float *G, float *B);
int vips_col_scRGB2XYZ(float R, float G, float B, float *X, float *Y, float *Z);
int vips_col_XYZ2scRGB(float X, float Y, float Z, float *R, float *G, float *B);
int vips_col_scRGB2sRGB_8(float R, float G, float B, int *r, int *g, int *b, int *og);
int vips_col_scRGB2sRGB_16(float R, float G, float B, int *r, int *g, int *b, int *og);
typedef int (*VipsColourTransformFn)(VipsImage *in, VipsImage **out, ...);


/* ---- include/vips/interpolate.h, include/vips/transform.h */
typedef struct _VipsInterpolate { VipsObject parent_object; } VipsInterpolate;
typedef void (*VipsInterpolateMethod)(VipsInterpolate *interpolate, void *out, VipsRegion *in, double x, double y);
typedef struct _VipsInterpolateClass {
	VipsObjectClass parent_class;
	VipsInterpolateMethod interpolate;
	int (*get_window_size)(VipsInterpolate *interpolate);
	int window_size;
	int (*get_window_offset)(VipsInterpolate *interpolate);
	int window_offset;
} VipsInterpolateClass;
#define VIPS_TYPE_INTERPOLATE (vips_interpolate_get_type())
#define VIPS_INTERPOLATE(obj) ((VipsInterpolate *) (obj))
#define VIPS_INTERPOLATE_CLASS(klass) ((VipsInterpolateClass *) (klass))
#define VIPS_INTERPOLATE_GET_CLASS(obj) ((VipsInterpolateClass *) ((GObject *) (obj))->klass)
GType vips_interpolate_get_type(void);
void vips_interpolate(VipsInterpolate *interpolate, void *out, VipsRegion *in, double x, double y);
VipsInterpolateMethod vips_interpolate_get_method(VipsInterpolate *interpolate);
int vips_interpolate_get_window_size(VipsInterpolate *interpolate);
int vips_interpolate_get_window_offset(VipsInterpolate *interpolate);
VipsInterpolate *vips_interpolate_new(const char *nickname);
VipsInterpolate *vips_interpolate_nearest_static(void);
VipsInterpolate *vips_interpolate_bilinear_static(void);
void vips__interpolate_init(void);
typedef struct {
	VipsRect iarea;
	VipsRect oarea;
	double a, b, c, d;
	double idx, idy;
	double odx, ody;
	double ia, ib, ic, id;
} VipsTransformation;
void vips__transform_init(VipsTransformation *trn);
int vips__transform_calc_inverse(VipsTransformation *trn);
int vips__transform_isidentity(const VipsTransformation *trn);
int vips__transform_add(const VipsTransformation *in1, const VipsTransformation *in2, VipsTransformation *out);
void vips__transform_print(const VipsTransformation *trn);
void vips__transform_forward_point(const VipsTransformation *trn, const double x, const double y, double *ox, double *oy);
void vips__transform_invert_point(const VipsTransformation *trn, const double x, const double y, double *ox, double *oy);
void vips__transform_forward_rect(const VipsTransformation *trn, const VipsRect *in, VipsRect *out);
void vips__transform_invert_rect(const VipsTransformation *trn, const VipsRect *in, VipsRect *out);
void vips__transform_set_area(VipsTransformation *);
#define VIPS_ROUND_INT(R) ((int) ((R) > 0 ? ((R) + 0.5) : ((R) -0.5)))
#define VIPS_AREA(X) ((VipsArea *) (X))
void vips_rect_marginadjust(VipsRect *r, int n);
void vips_rect_intersectrect(const VipsRect *r1, const VipsRect *r2, VipsRect *out);
gboolean vips_rect_isempty(const VipsRect *r);
void vips_region_paint_pel(VipsRegion *reg, const VipsRect *r, const VipsPel *ink);
VipsPel *vips__vector_to_ink(const char *domain, VipsImage *im, double *real, double *imag, int n);
void *g_object_ref(void *p);
VipsArrayDouble *vips_array_double_newv(int n, ...);
void vips_area_unref(VipsArea *area);
double *vips_array_double_get(VipsArrayDouble *array, int *n);
double vips_image_get_format_max(VipsBandFormat format);
void vips_object_set_static(VipsObject *object, gboolean static_object);
#define DBL_MIN_SHIM 2.2250738585072014e-308

/* shim: the sink.  Evaluate a lazy image into packed memory with the tile
 * geometry vips_get_tile_size() would pick from its demand hint
 * (iofuncs/thread.c:288-325) or an explicit one.
 */
int vips__shim_write_to_memory(VipsImage *im, void *out, int tile_w, int tile_h);
VipsImage *vips__shim_image_from_memory(const void *data, int w, int h, int bands, VipsBandFormat fmt,
	VipsInterpretation type);
void vips__shim_tile_size(VipsImage *im, int *tile_w, int *tile_h);
const char *vips__shim_error(void);

#ifdef __cplusplus
}
#endif

#endif
