/* ref_thumbnail.c -- the reference's resample/thumbnail.c compiled in place. TEST INFRASTRUCTURE ONLY.
 *
 * Only the size arithmetic is reached: vips_thumbnail_calculate_shrink (thumbnail.c:412-466),
 * vips_thumbnail_calculate_common_shrink (:471-486) and vips_thumbnail_find_jpegshrink (:490-517), all file-static, so
 * the wrappers live in this translation unit.  The loaders, ICC handling and graph building the rest of the file
 * mentions resolve to the Makefile's abort() stubs.
 */
#include <stdarg.h>
#include <vips/vips.h>

/* the enums thumbnail.c needs that the shim header has no other use for; orders as in include/vips/resample.h:53-59,
 * conversion.h:64-68 and :97-107, foreign.h:121-127, image.h:143-149 */
typedef enum { VIPS_SIZE_BOTH, VIPS_SIZE_UP, VIPS_SIZE_DOWN, VIPS_SIZE_FORCE, VIPS_SIZE_LAST } VipsSize;
typedef enum { VIPS_DIRECTION_HORIZONTAL, VIPS_DIRECTION_VERTICAL, VIPS_DIRECTION_LAST } VipsDirection;
typedef enum {
	VIPS_INTERESTING_NONE, VIPS_INTERESTING_CENTRE, VIPS_INTERESTING_ENTROPY, VIPS_INTERESTING_ATTENTION,
	VIPS_INTERESTING_LOW, VIPS_INTERESTING_HIGH, VIPS_INTERESTING_ALL, VIPS_INTERESTING_SPECIFIC, VIPS_INTERESTING_LAST
} VipsInteresting;
typedef enum { VIPS_FAIL_ON_NONE, VIPS_FAIL_ON_TRUNCATED, VIPS_FAIL_ON_ERROR, VIPS_FAIL_ON_WARNING, VIPS_FAIL_ON_LAST } VipsFailOn;
typedef enum { VIPS_ACCESS_RANDOM, VIPS_ACCESS_SEQUENTIAL, VIPS_ACCESS_SEQUENTIAL_UNBUFFERED, VIPS_ACCESS_LAST } VipsAccess;
typedef struct _VipsBlob VipsBlob;
typedef struct _VipsSource VipsSource;
#define VIPS_META_ICC_NAME "icc-profile-data"
#define VIPS_META_PAGE_HEIGHT "page-height"
#define VIPS_META_ORIENTATION "orientation"

#include "thumbnail.c"

static void
ref__thumbnail_fill(VipsThumbnail *t, int width, int height, int size, int crop, int linear)
{
	memset(t, 0, sizeof(*t));
	t->width = width;
	t->height = height;
	t->size = (VipsSize) size;
	t->crop = (VipsInteresting) crop;
	t->linear = linear;
}

void
ref_thumbnail_calculate_shrink(int input_width, int input_height, int width, int height, int size, int crop,
	double *hshrink, double *vshrink)
{
	VipsThumbnail t;

	ref__thumbnail_fill(&t, width, height, size, crop, 0);
	vips_thumbnail_calculate_shrink(&t, input_width, input_height, hshrink, vshrink);
}

int
ref_thumbnail_find_jpegshrink(int input_width, int input_height, int width, int height, int size, int crop, int linear)
{
	VipsThumbnail t;

	ref__thumbnail_fill(&t, width, height, size, crop, linear);
	return vips_thumbnail_find_jpegshrink(&t, input_width, input_height);
}
