/* shim: everything lives in vips/vips.h */
#include <vips/vips.h>
