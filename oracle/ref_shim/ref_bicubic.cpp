/* ref_bicubic.cpp -- the reference's resample/bicubic.cpp compiled in place. TEST INFRASTRUCTURE ONLY. */
#include <vips/vips.h>
#define vips_reduce_make_mask vips_reduce_make_mask__bicubic_copy
#include "bicubic.cpp"
