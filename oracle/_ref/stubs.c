#include <stdio.h>
#include <stdlib.h>
void vips_affine(void) { fputs("ref shim: vips_affine() is not available", stderr); abort(); }
void vips_call_split(void) { fputs("ref shim: vips_call_split() is not available", stderr); abort(); }
void vips_interpolate_new(void) { fputs("ref shim: vips_interpolate_new() is not available", stderr); abort(); }
void vips_sequential(void) { fputs("ref shim: vips_sequential() is not available", stderr); abort(); }
void vips_subsample(void) { fputs("ref shim: vips_subsample() is not available", stderr); abort(); }
void vips_zoom(void) { fputs("ref shim: vips_zoom() is not available", stderr); abort(); }
