#include <stdio.h>
#include <stdlib.h>
void vips_call_split(void) { fputs("ref shim: vips_call_split() is not available", stderr); abort(); }
void vips_colour_code_get_type(void) { fputs("ref shim: vips_colour_code_get_type() is not available", stderr); abort(); }
void vips_colour_transform_get_type(void) { fputs("ref shim: vips_colour_transform_get_type() is not available", stderr); abort(); }
void vips_conva(void) { fputs("ref shim: vips_conva() is not available", stderr); abort(); }
void vips_convasep(void) { fputs("ref shim: vips_convasep() is not available", stderr); abort(); }
void vips_interpolate_lbb_get_type(void) { fputs("ref shim: vips_interpolate_lbb_get_type() is not available", stderr); abort(); }
void vips_interpolate_nohalo_get_type(void) { fputs("ref shim: vips_interpolate_nohalo_get_type() is not available", stderr); abort(); }
void vips_interpolate_vsqbs_get_type(void) { fputs("ref shim: vips_interpolate_vsqbs_get_type() is not available", stderr); abort(); }
void vips_object_new(void) { fputs("ref shim: vips_object_new() is not available", stderr); abort(); }
void vips_sequential(void) { fputs("ref shim: vips_sequential() is not available", stderr); abort(); }
void vips_subsample(void) { fputs("ref shim: vips_subsample() is not available", stderr); abort(); }
void vips_type_find(void) { fputs("ref shim: vips_type_find() is not available", stderr); abort(); }
void vips_zoom(void) { fputs("ref shim: vips_zoom() is not available", stderr); abort(); }
