/* div_const_check.c -- the 3-instruction constant division of libvips_b200/csrc/colour.cu
 * (div_const: q0 = x * RN(1/d); r = fma(-q0, d, x); q = fma(r, RN(1/d), q0)) against the
 * hardware quotient, for every float mantissa (both signs; also as float * 100000 products)
 * and 2 * 10^6 random doubles, for each divisor the colour kernels use.  Built and run by
 * tests/test_div_const.py; exit status 0 = identical everywhere.
 */
#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
static inline double mk(double x, double d, double r){ double q0 = x*r; double rem = fma(-q0, d, x); if (rem == 0.0 || !(fabs(q0) < INFINITY)) return q0; return fma(rem, r, q0);} 
int main(){
  double ds[] = {100.0, 95.0470, 108.8827, 903.3, 116.0, 500.0, 200.0, 7.787, 32767.0/100.0, 32768.0/128.0};
  long bad=0;
  for (int k=0;k<10;k++){ double d=ds[k], r=1.0/d; long b=0;
    for (uint32_t m=0;m<(1u<<23);m++){ uint32_t bits = 0x3f800000u | m; float f; memcpy(&f,&bits,4); 
      for (int s=-1;s<=1;s+=2){ double x=(double)f*s; double a=x/d, c=mk(x,d,r); if (memcmp(&a,&c,8)) b++; 
        double x2 = x*100000.0f; /* float product cases */ a=x2/d; c=mk(x2,d,r); if (memcmp(&a,&c,8)) b++; }
    }
    /* random doubles */
    uint64_t st=88172645463325252ull;
    for (long i=0;i<2000000;i++){ st^=st<<13; st^=st>>7; st^=st<<17; uint64_t mb=(st>>12)|0x3ff0000000000000ull; double x; memcpy(&x,&mb,8); if (st&1) x=-x; x*= (double)(1<<(st%20)); double a=x/d,c=mk(x,d,r); if (memcmp(&a,&c,8)) b++; }
    printf("d=%g mismatches %ld\n", d, b); bad+=b; }
  double z=-0.0; double a=z/100.0, c=mk(z,100.0,0.01); printf("-0: %d inf: %g nan: %g\n", signbit(c)==signbit(a), mk(INFINITY,100.0,0.01), mk(NAN,100,0.01));
  return bad!=0; }
