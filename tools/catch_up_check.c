#include <stdio.h>
#include <stdlib.h>
#include <string.h>
#include <math.h>
#include <stdint.h>
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
/* u (+) w repeated n times, w = -2^k */
static float catchup(float u, float w, int k, int n, int *iters) {
  while (n > 0) {
    (*iters)++;
    const uint32_t bits = f2u(u);
    const int expf = (bits >> 23) & 0xff;
    if (expf == 0) {              /* zero or denormal: one plain step */
      u = u + w; n--; continue;
    }
    const float B = u2f((bits & 0x7f800000u) + 0x00800000u);  /* 2^(e+1) */
    const int shift = k + 150 - expf;                          /* log2(|w| / ulp(u)) */
    if (shift < 0) { u = u + w; n--; continue; }              /* w finer than u's grid: every step rounds */
    const float r = fmaf(w, (float)n, u);
    if (fabsf(r) <= B) return r;
    const int32_t m = (int32_t)((bits & 0x7fffffu) | 0x800000u);
    const int32_t Ug = (bits >> 31) ? -m : m;
    const int32_t num = Ug + (1 << 24);
    const int J = shift >= 25 ? 0 : (num >> shift);
    /* n > J here (else r would have been within B) */
    u = fmaf(w, (float)(J + 1), u);
    n -= J + 1;
  }
  return u;
}
/* tools/catch_up_check.c [cases]: the closed form of k_dp_regtu's catch_up (rnacode_b200/csrc/rc_kernels.cuh -- the same statements
 * in C) against the step-by-step loop `u = u + w`, bit for bit, on random (value, k, n).  gcc -O2 -ffp-contract=off ... -lm */
int main(int argc, char **argv){
  const long cases = argc > 1 ? atol(argv[1]) : 20000000;
  srand48(12345);
  long bad=0, tot=0, it=0, maxit=0;
  const int ks[]={1,2,0,-1,3,-3,5};
  for (long t=0;t<cases;t++){
    int k=ks[lrand48()%7]; float w=-ldexpf(1.0f,k);
    int mode=lrand48()%6; float u;
    if(mode==0) u=(float)((drand48()-0.5)*20000.0);
    else if(mode==1) u=(float)((drand48()-0.5)*40.0);
    else if(mode==2) u=(float)(-(drand48())*5000.0);
    else if(mode==3) u=(float)((drand48()-0.5)*1e-3);
    else if(mode==4) u=ldexpf((float)(drand48()+1.0), (int)(lrand48()%30)-10) * (lrand48()&1?-1:1);
    else u = (float)((long)(drand48()*64)-32) * 0.25f;   /* exact grid values incl 0 */
    int n=(int)(lrand48()% (mode==3?50:1700));
    float s=u; for(int i=0;i<n;i++){ volatile float x=s+w; s=x; }
    int iters=0; float c=catchup(u,w,k,n,&iters);
    it+=iters; if(iters>maxit)maxit=iters;
    tot++;
    if (f2u(s)!=f2u(c)) { if(bad<10) printf("MISMATCH u=%a w=%a n=%d seq=%a got=%a\n",u,w,n,s,c); bad++; }
  }
  printf("tested %ld mismatches %ld avg iters %.2f max %ld\n",tot,bad,(double)it/tot,maxit);
  return bad!=0;
}
