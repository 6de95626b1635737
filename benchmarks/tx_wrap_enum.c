/* tx_wrap_enum.c -- CPU proof by enumeration for the modulator's fp32 phase wrap (m17_sdr_b200/csrc/mod.cuh, tx_wrap):
 * for every float with biased exponent 60..149, both signs, the fp32 form (Markstein-corrected quotient with tuned constants,
 * floor of |q| by an add that rounds toward zero -- floorf here --, error-free product, sign of the quotient ORed in) against
 * the reference's (float)((double)x / (2 pi)), modf, (float)((double)f * 2.0 * M_PI)  (m17_modulate.cpp:33-37).
 * Build: gcc -O2 -mfma -ffp-contract=off -fopenmp -o tx_wrap_enum tx_wrap_enum.c -lm        Result on record: total mismatches 0.
 * The GPU repeats the proof on all 2^32 bit patterns with the real instruction sequence (m17b_selftest_tx_wrap, a -m gpu test). */
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
#include <fenv.h>
#include <omp.h>
static inline float u2f(uint32_t u){float f;memcpy(&f,&u,4);return f;}
static inline uint32_t f2u(float f){uint32_t u;memcpy(&u,&f,4);return u;}
int main(){
  const double C = 2.0*M_PI;
  const float r=u2f(0x3E22F983u), r2=u2f(0x3E22F984u), chi=u2f(0x40C90FDBu), cloq=u2f(0xB43BBD2Du), clo=u2f(0xB43BBD2Eu);
  long tot=0;
  for (int e=60;e<150;e++){
    long mis=0;
    #pragma omp parallel for reduction(+:mis)
    for (uint32_t m=0;m<(1u<<23);m++) for(int s=0;s<2;s++){
      float x=u2f(((uint32_t)e<<23)|m|((uint32_t)s<<31));
      float a1=(float)((double)x/C); double ip; float fr=(float)modf((double)a1,&ip); float refw=(float)((double)fr*2.0*M_PI);
      float q0=x*r; float e1=fmaf(-q0,chi,x); float e2=fmaf(-q0,cloq,e1); float q1=fmaf(e2,r2,q0);
      float aq=fabsf(q1); volatile float t1=aq+8388608.0f; /* RN here; emulate RZ: */ float fl=floorf(aq);
      float f=aq-fl; float p=f*chi; float ep=fmaf(f,chi,-p); float t=fmaf(f,clo,ep);
      uint32_t out=f2u(p+t)|(f2u(q1)&0x80000000u);
      if(out!=f2u(refw)) mis++;
    }
    tot+=mis; if(mis) printf("e=%d mis=%ld\n",e,mis);
  }
  printf("total mismatches %ld\n",tot);
}
