// Exhaustive check (all 2^31 non-negative binary32 values, ~1 min on 8 cores) that
//   q = s*c; r = fmaf(-q, 9, s); q2 = fmaf(r, c, q),  c = fl32(1/9)
// equals the IEEE division s / 9.0f.  Build: gcc -O2 -march=x86-64-v3 -ffp-contract=off -fopenmp exhaustive_div9.c -lm
// Result recorded in DESIGN.md: bad=0.
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <math.h>
int main(){
  const float c = 1.0f/9.0f;
  uint64_t bad=0, first=0; 
  #pragma omp parallel for reduction(+:bad)
  for (int64_t u=0; u<0x7f800000LL; ++u){
    uint32_t uu=(uint32_t)u; float s; memcpy(&s,&uu,4);
    float ref = s/9.0f;
    float q = s*c; float r = fmaf(-q,9.0f,s); float q2 = fmaf(r,c,q);
    if (memcmp(&ref,&q2,4)) { bad++; if(!first){first=u;} }
  }
  printf("bad=%llu first=%llx\n",(unsigned long long)bad,(unsigned long long)first);
  // find range of bad
  uint32_t lo=0xffffffff, hi=0;
  for (int64_t u=0; u<0x7f800000LL; ++u){
    uint32_t uu=(uint32_t)u; float s; memcpy(&s,&uu,4);
    float ref = s/9.0f;
    float q = s*c; float r = fmaf(-q,9.0f,s); float q2 = fmaf(r,c,q);
    if (memcmp(&ref,&q2,4)) { if(uu<lo)lo=uu; if(uu>hi)hi=uu; }
  }
  float flo,fhi; memcpy(&flo,&lo,4); memcpy(&fhi,&hi,4);
  printf("bad range %x (%g) .. %x (%g)\n",lo,flo,hi,fhi);
}
