#include <math.h>
#include <stdio.h>
#include <stdint.h>
#include <string.h>
#include <omp.h>
int main() {
  float ds[2] = {50.f, 3600.f};
  for (int k = 0; k < 2; ++k) {
    const float d = ds[k];
    const float r = (float)(1.0 / (double)d);
    uint64_t bad = 0; uint32_t firstbad = 0, lastbad = 0;
    #pragma omp parallel for reduction(+:bad) schedule(static)
    for (uint64_t u = 0; u <= 0x7f7fffffull; ++u) {
      uint32_t b = (uint32_t)u; float x; memcpy(&x, &b, 4);
      float q0 = x * r;
      float e = fmaf(-d, q0, x);
      float q = fmaf(e, r, q0);
      float t = x / d;
      if (memcmp(&q, &t, 4) != 0) { bad++; 
        #pragma omp critical
        { if (!firstbad || b < firstbad) firstbad = b; if (b > lastbad) lastbad = b; } }
    }
    float fb, lb; memcpy(&fb, &firstbad, 4); memcpy(&lb, &lastbad, 4);
    printf("d=%g r=%a mismatches=%llu first=%g (0x%08x) last=%g (0x%08x)\n", d, r, (unsigned long long)bad, fb, firstbad, lb, lastbad);
  }
  return 0;
}
