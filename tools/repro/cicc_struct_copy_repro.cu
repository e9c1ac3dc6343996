#include <stdint.h>
struct alignas(16) S { uint32_t l[16]; };
struct P { S c0, c1; };
struct H { P c0, c1, c2; };
struct T { H c0, c1; };
__device__ __forceinline__ P* pick(T& f, int k) {
  H& h = (k & 1) ? f.c1 : f.c0;
  return (k >> 1) == 0 ? &h.c0 : (k >> 1) == 1 ? &h.c1 : &h.c2;
}
__device__ __forceinline__ void make(S& r, const uint8_t* b) {
#pragma unroll
  for (int i = 0; i < 14; i++) r.l[i] = (uint32_t)b[i] * 3u + b[i + 1];
  r.l[14] = r.l[15] = 0;
}
__global__ void k(const uint8_t* __restrict__ in, T* __restrict__ out) {
  T f;
  for (int k = 0; k < 6; k++) {
    P c;
    for (int h = 0; h < 2; h++) {
      uint8_t b[48];
      for (int j = 0; j < 48; j++) b[j] = in[96 * k + 48 * h + j];
      make(h ? c.c1 : c.c0, b);
    }
    *pick(f, k) = c;
  }
  out[threadIdx.x] = f;
}
