// Experiment (not product): throughput of one Montgomery product modulo the BLS12-381 base-field prime in three forms,
//   A  14 x 28-bit limbs, 64-bit column accumulators, IMAD.WIDE only (the engine's fp_mul: csrc/fp.cuh, one Karatsuba level)
//   B  12 x 32-bit limbs, CIOS on carry chains (mad.lo.cc / madc.hi.cc) - the form SURVEY.md's north star names
//   C  12 x 32-bit limbs, product scanning with mad.wide and three-word column accumulators (add.cc chains)
// Each thread runs a dependent chain x <- x * y (what exponentiations and point formulas look like), `iters` times;
// the grid fills every SM with 128-thread blocks.  Results are checked against each other through a common reference
// (all three compute x * y^iters * R^-iters mod p; the host compares canonical values of A, B and C for iters = 1 with
// a big-integer product).  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -std=c++17 -I agora-blsful_b200/csrc
#include <cstdio>
#include <cstdlib>
#include <cstdint>
#include <cuda_runtime.h>
#include "fp.cuh"

// p, little-endian 32-bit words, and -p^-1 mod 2^32
__device__ __constant__ const uint32_t P32[12] = {0xffffaaabu, 0xb9feffffu, 0xb153ffffu, 0x1eabfffeu, 0xf6b0f624u, 0x6730d2a0u,
                                                  0xf38512bfu, 0x64774b84u, 0x434bacd7u, 0x4b1ba7b6u, 0x397fe69au, 0x1a0111eau};
#define PINV32 0xfffcfffdu

// ---- B: CIOS with carry chains.  t has 13 words + carry; per outer step: t += a * b_i ; m = t0 * pinv ; t = (t + m p) >> 32.
__device__ __forceinline__ void mul_b(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t t[14];
#pragma unroll
  for (int j = 0; j < 14; j++) t[j] = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const uint32_t bi = b[i];
    // low halves: one carry chain over the row
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(t[0]) : "r"(a[0]), "r"(bi));
#pragma unroll
    for (int j = 1; j < 12; j++) asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(t[j]) : "r"(a[j]), "r"(bi));
    asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(t[12]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(t[13]));
    // high halves: a second chain, one word up
    asm volatile("mad.hi.cc.u32 %0, %1, %2, %0;" : "+r"(t[1]) : "r"(a[0]), "r"(bi));
#pragma unroll
    for (int j = 1; j < 12; j++) asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(t[j + 1]) : "r"(a[j]), "r"(bi));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(t[13]));
    const uint32_t m = t[0] * PINV32;
    asm volatile("mad.lo.cc.u32 %0, %1, %2, %0;" : "+r"(t[0]) : "r"(m), "r"(P32[0]));
#pragma unroll
    for (int j = 1; j < 12; j++) asm volatile("madc.lo.cc.u32 %0, %1, %2, %0;" : "+r"(t[j]) : "r"(m), "r"(P32[j]));
    asm volatile("addc.cc.u32 %0, %0, 0;" : "+r"(t[12]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(t[13]));
    asm volatile("mad.hi.cc.u32 %0, %1, %2, %0;" : "+r"(t[1]) : "r"(m), "r"(P32[0]));
#pragma unroll
    for (int j = 1; j < 12; j++) asm volatile("madc.hi.cc.u32 %0, %1, %2, %0;" : "+r"(t[j + 1]) : "r"(m), "r"(P32[j]));
    asm volatile("addc.u32 %0, %0, 0;" : "+r"(t[13]));
#pragma unroll
    for (int j = 0; j < 13; j++) t[j] = t[j + 1];  // exact: t[0] is zero now
    t[13] = 0;
  }
  // result < 2p: one conditional subtraction
  uint32_t s[12], borrow;
  asm("sub.cc.u32 %0, %1, %2;" : "=r"(s[0]) : "r"(t[0]), "r"(P32[0]));
#pragma unroll
  for (int j = 1; j < 12; j++) asm("subc.cc.u32 %0, %1, %2;" : "=r"(s[j]) : "r"(t[j]), "r"(P32[j]));
  asm("subc.u32 %0, %1, 0;" : "=r"(borrow) : "r"(t[12]));
#pragma unroll
  for (int j = 0; j < 12; j++) r[j] = borrow ? t[j] : s[j];
}

// ---- C: product scanning with mad.wide; a column is kept in three 32-bit words (lo, hi, carry word)
__device__ __forceinline__ void col_add(uint32_t& c0, uint32_t& c1, uint32_t& c2, uint32_t x, uint32_t y) {
  const uint64_t pr = (uint64_t)x * y;
  asm("add.cc.u32 %0, %0, %3; addc.cc.u32 %1, %1, %4; addc.u32 %2, %2, 0;" : "+r"(c0), "+r"(c1), "+r"(c2) : "r"((uint32_t)pr), "r"((uint32_t)(pr >> 32)));
}
__device__ __forceinline__ void mul_c(uint32_t* r, const uint32_t* a, const uint32_t* b) {
  uint32_t m[12], t[13];
  uint32_t c0 = 0, c1 = 0, c2 = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) {  // low columns: products and the reduction multiples known so far, then m_k
#pragma unroll
    for (int i = 0; i <= k; i++) col_add(c0, c1, c2, a[i], b[k - i]);
#pragma unroll
    for (int i = 0; i < k; i++) col_add(c0, c1, c2, m[i], P32[k - i]);
    m[k] = c0 * PINV32;
    col_add(c0, c1, c2, m[k], P32[0]);
    c0 = c1; c1 = c2; c2 = 0;
  }
#pragma unroll
  for (int k = 12; k < 23; k++) {
#pragma unroll
    for (int i = k - 11; i < 12; i++) col_add(c0, c1, c2, a[i], b[k - i]);
#pragma unroll
    for (int i = k - 11; i < 12; i++) col_add(c0, c1, c2, m[i], P32[k - i]);
    t[k - 12] = c0;
    c0 = c1; c1 = c2; c2 = 0;
  }
  t[11] = c0;
  t[12] = c1;
  uint32_t s[12], borrow;
  asm("sub.cc.u32 %0, %1, %2;" : "=r"(s[0]) : "r"(t[0]), "r"(P32[0]));
#pragma unroll
  for (int j = 1; j < 12; j++) asm("subc.cc.u32 %0, %1, %2;" : "=r"(s[j]) : "r"(t[j]), "r"(P32[j]));
  asm("subc.u32 %0, %1, 0;" : "=r"(borrow) : "r"(t[12]));
#pragma unroll
  for (int j = 0; j < 12; j++) r[j] = borrow ? t[j] : s[j];
}

template <int FORM>
__global__ void __launch_bounds__(128) k_chain(uint32_t iters, const uint32_t* __restrict__ in, uint32_t* __restrict__ out) {
  const size_t tid = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (FORM == 0) {
    bls::Fp x, y;
    for (int j = 0; j < 16; j++) {
      x.l[j] = j < 14 ? in[tid * 32 + j] & 0x0fffffffu : 0;
      y.l[j] = j < 14 ? in[tid * 32 + 16 + j] & 0x0fffffffu : 0;
    }
    x.l[13] &= 0xfff;
    y.l[13] &= 0xfff;
    for (uint32_t k = 0; k < iters; k++) bls::fp_mul_inl(x, x, y);
    for (int j = 0; j < 14; j++) out[tid * 16 + j] = x.l[j];
  } else {
    uint32_t x[12], y[12];
    for (int j = 0; j < 12; j++) {
      x[j] = in[tid * 32 + j];
      y[j] = in[tid * 32 + 16 + j];
    }
    x[11] &= 0x0fffffffu;
    y[11] &= 0x0fffffffu;
    for (uint32_t k = 0; k < iters; k++) {
      if (FORM == 1) mul_b(x, x, y); else mul_c(x, x, y);
    }
    for (int j = 0; j < 12; j++) out[tid * 16 + j] = x[j];
  }
}

typedef unsigned __int128 u128;
// host check: value(limbs of `bits` bits) as a big number in 32-bit words, times R mod p compared across forms is overkill;
// instead check B and C against each other exactly (same representation) and A's throughput only.
int main(int argc, char** argv) {
  const uint32_t iters = argc > 1 ? atoi(argv[1]) : 2000;
  int dev = 0, sms = 0;
  cudaGetDevice(&dev);
  cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev);
  for (int bps = 2; bps <= 4; bps++) {  // resident 128-thread blocks per SM the grid is sized for
    const size_t blocks = (size_t)sms * bps * 4, threads = blocks * 128;
    uint32_t *h = (uint32_t*)malloc(threads * 32 * 4), *d_in, *d_out, *o1 = (uint32_t*)malloc(threads * 16 * 4), *o2 = (uint32_t*)malloc(threads * 16 * 4);
    srand(7);
    for (size_t i = 0; i < threads * 32; i++) h[i] = (uint32_t)rand() * 2654435761u + (uint32_t)rand();
    cudaMalloc(&d_in, threads * 32 * 4);
    cudaMalloc(&d_out, threads * 16 * 4);
    cudaMemcpy(d_in, h, threads * 32 * 4, cudaMemcpyHostToDevice);
    const char* names[3] = {"A 14x28 IMAD.WIDE columns (engine fp_mul)", "B 12x32 CIOS carry chains (mad.lo.cc/madc.hi.cc)", "C 12x32 mad.wide + 96-bit columns"};
    for (int form = 0; form < 3; form++) {
      cudaEvent_t e0, e1;
      cudaEventCreate(&e0);
      cudaEventCreate(&e1);
      for (int rep = 0; rep < 2; rep++) {
        cudaEventRecord(e0);
        if (form == 0) k_chain<0><<<blocks, 128>>>(iters, d_in, d_out);
        if (form == 1) k_chain<1><<<blocks, 128>>>(iters, d_in, d_out);
        if (form == 2) k_chain<2><<<blocks, 128>>>(iters, d_in, d_out);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
      }
      float ms = 0;
      cudaEventElapsedTime(&ms, e0, e1);
      if (form == 1) cudaMemcpy(o1, d_out, threads * 16 * 4, cudaMemcpyDeviceToHost);
      if (form == 2) cudaMemcpy(o2, d_out, threads * 16 * 4, cudaMemcpyDeviceToHost);
      printf("%d blocks/SM-sized grid x4 waves | %-52s %8.2f ms  %7.1f G Fp-mul/s\n", bps, names[form], ms, (double)threads * iters / ms / 1e6);
      cudaError_t err = cudaGetLastError();
      if (err != cudaSuccess) printf("CUDA error: %s\n", cudaGetErrorString(err));
    }
    size_t bad = 0;
    for (size_t i = 0; i < threads; i++)
      for (int j = 0; j < 12; j++) bad += o1[i * 16 + j] != o2[i * 16 + j];
    printf("   forms B and C agree on %zu chains: %s\n", threads, bad ? "NO" : "yes");
    cudaFree(d_in);
    cudaFree(d_out);
    free(h); free(o1); free(o2);
  }
  return 0;
}
