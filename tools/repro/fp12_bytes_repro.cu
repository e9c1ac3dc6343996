// Reproducer for the Fp12 byte-conversion anomaly seen in round 2 (k_fp12_from_bytes returned garbage in the first
// coefficient and zeros elsewhere on the device while its PTX reads correct).  Standalone: nvcc ... -o repro; ./repro
#include <cstdio>
#include <cstring>
#include "../../agora-blsful_b200/csrc/kernels.cuh"
using namespace bls;

// variant 2: one thread per Fp (12 n threads), result written straight to global memory
__global__ void k_from_bytes_v2(size_t n, const uint8_t* in, Fp12* out, uint8_t* bad) {
  size_t t = BLS_TID();
  if (t >= 12 * n) return;
  const size_t i = t / 12;
  const int k = (int)((t % 12) / 2), h = (int)(t & 1);
  uint8_t b[48];
  for (int j = 0; j < 48; j++) b[j] = in[576 * i + 96 * k + 48 * h + j];
  uint8_t flag = (b[0] & 0xe0) ? 1 : 0;
  b[0] &= 0x1f;
  Fp raw, m;
  if (!fp_from_be48_raw(raw, b)) flag = 1;
  fp_to_mont(m, raw);
  Fp2* c = fp12_coeff(out[i], k);
  if (h) c->c1 = m; else c->c0 = m;
  if (flag) bad[i] = 1;
}

int main() {
  uint8_t h_in[576], h_out[576];
  for (int i = 0; i < 576; i++) h_in[i] = (uint8_t)(i * 37 + 11);
  for (int c = 0; c < 12; c++) h_in[48 * c] &= 0x0f;  // < p
  uint8_t *d_in, *d_out, *d_bad;
  Fp12* d_f;
  cudaMalloc(&d_in, 576); cudaMalloc(&d_out, 576); cudaMalloc(&d_bad, 16); cudaMalloc(&d_f, sizeof(Fp12));
  cudaMemcpy(d_in, h_in, 576, cudaMemcpyHostToDevice);
  for (int variant = 0; variant < 2; variant++) {
    cudaMemset(d_f, 0xee, sizeof(Fp12)); cudaMemset(d_out, 0, 576); cudaMemset(d_bad, 0, 16);
    if (variant == 0) k_fp12_from_bytes<<<1, 128>>>(1, d_in, d_f, d_bad);
    else k_from_bytes_v2<<<1, 128>>>(1, d_in, d_f, d_bad);
    k_fp12_to_bytes<<<1, 32>>>(1, d_f, d_out);
    cudaError_t e = cudaDeviceSynchronize();
    cudaMemcpy(h_out, d_out, 576, cudaMemcpyDeviceToHost);
    uint8_t bad = 0; cudaMemcpy(&bad, d_bad, 1, cudaMemcpyDeviceToHost);
    int diff = 0, first = -1;
    for (int i = 0; i < 576; i++) if (h_in[i] != h_out[i]) { diff++; if (first < 0) first = i; }
    printf("variant %d: err=%s bad=%d round-trip diffs=%d first=%d\n", variant, cudaGetErrorString(e), bad, diff, first);
    if (diff) { for (int i = 0; i < 96; i++) printf("%02x", h_out[i]); printf("\n"); for (int i = 0; i < 96; i++) printf("%02x", h_in[i]); printf("\n"); }
  }
  return 0;
}
