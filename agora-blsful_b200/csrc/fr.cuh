// Fr: the scalar field of BLS12-381 (group order r, 255 bits), 8 x 32-bit limbs, Montgomery form with R = 2^256.
// Used for the Lagrange coefficients of threshold-share combination (`shares.combine()` of vsss-rs behind
// `Signature::from_shares` / `PublicKey::from_shares`, reference src/signature.rs:151-165, src/traits/sig_core.rs:92-105):
// a few hundred multiplications per share next to a 255-bit point multiplication, so this is plain CIOS code, not tuned.
#pragma once
#include "fp.cuh"

namespace bls {

struct Fr {
  uint32_t l[8];
};
BLS_CONST uint32_t K_FR_R2[8] = {0xf3f29c6du, 0xc999e990u, 0x87925c23u, 0x2b6cedcbu, 0x7254398fu, 0x05d31496u, 0x9f59ff11u, 0x0748d9d9u};
BLS_CONST uint32_t K_FR_ONE[8] = {0xfffffffeu, 0x00000001u, 0x00034802u, 0x5884b7fau, 0xecbc4ff5u, 0x998c4fefu, 0xacc5056fu, 0x1824b159u};
BLS_CONST uint32_t K_FR_RM2[8] = {0xffffffffu, 0xfffffffeu, 0xfffe5bfeu, 0x53bda402u, 0x09a1d805u, 0x3339d808u, 0x299d7d48u, 0x73eda753u};
constexpr uint32_t K_FR_NINV = 0xffffffffu;  // -r^-1 mod 2^32

// a >= r ?  (raw words)
BLS_HD bool fr_raw_ge_r(const uint32_t* a) {
  for (int i = 7; i >= 0; i--) {
    if (a[i] > K_R_ORDER[i]) return true;
    if (a[i] < K_R_ORDER[i]) return false;
  }
  return true;
}
BLS_HD void fr_cond_sub_r(uint32_t* t, uint32_t top) {
  uint32_t u[8];
  int64_t br = 0;
  for (int i = 0; i < 8; i++) {
    int64_t v = (int64_t)t[i] - (int64_t)K_R_ORDER[i] + br;
    u[i] = (uint32_t)v;
    br = v >> 32;
  }
  if (top || br == 0)
    for (int i = 0; i < 8; i++) t[i] = u[i];
}
// r = a * b / R mod r   (operand scanning, interleaved reduction; result < r)
BLS_HD void fr_mul(Fr& r, const Fr& a, const Fr& b) {
  uint32_t t[10];
  for (int i = 0; i < 10; i++) t[i] = 0;
  for (int i = 0; i < 8; i++) {
    uint64_t c = 0;
    for (int j = 0; j < 8; j++) {
      c += (uint64_t)a.l[j] * b.l[i] + t[j];
      t[j] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[8] = (uint32_t)c;
    t[9] = (uint32_t)(c >> 32);
    const uint32_t m = t[0] * K_FR_NINV;
    c = ((uint64_t)m * K_R_ORDER[0] + t[0]) >> 32;
    for (int j = 1; j < 8; j++) {
      c += (uint64_t)m * K_R_ORDER[j] + t[j];
      t[j - 1] = (uint32_t)c;
      c >>= 32;
    }
    c += t[8];
    t[7] = (uint32_t)c;
    t[8] = t[9] + (uint32_t)(c >> 32);
    t[9] = 0;
  }
  fr_cond_sub_r(t, t[8]);
  for (int i = 0; i < 8; i++) r.l[i] = t[i];
}
BLS_HD void fr_sub(Fr& r, const Fr& a, const Fr& b) {  // a - b mod r, inputs < r
  int64_t br = 0;
  uint32_t t[8];
  for (int i = 0; i < 8; i++) {
    int64_t v = (int64_t)a.l[i] - (int64_t)b.l[i] + br;
    t[i] = (uint32_t)v;
    br = v >> 32;
  }
  if (br) {
    uint64_t c = 0;
    for (int i = 0; i < 8; i++) {
      c += (uint64_t)t[i] + K_R_ORDER[i];
      t[i] = (uint32_t)c;
      c >>= 32;
    }
  }
  for (int i = 0; i < 8; i++) r.l[i] = t[i];
}
BLS_HD bool fr_is_zero(const Fr& a) {
  uint32_t t = 0;
  for (int i = 0; i < 8; i++) t |= a.l[i];
  return t == 0;
}
BLS_HD void fr_from_raw(Fr& r, const uint32_t* raw) {  // raw < r
  Fr x, r2;
  for (int i = 0; i < 8; i++) {
    x.l[i] = raw[i];
    r2.l[i] = K_FR_R2[i];
  }
  fr_mul(r, x, r2);
}
BLS_HD void fr_to_raw(uint32_t* raw, const Fr& a) {
  Fr one, t;
  for (int i = 0; i < 8; i++) one.l[i] = i == 0 ? 1u : 0u;
  fr_mul(t, a, one);
  for (int i = 0; i < 8; i++) raw[i] = t.l[i];
}
BLS_FN void fr_inv(Fr& r, const Fr& a) {  // a^(r-2); a != 0
  Fr acc;
  for (int i = 0; i < 8; i++) acc.l[i] = K_FR_ONE[i];
  for (int w = 7; w >= 0; w--)
    for (int b = 31; b >= 0; b--) {
      fr_mul(acc, acc, acc);
      if ((K_FR_RM2[w] >> b) & 1u) fr_mul(acc, acc, a);
    }
  r = acc;
}
// 32 big-endian bytes -> raw little-endian words
BLS_HD void fr_raw_from_be32(uint32_t* raw, const uint8_t* b) {
  for (int w = 0; w < 8; w++) {
    const uint8_t* p = b + 28 - 4 * w;
    raw[w] = ((uint32_t)p[0] << 24) | ((uint32_t)p[1] << 16) | ((uint32_t)p[2] << 8) | p[3];
  }
}

// Lagrange basis at zero for share i of a set with identifiers ids[0..m) (raw words, all < r, non-zero, distinct):
//   lambda_i = prod_{j != i} x_j / (x_j - x_i)      -> raw words.  Returns false if some x_j == x_i (duplicate identifier).
BLS_FN bool fr_lagrange_at_zero(uint32_t* out_raw, const uint32_t* ids, uint32_t m, uint32_t i) {
  Fr xi, num, den;
  fr_from_raw(xi, ids + 8 * (size_t)i);
  for (int k = 0; k < 8; k++) num.l[k] = den.l[k] = K_FR_ONE[k];
  for (uint32_t j = 0; j < m; j++) {
    if (j == i) continue;
    Fr xj, d;
    fr_from_raw(xj, ids + 8 * (size_t)j);
    fr_sub(d, xj, xi);
    if (fr_is_zero(d)) return false;
    fr_mul(num, num, xj);
    fr_mul(den, den, d);
  }
  Fr di, lam;
  fr_inv(di, den);
  fr_mul(lam, num, di);
  fr_to_raw(out_raw, lam);
  return true;
}

}  // namespace bls
