// RFC 9380 hash_to_curve, suites BLS12381G2_XMD:SHA-256_SSWU_RO_ and BLS12381G1_XMD:SHA-256_SSWU_RO_, one message per
// thread.  Replaces `G::hash::<ExpandMsgXmd<Sha256>>(m, dst)` (reference src/impls/g2.rs:15-17, src/impls/g1.rs:17-19)
// called from core_verify / core_aggregate_verify (src/traits/sig_core.rs:136,168).
#pragma once
#include "curve.cuh"

namespace bls {

// ------------------------------------------------------------------------------------------------ SHA-256
BLS_CONST uint32_t K_SHA256[64] = {
    0x428a2f98, 0x71374491, 0xb5c0fbcf, 0xe9b5dba5, 0x3956c25b, 0x59f111f1, 0x923f82a4, 0xab1c5ed5, 0xd807aa98, 0x12835b01,
    0x243185be, 0x550c7dc3, 0x72be5d74, 0x80deb1fe, 0x9bdc06a7, 0xc19bf174, 0xe49b69c1, 0xefbe4786, 0x0fc19dc6, 0x240ca1cc,
    0x2de92c6f, 0x4a7484aa, 0x5cb0a9dc, 0x76f988da, 0x983e5152, 0xa831c66d, 0xb00327c8, 0xbf597fc7, 0xc6e00bf3, 0xd5a79147,
    0x06ca6351, 0x14292967, 0x27b70a85, 0x2e1b2138, 0x4d2c6dfc, 0x53380d13, 0x650a7354, 0x766a0abb, 0x81c2c92e, 0x92722c85,
    0xa2bfe8a1, 0xa81a664b, 0xc24b8b70, 0xc76c51a3, 0xd192e819, 0xd6990624, 0xf40e3585, 0x106aa070, 0x19a4c116, 0x1e376c08,
    0x2748774c, 0x34b0bcb5, 0x391c0cb3, 0x4ed8aa4a, 0x5b9cca4f, 0x682e6ff3, 0x748f82ee, 0x78a5636f, 0x84c87814, 0x8cc70208,
    0x90befffa, 0xa4506ceb, 0xbef9a3f7, 0xc67178f2};

BLS_HD uint32_t rotr32(uint32_t x, int n) { return (x >> n) | (x << (32 - n)); }

struct Sha256 {
  uint32_t h[8];
  uint8_t buf[64];
  uint32_t fill;
  uint64_t total;
};

BLS_FN void sha256_compress(uint32_t* h, const uint8_t* blk) {
  uint32_t w[64];
  for (int i = 0; i < 16; i++)
    w[i] = ((uint32_t)blk[4 * i] << 24) | ((uint32_t)blk[4 * i + 1] << 16) | ((uint32_t)blk[4 * i + 2] << 8) | blk[4 * i + 3];
  for (int i = 16; i < 64; i++) {
    uint32_t s0 = rotr32(w[i - 15], 7) ^ rotr32(w[i - 15], 18) ^ (w[i - 15] >> 3);
    uint32_t s1 = rotr32(w[i - 2], 17) ^ rotr32(w[i - 2], 19) ^ (w[i - 2] >> 10);
    w[i] = w[i - 16] + s0 + w[i - 7] + s1;
  }
  uint32_t a = h[0], b = h[1], c = h[2], d = h[3], e = h[4], f = h[5], g = h[6], hh = h[7];
  for (int i = 0; i < 64; i++) {
    uint32_t S1 = rotr32(e, 6) ^ rotr32(e, 11) ^ rotr32(e, 25);
    uint32_t ch = (e & f) ^ (~e & g);
    uint32_t t1 = hh + S1 + ch + K_SHA256[i] + w[i];
    uint32_t S0 = rotr32(a, 2) ^ rotr32(a, 13) ^ rotr32(a, 22);
    uint32_t mj = (a & b) ^ (a & c) ^ (b & c);
    uint32_t t2 = S0 + mj;
    hh = g;
    g = f;
    f = e;
    e = d + t1;
    d = c;
    c = b;
    b = a;
    a = t1 + t2;
  }
  h[0] += a;
  h[1] += b;
  h[2] += c;
  h[3] += d;
  h[4] += e;
  h[5] += f;
  h[6] += g;
  h[7] += hh;
}
BLS_HD void sha256_init(Sha256& s) {
  s.h[0] = 0x6a09e667;
  s.h[1] = 0xbb67ae85;
  s.h[2] = 0x3c6ef372;
  s.h[3] = 0xa54ff53a;
  s.h[4] = 0x510e527f;
  s.h[5] = 0x9b05688c;
  s.h[6] = 0x1f83d9ab;
  s.h[7] = 0x5be0cd19;
  s.fill = 0;
  s.total = 0;
}
BLS_HD void sha256_update(Sha256& s, const uint8_t* p, uint32_t n) {
  s.total += n;
  for (uint32_t i = 0; i < n; i++) {
    s.buf[s.fill++] = p[i];
    if (s.fill == 64) {
      sha256_compress(s.h, s.buf);
      s.fill = 0;
    }
  }
}
BLS_HD void sha256_update_zero(Sha256& s, uint32_t n) {
  s.total += n;
  for (uint32_t i = 0; i < n; i++) {
    s.buf[s.fill++] = 0;
    if (s.fill == 64) {
      sha256_compress(s.h, s.buf);
      s.fill = 0;
    }
  }
}
BLS_HD void sha256_final(Sha256& s, uint8_t* out) {
  uint64_t bits = s.total * 8;
  s.buf[s.fill++] = 0x80;
  if (s.fill > 56) {
    while (s.fill < 64) s.buf[s.fill++] = 0;
    sha256_compress(s.h, s.buf);
    s.fill = 0;
  }
  while (s.fill < 56) s.buf[s.fill++] = 0;
  for (int i = 0; i < 8; i++) s.buf[56 + i] = (uint8_t)(bits >> (56 - 8 * i));
  sha256_compress(s.h, s.buf);
  for (int i = 0; i < 8; i++) {
    out[4 * i] = (uint8_t)(s.h[i] >> 24);
    out[4 * i + 1] = (uint8_t)(s.h[i] >> 16);
    out[4 * i + 2] = (uint8_t)(s.h[i] >> 8);
    out[4 * i + 3] = (uint8_t)s.h[i];
  }
}

// expand_message_xmd (RFC 9380 5.3.1), SHA-256, message = prefix || msg (prefix is the MessageAugmentation pk bytes,
// reference src/traits/sig_aug.rs:20-24,41-47), dst_len <= 255, out_len = 32*ell <= 256.
BLS_FN void expand_message_xmd(uint8_t* out, uint32_t out_len, const uint8_t* prefix, uint32_t prefix_len, const uint8_t* msg,
                               uint32_t msg_len, const uint8_t* dst, uint32_t dst_len) {
  const uint32_t ell = out_len / 32;
  uint8_t b0[32], bi[32], t[32];
  uint8_t tail[3];
  uint8_t dl = (uint8_t)dst_len;
  Sha256 s;
  sha256_init(s);
  sha256_update_zero(s, 64);
  if (prefix_len) sha256_update(s, prefix, prefix_len);
  sha256_update(s, msg, msg_len);
  tail[0] = (uint8_t)(out_len >> 8);
  tail[1] = (uint8_t)out_len;
  tail[2] = 0;
  sha256_update(s, tail, 3);
  sha256_update(s, dst, dst_len);
  sha256_update(s, &dl, 1);
  sha256_final(s, b0);
  for (uint32_t i = 1; i <= ell; i++) {
    if (i == 1) {
      for (int k = 0; k < 32; k++) t[k] = b0[k];
    } else {
      for (int k = 0; k < 32; k++) t[k] = b0[k] ^ bi[k];
    }
    uint8_t ib = (uint8_t)i;
    sha256_init(s);
    sha256_update(s, t, 32);
    sha256_update(s, &ib, 1);
    sha256_update(s, dst, dst_len);
    sha256_update(s, &dl, 1);
    sha256_final(s, bi);
    for (int k = 0; k < 32; k++) out[32 * (i - 1) + k] = bi[k];
  }
}

// 64 big-endian bytes -> Fp (Montgomery), i.e. OS2IP(bytes) mod p
BLS_HD void fp_from_be64_mod(Fp& r, const uint8_t* b) {
  uint32_t wl[8], wh[8];
#pragma unroll
  for (int i = 0; i < 8; i++) {
    const uint8_t* q = b + 60 - 4 * i;  // low 256 bits: bytes 32..63
    wl[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | q[3];
    const uint8_t* qh = b + 28 - 4 * i;  // high 256 bits: bytes 0..31
    wh[i] = ((uint32_t)qh[0] << 24) | ((uint32_t)qh[1] << 16) | ((uint32_t)qh[2] << 8) | qh[3];
  }
  Fp hi, lo, c, t;
  fp_raw_from_words(lo, wl, 8);
  fp_raw_from_words(hi, wh, 8);
  fp_set(c, K_R2);
  fp_mul(lo, lo, c);
  fp_set(c, K_R2_256);
  fp_mul(t, hi, c);
  fp_add(r, lo, t);
  fp_norm(r, r);
}

// ------------------------------------------------------------------------------------------------ G2 suite
// Simplified SWU for E2': y^2 = x^3 + A x + B (A = 240u, B = 1012(1+u), Z = -(2+u)); returns the point as
// x = xn/xd together with the AFFINE y (needed for the sgn0 rule), no inversion.
BLS_FN void sswu_g2(Fp2& xn, Fp2& xd, Fp2& y, const Fp2& u) {
  Fp2 A, B, Z, zu2, tv1, x1n, gxn, gxd, t, xd2;
  fp2_set(A, K_SSWU2_A);
  fp2_set(B, K_SSWU2_B);
  fp2_set(Z, K_SSWU2_Z);
  fp2_sqr(t, u);
  fp2_mul(zu2, Z, t);  // Z u^2
  fp2_sqr(tv1, zu2);
  fadd(tv1, tv1, zu2);  // Z^2 u^4 + Z u^2
  Fp2 one;
  fone(one);
  fadd(x1n, tv1, one);
  fp2_mul(x1n, x1n, B);  // B (tv1 + 1)
  if (fis_zero(tv1)) {
    fp2_set(xd, K_SSWU2_ZA);  // x1 = B / (Z A)
  } else {
    fp2_mul(xd, A, tv1);
    fneg(xd, xd);  // -A tv1
  }
  fp2_sqr(xd2, xd);
  fp2_mul(gxd, xd2, xd);  // xd^3
  fp2_mul(t, A, xd2);
  fp2_sqr(gxn, x1n);
  fadd(gxn, gxn, t);
  fp2_mul(gxn, gxn, x1n);  // x1n^3 + A x1n xd^2
  fp2_mul(t, B, gxd);
  fadd(gxn, gxn, t);  // + B xd^3
  Fp2 root;
  bool sq = fp2_sqrt_ratio(root, gxn, gxd);
  if (sq) {
    xn = x1n;
    y = root;
  } else {
    // x2 = Z u^2 x1 ; y2 = sqrt(Z^3 u^6 gx1) = Z u^3 sqrt(Z gx1)
    fp2_mul(xn, zu2, x1n);
    fp2_mul(t, zu2, u);
    fp2_mul(y, t, root);
  }
  if (fp2_sgn0(u) != fp2_sgn0(y)) fneg(y, y);
}

// evaluates sum k_i xn^i xd^(deg-i) by Horner on the pair (xn, xd); pw[j] = xd^j precomputed
BLS_HD void iso_poly2(Fp2& r, const uint32_t (*k)[2][NL], int deg, const Fp2& xn, const Fp2* pw) {
  Fp2 acc, c, t;
  fp2_set(acc, k[deg]);
  for (int i = deg - 1; i >= 0; i--) {
    fp2_mul(acc, acc, xn);
    fp2_set(c, k[i]);
    fp2_mul(t, c, pw[deg - i]);
    fadd(acc, acc, t);
  }
  r = acc;
}

// 3-isogeny E2' -> E2 (RFC 9380 E.3) on x = xn/xd, affine y; Jacobian output
BLS_FN void iso3_map(G2Jac& r, const Fp2& xn, const Fp2& xd, const Fp2& y) {
  Fp2 pw[4];
  fone(pw[0]);
  pw[1] = xd;
  fp2_sqr(pw[2], xd);
  fp2_mul(pw[3], pw[2], xd);
  Fp2 XN, XD, YN, YD, W, t, yd2;
  iso_poly2(XN, K_ISO3_XN, 3, xn, pw);
  iso_poly2(XD, K_ISO3_XD, 2, xn, pw);
  iso_poly2(YN, K_ISO3_YN, 3, xn, pw);
  iso_poly2(YD, K_ISO3_YD, 3, xn, pw);
  // x' = XN / (XD xd) ; y' = y YN / YD ; Z' = W YD with W = XD xd
  fp2_mul(W, XD, xd);
  fp2_mul(r.Z, W, YD);
  fp2_sqr(yd2, YD);
  fp2_mul(t, XN, W);
  fp2_mul(r.X, t, yd2);  // XN W YD^2
  fp2_sqr(t, W);
  fp2_mul(t, t, W);
  fp2_mul(t, t, yd2);
  fp2_mul(t, t, YN);
  fp2_mul(r.Y, t, y);  // y YN W^3 YD^2
}

// h_eff multiplication by the psi method: [x^2-x-1]P + [x-1]psi(P) + psi^2(2P)   (RFC 9380 appendix G.3)
BLS_FN void g2_clear_cofactor(G2Jac& r, const G2Jac& p) {
  G2Jac t1, t2, t3, n;
  jac_mul_xabs(t1, p);
  jac_neg(t1, t1);  // x P
  g2_psi(t2, p);    // psi(P)
  jac_dbl(t3, p);
  g2_psi2(t3, t3);  // psi^2(2P)
  jac_neg(n, t2);
  jac_add(t3, t3, n);  // psi^2(2P) - psi(P)
  jac_add(t2, t1, t2);  // xP + psi(P)
  jac_mul_xabs(n, t2);
  jac_neg(n, n);  // x (xP + psi(P))
  jac_add(t3, t3, n);
  jac_neg(n, t1);
  jac_add(t3, t3, n);  // - xP
  jac_neg(n, p);
  jac_add(r, t3, n);  // - P
}

// msg' = prefix || msg.  The batch kernels run hash_to_curve as two launches - everything up to Q0 + Q1 (hash_map_g2: a
// point of E2, not yet of G2), then clear_cofactor - so that neither carries the other's stack frame (kernels.cuh k_hash /
// k_clear_cofactor).
BLS_FN void hash_field_g2(Fp2* u, const uint8_t* prefix, uint32_t prefix_len, const uint8_t* msg, uint32_t msg_len,
                          const uint8_t* dst, uint32_t dst_len) {
  uint8_t ub[256];
  expand_message_xmd(ub, 256, prefix, prefix_len, msg, msg_len, dst, dst_len);
  fp_from_be64_mod(u[0].c0, ub);
  fp_from_be64_mod(u[0].c1, ub + 64);
  fp_from_be64_mod(u[1].c0, ub + 128);
  fp_from_be64_mod(u[1].c1, ub + 192);
}
BLS_FN void map_to_curve_g2(G2Jac& r, const Fp2& u) {
  Fp2 xn, xd, y;
  sswu_g2(xn, xd, y, u);
  iso3_map(r, xn, xd, y);
}
BLS_FN void hash_map_g2(G2Jac& r, const uint8_t* prefix, uint32_t prefix_len, const uint8_t* msg, uint32_t msg_len,
                       const uint8_t* dst, uint32_t dst_len) {
  Fp2 u[2];
  hash_field_g2(u, prefix, prefix_len, msg, msg_len, dst, dst_len);
  G2Jac q0, q1;
  map_to_curve_g2(q0, u[0]);
  map_to_curve_g2(q1, u[1]);
  jac_add(r, q0, q1);
}
BLS_FN void hash_to_g2(G2Jac& r, const uint8_t* prefix, uint32_t prefix_len, const uint8_t* msg, uint32_t msg_len,
                       const uint8_t* dst, uint32_t dst_len) {
  G2Jac q;
  hash_map_g2(q, prefix, prefix_len, msg, msg_len, dst, dst_len);
  g2_clear_cofactor(r, q);
}

// ------------------------------------------------------------------------------------------------ G1 suite
BLS_FN void sswu_g1(Fp& xn, Fp& xd, Fp& y, const Fp& u) {
  Fp A, B, Z, zu2, tv1, x1n, gxn, gxd, t, xd2, one;
  fp_set(A, K_SSWU1_A);
  fp_set(B, K_SSWU1_B);
  fp_set(Z, K_SSWU1_Z);
  fp_one(one);
  fp_sqr(t, u);
  fp_mul(zu2, Z, t);
  fp_sqr(tv1, zu2);
  fp_add(tv1, tv1, zu2);
  fp_add(x1n, tv1, one);
  fp_mul(x1n, x1n, B);
  if (fp_is_zero(tv1)) {
    fp_set(xd, K_SSWU1_ZA);
  } else {
    fp_mul(xd, A, tv1);
    fp_neg(xd, xd);
  }
  fp_sqr(xd2, xd);
  fp_mul(gxd, xd2, xd);
  fp_mul(t, A, xd2);
  fp_sqr(gxn, x1n);
  fp_add(gxn, gxn, t);
  fp_mul(gxn, gxn, x1n);
  fp_mul(t, B, gxd);
  fp_add(gxn, gxn, t);
  // sqrt(gxn/gxd) = gxn * (gxn gxd)^((p-3)/4) when square; else sqrt(Z gxn/gxd) = sqrt(-11) * (that value)
  Fp s, e, root, chk;
  fp_mul(s, gxn, gxd);
  fp_isqrt_pow(e, s);
  fp_mul(root, gxn, e);
  fp_sqr(chk, e);
  fp_mul(chk, chk, s);
  if (fp_eq(chk, one) || fp_is_zero(s)) {
    xn = x1n;
    y = root;
  } else {
    Fp c;
    fp_set(c, K_SQRT_M11);
    fp_mul(root, root, c);  // sqrt(Z gx1)
    fp_mul(xn, zu2, x1n);
    fp_mul(t, zu2, u);
    fp_mul(y, t, root);
  }
  if (fp_sgn0(u) != fp_sgn0(y)) fp_neg(y, y);
}

BLS_HD void iso_poly1(Fp& r, const uint32_t (*k)[NL], int deg, const Fp& xn, const Fp* pw) {
  Fp acc, c, t;
  fp_set(acc, k[deg]);
  for (int i = deg - 1; i >= 0; i--) {
    fp_mul(acc, acc, xn);
    fp_set(c, k[i]);
    fp_mul(t, c, pw[deg - i]);
    fp_add(acc, acc, t);
  }
  r = acc;
}

// 11-isogeny E1' -> E1 (RFC 9380 E.2): degrees x_num 11, x_den 10, y_num 15, y_den 15
BLS_FN void iso11_map(G1Jac& r, const Fp& xn, const Fp& xd, const Fp& y) {
  Fp pw[16];
  fp_one(pw[0]);
  pw[1] = xd;
  for (int i = 2; i < 16; i++) fp_mul(pw[i], pw[i - 1], xd);
  Fp XN, XD, YN, YD, W, t, yd2;
  iso_poly1(XN, K_ISO11_XN, 11, xn, pw);
  iso_poly1(XD, K_ISO11_XD, 10, xn, pw);
  iso_poly1(YN, K_ISO11_YN, 15, xn, pw);
  iso_poly1(YD, K_ISO11_YD, 15, xn, pw);
  // x' = XN / (XD xd) ; y' = y YN / YD
  fp_mul(W, XD, xd);
  fp_mul(r.Z, W, YD);
  fp_sqr(yd2, YD);
  fp_mul(t, XN, W);
  fp_mul(r.X, t, yd2);
  fp_sqr(t, W);
  fp_mul(t, t, W);
  fp_mul(t, t, yd2);
  fp_mul(t, t, YN);
  fp_mul(r.Y, t, y);
}

BLS_FN void hash_field_g1(Fp* u, const uint8_t* prefix, uint32_t prefix_len, const uint8_t* msg, uint32_t msg_len,
                          const uint8_t* dst, uint32_t dst_len) {
  uint8_t ub[128];
  expand_message_xmd(ub, 128, prefix, prefix_len, msg, msg_len, dst, dst_len);
  fp_from_be64_mod(u[0], ub);
  fp_from_be64_mod(u[1], ub + 64);
}
BLS_FN void map_to_curve_g1(G1Jac& r, const Fp& u) {
  Fp xn, xd, y;
  sswu_g1(xn, xd, y, u);
  iso11_map(r, xn, xd, y);
}
BLS_FN void hash_map_g1(G1Jac& r, const uint8_t* prefix, uint32_t prefix_len, const uint8_t* msg, uint32_t msg_len,
                       const uint8_t* dst, uint32_t dst_len) {
  Fp u[2];
  hash_field_g1(u, prefix, prefix_len, msg, msg_len, dst, dst_len);
  G1Jac q0, q1;
  map_to_curve_g1(q0, u[0]);
  map_to_curve_g1(q1, u[1]);
  jac_add(r, q0, q1);
}
// h_eff = 1 - x = 1 + |x|
BLS_FN void g1_clear_cofactor(G1Jac& r, const G1Jac& q) {
  G1Jac t;
  jac_mul_xabs(t, q);
  jac_add(r, t, q);
}
BLS_FN void hash_to_g1(G1Jac& r, const uint8_t* prefix, uint32_t prefix_len, const uint8_t* msg, uint32_t msg_len,
                       const uint8_t* dst, uint32_t dst_len) {
  G1Jac q;
  hash_map_g1(q, prefix, prefix_len, msg, msg_len, dst, dst_len);
  g1_clear_cofactor(r, q);
}

}  // namespace bls
