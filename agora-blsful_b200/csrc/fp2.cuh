// Fp2 = Fp[u]/(u^2+1).  Overloaded f* helpers give Fp and Fp2 one vocabulary so the curve code is generic.
#pragma once
#include "fp.cuh"

namespace bls {

struct Fp2 {
  Fp c0, c1;
};

// ---- generic vocabulary for Fp (lazy additive ops: see the bound discipline in fp.cuh)
BLS_HD void fadd(Fp& r, const Fp& a, const Fp& b) { fp_add(r, a, b); }
template <int K>
BLS_HD void fsub_k(Fp& r, const Fp& a, const Fp& b) { fp_sub_k<K>(r, a, b); }
template <int K>
BLS_HD void fneg_k(Fp& r, const Fp& a) { fp_neg_k<K>(r, a); }
// defaults: subtrahend value < 16, result normalised (the raw lazy forms are the _k templates)
BLS_HD void fsub(Fp& r, const Fp& a, const Fp& b) {
  fp_sub_k<16>(r, a, b);
  fp_norm(r, r);
}
BLS_HD void fneg(Fp& r, const Fp& a) {
  fp_neg_k<16>(r, a);
  fp_norm(r, r);
}
BLS_HD void fdbl(Fp& r, const Fp& a) { fp_add(r, a, a); }
BLS_HD void fnorm(Fp& r, const Fp& a) { fp_norm(r, a); }
BLS_HD void fred(Fp& r, const Fp& a) { fp_red(r, a); }
BLS_HD void fmul(Fp& r, const Fp& a, const Fp& b) { fp_mul(r, a, b); }
BLS_HD void fsqr(Fp& r, const Fp& a) { fp_sqr(r, a); }
BLS_HD bool fis_zero(const Fp& a) { return fp_is_zero(a); }
BLS_HD bool feq(const Fp& a, const Fp& b) { return fp_eq(a, b); }
BLS_HD void fzero(Fp& r) { fp_zero(r); }
BLS_HD void fone(Fp& r) { fp_one(r); }
BLS_HD void fselect(Fp& r, bool c, const Fp& a, const Fp& b) { fp_select(r, c, a, b); }
BLS_HD void finv(Fp& r, const Fp& a) { fp_inv(r, a); }

// ---- Fp2
BLS_HD void fp2_set(Fp2& r, const uint32_t (*c)[NL]) {
  fp_set(r.c0, c[0]);
  fp_set(r.c1, c[1]);
}
BLS_HD void fadd(Fp2& r, const Fp2& a, const Fp2& b) {
  fp_add(r.c0, a.c0, b.c0);
  fp_add(r.c1, a.c1, b.c1);
}
BLS_HD void fnorm(Fp2& r, const Fp2& a) {
  fp_norm(r.c0, a.c0);
  fp_norm(r.c1, a.c1);
}
BLS_HD void fred(Fp2& r, const Fp2& a) {
  fp_red(r.c0, a.c0);
  fp_red(r.c1, a.c1);
}
template <int K>
BLS_HD void fsub_k(Fp2& r, const Fp2& a, const Fp2& b) {
  fp_sub_k<K>(r.c0, a.c0, b.c0);
  fp_sub_k<K>(r.c1, a.c1, b.c1);
}
template <int K>
BLS_HD void fneg_k(Fp2& r, const Fp2& a) {
  fp_neg_k<K>(r.c0, a.c0);
  fp_neg_k<K>(r.c1, a.c1);
}
BLS_HD void fsub(Fp2& r, const Fp2& a, const Fp2& b) {
  fsub_k<16>(r, a, b);
  fnorm(r, r);
}
BLS_HD void fneg(Fp2& r, const Fp2& a) {
  fneg_k<16>(r, a);
  fnorm(r, r);
}
BLS_HD void fdbl(Fp2& r, const Fp2& a) {
  fp_add(r.c0, a.c0, a.c0);
  fp_add(r.c1, a.c1, a.c1);
}
template <int K>
BLS_HD void fp2_conj_k(Fp2& r, const Fp2& a) {
  r.c0 = a.c0;
  fp_neg_k<K>(r.c1, a.c1);
}
BLS_HD void fp2_conj(Fp2& r, const Fp2& a) {
  fp2_conj_k<16>(r, a);
  fp_norm(r.c1, r.c1);
}
BLS_HD bool fis_zero(const Fp2& a) { return fp_is_zero(a.c0) && fp_is_zero(a.c1); }
BLS_HD bool feq(const Fp2& a, const Fp2& b) { return fp_eq(a.c0, b.c0) && fp_eq(a.c1, b.c1); }
BLS_HD void fzero(Fp2& r) {
  fp_zero(r.c0);
  fp_zero(r.c1);
}
BLS_HD void fone(Fp2& r) {
  fp_one(r.c0);
  fp_zero(r.c1);
}
BLS_HD void fselect(Fp2& r, bool c, const Fp2& a, const Fp2& b) {
  fp_select(r.c0, c, a.c0, b.c0);
  fp_select(r.c1, c, a.c1, b.c1);
}

// Karatsuba: 3 Fp products (calls of the one resident fp_mul instance: the hot instruction footprint of every kernel is
// fp_mul + fp_sqr, ~12 KB, which is what keeps the SM's instruction cache from thrashing).
// Inputs: limbs <= 2^29+64 (a sum of two normalised values), value bounds with (vb(a0)+vb(a1)) * (vb(b0)+vb(b1)) <= 2000.
// Output: limbs <= 2^28+11, value bounds c0 <= 6, c1 <= 10.
BLS_FN void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b) {
  Fp t0, t1, sa, sb, t2;
  fp_add(sa, a.c0, a.c1);
  fp_add(sb, b.c0, b.c1);
  fp_norm(sa, sa);
  fp_norm(sb, sb);
  fp_mul(t0, a.c0, b.c0);
  fp_mul(t1, a.c1, b.c1);
  fp_mul(t2, sa, sb);
  fp_sub_k<4>(t2, t2, t0);
  fp_sub_k<4>(t2, t2, t1);
  fp_sub_k<4>(t0, t0, t1);
  fp_norm(r.c1, t2);
  fp_norm(r.c0, t0);
}
// (a0+a1)(a0-a1), 2 a0 a1: 2 Fp products.  Input value bounds <= 22 each; output c0 <= 2, c1 <= 4.
BLS_FN void fp2_sqr(Fp2& r, const Fp2& a) {
  Fp s, d, m;
  fp_add(s, a.c0, a.c1);
  fp_sub_k<32>(d, a.c0, a.c1);
  fp_norm(s, s);
  fp_norm(d, d);
  fp_mul(m, a.c0, a.c1);
  fp_mul(r.c0, s, d);
  fp_add(r.c1, m, m);
}
BLS_HD void fmul(Fp2& r, const Fp2& a, const Fp2& b) { fp2_mul(r, a, b); }
BLS_HD void fsqr(Fp2& r, const Fp2& a) { fp2_sqr(r, a); }

BLS_HD void fp2_mul_fp(Fp2& r, const Fp2& a, const Fp& k) {
  fp_mul(r.c0, a.c0, k);
  fp_mul(r.c1, a.c1, k);
}
// multiply by xi = 1 + u : (c0 - c1) + (c0 + c1) u ; K bounds the value of a.c1
template <int K>
BLS_HD void fp2_mul_xi_k(Fp2& r, const Fp2& a) {
  Fp t;
  fp_sub_k<K>(t, a.c0, a.c1);
  fp_add(r.c1, a.c0, a.c1);
  r.c0 = t;
}
BLS_HD void fp2_mul_xi(Fp2& r, const Fp2& a) {
  fp2_mul_xi_k<16>(r, a);
  fnorm(r, r);
}
// norm a0^2 + a1^2
BLS_HD void fp2_norm(Fp& r, const Fp2& a) {
  Fp t0, t1;
  fp_sqr(t0, a.c0);
  fp_sqr(t1, a.c1);
  fp_add(r, t0, t1);
}
BLS_HD void fp2_inv(Fp2& r, const Fp2& a) {
  Fp n, ni;
  fp2_norm(n, a);
  fp_inv(ni, n);
  fp_mul(r.c0, a.c0, ni);
  fp_mul(n, a.c1, ni);
  fp_neg_k<4>(r.c1, n);
  fp_norm(r.c1, r.c1);
}
BLS_HD void finv(Fp2& r, const Fp2& a) { fp2_inv(r, a); }

// sgn0 (RFC 9380 4.1) on Montgomery inputs
BLS_HD uint32_t fp_sgn0(const Fp& a) {
  Fp r;
  fp_from_mont(r, a);
  return r.l[0] & 1u;
}
BLS_HD bool raw_is_zero(const Fp& raw) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) t |= raw.l[i];
  return t == 0;
}
BLS_HD uint32_t fp2_sgn0(const Fp2& a) {
  Fp r0, r1;
  fp_from_mont(r0, a.c0);
  fp_from_mont(r1, a.c1);
  uint32_t s0 = r0.l[0] & 1u, z0 = raw_is_zero(r0) ? 1u : 0u;
  return s0 | (z0 & (r1.l[0] & 1u));
}
// "lexicographically largest" flag of the compressed encodings (c1 first, then c0)
BLS_HD bool fp_lex_largest(const Fp& a) {
  Fp r;
  fp_from_mont(r, a);
  return raw_gt_half(r);
}
BLS_HD bool fp2_lex_largest(const Fp2& a) {
  Fp r0, r1;
  fp_from_mont(r1, a.c1);
  if (!raw_is_zero(r1)) return raw_gt_half(r1);
  fp_from_mont(r0, a.c0);
  return raw_gt_half(r0);
}

// Square root of the RATIO num/den in Fp2 with two Fp exponentiations and no inversion (complex method with the
// inverse-square-root trick).  Returns true and r = sqrt(num/den) if the ratio is a square; otherwise returns false and
// r = sqrt(Z * num/den) with Z = -(2+u) (the SSWU constant of the G2 suite; norm(Z) = 5, sqrt(-5) = K_SQRT_M5).
// den must be non-zero.  Used by G2 decompression (den = 1) and by the SSWU map (RFC 9380 6.6.2 sqrt_ratio contract).
BLS_FN bool fp2_sqrt_ratio(Fp2& r, const Fp2& num, const Fp2& den) {
  Fp2 M, dc;
  Fp d, nm, t1, n, chk, one, delta, e, t, half, x0, x1, tmp;
  fp_one(one);
  fp_set(half, K_HALF);
  fp2_conj(dc, den);
  fp2_mul(M, num, dc);  // ratio = M / d
  fp2_norm(d, den);
  if (fis_zero(M)) {
    fzero(r);
    return true;
  }
  fp2_norm(nm, M);
  fp_isqrt_pow(t1, nm);
  fp_mul(n, t1, nm);  // candidate sqrt(nm)
  fp_sqr(chk, n);
  bool is_sq = fp_eq(chk, nm);
  if (!is_sq) {
    // ratio is a non-square: switch to Z*ratio.  M' = Z*M, norm(M') = 5*nm, sqrt(5 nm) = sqrt(-5) * (t1*nm)
    // because (t1*nm)^2 = -nm when nm is a non-residue.
    Fp2 Z;
    fp2_set(Z, K_SSWU2_Z);
    fp2_mul(M, M, Z);
    Fp s5;
    fp_set(s5, K_SQRT_M5);
    fp_mul(n, n, s5);
  }
  // delta' = (M0 + n)/2 ; if zero take (M0 - n)/2
  fp_add(delta, M.c0, n);
  if (fp_is_zero(delta)) fp_sub(delta, M.c0, n);
  fp_mul(delta, delta, half);
  fp_mul(e, delta, d);
  fp_isqrt_pow(t, e);  // 1/sqrt(delta' d) when that is a residue
  fp_sqr(chk, t);
  fp_mul(chk, chk, e);
  fp_mul(tmp, M.c1, t);
  fp_mul(tmp, tmp, half);  // M1 t / 2
  if (fp_eq(chk, one)) {
    fp_mul(x0, delta, t);
    x1 = tmp;
  } else {
    x0 = tmp;
    fp_mul(x1, delta, t);
    fp_neg_k<4>(x1, x1);
    fp_norm(x1, x1);
  }
  r.c0 = x0;
  r.c1 = x1;
  return is_sq;
}

}  // namespace bls
