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

// ---- fused Fp2 products --------------------------------------------------------------------------------------------------
// Round 1 built an Fp2 product from three calls of fp_mul plus five additive steps, every operand and every partial result
// travelling through local memory (64 local loads + 32 local stores of 16 bytes per product: k_clear_cofactor moved 424 KB
// of DRAM traffic per signature).  Now the three integer products (Karatsuba over Fp2, and one Karatsuba level over the limbs
// inside each: 3 x 147 multiplies) accumulate in 64-bit columns, and each output half gets ONE Montgomery reduction
// (2 x 210): 861 multiply-accumulates instead of 1,071, 16 local loads + 8 stores, no glue.
//   c1 = a0 b1 + a1 b0 = (a0 + a1)(b0 + b1) - a0 b0 - a1 b1          (a non-negative integer)
//   c0 = a0 b0 - a1 b1 + R p                                          (R p = p in columns 14..27: keeps the integer >= 0)
// Columns may wrap modulo 2^64 while they are being combined; the TRUE column values fit a signed 64-bit word, which is what
// the BLS_TRACK build checks for every call site.
// MEASURED (B200, 1M signatures, profiles/fp2_fused_r2.txt): correct and bound-checked, but SLOWER than three fp_mul calls in
// every register configuration tried - k_clear_cofactor 159 ms unfused vs 168 (246 registers, 2 blocks/SM), 174 (168 registers),
// 169 (128 registers); k_hash 93 vs 99.  The fused body needs ~190 live registers (27 + 27 columns, operands, limb-Karatsuba
// temporaries): under the kernels' caps it spills, and with the cap lifted it loses the warps the multiplier needs.  Off by default.
#if !defined(FP2_FUSED)
#define FP2_FUSED 0
#endif

// t[i + j] += a[i] * b[j], i, j < 14 (unsigned limbs < 2^31; one Karatsuba level over the limbs; columns modulo 2^64)
BLS_HD void fp_cols_acc(uint64_t* t, const uint32_t* a, const uint32_t* b) {
  constexpr int HL = NL / 2;
  uint64_t C[2 * HL - 1];
#pragma unroll
  for (int part = 0; part < 2; part++) {
#pragma unroll
    for (int i = 0; i < 2 * HL - 1; i++) C[i] = 0;
#pragma unroll
    for (int i = 0; i < HL; i++) {
#pragma unroll
      for (int j = 0; j < HL; j++) C[i + j] += (uint64_t)a[part * HL + i] * b[part * HL + j];
    }
#pragma unroll
    for (int i = 0; i < 2 * HL - 1; i++) {
      t[2 * part * HL + i] += C[i];
      t[HL + i] -= C[i];
    }
  }
  uint32_t sa[HL], sb[HL];
#pragma unroll
  for (int i = 0; i < HL; i++) {
    sa[i] = a[i] + a[HL + i];
    sb[i] = b[i] + b[HL + i];
  }
#pragma unroll
  for (int i = 0; i < HL; i++) {
#pragma unroll
    for (int j = 0; j < HL; j++) t[HL + i + j] += (uint64_t)sa[i] * sb[j];
  }
}
// Montgomery reduction of 27 SIGNED columns whose total is a non-negative integer below R * 2^388 -> 14 limbs < 2^28 (+ top)
BLS_HD void fp_cols_redc(uint32_t* rl, uint64_t* t) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const uint32_t m = opaque32(((uint32_t)t[i] * K_PINV28) & M28);
#pragma unroll
    for (int j = 0; j < NL; j++) t[i + j] += (uint64_t)m * p28(j);
    t[i + 1] += (uint64_t)((int64_t)t[i] >> 28);  // exact (the low 28 bits are zero now); arithmetic: a column may be negative
  }
  int64_t c = 0;
#pragma unroll
  for (int j = 0; j < NL - 1; j++) {
    c += (int64_t)t[NL + j];
    rl[j] = (uint32_t)c & M28;
    c >>= 28;
  }
  rl[NL - 1] = (uint32_t)(c + (int64_t)t[2 * NL - 1]);
}
// 28 x u64 parked in local memory with 128-bit accesses (kept out of the register file while the next product runs)
struct alignas(16) Fp2Keep {
  uint64_t v[2 * NL];
};
BLS_HD void fp2_keep_st(Fp2Keep& k, int i, uint64_t a, uint64_t b) {
#if defined(__CUDA_ARCH__)
  asm volatile("st.local.v2.u64 [%0], {%1, %2};" ::"l"(__cvta_generic_to_local(&k.v[2 * i])), "l"(a), "l"(b) : "memory");
#else
  k.v[2 * i] = a;
  k.v[2 * i + 1] = b;
#endif
}
BLS_HD void fp2_keep_ld(const Fp2Keep& k, int i, uint64_t& a, uint64_t& b) {
#if defined(__CUDA_ARCH__)
  asm volatile("ld.local.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(__cvta_generic_to_local(&k.v[2 * i])) : "memory");
#else
  a = k.v[2 * i];
  b = k.v[2 * i + 1];
#endif
}

#if FP2_FUSED
// Inputs: limbs <= 2^29+64, value bounds with (vb(a0)+vb(a1)) * (vb(b0)+vb(b1)) <= 2000.  Output: limbs < 2^28, values <= 3.
BLS_FN void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b) {
#if defined(BLS_TRACK)
  {
    const double la = (double)(a.c0.lb > a.c1.lb ? a.c0.lb : a.c1.lb), lb = (double)(b.c0.lb > b.c1.lb ? b.c0.lb : b.c1.lb);
    BLS_REQ(2.0 * 14.0 * la * lb + 15.0 * 72057594037927936.0 < 9223372036854775808.0, "fp2_mul column overflow");
    BLS_REQ(la < 1073741824.0 + 256.0 && lb < 1073741824.0 + 256.0, "fp2_mul operand half sums");
    BLS_REQ((a.c0.vb + a.c1.vb) * (b.c0.vb + b.c1.vb) <= 2000.0, "fp2_mul value bound");
  }
#endif
  uint64_t T[2 * NL];
  Fp2Keep keep;
  uint32_t x[NL], y[NL], rl[NL];
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
  fp_ld(x, a.c0);
  fp_ld(y, b.c0);
  fp_cols_acc(T, x, y);  // P0
#pragma unroll
  for (int i = 0; i < NL; i++) {
    fp2_keep_st(keep, i, T[2 * i], T[2 * i + 1]);
    T[2 * i] = T[2 * i + 1] = 0;
  }
  fp_ld(x, a.c1);
  fp_ld(y, b.c1);
  fp_cols_acc(T, x, y);  // P1
#pragma unroll
  for (int i = 0; i < NL; i++) {
    uint64_t u0, u1;
    fp2_keep_ld(keep, i, u0, u1);
    const uint64_t v0 = T[2 * i], v1 = T[2 * i + 1];
    fp2_keep_st(keep, i, u0 + v0, u1 + v1);  // P0 + P1
    T[2 * i] = u0 - v0;                      // P0 - P1
    T[2 * i + 1] = u1 - v1;
  }
#pragma unroll
  for (int j = 0; j < NL; j++) T[NL + j] += p28(j);  // + R p
  T[2 * NL - 1] = p28(NL - 1);                       // (column 27 carries nothing else)
  fp_cols_redc(rl, T);
  // the operand sums of the third product while c0 is still in registers: r may alias a or b
  {
    uint32_t x1[NL], y1[NL];
    fp_ld(x, a.c0);
    fp_ld(x1, a.c1);
    fp_ld(y, b.c0);
    fp_ld(y1, b.c1);
#pragma unroll
    for (int i = 0; i < NL; i++) {
      x[i] += x1[i];
      y[i] += y1[i];
    }
  }
  fp_st(r.c0, rl);
  TRK(r.c0, 3.0, M28);
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
  fp_cols_acc(T, x, y);  // P2
#pragma unroll
  for (int i = 0; i < NL; i++) {
    uint64_t u0, u1;
    fp2_keep_ld(keep, i, u0, u1);
    T[2 * i] -= u0;
    T[2 * i + 1] -= u1;
  }
  T[2 * NL - 1] = 0;
  fp_cols_redc(rl, T);
  fp_st(r.c1, rl);
  TRK(r.c1, 2.0, M28);
}
// c0 = (a0 + a1)(a0 - a1 + 32 p), c1 = a0 * 2 a1: two products, two reductions.  Input value bounds <= 22 each, limbs <= 2^29+64;
// output limbs < 2^28, values <= 2.
BLS_FN void fp2_sqr(Fp2& r, const Fp2& a) {
  Fp s, d;
  fp_add(s, a.c0, a.c1);
  fp_sub_k<32>(d, a.c0, a.c1);
  fp_norm(s, s);
  fp_norm(d, d);
#if defined(BLS_TRACK)
  BLS_REQ(14.0 * (double)s.lb * (double)d.lb + 15.0 * 72057594037927936.0 < 9223372036854775808.0, "fp2_sqr column overflow");
  BLS_REQ(14.0 * 2.0 * (double)a.c0.lb * (double)a.c1.lb + 15.0 * 72057594037927936.0 < 9223372036854775808.0, "fp2_sqr column overflow");
  BLS_REQ(s.vb * d.vb <= 2000.0 && 2.0 * a.c0.vb * a.c1.vb <= 2000.0, "fp2_sqr value bound");
#endif
  uint64_t T[2 * NL];
  uint32_t x[NL], y[NL], r0[NL], r1[NL];
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
  fp_ld(x, a.c0);
  fp_ld(y, a.c1);
#pragma unroll
  for (int i = 0; i < NL; i++) y[i] <<= 1;
  fp_cols_acc(T, x, y);
  T[2 * NL - 1] = 0;
  fp_cols_redc(r1, T);
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
  fp_ld(x, s);
  fp_ld(y, d);
  fp_cols_acc(T, x, y);
  T[2 * NL - 1] = 0;
  fp_cols_redc(r0, T);
  fp_st(r.c0, r0);
  fp_st(r.c1, r1);
  TRK(r.c0, 2.0, M28);
  TRK(r.c1, 2.0, M28);
}
#else
// Karatsuba: 3 Fp products (calls of the one resident fp_mul instance).
// Inputs: limbs <= 2^29+64 (a sum of two normalised values), value bounds with (vb(a0)+vb(a1)) * (vb(b0)+vb(b1)) <= 2000.
// Output: limbs <= 2^28+11, value bounds c0 <= 6, c1 <= 10.
BLS_FN void fp2_mul(Fp2& r, const Fp2& a, const Fp2& b) {
  Fp sa, sb, t2;  // three stack slots (they are the hot working set of every G2 kernel): the sums are dead after the first product
  fp_add(sa, a.c0, a.c1);
  fp_add(sb, b.c0, b.c1);
  fp_norm(sa, sa);
  fp_norm(sb, sb);
  fp_mul(t2, sa, sb);
  fp_mul(sa, a.c0, b.c0);
  fp_mul(sb, a.c1, b.c1);
  fp_sub_k<4>(t2, t2, sa);
  fp_sub_k<4>(t2, t2, sb);
  fp_sub_k<4>(sa, sa, sb);
  fp_norm(r.c1, t2);
  fp_norm(r.c0, sa);
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
#endif
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
