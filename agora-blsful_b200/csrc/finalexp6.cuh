// Cooperative final exponentiation: one Fp12 value spread over SIX lanes (lane k owns the coefficient of w^k, as in the
// cooperative Miller loop of miller6.cuh), every Fp12 product / squaring one call of the fused sum-of-products unit per lane.
//
// Why: the final exponentiation is ~8,400 Fp multiplications that depend on one another.  On ONE thread (round 1:
// k_probe_fin<<<1,32>>>) that is 14 ms of pure latency - nothing at 1M signatures, but the whole cost of a small batch, of
// every bisection probe and of a 1/8 slice of a batch on 8 GPUs.  Six lanes of one warp execute the six coefficients of a
// product in the time of one, so the chain shortens ~6x (the Fp12 inversion of the easy part stays on one lane: ~500
// multiplications).  Replaces `.final_exponentiation()` + `Gt::is_identity` (reference src/helpers.rs:50,62,
// src/traits/sig_core.rs:138-145) for the batch checks; fp12.cuh's one-thread version stays for per-item kernels and as
// the cross-check in the host-emulation tests.
//
// Exponent: 3 (p^12 - 1)/r as in fp12.cuh (the cube of the canonical pairing value: "== 1" is unchanged because 3 does
// not divide r).  easy: f^((p^6-1)(p^2+1));  hard: (x-1)^2 (x+p)(x^2+p^2-1) + 3.
#pragma once
#include "miller6.cuh"

namespace bls {

// registers of the exponentiation program: NREG Fp12 values = NREG x 6 expanded records per group (shared memory)
constexpr int FE6_NREG = 6;
struct Fe6 {
  SAccRec* R;      // R[reg * 6 + k]
  Fp12* scratch;   // one Fp12 in unsigned form (global or shared): the inversion's way in and out
  int k;           // this lane's coefficient (device); the host emulation loops over 0..5
  unsigned mask;   // the lanes of the group (device barriers)
};
BLS_HD SAccRec* fe6_reg(const Fe6& c, int reg) { return c.R + 6 * reg; }

#if defined(__CUDA_ARCH__)
#define FE6_SYNC(c) __syncwarp((c).mask)
#define FE6_EACH_LANE(c, k) for (int k = (c).k, once_ = 1; once_; once_ = 0)
#else
#define FE6_SYNC(c) \
  do {              \
  } while (0)
#define FE6_EACH_LANE(c, k) for (int k = 0; k < 6; k++)
#endif

// coefficient k of X * Y (general product): sum_j X[(k - j) mod 6] * Y[j] * (xi if j > k).  dst must not alias X or Y.
BLS_FN void fe6_mul_lane(SAccRec* dst, const SAccRec* X, const SAccRec* Y, int k) {
  uint64_t T[2 * NL];
  SopKeep keep;
  int32_t res[2 * NL], x[NL], y[NL];
  SopI4 ra[4], rb[4];
#if defined(BLS_TRACK)
  double col = 0, vsum = 0;
  for (int j = 0; j < 6; j++) {
    const int xi = j > k;
    const SAccRec& A = X[xi ? k - j + 6 : k - j];
    const double la = A.lb * (xi ? 2.0 : 1.0);
    BLS_REQ(la < 1073741824.0 && Y[j].lb < 1073741824.0, "fe6_mul operand");
    col += 14.0 * 2.0 * la * Y[j].lb;
    vsum += 2.0 * A.vb * (xi ? 2.0 : 1.0) * Y[j].vb;
  }
  {
    const double lim = 9223372036854775808.0 - 15.0 * 72057594037927936.0;
    BLS_REQ(col < lim, "fe6_mul column overflow");
    BLS_REQ(vsum / 2500.0 + 1.0 < 16.0, "fe6_mul result value bound");
  }
#endif
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
#pragma unroll 1
  for (int pass = 0; pass < 3; pass++) {
#pragma unroll 1
    for (int j = 0; j < 6; j++) {
      const int xi = j > k;
      const int ia = xi ? k - j + 6 : k - j;
      const int run = xi ? (pass == 0 ? 3 : pass == 1 ? 2 : 0) : pass;  // xi a: halves a0 - a1, a0 + a1, their sum 2 a0
      sopw_fetch(ra, X[ia].w + 16 * run);
      sopw_fetch(rb, Y[j].w + 16 * pass);
      sopw_take(x, ra, xi & (pass == 2));
      sopw_take(y, rb, 0);
      sop_acc(T, x, y);
    }
    if (pass == 0) {
#pragma unroll
      for (int i = 0; i < NL; i++) {
        sop_keep_st(keep, i, T[2 * i], T[2 * i + 1]);
        T[2 * i] = T[2 * i + 1] = 0;
      }
    } else {
#pragma unroll
      for (int i = 0; i < NL; i++) {
        uint64_t u0, u1;
        sop_keep_ld(keep, i, u0, u1);
        const uint64_t v0 = T[2 * i], v1 = T[2 * i + 1];
        if (pass == 1) sop_keep_st(keep, i, u0 + v0, u1 + v1);
        T[2 * i] = pass == 1 ? u0 - v0 : v0 - u0;
        T[2 * i + 1] = pass == 1 ? u1 - v1 : v1 - u1;
      }
      T[2 * NL - 1] = 0;
      int32_t c[NL];
      sop_redc(c, T);
#pragma unroll
      for (int i = 0; i < NL; i++) {
        if (pass == 1) res[i] = c[i]; else res[NL + i] = c[i];
      }
#pragma unroll
      for (int i = 0; i < 2 * NL; i++) T[i] = 0;
    }
  }
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int32_t a0 = i < NL ? res[i] : 0, a1 = i < NL ? res[NL + i] : 0;
    dst->w[i] = a0;
    dst->w[16 + i] = a1;
    dst->w[32 + i] = a0 + a1;
    dst->w[48 + i] = a0 - a1;
  }
#if defined(BLS_TRACK)
  STRK(*dst, vsum / 2500.0 + 1.0, 134217728.0);
#endif
}

// lane-local pieces -------------------------------------------------------------------------------------------------------
BLS_HD void sacc_neg(SAccRec& r, const SAccRec& a) {
#pragma unroll
  for (int i = 0; i < 64; i++) r.w[i] = -a.w[i];
  STRK(r, a.vb, a.lb);
}
// complex conjugate of the Fp2 coefficient in expanded form: (a0, -a1, a0 - a1, a0 + a1)
BLS_HD void sacc_conj2(SAccRec& r, const SAccRec& a) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int32_t a0 = a.w[i], a1 = a.w[16 + i], s = a.w[32 + i], d = a.w[48 + i];
    r.w[i] = a0;
    r.w[16 + i] = -a1;
    r.w[32 + i] = d;
    r.w[48 + i] = s;
  }
  STRK(r, a.vb, a.lb);
}
// dst = A * B for two single records (one Karatsuba product, two reductions): the Frobenius coefficient multiplications
BLS_FN void sacc_mul1(SAccRec* dst, const SAccRec* A, const SAccRec* B) {
  uint64_t P0[2 * NL], P1[2 * NL], P2[2 * NL];
  int32_t x[NL], y[NL], c0[NL], c1[NL];
  SopI4 ra[4], rb[4];
#if defined(BLS_TRACK)
  BLS_REQ(A->lb < 1073741824.0 && B->lb < 1073741824.0, "sacc_mul1 operand");
  BLS_REQ(14.0 * 2.0 * A->lb * B->lb < 9223372036854775808.0 - 15.0 * 72057594037927936.0, "sacc_mul1 column overflow");
#endif
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) P0[i] = P1[i] = P2[i] = 0;
  sopw_fetch(ra, A->w); sopw_fetch(rb, B->w); sopw_take(x, ra, 0); sopw_take(y, rb, 0); sop_acc(P0, x, y);
  sopw_fetch(ra, A->w + 16); sopw_fetch(rb, B->w + 16); sopw_take(x, ra, 0); sopw_take(y, rb, 0); sop_acc(P1, x, y);
  sopw_fetch(ra, A->w + 32); sopw_fetch(rb, B->w + 32); sopw_take(x, ra, 0); sopw_take(y, rb, 0); sop_acc(P2, x, y);
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) {
    const uint64_t s = P0[i] + P1[i];
    P0[i] = P0[i] - P1[i];
    P2[i] = P2[i] - s;
  }
  P0[2 * NL - 1] = 0;
  P2[2 * NL - 1] = 0;
  sop_redc(c0, P0);
  sop_redc(c1, P2);
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int32_t a0 = i < NL ? c0[i] : 0, a1 = i < NL ? c1[i] : 0;
    dst->w[i] = a0;
    dst->w[16 + i] = a1;
    dst->w[32 + i] = a0 + a1;
    dst->w[48 + i] = a0 - a1;
  }
#if defined(BLS_TRACK)
  STRK(*dst, 2.0 * A->vb * B->vb / 2500.0 + 1.0, 134217728.0);
#endif
}
// unsigned tower form <-> lane records
BLS_HD void fe6_load_coeff(SAccRec& r, const Fp2& c) {
  SFp2 s;
  sfp2_from_fp2(s, c);
  sfp2_lin(s, &s, 1, 0, nullptr, 0, nullptr, 0);  // balanced limbs ([-2^27, 2^27)): what the six-term column bound assumes
  sacc_from_sfp2(r, s);
}
BLS_HD void fe6_store_coeff(Fp2& out, const SAccRec& a) {
  SFp2 s;
  sfp2_from_sacc(s, a);
  Fp2 u;
  fp2_from_sfp2(u, s);
  fred(out, u);
}

// whole-value operations: every lane of the group calls them with the same arguments ---------------------------------------
BLS_HD void fe6_mul(const Fe6& c, int d, int a, int b) {
  FE6_EACH_LANE(c, k) fe6_mul_lane(fe6_reg(c, d) + k, fe6_reg(c, a), fe6_reg(c, b), k);
  FE6_SYNC(c);
}
BLS_HD void fe6_sqr(const Fe6& c, int d, int a) {  // d != a
  FE6_EACH_LANE(c, k) sopw(fe6_reg(c, d) + k, K_M6_SQR[k], 4, fe6_reg(c, a), nullptr, k, c.mask);
  FE6_SYNC(c);
}
BLS_HD void fe6_copy(const Fe6& c, int d, int a) {
  FE6_EACH_LANE(c, k) fe6_reg(c, d)[k] = fe6_reg(c, a)[k];
  FE6_SYNC(c);
}
// a^(p^6): w -> -w   (the inverse, for values in the cyclotomic subgroup).  d may equal a.
BLS_HD void fe6_conj(const Fe6& c, int d, int a) {
  FE6_EACH_LANE(c, k) {
    if (k & 1) sacc_neg(fe6_reg(c, d)[k], fe6_reg(c, a)[k]); else fe6_reg(c, d)[k] = fe6_reg(c, a)[k];
  }
  FE6_SYNC(c);
}
// a^p: coefficient k -> conj(c_k) * K_FROB1[k];   a^(p^2): c_k * K_FROB2[k]   (d != a)
BLS_HD void fe6_frob(const Fe6& c, int d, int a, int power) {
  FE6_EACH_LANE(c, k) {
    SAccRec t, g;
    if (power == 1) {
      sacc_conj2(t, fe6_reg(c, a)[k]);
      Fp2 cg;
      fp2_set(cg, K_FROB1[k]);
      fe6_load_coeff(g, cg);
    } else {
      t = fe6_reg(c, a)[k];
      Fp2 cg;
      fp_set(cg.c0, K_FROB2[k]);
      fzero(cg.c1);
      fe6_load_coeff(g, cg);
    }
    if (k == 0) fe6_reg(c, d)[k] = t; else sacc_mul1(fe6_reg(c, d) + k, &t, &g);  // K_FROB*[0] = 1
  }
  FE6_SYNC(c);
}
// d = 1 / a through the tower code of fp12.cuh on lane 0 (the one serial piece: ~500 multiplications)
BLS_HD void fe6_inv(const Fe6& c, int d, int a) {
  FE6_EACH_LANE(c, k) fe6_store_coeff(*fp12_coeff(*c.scratch, k), fe6_reg(c, a)[k]);
  FE6_SYNC(c);
#if defined(__CUDA_ARCH__)
  if (c.k == 0)
#endif
  {
    Fp12 t = *c.scratch, r;
    fp12_inv(r, t);
    *c.scratch = r;
  }
  FE6_SYNC(c);
  FE6_EACH_LANE(c, k) fe6_load_coeff(fe6_reg(c, d)[k], *fp12_coeff(*c.scratch, k));
  FE6_SYNC(c);
}
// d = a^x (x = -|x|) for a in the cyclotomic subgroup; uses registers t0, t1 (all four distinct)
BLS_HD void fe6_pow_x(const Fe6& c, int d, int a, int t0, int t1) {
  fe6_copy(c, t0, a);
  int cur = t0, oth = t1;
  const uint64_t e = K_X_ABS;
#pragma unroll 1
  for (int i = 62; i >= 0; i--) {
    fe6_sqr(c, oth, cur);
    int s = cur; cur = oth; oth = s;
    if ((e >> i) & 1) {
      fe6_mul(c, oth, cur, a);
      s = cur; cur = oth; oth = s;
    }
  }
  fe6_conj(c, d, cur);
}
// register 0 holds f on entry; on exit register 0 holds f^(3 (p^12-1)/r).  Registers 1..5 are scratch.
BLS_HD void fe6_final_exponentiation(const Fe6& c) {
  // easy part: m = f^((p^6-1)(p^2+1))
  fe6_inv(c, 1, 0);
  fe6_conj(c, 2, 0);
  fe6_mul(c, 3, 2, 1);      // f^(p^6-1)
  fe6_frob(c, 1, 3, 2);
  fe6_mul(c, 0, 1, 3);      // m                                   (reg 0)
  // a = m^((x-1)^2)
  fe6_pow_x(c, 1, 0, 2, 3); // m^x
  fe6_conj(c, 2, 0);
  fe6_mul(c, 3, 1, 2);      // m^(x-1)                             (reg 3)
  fe6_pow_x(c, 1, 3, 2, 4); // (m^(x-1))^x
  fe6_conj(c, 2, 3);
  fe6_mul(c, 4, 1, 2);      // a                                   (reg 4)
  // b = a^(x+p)
  fe6_pow_x(c, 1, 4, 2, 3);
  fe6_frob(c, 2, 4, 1);
  fe6_mul(c, 5, 1, 2);      // b                                   (reg 5)
  // c = b^(x^2+p^2-1)
  fe6_pow_x(c, 1, 5, 2, 3);
  fe6_pow_x(c, 4, 1, 2, 3); // b^(x^2)
  fe6_frob(c, 1, 5, 2);
  fe6_mul(c, 2, 4, 1);
  fe6_conj(c, 1, 5);
  fe6_mul(c, 3, 2, 1);      // c                                   (reg 3)
  // result = c * m^3
  fe6_sqr(c, 1, 0);
  fe6_mul(c, 2, 1, 0);
  fe6_mul(c, 0, 3, 2);
}
// lane k's share of "register 0 == 1 ?"
BLS_HD bool fe6_lane_is_one(const Fe6& c, int k) {
  Fp2 v, one;
  fe6_store_coeff(v, fe6_reg(c, 0)[k]);
  if (k == 0) {
    fone(one);
    return feq(v, one);
  }
  return fis_zero(v);
}

}  // namespace bls
