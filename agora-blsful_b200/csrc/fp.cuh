// Fp: the BLS12-381 base field in a carry-free redundant radix: 14 limbs of 28 bits in 32-bit words,
// Montgomery form with R = 2^392.
//
// Replaces (for public-input work) the Fp arithmetic the reference obtains from blstrs_plus/blst
// (reference src/impls.rs:185-215 re-exports the types; there is no arithmetic in the reference tree).
//
// Why this shape on sm_100a: every 28x28-bit product is < 2^56, so a whole column of a 14-limb product (and of the
// Montgomery m*p correction) accumulates in one 64-bit register pair with plain IMAD.WIDE.U32 - no carry flag, no
// predicate chains.  (A 12x32-bit CIOS built on mad.lo.cc/madc.hi.cc was tried first: ptxas interleaves the independent
// carry chains until it runs out of predicate registers and then spills carries through P2R/LOP3 - ~1000 LOP3 per 828
// IMAD.WIDE in fp2_mul)  The 4 spare bits per word make additions and subtractions 14
// independent IADD3 with no reduction at all ("lazy"); values are only brought back below 2p by the next multiplication.
//
// Bound discipline (checked mechanically by the host test harness when BLS_TRACK is defined - test infrastructure only):
//   vb = value bound in units of p, lb = bound on limbs 0..12.
//   fp_mul/fp_sqr : need 14*lb_a*lb_b + 14*2^56 < 2^63 and vb_a*vb_b <= 2000; give lb < 2^28, vb <= 2
//   fp_add        : limbwise, lb and vb add up
//   fp_sub<K>     : a + spread(K*p) - b, needs vb_b < K and lb_b <= 2^30-4; vb_out = vb_a + K
//   fp_norm       : one parallel carry pass, lb -> 2^28 + (lb >> 28)
//   fp_canon      : the unique representative in [0,p) with 28-bit limbs (for ==, is_zero, sgn0, serialisation)
#pragma once
#include <stdint.h>

#if defined(__CUDACC__)
#define BLS_HD __device__ __forceinline__
#define BLS_FN __device__ __noinline__
#define BLS_CONST __device__ __constant__ const
#else
#define BLS_HD static inline
#define BLS_FN static
#define BLS_CONST static const
#endif

#include "consts_gen.h"

#if defined(BLS_TRACK)
#include <cstdio>
#include <cstdlib>
#include <execinfo.h>
#define BLS_REQ(cond, what)                                                                  \
  do {                                                                                       \
    if (!(cond)) {                                                                           \
      fprintf(stderr, "BLS_TRACK bound violation: %s (%s:%d)\n", what, __FILE__, __LINE__);  \
      void* bt_[24];                                                                         \
      int n_ = backtrace(bt_, 24);                                                           \
      backtrace_symbols_fd(bt_, n_, 2);                                                      \
      abort();                                                                               \
    }                                                                                        \
  } while (0)
#endif

namespace bls {

constexpr int NL = 14;
constexpr uint32_t M28 = 0x0fffffffu;

// 14 limbs + 2 zero pad words: 64 bytes, 16-byte aligned, so that every load/store of a field element in local or global
// memory is four 128-bit transactions.
struct alignas(16) Fp {
  uint32_t l[NL + 2];
#if defined(BLS_TRACK)
  double vb;
  uint64_t lb;
#endif
};

#if defined(BLS_TRACK)
#define TRK(r, v, b) \
  do {               \
    (r).vb = (v);    \
    (r).lb = (b);    \
  } while (0)
#else
#define TRK(r, v, b) \
  do {               \
  } while (0)
#endif

// p in radix 2^28
BLS_HD constexpr uint32_t p28(int i) {
  constexpr uint32_t t[NL] = {K_P28_LIMBS};
  return t[i];
}
// spread(K*p): K*p written with every limb 0..12 raised by 2^30 (borrowed from the limb above) so that a limbwise
// "a + spread - b" never goes negative for b with limbs <= 2^30-4 and value < K*p
template <int K>
BLS_HD constexpr uint32_t kps(int i) {
  static_assert(K == 4 || K == 8 || K == 16 || K == 32 || K == 64 || K == 128 || K == 256 || K == 512, "no spread constant");
  constexpr uint32_t t4[NL] = {K_KPS_4}, t8[NL] = {K_KPS_8}, t16[NL] = {K_KPS_16}, t32[NL] = {K_KPS_32}, t64[NL] = {K_KPS_64},
                     t128[NL] = {K_KPS_128}, t256[NL] = {K_KPS_256}, t512[NL] = {K_KPS_512};
  return K == 4 ? t4[i] : K == 8 ? t8[i] : K == 16 ? t16[i] : K == 32 ? t32[i] : K == 64 ? t64[i] : K == 128 ? t128[i]
       : K == 256 ? t256[i] : t512[i];
}

BLS_HD void fp_set(Fp& r, const uint32_t* c) {
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = c[i];
  r.l[NL] = r.l[NL + 1] = 0;
  TRK(r, 1.0, M28);
}
BLS_HD void fp_zero(Fp& r) {
#pragma unroll
  for (int i = 0; i < NL + 2; i++) r.l[i] = 0;
  TRK(r, 0.0, 0);
}
BLS_HD void fp_one(Fp& r) { fp_set(r, K_ONE); }

// 128-bit moves of a whole element (the records are 64 bytes, 16-byte aligned): left to itself the compiler loads the
// operands of fp_mul with 28 scalar loads and stores the result with 8 64-bit stores
struct alignas(16) __attribute__((may_alias)) FpQuad {  // may_alias: it overlays the uint32_t limbs
  uint32_t x, y, z, w;
};
BLS_HD void fp_ld(uint32_t* v, const Fp& a) {
  const FpQuad* q = reinterpret_cast<const FpQuad*>(a.l);
  const FpQuad q0 = q[0], q1 = q[1], q2 = q[2], q3 = q[3];
  v[0] = q0.x; v[1] = q0.y; v[2] = q0.z; v[3] = q0.w;
  v[4] = q1.x; v[5] = q1.y; v[6] = q1.z; v[7] = q1.w;
  v[8] = q2.x; v[9] = q2.y; v[10] = q2.z; v[11] = q2.w;
  v[12] = q3.x; v[13] = q3.y;
}
BLS_HD void fp_st(Fp& r, const uint32_t* v) {
  FpQuad* q = reinterpret_cast<FpQuad*>(r.l);
  FpQuad q0, q1, q2, q3;
  q0.x = v[0]; q0.y = v[1]; q0.z = v[2]; q0.w = v[3];
  q1.x = v[4]; q1.y = v[5]; q1.z = v[6]; q1.w = v[7];
  q2.x = v[8]; q2.y = v[9]; q2.z = v[10]; q2.w = v[11];
  q3.x = v[12]; q3.y = v[13]; q3.z = 0; q3.w = 0;
  q[0] = q0; q[1] = q1; q[2] = q2; q[3] = q3;
}

// ---- lazy additive operations --------------------------------------------------------------------------------------
// a + b issued on the ALU pipe.  ptxas is free to emit a 32-bit addition as IMAD.IADD on the multiplier pipe, and in these
// kernels it does so for the operand sums of the Karatsuba products - where the multiplier is the one pipe that has no time to
// spare.  VIADDMNMX (max(a + b, c)) exists only on the ALU pipe: with c = the smallest value it IS the sum.
// Measured at 1M (round 2): the Karatsuba operand sums this way: 776 -> 768 ms over the step.  NOT for everything: the same
// trick on fp_add cost 3.6 ms and on fp_norm 7 ms (they break ptxas's three-input IADD3 fusions), and on the operand preparation of the
// lines kernel 13 ms (that kernel is bound by issue slots and the ALU pipe, not by the multiplier alone).
#if !defined(BLS_ALU_ADD)
#define BLS_ALU_ADD 1
#endif
BLS_HD uint32_t alu_add_u32(uint32_t a, uint32_t b) {
#if defined(__CUDA_ARCH__) && BLS_ALU_ADD
  return __viaddmax_u32(a, b, 0u);
#else
  return a + b;
#endif
}
BLS_HD int32_t alu_add_s32(int32_t a, int32_t b) {
#if defined(__CUDA_ARCH__) && BLS_ALU_ADD
  return __viaddmax_s32(a, b, (int32_t)0x80000000);
#else
  return a + b;
#endif
}
BLS_HD void fp_add(Fp& r, const Fp& a, const Fp& b) {
#if defined(BLS_TRACK)
  BLS_REQ(a.lb + b.lb < (1ull << 32), "fp_add limb overflow");
  BLS_REQ(a.vb + b.vb < 2000.0, "fp_add value overflow");
  double v_ = a.vb + b.vb;
  uint64_t l_ = a.lb + b.lb;
#endif
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = a.l[i] + b.l[i];
  r.l[NL] = r.l[NL + 1] = 0;
  TRK(r, v_, l_);
}
BLS_HD void fp_dbl(Fp& r, const Fp& a) { fp_add(r, a, a); }

template <int K>
BLS_HD void fp_sub_k(Fp& r, const Fp& a, const Fp& b) {
#if defined(BLS_TRACK)
  BLS_REQ(b.vb < (double)K, "fp_sub: subtrahend value bound >= K");
  BLS_REQ(b.lb <= (1ull << 30) - 4, "fp_sub: subtrahend limbs too large");
  BLS_REQ(a.lb + (1ull << 30) + (1ull << 28) < (1ull << 32), "fp_sub limb overflow");
  BLS_REQ(a.vb + K < 2000.0, "fp_sub value overflow");
  double v_ = a.vb + K;
  uint64_t l_ = a.lb + (1ull << 30) + (1ull << 28);
#endif
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = a.l[i] + kps<K>(i) - b.l[i];
  r.l[NL] = r.l[NL + 1] = 0;
  TRK(r, v_, l_);
}
template <int K>
BLS_HD void fp_neg_k(Fp& r, const Fp& a) {
#if defined(BLS_TRACK)
  BLS_REQ(a.vb < (double)K, "fp_neg: value bound >= K");
  BLS_REQ(a.lb <= (1ull << 30) - 4, "fp_neg: limbs too large");
  double v_ = K;
#endif
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = kps<K>(i) - a.l[i];
  r.l[NL] = r.l[NL + 1] = 0;
  TRK(r, v_, (1ull << 30) + (1ull << 28));
}

// one parallel carry pass: limbs 0..12 end up <= 2^28 + (lb >> 28); the top limb absorbs the rest
BLS_HD void fp_norm(Fp& r, const Fp& a) {
  uint32_t c[NL];
#pragma unroll
  for (int i = 0; i < NL - 1; i++) c[i] = a.l[i] >> 28;
  uint32_t top = a.l[NL - 1] + c[NL - 2];
#pragma unroll
  for (int i = NL - 2; i >= 1; i--) r.l[i] = (a.l[i] & M28) + c[i - 1];
  r.l[0] = a.l[0] & M28;
  r.l[NL - 1] = top;
  r.l[NL] = r.l[NL + 1] = 0;
#if defined(BLS_TRACK)
  double v_ = a.vb;
  uint64_t l_ = M28 + (a.lb >> 28);
#endif
  TRK(r, v_, l_);
}

// default flavours: subtrahend value < 16, result normalised (the raw lazy forms are the _k templates)
BLS_HD void fp_sub(Fp& r, const Fp& a, const Fp& b) {
  fp_sub_k<16>(r, a, b);
  fp_norm(r, r);
}
BLS_HD void fp_neg(Fp& r, const Fp& a) {
  fp_neg_k<16>(r, a);
  fp_norm(r, r);
}

// ---- Montgomery multiplication ---------------------------------------------------------------------------------------
// Keeps a value in a 32-bit register across the optimiser: without it LLVM widens `m` to 64 bits and ptxas lowers each
// m*p_j into IMAD.WIDE plus a dead high-word IADD3 (one extra ALU instruction per multiply-accumulate).
BLS_HD uint32_t opaque32(uint32_t x) {
#if defined(__CUDA_ARCH__)
  asm("" : "+r"(x));
#endif
  return x;
}

// r = a*b/R mod p.  One level of Karatsuba on the 14 x 14 limb product (7 + 7 limbs): three 7 x 7 products instead of
// four, 147 multiply-accumulates instead of 196, then a 14-row Montgomery reduction (196 + 14): 357 IMAD in all where
// the operand-scanning form needs 406.  The kernels built on this function run at ~90% of the IMAD.WIDE issue rate and a
// fifth of the ALU rate, so trading 49 multiplies for ~110 additions is a net gain.  The middle product
// (a_lo + a_hi)(b_lo + b_hi) may wrap modulo 2^64; (M - L - H) is exact because the true cross sum fits (same column bound
// as before: 14 lb_a lb_b + 14 * 2^56 < 2^63).
BLS_HD void fp_mul_inl(Fp& r, const Fp& a, const Fp& b) {
#if defined(BLS_TRACK)
  BLS_REQ((double)a.lb * (double)b.lb * 14.0 + 14.0 * 72057594037927936.0 < 9.2e18, "fp_mul column overflow");
  BLS_REQ(a.vb * b.vb <= 2000.0, "fp_mul value bound");
  BLS_REQ(a.lb < (1ull << 31) && b.lb < (1ull << 31), "fp_mul operand half sums");
#endif
  constexpr int HL = NL / 2;
  uint64_t L[2 * HL - 1], H[2 * HL - 1], M[2 * HL - 1];
  uint32_t al[NL], bl[NL], sa[HL], sb[HL], rl[NL];
  fp_ld(al, a);
  fp_ld(bl, b);
#pragma unroll
  for (int i = 0; i < HL; i++) {
    sa[i] = alu_add_u32(al[i], al[HL + i]);
    sb[i] = alu_add_u32(bl[i], bl[HL + i]);
  }
#pragma unroll
  for (int i = 0; i < 2 * HL - 1; i++) L[i] = H[i] = M[i] = 0;
#pragma unroll
  for (int i = 0; i < HL; i++) {
#pragma unroll
    for (int j = 0; j < HL; j++) {
      L[i + j] += (uint64_t)al[i] * bl[j];
      H[i + j] += (uint64_t)al[HL + i] * bl[HL + j];
      M[i + j] += (uint64_t)sa[i] * sb[j];
    }
  }
  uint64_t t[2 * NL];
#pragma unroll
  for (int i = 0; i < 2 * HL - 1; i++) {
    t[i] = L[i];
    t[NL + i] = H[i];
  }
  t[2 * HL - 1] = 0;
  t[2 * NL - 1] = 0;
#pragma unroll
  for (int i = 0; i < 2 * HL - 1; i++) t[HL + i] += M[i] - L[i] - H[i];
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const uint32_t m = opaque32(((uint32_t)t[i] * K_PINV28) & M28);
#pragma unroll
    for (int j = 0; j < NL; j++) t[i + j] += (uint64_t)m * p28(j);
    t[i + 1] += t[i] >> 28;
  }
  uint64_t c = 0;
#pragma unroll
  for (int j = 0; j < NL - 1; j++) {
    c += t[NL + j];
    rl[j] = (uint32_t)c & M28;
    c >>= 28;
  }
  rl[NL - 1] = (uint32_t)(c + t[2 * NL - 1]);
  fp_st(r, rl);
  TRK(r, 2.0, M28);
}

// squaring: the same Karatsuba level with three 7-limb squarings (cross products once, with a doubled operand):
// 3 * 28 + 196 + 14 = 294 multiply-accumulates (plain squaring: 315, plain product: 406).  The square-root and inversion
// chains of the decode and hash kernels are ~80% squarings.
BLS_HD void fp_sqr_half(uint64_t* T, const uint32_t* a) {  // T[0..12] = (7 limbs)^2
  constexpr int HL = NL / 2;
  uint32_t a2[HL];
#pragma unroll
  for (int i = 0; i < HL; i++) a2[i] = a[i] << 1;
#pragma unroll
  for (int i = 0; i < 2 * HL - 1; i++) T[i] = 0;
#pragma unroll
  for (int i = 0; i < HL; i++) {
    T[2 * i] += (uint64_t)a[i] * a[i];
#pragma unroll
    for (int j = i + 1; j < HL; j++) T[i + j] += (uint64_t)a2[i] * a[j];
  }
}
BLS_HD void fp_sqr_inl(Fp& r, const Fp& a) {
#if defined(BLS_TRACK)
  BLS_REQ((double)a.lb * (double)a.lb * 2.0 * 14.0 + 14.0 * 72057594037927936.0 < 9.2e18, "fp_sqr column overflow");
  BLS_REQ(a.vb * a.vb <= 2000.0, "fp_sqr value bound");
  BLS_REQ(a.lb < (1ull << 30), "fp_sqr operand half sums");
#endif
  constexpr int HL = NL / 2;
  uint64_t L[2 * HL - 1], H[2 * HL - 1], M[2 * HL - 1];
  uint32_t al[NL], sa[HL], rl[NL];
  fp_ld(al, a);
#pragma unroll
  for (int i = 0; i < HL; i++) sa[i] = alu_add_u32(al[i], al[HL + i]);
  fp_sqr_half(L, al);
  fp_sqr_half(H, al + HL);
  fp_sqr_half(M, sa);
  uint64_t t[2 * NL];
#pragma unroll
  for (int i = 0; i < 2 * HL - 1; i++) {
    t[i] = L[i];
    t[NL + i] = H[i];
  }
  t[2 * HL - 1] = 0;
  t[2 * NL - 1] = 0;
#pragma unroll
  for (int i = 0; i < 2 * HL - 1; i++) t[HL + i] += M[i] - L[i] - H[i];
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const uint32_t m = opaque32(((uint32_t)t[i] * K_PINV28) & M28);
#pragma unroll
    for (int j = 0; j < NL; j++) t[i + j] += (uint64_t)m * p28(j);
    t[i + 1] += t[i] >> 28;
  }
  uint64_t c = 0;
#pragma unroll
  for (int j = 0; j < NL - 1; j++) {
    c += t[NL + j];
    rl[j] = (uint32_t)c & M28;
    c >>= 28;
  }
  rl[NL - 1] = (uint32_t)(c + t[2 * NL - 1]);
  fp_st(r, rl);
  TRK(r, 2.0, M28);
}

// out-of-line instances for everything that is not an innermost loop (keeps the instruction footprint bounded)
BLS_FN void fp_mul(Fp& r, const Fp& a, const Fp& b) { fp_mul_inl(r, a, b); }
BLS_FN void fp_sqr(Fp& r, const Fp& a) { fp_sqr_inl(r, a); }

// ---- reductions ---------------------------------------------------------------------------------------------------------
// exact sequential normalisation followed by one quotient-estimate subtraction: limbs < 2^28, value < 3p.
// Accepts any lazily reduced input (limbs < 2^32, value < 2000 p).  ~110 instructions: used where bounds must be reset
// without a multiplication.
BLS_HD void fp_red_inl(uint32_t* x, const Fp& a) {
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < NL - 1; i++) {
    c += a.l[i];
    x[i] = (uint32_t)c & M28;
    c >>= 28;
  }
  c += a.l[NL - 1];
  x[NL - 1] = (uint32_t)c;  // value < 2^392 keeps this below 2^28
  // quotient estimate from the top limb: p >> 364 = 0x1a011; q <= floor(value / p) <= q + 2
  uint32_t q = (uint32_t)(((uint64_t)x[NL - 1] * 40322ull) >> 32);
  int64_t s = 0;
#pragma unroll
  for (int i = 0; i < NL - 1; i++) {
    s += (int64_t)x[i] - (int64_t)((uint64_t)q * p28(i));
    x[i] = (uint32_t)s & M28;
    s >>= 28;  // arithmetic shift: borrows propagate as negative carries
  }
  s += (int64_t)x[NL - 1] - (int64_t)((uint64_t)q * p28(NL - 1));
  x[NL - 1] = (uint32_t)s;
}
BLS_FN void fp_red(Fp& r, const Fp& a) {
  uint32_t x[NL];
  fp_red_inl(x, a);
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = x[i];
  r.l[NL] = r.l[NL + 1] = 0;
  TRK(r, 3.0, M28);
}
// r = the representative of a in [0,p), limbs < 2^28 (for ==, is_zero, sgn0, serialisation)
BLS_FN void fp_canon(Fp& r, const Fp& a) {
  uint32_t x[NL];
  fp_red_inl(x, a);
  // at most three conditional subtractions of p
  for (int rep = 0; rep < 3; rep++) {
    uint32_t d[NL];
    int32_t br = 0;
#pragma unroll
    for (int i = 0; i < NL; i++) {
      int32_t t = (int32_t)x[i] - (int32_t)p28(i) + br;
      br = t >> 31;  // 0 or -1
      d[i] = (uint32_t)t & M28;
    }
    if (br == 0) {
#pragma unroll
      for (int i = 0; i < NL; i++) x[i] = d[i];
    }
  }
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = x[i];
  r.l[NL] = r.l[NL + 1] = 0;
  TRK(r, 1.0, M28);
}

BLS_HD bool fp_is_zero(const Fp& a) {
  Fp c;
  fp_canon(c, a);
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) t |= c.l[i];
  return t == 0;
}
BLS_HD bool fp_eq(const Fp& a, const Fp& b) {
  Fp ca, cb;
  fp_canon(ca, a);
  fp_canon(cb, b);
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) t |= ca.l[i] ^ cb.l[i];
  return t == 0;
}
// r = c ? a : b
BLS_HD void fp_select(Fp& r, bool c, const Fp& a, const Fp& b) {
#pragma unroll
  for (int i = 0; i < NL; i++) r.l[i] = c ? a.l[i] : b.l[i];
  r.l[NL] = r.l[NL + 1] = 0;
#if defined(BLS_TRACK)
  r.vb = a.vb > b.vb ? a.vb : b.vb;
  r.lb = a.lb > b.lb ? a.lb : b.lb;
#endif
}

// Montgomery <-> plain integers.  "raw" Fp values hold a plain integer < p in canonical limbs.
BLS_HD void fp_to_mont(Fp& r, const Fp& raw) {
  Fp r2;
  fp_set(r2, K_R2);
  fp_mul(r, raw, r2);
}
// canonical plain integer of a Montgomery value
BLS_HD void fp_from_mont(Fp& raw, const Fp& a) {
  Fp one, t;
  fp_zero(one);
  one.l[0] = 1;
  TRK(one, 1.0, 1);
  fp_mul(t, a, one);
  fp_canon(raw, t);
}

// raw (canonical plain) value > (p-1)/2 ?
BLS_HD bool raw_gt_half(const Fp& raw) {
  int32_t br = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    int32_t t = (int32_t)K_HALF_P28[i] - (int32_t)raw.l[i] + br;
    br = t >> 31;
  }
  return br != 0;
}

// r = a^e for a fixed exponent given as 12 little-endian 32-bit words (4-bit fixed window).  Variable time on e only.
BLS_FN void fp_pow_fixed(Fp& r, const Fp& a, const uint32_t* e) {
  Fp tbl[16];
  fp_one(tbl[0]);
  fp_norm(tbl[1], a);
  for (int i = 2; i < 16; i++) fp_mul(tbl[i], tbl[i - 1], tbl[1]);
  Fp acc;
  fp_one(acc);
  bool started = false;
  for (int limb = 11; limb >= 0; limb--) {
    uint32_t w = e[limb];
    for (int nib = 7; nib >= 0; nib--) {
      uint32_t d = (w >> (4 * nib)) & 15u;
      if (started) {
#pragma unroll 1
        for (int k = 0; k < 4; k++) fp_sqr_inl(acc, acc);
        if (d) fp_mul_inl(acc, acc, tbl[d]);
      } else if (d) {
        acc = tbl[d];
        started = true;
      }
    }
  }
  r = acc;
}

BLS_HD void fp_inv(Fp& r, const Fp& a) { fp_pow_fixed(r, a, K_EXP_PM2); }

// t = a^((p-3)/4).  If a is a non-zero square: t = 1/sqrt(a) and a*t = sqrt(a).  If a is a non-residue: t^2 * a = -1.
BLS_HD void fp_isqrt_pow(Fp& t, const Fp& a) { fp_pow_fixed(t, a, K_EXP_PM3D4); }

// sqrt; returns false if a is not a square
BLS_HD bool fp_sqrt(Fp& r, const Fp& a) {
  Fp t, s, chk;
  fp_isqrt_pow(t, a);
  fp_mul(s, t, a);
  fp_sqr(chk, s);
  r = s;
  return fp_eq(chk, a);
}

// 48 big-endian bytes (top three bits already masked by the caller) -> raw value; returns false if >= p
BLS_HD bool fp_from_be48_raw(Fp& raw, const uint8_t* b) {
  // bit k of the integer lives in byte 47 - k/8
#pragma unroll
  for (int i = 0; i < NL; i++) {
    uint64_t v = 0;
#pragma unroll
    for (int k = 0; k < 5; k++) {  // 5 bytes cover 28 bits at any alignment
      int byte_index = (28 * i) / 8 + k;
      if (byte_index < 48) v |= (uint64_t)b[47 - byte_index] << (8 * k);
    }
    raw.l[i] = (uint32_t)(v >> ((28 * i) % 8)) & M28;
  }
  raw.l[NL] = raw.l[NL + 1] = 0;
  TRK(raw, 1.0, M28);
  // < p ?
  int32_t br = 0;
#pragma unroll
  for (int i = 0; i < NL; i++) {
    int32_t t = (int32_t)raw.l[i] - (int32_t)p28(i) + br;
    br = t >> 31;
  }
  return br != 0;
}
BLS_HD void fp_to_be48_raw(uint8_t* b, const Fp& raw) {
#pragma unroll
  for (int byte_index = 0; byte_index < 48; byte_index++) {
    int bit = 8 * byte_index;
    int i = bit / 28, sh = bit % 28;
    uint32_t v = raw.l[i] >> sh;
    if (sh > 20 && i + 1 < NL) v |= raw.l[i + 1] << (28 - sh);
    b[47 - byte_index] = (uint8_t)v;
  }
}
// little-endian 32-bit words (n <= 12) of a plain integer < 2^(32 n) -> raw limbs (used by hash_to_field)
BLS_HD void fp_raw_from_words(Fp& raw, const uint32_t* w, int n) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    int bit = 28 * i;
    int wi = bit / 32, sh = bit % 32;
    uint64_t v = 0;
    if (wi < n) v = w[wi];
    if (wi + 1 < n) v |= (uint64_t)w[wi + 1] << 32;
    raw.l[i] = (uint32_t)(v >> sh) & M28;
  }
  raw.l[NL] = raw.l[NL + 1] = 0;
  TRK(raw, 1.0, M28);
}

}  // namespace bls
