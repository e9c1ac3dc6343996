// Fp: the BLS12-381 base field, 12 x 32-bit little-endian limbs, Montgomery form (R = 2^384), values kept in [0, p).
//
// Replaces (for public-input work) the Fp arithmetic the reference obtains from blstrs_plus/blst
// (reference src/impls.rs:185-215 re-exports the types; there is no arithmetic in the reference tree).
//
// Multiplication is the hot instruction stream of the whole engine: 288 IMAD.WIDE.U32 per product, arranged as
// two carry chains per row (even / odd columns) so that every 32x32->64 product is accumulated 64-bit aligned and
// the carries ride the CC flag (mad.lo.cc / madc.hi.cc pairs, which ptxas fuses into IMAD.WIDE.U32[.X]).
// The same row schedule is also expressed with portable C "chains" (BLS_PORTABLE_CHAINS) so the logic can be
// exercised by the CPU test harness (tests/hostemu) - that harness is test infrastructure, never a fallback.
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

#if defined(__CUDA_ARCH__) && !defined(BLS_PORTABLE_CHAINS)
#define BLS_ASM_CHAINS 1
#endif

namespace bls {

struct Fp {
  uint32_t l[12];
};

// p as immediates (folded by the compiler after unrolling)
BLS_HD constexpr uint32_t p_limb(int i) {
  return i == 0 ? 0xffffaaabu : i == 1 ? 0xb9feffffu : i == 2 ? 0xb153ffffu : i == 3 ? 0x1eabfffeu
       : i == 4 ? 0xf6b0f624u : i == 5 ? 0x6730d2a0u : i == 6 ? 0xf38512bfu : i == 7 ? 0x64774b84u
       : i == 8 ? 0x434bacd7u : i == 9 ? 0x4b1ba7b6u : i == 10 ? 0x397fe69au : 0x1a0111eau;
}

BLS_HD void fp_set(Fp& r, const uint32_t* c) {
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = c[i];
}
BLS_HD void fp_zero(Fp& r) {
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = 0;
}
BLS_HD void fp_one(Fp& r) { fp_set(r, K_ONE); }
BLS_HD bool fp_is_zero(const Fp& a) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) t |= a.l[i];
  return t == 0;
}
BLS_HD bool fp_eq(const Fp& a, const Fp& b) {
  uint32_t t = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) t |= a.l[i] ^ b.l[i];
  return t == 0;
}
// r = c ? a : b
BLS_HD void fp_select(Fp& r, bool c, const Fp& a, const Fp& b) {
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = c ? a.l[i] : b.l[i];
}

// returns borrow of (a - p); r = a - p (mod 2^384)
BLS_HD uint32_t fp_sub_p_raw(uint32_t* r, const uint32_t* a) {
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)a[i] - p_limb(i) - br;
    r[i] = (uint32_t)t;
    br = (t >> 32) & 1;
  }
  return (uint32_t)br;
}

BLS_HD void fp_add(Fp& r, const Fp& a, const Fp& b) {
  uint32_t s[12], d[12];
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    c += (uint64_t)a.l[i] + b.l[i];
    s[i] = (uint32_t)c;
    c >>= 32;
  }
  uint32_t br = fp_sub_p_raw(d, s);
#pragma unroll
  for (int i = 0; i < 12; i++) r.l[i] = br ? s[i] : d[i];
}

BLS_HD void fp_sub(Fp& r, const Fp& a, const Fp& b) {
  uint32_t d[12];
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)a.l[i] - b.l[i] - br;
    d[i] = (uint32_t)t;
    br = (t >> 32) & 1;
  }
  uint32_t mask = 0u - (uint32_t)br;
  uint64_t c = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    c += (uint64_t)d[i] + (p_limb(i) & mask);
    r.l[i] = (uint32_t)c;
    c >>= 32;
  }
}

BLS_HD void fp_neg(Fp& r, const Fp& a) {
  uint32_t nz = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) nz |= a.l[i];
  uint32_t mask = nz ? 0xffffffffu : 0u;
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)p_limb(i) - a.l[i] - br;
    r.l[i] = (uint32_t)t & mask;
    br = (t >> 32) & 1;
  }
}

BLS_HD void fp_dbl(Fp& r, const Fp& a) { fp_add(r, a, a); }

// ---------------------------------------------------------------------------------------------------------
// carry chains.  X[0..11] is a 12-limb accumulator; "pairs" are (X[2k], X[2k+1]).
//   chain_mul6 : X = {a[0],a[2],..,a[10]} * b                       (no carries between pairs needed)
//   chain_mad6 : X += {a[0],a[2],..} * b, carry out added into top
//   chain_mad6_shift : X0 += Z[1]; Z[k] = Z[k+2] + {a[0],a[2],..}*b for k=0..9 (carry-in from the X0 add),
//                      (Z[10],Z[11]) = a[10]*b + carry
// a points at the first of 6 limbs taken with stride 2.
// ---------------------------------------------------------------------------------------------------------
#if defined(BLS_ASM_CHAINS)
BLS_HD void chain_mul6(uint32_t* X, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int k = 0; k < 6; k++)
    asm("mul.lo.u32 %0, %2, %3;\n\tmul.hi.u32 %1, %2, %3;" : "=&r"(X[2 * k]), "=r"(X[2 * k + 1]) : "r"(a[2 * k]), "r"(b));  // & : %0 is written before the inputs are dead
}
BLS_HD void chain_mad6(uint32_t* X, const uint32_t* a, uint32_t b, uint32_t& top) {
  asm("mad.lo.cc.u32 %0, %13, %19, %0;\n\t"
      "madc.hi.cc.u32 %1, %13, %19, %1;\n\t"
      "madc.lo.cc.u32 %2, %14, %19, %2;\n\t"
      "madc.hi.cc.u32 %3, %14, %19, %3;\n\t"
      "madc.lo.cc.u32 %4, %15, %19, %4;\n\t"
      "madc.hi.cc.u32 %5, %15, %19, %5;\n\t"
      "madc.lo.cc.u32 %6, %16, %19, %6;\n\t"
      "madc.hi.cc.u32 %7, %16, %19, %7;\n\t"
      "madc.lo.cc.u32 %8, %17, %19, %8;\n\t"
      "madc.hi.cc.u32 %9, %17, %19, %9;\n\t"
      "madc.lo.cc.u32 %10, %18, %19, %10;\n\t"
      "madc.hi.cc.u32 %11, %18, %19, %11;\n\t"
      "addc.u32 %12, %12, 0;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8]),
        "+r"(X[9]), "+r"(X[10]), "+r"(X[11]), "+r"(top)
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(b));
}
// same without carry out (the caller knows it is zero)
BLS_HD void chain_mad6_nc(uint32_t* X, const uint32_t* a, uint32_t b) {
  asm("mad.lo.cc.u32 %0, %12, %18, %0;\n\t"
      "madc.hi.cc.u32 %1, %12, %18, %1;\n\t"
      "madc.lo.cc.u32 %2, %13, %18, %2;\n\t"
      "madc.hi.cc.u32 %3, %13, %18, %3;\n\t"
      "madc.lo.cc.u32 %4, %14, %18, %4;\n\t"
      "madc.hi.cc.u32 %5, %14, %18, %5;\n\t"
      "madc.lo.cc.u32 %6, %15, %18, %6;\n\t"
      "madc.hi.cc.u32 %7, %15, %18, %7;\n\t"
      "madc.lo.cc.u32 %8, %16, %18, %8;\n\t"
      "madc.hi.cc.u32 %9, %16, %18, %9;\n\t"
      "madc.lo.cc.u32 %10, %17, %18, %10;\n\t"
      "madc.hi.u32 %11, %17, %18, %11;"
      : "+r"(X[0]), "+r"(X[1]), "+r"(X[2]), "+r"(X[3]), "+r"(X[4]), "+r"(X[5]), "+r"(X[6]), "+r"(X[7]), "+r"(X[8]),
        "+r"(X[9]), "+r"(X[10]), "+r"(X[11])
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(b));
}
BLS_HD void chain_mad6_shift(uint32_t& X0, uint32_t* Z, const uint32_t* a, uint32_t b) {
  asm("add.cc.u32 %12, %12, %1;\n\t"
      "madc.lo.cc.u32 %0, %13, %19, %2;\n\t"
      "madc.hi.cc.u32 %1, %13, %19, %3;\n\t"
      "madc.lo.cc.u32 %2, %14, %19, %4;\n\t"
      "madc.hi.cc.u32 %3, %14, %19, %5;\n\t"
      "madc.lo.cc.u32 %4, %15, %19, %6;\n\t"
      "madc.hi.cc.u32 %5, %15, %19, %7;\n\t"
      "madc.lo.cc.u32 %6, %16, %19, %8;\n\t"
      "madc.hi.cc.u32 %7, %16, %19, %9;\n\t"
      "madc.lo.cc.u32 %8, %17, %19, %10;\n\t"
      "madc.hi.cc.u32 %9, %17, %19, %11;\n\t"
      "madc.lo.cc.u32 %10, %18, %19, 0;\n\t"
      "madc.hi.u32 %11, %18, %19, 0;"
      : "+r"(Z[0]), "+r"(Z[1]), "+r"(Z[2]), "+r"(Z[3]), "+r"(Z[4]), "+r"(Z[5]), "+r"(Z[6]), "+r"(Z[7]), "+r"(Z[8]),
        "+r"(Z[9]), "+r"(Z[10]), "+r"(Z[11]), "+r"(X0)
      : "r"(a[0]), "r"(a[2]), "r"(a[4]), "r"(a[6]), "r"(a[8]), "r"(a[10]), "r"(b));
}
#else
BLS_HD void chain_mul6(uint32_t* X, const uint32_t* a, uint32_t b) {
#pragma unroll
  for (int k = 0; k < 6; k++) {
    uint64_t t = (uint64_t)a[2 * k] * b;
    X[2 * k] = (uint32_t)t;
    X[2 * k + 1] = (uint32_t)(t >> 32);
  }
}
BLS_HD uint32_t chain_mad6_core(uint32_t* X, const uint32_t* a, uint32_t b) {
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    uint64_t pr = (uint64_t)a[2 * k] * b;
    uint64_t lo = (uint64_t)X[2 * k] + (uint32_t)pr + c;
    X[2 * k] = (uint32_t)lo;
    uint64_t hi = (uint64_t)X[2 * k + 1] + (uint32_t)(pr >> 32) + (lo >> 32);
    X[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
  return (uint32_t)c;
}
BLS_HD void chain_mad6(uint32_t* X, const uint32_t* a, uint32_t b, uint32_t& top) { top += chain_mad6_core(X, a, b); }
BLS_HD void chain_mad6_nc(uint32_t* X, const uint32_t* a, uint32_t b) { (void)chain_mad6_core(X, a, b); }
BLS_HD void chain_mad6_shift(uint32_t& X0, uint32_t* Z, const uint32_t* a, uint32_t b) {
  uint64_t c = (uint64_t)X0 + Z[1];
  X0 = (uint32_t)c;
  c >>= 32;
#pragma unroll
  for (int k = 0; k < 6; k++) {
    uint64_t pr = (uint64_t)a[2 * k] * b;
    uint32_t in_lo = (k < 5) ? Z[2 * k + 2] : 0u;
    uint32_t in_hi = (k < 5) ? Z[2 * k + 3] : 0u;
    uint64_t lo = (uint64_t)in_lo + (uint32_t)pr + c;
    uint64_t hi = (uint64_t)in_hi + (uint32_t)(pr >> 32) + (lo >> 32);
    Z[2 * k] = (uint32_t)lo;
    Z[2 * k + 1] = (uint32_t)hi;
    c = hi >> 32;
  }
}
#endif

// One Montgomery row: (X,Y) hold the running value V = X + Y*2^32 (X[k] at limb k, Y[k] at limb k+1).
// FIRST row: V = a*b0.  Other rows: Z is the previous row's X array (its limb 0 is zero), X is the previous Y;
// the value is shifted down one limb while a*bi is added.  Then m*p is added so that X[0] becomes 0.
template <bool FIRST>
BLS_HD void mont_row(uint32_t* X, uint32_t* Z, const uint32_t* a, uint32_t bi) {
  const uint32_t pl[12] = {p_limb(0), p_limb(1), p_limb(2), p_limb(3), p_limb(4), p_limb(5),
                           p_limb(6), p_limb(7), p_limb(8), p_limb(9), p_limb(10), p_limb(11)};
  if (FIRST) {
    chain_mul6(X, a, bi);
    chain_mul6(Z, a + 1, bi);
  } else {
    chain_mad6_shift(X[0], Z, a + 1, bi);
    chain_mad6(X, a, bi, Z[11]);
  }
  uint32_t m = X[0] * K_PINV32;
  chain_mad6_nc(Z, pl + 1, m);
  chain_mad6(X, pl, m, Z[11]);
}

// r = a*b*R^-1 mod p, inputs < p (or any a,b with a*b < p*R), output in [0,p)
BLS_HD void fp_mul_inl(Fp& r, const Fp& a, const Fp& b) {
  uint32_t U[12], V[12];
  mont_row<true>(U, V, a.l, b.l[0]);
#pragma unroll
  for (int i = 1; i < 12; i += 2) {
    mont_row<false>(V, U, a.l, b.l[i]);
    if (i + 1 < 12) mont_row<false>(U, V, a.l, b.l[i + 1]);
  }
  // after row 11: X = V (limb 0 zero), Y = U.  result limb k = X[k+1] + Y[k]
  uint32_t s[12], d[12];
  uint64_t c = 0;
#pragma unroll
  for (int k = 0; k < 12; k++) {
    c += (uint64_t)U[k] + (k < 11 ? V[k + 1] : 0u);
    s[k] = (uint32_t)c;
    c >>= 32;
  }
  uint32_t br = fp_sub_p_raw(d, s);
#pragma unroll
  for (int k = 0; k < 12; k++) r.l[k] = br ? s[k] : d[k];
}

BLS_HD void fp_sqr_inl(Fp& r, const Fp& a) { fp_mul_inl(r, a, a); }
// out-of-line instances for everything that is not an innermost loop (keeps the instruction footprint bounded)
BLS_FN void fp_mul(Fp& r, const Fp& a, const Fp& b) { fp_mul_inl(r, a, b); }
BLS_FN void fp_sqr(Fp& r, const Fp& a) { fp_mul_inl(r, a, a); }

// plain reference multiplication (CIOS with 64-bit temporaries); used by parity kernels to cross-check fp_mul
BLS_HD void fp_mul_cios(Fp& r, const Fp& a, const Fp& b) {
  uint32_t t[14];
#pragma unroll
  for (int i = 0; i < 14; i++) t[i] = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t c = 0;
#pragma unroll
    for (int j = 0; j < 12; j++) {
      c += (uint64_t)a.l[j] * b.l[i] + t[j];
      t[j] = (uint32_t)c;
      c >>= 32;
    }
    c += t[12];
    t[12] = (uint32_t)c;
    t[13] = (uint32_t)(c >> 32);
    uint32_t m = t[0] * K_PINV32;
    c = ((uint64_t)m * p_limb(0) + t[0]) >> 32;
#pragma unroll
    for (int j = 1; j < 12; j++) {
      c += (uint64_t)m * p_limb(j) + t[j];
      t[j - 1] = (uint32_t)c;
      c >>= 32;
    }
    c += t[12];
    t[11] = (uint32_t)c;
    t[12] = t[13] + (uint32_t)(c >> 32);
  }
  uint32_t d[12];
  uint32_t br = fp_sub_p_raw(d, t);
  bool keep = br && t[12] == 0;
#pragma unroll
  for (int k = 0; k < 12; k++) r.l[k] = keep ? t[k] : d[k];
}

// Montgomery <-> canonical
BLS_HD void fp_to_mont(Fp& r, const Fp& a) {
  Fp r2;
  fp_set(r2, K_R2);
  fp_mul(r, a, r2);
}
BLS_HD void fp_from_mont(Fp& r, const Fp& a) {
  Fp one;
  fp_zero(one);
  one.l[0] = 1;
  fp_mul(r, a, one);
}

// raw integer compare helpers on canonical (non-Montgomery) limbs
BLS_HD bool raw_ge_p(const uint32_t* a) {
  uint32_t d[12];
  return fp_sub_p_raw(d, a) == 0;
}
// canonical a > (p-1)/2 ?
BLS_HD bool raw_gt_half(const uint32_t* a) {
  uint64_t br = 0;
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint64_t t = (uint64_t)K_HALF_P[i] - a[i] - br;
    br = (t >> 32) & 1;
  }
  return br != 0;
}

// r = a^e for a fixed 384-bit exponent given as RAW limbs (4-bit fixed window).  Variable time on e only.
BLS_FN void fp_pow_fixed(Fp& r, const Fp& a, const uint32_t* e) {
  Fp tbl[16];
  fp_one(tbl[0]);
  tbl[1] = a;
  for (int i = 2; i < 16; i++) fp_mul(tbl[i], tbl[i - 1], a);
  Fp acc;
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

// 48 big-endian bytes (top three bits already masked by the caller) -> canonical limbs; returns false if >= p
BLS_HD bool fp_from_be48_raw(uint32_t* raw, const uint8_t* b) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    const uint8_t* q = b + 44 - 4 * i;
    raw[i] = ((uint32_t)q[0] << 24) | ((uint32_t)q[1] << 16) | ((uint32_t)q[2] << 8) | (uint32_t)q[3];
  }
  return !raw_ge_p(raw);
}
BLS_HD void fp_to_be48_raw(uint8_t* b, const uint32_t* raw) {
#pragma unroll
  for (int i = 0; i < 12; i++) {
    uint8_t* q = b + 44 - 4 * i;
    q[0] = (uint8_t)(raw[i] >> 24);
    q[1] = (uint8_t)(raw[i] >> 16);
    q[2] = (uint8_t)(raw[i] >> 8);
    q[3] = (uint8_t)raw[i];
  }
}

}  // namespace bls
