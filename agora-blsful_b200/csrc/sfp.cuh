// S-form Fp2 and the fused sum-of-products unit `sop2s`:
//
//     r = ( sum_t  A_t * B_t ) / R            (Fp2, Montgomery form, R = 2^392, up to SOP_MAX_TERMS terms)
//
// with A_t = sa*a + sa2*a2 (optionally conjugated, optionally times xi = 1+u) and B_t = sb*b + sb2*b2 read from memory, and
// ONE Montgomery reduction per output coefficient: the 64-bit column accumulators of all terms are summed before
// reducing (lazy reduction), so a coefficient of an Fp12 product that is a sum of three or four Fp2 products costs
// 3 integer products per term (Karatsuba) plus 2 reductions, and no additive "glue" pass through local memory.
// This is the arithmetic under the cooperative Miller loop (miller6.cuh): it replaces what blsful takes from
// blstrs_plus/blst below `multi_miller_loop` (reference src/helpers.rs:50,62; there is no arithmetic in the reference tree).
//
// S-form: 14 SIGNED limbs of 28 bits per Fp ("balanced": stored limbs 0..12 lie in [-2^27, 2^27), limb 13 carries the
// rest and the sign), values are signed too (|value| < ~2p).  Consequences:
//   * products of balanced limbs are < 2^54, so a column of 14 of them is < 2^57.8: a sum of EIGHT Fp2 products with
//     doubled / xi-multiplied operands still fits a signed 64-bit column (fp.cuh's unsigned form allows one);
//   * differences and negations are plain limb-wise IADDs: no "spread" multiples of p, no renormalisation;
//   * the reduction output range is (-e p, (1+e) p) with e = sum |A||B| / (R p) << 1: bounds never compound.
// Intermediate Karatsuba columns may wrap modulo 2^64; only the true final column values must fit, and that is what the
// BLS_TRACK build checks for every call (worst case over ALL inputs, see tests/test_hostemu.py).
#pragma once
#include "fp2.cuh"

namespace bls {

struct alignas(16) SFp2 {
  int32_t w[2 * NL];  // c0 = w[0..13], c1 = w[14..27]; limb j has weight 2^(28 j)
#if defined(BLS_TRACK)
  double vb;  // |value| <= vb * p (both coefficients)
  double lb;  // |limb| <= lb
#endif
};

#if defined(BLS_TRACK)
#define STRK(r, v, l) \
  do {                \
    (r).vb = (v);     \
    (r).lb = (l);     \
  } while (0)
#else
#define STRK(r, v, l) \
  do {                \
  } while (0)
#endif

enum : uint32_t {
  SOP_XI = 1u,    // A operand times xi = 1 + u
  SOP_CONJ = 2u,  // A operand conjugated (applied before xi)
  SOP_BFP = 4u,   // B is an Fp scalar: its c1 is ignored (taken as zero): 2 integer products instead of 3
};
constexpr int SOP_MAX_TERMS = 8;

struct SopT {
  const SFp2 *a, *a2, *b, *b2;  // a2 / b2 may be nullptr
  int32_t sa, sa2, sb, sb2;     // small integer scales
  uint32_t fl;
};
BLS_HD SopT sop_t(const SFp2* a, const SFp2* b, int32_t sa = 1, uint32_t fl = 0) {
  SopT t;
  t.a = a;
  t.a2 = nullptr;
  t.b = b;
  t.b2 = nullptr;
  t.sa = sa;
  t.sa2 = 0;
  t.sb = 1;
  t.sb2 = 0;
  t.fl = fl;
  return t;
}
BLS_HD SopT sop_t2(const SFp2* a, int32_t sa, const SFp2* a2, int32_t sa2, const SFp2* b, int32_t sb, const SFp2* b2, int32_t sb2,
                   uint32_t fl = 0) {
  SopT t;
  t.a = a;
  t.a2 = a2;
  t.b = b;
  t.b2 = b2;
  t.sa = sa;
  t.sa2 = sa2;
  t.sb = sb;
  t.sb2 = sb2;
  t.fl = fl;
  return t;
}

// ---- operand load -------------------------------------------------------------------------------------------------
struct SopBnd {  // BLS_TRACK only: limb magnitude bounds of the prepared operand halves and its value bound
  double l0, l1, vb;
};

struct alignas(16) SopI4 {
  int32_t x, y, z, w;
};
BLS_HD void sop_ld28(int32_t* t, const SFp2* p) {
  const SopI4* q = reinterpret_cast<const SopI4*>(p->w);  // 128-bit loads: the records are 16-byte aligned
#pragma unroll
  for (int i = 0; i < 7; i++) {
    SopI4 v = q[i];
    t[4 * i] = v.x;
    t[4 * i + 1] = v.y;
    t[4 * i + 2] = v.z;
    t[4 * i + 3] = v.w;
  }
}

// The operand of integer product number `pass` (0: c0 half, 1: c1 half, 2: c0 + c1) of  s*p + s2*p2  [conj] [* xi].
// fp_only: the c1 half is taken as zero.
BLS_HD SopBnd sop_operand(int32_t* x, int pass, const SFp2* p, int32_t s, const SFp2* p2, int32_t s2, uint32_t fl, bool fp_only) {
  SopBnd bd;
  bd.l0 = bd.l1 = bd.vb = 0;
  int32_t t[2 * NL];
  sop_ld28(t, p);
  if (s != 1) {
#pragma unroll
    for (int i = 0; i < 2 * NL; i++) t[i] *= s;
  }
#if defined(BLS_TRACK)
  {
    double as = s < 0 ? -(double)s : (double)s;
    bd.l0 = bd.l1 = as * p->lb;
    bd.vb = as * p->vb;
  }
#endif
  if (p2 != nullptr) {
    int32_t u[2 * NL];
    sop_ld28(u, p2);
#pragma unroll
    for (int i = 0; i < 2 * NL; i++) t[i] += u[i] * s2;
#if defined(BLS_TRACK)
    {
      double as = s2 < 0 ? -(double)s2 : (double)s2;
      bd.l0 += as * p2->lb;
      bd.l1 += as * p2->lb;
      bd.vb += as * p2->vb;
    }
#endif
  }
  if (fp_only) {
#pragma unroll
    for (int i = 0; i < NL; i++) t[NL + i] = 0;
#if defined(BLS_TRACK)
    bd.l1 = 0;
#endif
  }
  if (fl & SOP_CONJ) {
#pragma unroll
    for (int i = 0; i < NL; i++) t[NL + i] = -t[NL + i];
  }
  if (fl & SOP_XI) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
      int32_t t0 = t[i] - t[NL + i], t1 = t[i] + t[NL + i];
      t[i] = t0;
      t[NL + i] = t1;
    }
#if defined(BLS_TRACK)
    bd.l0 = bd.l1 = bd.l0 + bd.l1;
    bd.vb *= 2.0;
#endif
  }
#if defined(BLS_TRACK)
  BLS_REQ(bd.l0 < 1073741824.0 && bd.l1 < 1073741824.0, "sop operand limb overflow (>= 2^30)");
#endif
#pragma unroll
  for (int i = 0; i < NL; i++) x[i] = pass == 0 ? t[i] : pass == 1 ? t[NL + i] : t[i] + t[NL + i];
  return bd;
}

// T[i+j] += a[i] * b[j]   (196 signed IMAD.WIDE; columns wrap modulo 2^64 by design)
BLS_HD void sop_acc(uint64_t* T, const int32_t* a, const int32_t* b) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) T[i + j] += (uint64_t)((int64_t)a[i] * (int64_t)b[j]);
  }
}

// Montgomery reduction of 27 signed columns (T[27] must be 0 on entry) -> 14 balanced limbs.  210 IMAD.
// Result value in ( -|T|/R , |T|/R + p ).
BLS_HD void sop_redc(int32_t* out, uint64_t* T) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const uint32_t m = opaque32(((uint32_t)T[i] * K_PINV28) & M28);
#pragma unroll
    for (int j = 0; j < NL; j++) T[i + j] += (uint64_t)m * p28(j);
    T[i + 1] += (uint64_t)((int64_t)T[i] >> 28);  // exact: the low 28 bits are zero now; arithmetic shift keeps the sign
  }
  int64_t c = 0;
#pragma unroll
  for (int j = 0; j < NL - 1; j++) {
    c += (int64_t)T[NL + j];
    const int64_t t = c + (1 << 27);
    out[j] = (int32_t)((uint32_t)t & M28) - (1 << 27);
    c = t >> 28;
  }
  out[NL - 1] = (int32_t)(c + (int64_t)T[2 * NL - 1]);
}

// A 28 x u64 scratch record that stays in LOCAL memory and is moved with 128-bit local accesses (a plain array would be
// promoted to 56 registers; a laundered generic pointer costs 64-bit generic accesses plus address bookkeeping).
struct alignas(16) SopKeep {
  uint64_t v[2 * NL];
};
BLS_HD void sop_keep_st(SopKeep& k, int i, uint64_t a, uint64_t b) {  // v[2i], v[2i+1]
#if defined(__CUDA_ARCH__)
  asm volatile("st.local.v2.u64 [%0], {%1, %2};" ::"l"(__cvta_generic_to_local(&k.v[2 * i])), "l"(a), "l"(b) : "memory");
#else
  k.v[2 * i] = a;
  k.v[2 * i + 1] = b;
#endif
}
BLS_HD void sop_keep_ld(const SopKeep& k, int i, uint64_t& a, uint64_t& b) {
#if defined(__CUDA_ARCH__)
  asm volatile("ld.local.v2.u64 {%0, %1}, [%2];" : "=l"(a), "=l"(b) : "l"(__cvta_generic_to_local(&k.v[2 * i])) : "memory");
#else
  a = k.v[2 * i];
  b = k.v[2 * i + 1];
#endif
}

// The unit.  r may alias any operand: results are written after the last operand read.
// Code-size discipline (the SM's instruction caches are 6 KB / 32 KB): ONE product body, ONE
// reduction body; the three Karatsuba products P0 = sum a0 b0, P1 = sum a1 b1, P2 = sum (a0+a1)(b0+b1) are three trips
// through the same loop, P0 (then P0 + P1) waits in local memory meanwhile.
//   real = P0 - P1 ,  imag = P2 - (P0 + P1)
BLS_FN void sop2s(SFp2& r, const SopT* t, int nt) {
  SopKeep keep;
  int32_t res[2 * NL];
#if defined(BLS_TRACK)
  double col_re = 0, col_im = 0, vsum = 0;
  BLS_REQ(nt >= 1 && nt <= SOP_MAX_TERMS, "sop2s term count");
#endif
#pragma unroll 1
  for (int pass = 0; pass < 3; pass++) {
    uint64_t T[2 * NL];
#pragma unroll
    for (int i = 0; i < 2 * NL; i++) T[i] = 0;
#pragma unroll 1
    for (int k = 0; k < nt; k++) {
      const bool bfp = (t[k].fl & SOP_BFP) != 0;
      if (bfp && pass == 1) continue;
      int32_t x[NL], y[NL];
      SopBnd ba = sop_operand(x, pass, t[k].a, t[k].sa, t[k].a2, t[k].sa2, t[k].fl, false);
      SopBnd bb = sop_operand(y, pass, t[k].b, t[k].sb, t[k].b2, t[k].sb2, 0, bfp);
      sop_acc(T, x, y);
#if defined(BLS_TRACK)
      if (pass == 0) {
        col_re += 14.0 * (ba.l0 * bb.l0 + ba.l1 * bb.l1);
        col_im += 14.0 * (ba.l0 * bb.l1 + ba.l1 * bb.l0);
        vsum += 2.0 * ba.vb * bb.vb;
      }
#else
      (void)ba;
      (void)bb;
#endif
    }
    if (pass == 0) {
#pragma unroll
      for (int i = 0; i < NL; i++) sop_keep_st(keep, i, T[2 * i], T[2 * i + 1]);
      continue;
    }
#if defined(BLS_TRACK)
    {
      // true column values + the reduction's own growth (14 * 2^56 for m*p, < 2^36 of carries) must fit int64
      const double lim = 9223372036854775808.0 - 15.0 * 72057594037927936.0;
      BLS_REQ(col_re < lim && col_im < lim, "sop2s column overflow");
      BLS_REQ(vsum / 2500.0 + 1.0 < 16.0, "sop2s result value bound");
    }
#endif
    // pass 1: U = P0 - P1, keep <- P0 + P1 ;  pass 2: U = P2 - keep
#pragma unroll
    for (int i = 0; i < NL; i++) {
      uint64_t s0, s1;
      sop_keep_ld(keep, i, s0, s1);
      const uint64_t v0 = T[2 * i], v1 = T[2 * i + 1];
      if (pass == 1) sop_keep_st(keep, i, s0 + v0, s1 + v1);
      T[2 * i] = pass == 1 ? s0 - v0 : v0 - s0;
      T[2 * i + 1] = pass == 1 ? s1 - v1 : v1 - s1;
    }
    T[2 * NL - 1] = 0;
    int32_t c[NL];
    sop_redc(c, T);
#pragma unroll
    for (int i = 0; i < NL; i++) {
      if (pass == 1) res[i] = c[i]; else res[NL + i] = c[i];
    }
  }
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) r.w[i] = res[i];
#if defined(BLS_TRACK)
  STRK(r, vsum / 2500.0 + 1.0, 134217728.0);
#endif
}

// r = sx * [xi] x + sy * y + sz * z  (y, z optional), limbs renormalised (balanced) - no multiplication, no reduction.
// r may alias the inputs.
BLS_FN void sfp2_lin(SFp2& r, const SFp2* x, int32_t sx, uint32_t flx, const SFp2* y, int32_t sy, const SFp2* z, int32_t sz) {
  int32_t t[2 * NL];
  int64_t v[2 * NL];
  sop_ld28(t, x);
  if (flx & SOP_XI) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
      const int32_t d = t[i] - t[NL + i], e = t[i] + t[NL + i];  // limbs < 2^30 (checked): no overflow
      v[i] = (int64_t)sx * (int64_t)d;
      v[NL + i] = (int64_t)sx * (int64_t)e;
    }
  } else {
#pragma unroll
    for (int i = 0; i < 2 * NL; i++) v[i] = (int64_t)sx * (int64_t)t[i];
  }
#if defined(BLS_TRACK)
  double vb = (sx < 0 ? -(double)sx : (double)sx) * x->vb * ((flx & SOP_XI) ? 2.0 : 1.0);
  BLS_REQ(x->lb < 1073741824.0, "sfp2_lin limb");
#endif
  if (y != nullptr) {
    sop_ld28(t, y);
#pragma unroll
    for (int i = 0; i < 2 * NL; i++) v[i] += (int64_t)sy * (int64_t)t[i];
#if defined(BLS_TRACK)
    vb += (sy < 0 ? -(double)sy : (double)sy) * y->vb;
#endif
  }
  if (z != nullptr) {
    sop_ld28(t, z);
#pragma unroll
    for (int i = 0; i < 2 * NL; i++) v[i] += (int64_t)sz * (int64_t)t[i];
#if defined(BLS_TRACK)
    vb += (sz < 0 ? -(double)sz : (double)sz) * z->vb;
#endif
  }
#pragma unroll
  for (int h = 0; h < 2; h++) {
    int64_t c = 0;
#pragma unroll
    for (int j = 0; j < NL - 1; j++) {
      c += v[h * NL + j];
      const int64_t u = c + (1 << 27);
      r.w[h * NL + j] = (int32_t)((uint32_t)u & M28) - (1 << 27);
      c = u >> 28;
    }
    r.w[h * NL + NL - 1] = (int32_t)(c + v[h * NL + NL - 1]);
  }
#if defined(BLS_TRACK)
  BLS_REQ(vb < 1000.0, "sfp2_lin value bound");  // keeps the top limb below 2^27
  STRK(r, vb, 134217728.0);
#endif
}

// ---- conversions between the unsigned lazy form (fp.cuh) and S-form -----------------------------------------------------
// Unsigned limbs are taken as they are (they must be < 2^30); bounds carry over.
BLS_HD void sfp2_from_fp2(SFp2& r, const Fp2& a) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    r.w[i] = (int32_t)a.c0.l[i];
    r.w[NL + i] = (int32_t)a.c1.l[i];
  }
#if defined(BLS_TRACK)
  BLS_REQ(a.c0.lb < (1u << 30) && a.c1.lb < (1u << 30), "sfp2_from_fp2: limbs too large");
  STRK(r, a.c0.vb > a.c1.vb ? a.c0.vb : a.c1.vb, (double)(a.c0.lb > a.c1.lb ? a.c0.lb : a.c1.lb));
#endif
}
// An Fp scalar as the c0 of an S-form record (c1 = 0): the B operand of SOP_BFP terms
BLS_HD void sfp2_from_fp(SFp2& r, const Fp& a) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    r.w[i] = (int32_t)a.l[i];
    r.w[NL + i] = 0;
  }
#if defined(BLS_TRACK)
  BLS_REQ(a.lb < (1u << 30), "sfp2_from_fp: limbs too large");
  STRK(r, a.vb, (double)a.lb);
#endif
}
// value + 4p with limbs in [0, 2^28): a normalised unsigned element of value < 4 + vb (needs vb < 4)
BLS_HD void sfp_to_fp(Fp& r, const int32_t* w) {
  int64_t c = 0;
#pragma unroll
  for (int j = 0; j < NL - 1; j++) {
    c += (int64_t)w[j] + 4 * (int64_t)p28(j);
    r.l[j] = (uint32_t)c & M28;
    c >>= 28;
  }
  c += (int64_t)w[NL - 1] + 4 * (int64_t)p28(NL - 1);
  r.l[NL - 1] = (uint32_t)c;
  r.l[NL] = r.l[NL + 1] = 0;
}
BLS_HD void fp2_from_sfp2(Fp2& r, const SFp2& a) {
#if defined(BLS_TRACK)
  BLS_REQ(a.vb < 4.0 && a.lb < 2147483648.0, "fp2_from_sfp2: value bound >= 4");
#endif
  sfp_to_fp(r.c0, a.w);
  sfp_to_fp(r.c1, a.w + NL);
  TRK(r.c0, 4.0 + a.vb, M28);
  TRK(r.c1, 4.0 + a.vb, M28);
}
BLS_HD void sfp2_zero(SFp2& r) {
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) r.w[i] = 0;
  STRK(r, 0.0, 0.0);
}
BLS_HD void sfp2_one(SFp2& r) {
  Fp2 o;
  fone(o);
  sfp2_from_fp2(r, o);
}
BLS_HD void sfp2_neg(SFp2& r, const SFp2& a) {
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) r.w[i] = -a.w[i];
  STRK(r, a.vb, a.lb);
}

}  // namespace bls
