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
BLS_HD void sop_ld28(int32_t* x0, int32_t* x1, const SFp2* p, bool fp_only) {
  const SopI4* q = reinterpret_cast<const SopI4*>(p->w);  // 128-bit loads: the records are 16-byte aligned
  int32_t t[2 * NL];
#pragma unroll
  for (int i = 0; i < (fp_only ? 4 : 7); i++) {
    SopI4 v = q[i];
    t[4 * i] = v.x;
    t[4 * i + 1] = v.y;
    t[4 * i + 2] = v.z;
    t[4 * i + 3] = v.w;
  }
#pragma unroll
  for (int i = 0; i < NL; i++) {
    x0[i] = t[i];
    x1[i] = fp_only ? 0 : t[NL + i];
  }
}

// x = s*p + s2*p2, then CONJ, then XI
BLS_HD SopBnd sop_load(int32_t* x0, int32_t* x1, const SFp2* p, int32_t s, const SFp2* p2, int32_t s2, uint32_t fl, bool fp_only) {
  SopBnd bd;
  bd.l0 = bd.l1 = bd.vb = 0;
  sop_ld28(x0, x1, p, fp_only);
  if (s != 1) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
      x0[i] *= s;
      x1[i] *= s;
    }
  }
#if defined(BLS_TRACK)
  {
    double as = s < 0 ? -(double)s : (double)s;
    bd.l0 = as * p->lb;
    bd.l1 = fp_only ? 0.0 : as * p->lb;
    bd.vb = as * p->vb;
  }
#endif
  if (p2 != nullptr) {
    int32_t y0[NL], y1[NL];
    sop_ld28(y0, y1, p2, fp_only);
#pragma unroll
    for (int i = 0; i < NL; i++) {
      x0[i] += y0[i] * s2;
      x1[i] += y1[i] * s2;
    }
#if defined(BLS_TRACK)
    {
      double as = s2 < 0 ? -(double)s2 : (double)s2;
      bd.l0 += as * p2->lb;
      bd.l1 += fp_only ? 0.0 : as * p2->lb;
      bd.vb += as * p2->vb;
    }
#endif
  }
  if (fl & SOP_CONJ) {
#pragma unroll
    for (int i = 0; i < NL; i++) x1[i] = -x1[i];
  }
  if (fl & SOP_XI) {
#pragma unroll
    for (int i = 0; i < NL; i++) {
      int32_t t0 = x0[i] - x1[i], t1 = x0[i] + x1[i];
      x0[i] = t0;
      x1[i] = t1;
    }
#if defined(BLS_TRACK)
    bd.l0 = bd.l1 = bd.l0 + bd.l1;
    bd.vb *= 2.0;
#endif
  }
#if defined(BLS_TRACK)
  BLS_REQ(bd.l0 < 1073741824.0 && bd.l1 < 1073741824.0, "sop operand limb overflow (>= 2^30)");
#endif
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

// The unit.  r may alias any operand: results are written after the last operand read.
BLS_FN void sop2s(SFp2& r, const SopT* t, int nt) {
  uint64_t A0[2 * NL], A1[2 * NL];
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) A0[i] = A1[i] = 0;
#if defined(BLS_TRACK)
  double col_re = 0, col_im = 0, vsum = 0;
  BLS_REQ(nt >= 1 && nt <= SOP_MAX_TERMS, "sop2s term count");
#endif
  // ---- pass 1: A0 += a0 b0 ; A1 += a1 b1
#pragma unroll 1
  for (int k = 0; k < nt; k++) {
    const bool bfp = (t[k].fl & SOP_BFP) != 0;
    int32_t a0[NL], a1[NL], b0[NL], b1[NL];
    SopBnd ba = sop_load(a0, a1, t[k].a, t[k].sa, t[k].a2, t[k].sa2, t[k].fl, false);
    SopBnd bb = sop_load(b0, b1, t[k].b, t[k].sb, t[k].b2, t[k].sb2, 0, bfp);
    sop_acc(A0, a0, b0);
    if (!bfp) sop_acc(A1, a1, b1);
#if defined(BLS_TRACK)
    col_re += 14.0 * (ba.l0 * bb.l0 + ba.l1 * bb.l1);
    col_im += 14.0 * (ba.l0 * bb.l1 + ba.l1 * bb.l0);
    vsum += 2.0 * ba.vb * bb.vb;
#else
    (void)ba;
    (void)bb;
#endif
  }
#if defined(BLS_TRACK)
  {
    // true column values + the reduction's own growth (14 * 2^56 for m*p, < 2^36 of carries) must fit int64
    const double lim = 9223372036854775808.0 - 15.0 * 72057594037927936.0;
    BLS_REQ(col_re < lim && col_im < lim, "sop2s column overflow");
    BLS_REQ(vsum / 2500.0 + 1.0 < 8.0, "sop2s result value bound");
  }
#endif
  // real part D = A0 - A1 ; S = A0 + A1 is kept for the imaginary part
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) {
    const uint64_t x = A0[i], y = A1[i];
    A0[i] = x - y;
    A1[i] = x + y;
  }
  A0[2 * NL - 1] = 0;
  int32_t c0[NL];
  sop_redc(c0, A0);
  // ---- pass 2: A2 += (a0 + a1)(b0 + b1) ; E = A2 - S
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) A0[i] = 0;
#pragma unroll 1
  for (int k = 0; k < nt; k++) {
    const bool bfp = (t[k].fl & SOP_BFP) != 0;
    int32_t a0[NL], a1[NL], b0[NL], b1[NL];
    sop_load(a0, a1, t[k].a, t[k].sa, t[k].a2, t[k].sa2, t[k].fl, false);
    sop_load(b0, b1, t[k].b, t[k].sb, t[k].b2, t[k].sb2, 0, bfp);
#pragma unroll
    for (int i = 0; i < NL; i++) {
      a0[i] += a1[i];
      b0[i] += b1[i];
    }
    sop_acc(A0, a0, b0);
  }
#pragma unroll
  for (int i = 0; i < 2 * NL - 1; i++) A0[i] -= A1[i];
  A0[2 * NL - 1] = 0;
  int32_t c1[NL];
  sop_redc(c1, A0);
#pragma unroll
  for (int i = 0; i < NL; i++) {
    r.w[i] = c0[i];
    r.w[NL + i] = c1[i];
  }
#if defined(BLS_TRACK)
  STRK(r, vsum / 2500.0 + 1.0, 134217728.0);
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
