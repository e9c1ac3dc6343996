// S-form Fp2 and the fused sum-of-products unit `sop2f`:
//
//     r = ( sum_t  A_t * B_t ) / R            (Fp2, Montgomery form, R = 2^392, up to SOP_MAX_TERMS terms)
//
// with A_t = +-(a << sha) [* xi], B_t = b << shb records read from memory, and ONE Montgomery reduction per output
// coefficient: the 64-bit column accumulators of all terms are summed before reducing (lazy reduction), so a coefficient
// of an Fp12 product that is a sum of three or four Fp2 products costs 3 integer products per term (Karatsuba) plus
// 2 reductions, and no additive "glue" pass through local memory.
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

// 1: sop_acc multiplies with one Karatsuba level over the limbs (Miller stage 366 -> 357 ms at 1M); 0: schoolbook
#if !defined(SOP_KARATSUBA)
#define SOP_KARATSUBA 1
#endif

namespace bls {

struct alignas(16) SFp2 {
  int32_t w[2 * NL];  // c0 = w[0..13], c1 = w[14..27]; limb j has weight 2^(28 j)
#if defined(BLS_TRACK)
  double vb;  // |value| <= vb * p (both coefficients)
  double lb;  // |limb| <= lb
#endif
};

// "Expanded" records: the operand halves of the three Karatsuba products stored side by side, 16 words apiece, so that the
// consumer's operand of ANY product is one run of four 128-bit loads and nothing else (no masks, sums or differences in
// the consumer's loop).  The producer pays 28 additions and a few more stores once.
//   SLineRec  (B operands only):  b0 | b1 | b0 + b1                 192 bytes   - the line records streamed through HBM
//   SAccRec   (A and B operands): a0 | a1 | a0 + a1 | a0 - a1       256 bytes   - the accumulator coefficients
// (xi * a has the halves a0 - a1, a0 + a1 and their sum 2 a0: every one is a stored run, the last with a shift.)
struct alignas(16) SLineRec {
  int32_t w[48];
#if defined(BLS_TRACK)
  double vb, lb;  // bounds of the value and of the limbs of b0 / b1 (the sum run has twice the limb bound)
#endif
};
struct alignas(16) SAccRec {
  int32_t w[64];
  int32_t pad[4];  // 272-byte stride: the six lanes of a group (and the five groups of a warp) read DIFFERENT records of the
                   // accumulator at the same run offset; a 256-byte stride would put them all on the same four banks
#if defined(BLS_TRACK)
  double vb, lb;
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
  SOP_XI = 1u,        // A operand times xi = 1 + u
  SOP_NEG = 2u,       // A operand negated
  SOP_XI_LT2 = 8u,    // A operand times xi when the lane's coefficient index k < 2   (Fp12 wrap-around, miller6.cuh)
  SOP_XI_LT3 = 16u,   //                                                      k < 3
  SOP_SQR = 32u,      // a and b name the SAME record (a square, up to the shifts): the two-lane unit sop1 then needs ONE
                      // product per lane - (a0 + a1)(a0 - a1) | 2 a0 a1 - instead of two; other units may ignore the flag
};
constexpr int SOP_MAX_TERMS = 8;

// Operand references are one-byte indices into a few record spaces, so that a whole computation is a constant table
// (no descriptor is ever built in local memory; the table is read through the constant cache):
//   0..23   the pair's record file: reg[i * reg_stride]       24..31  read-only per-item records P[i - 24] (HBM)
//   32..34  the line record being produced (destination of the line programs)
//   48..53  accumulator coefficient F[i - 48]                  56..58  coefficient i - 56 of the line being multiplied in
//   64..69  accumulator coefficient F[(k - (i - 64)) mod 6], k = the lane's coefficient index        255 none
enum : uint8_t { SOPX_P = 24, SOPX_LINE = 32, SOPX_F = 48, SOPX_JL = 56, SOPX_FREL = 64, SOPX_NONE = 255 };
// one product term  +-(a << sha) [xi]  *  (b << shb)
struct SopTerm {
  uint8_t a, b, sha, shb, fl, pad0, pad1, pad2;
};
struct SopSpaces {
  SFp2* reg;
  int reg_stride;  // in records: 1 = the thread's records are contiguous; blockDim.x = record-major shared memory
  const SFp2* P;
  SLineRec* line;  // destination only
  const SFp2* F;
  const SFp2* jl;
  int k;
};
// record reference -> address (branch-free: the callers sit inside the software-pipelined loop)
BLS_HD SFp2* sop_rec(const SopSpaces& c, uint32_t i) {
  const int is_reg = i < SOPX_P, is_p = (i >= SOPX_P) & (i < SOPX_LINE), is_jl = (i >= SOPX_JL) & (i < SOPX_FREL), is_rel = i >= SOPX_FREL;
  int jr = c.k - ((int)i - SOPX_FREL);
  jr += jr < 0 ? 6 : 0;
  const int j = is_reg ? (int)i * c.reg_stride : is_p ? (int)i - SOPX_P : is_jl ? (int)i - SOPX_JL : is_rel ? jr : (int)i - SOPX_F;
  SFp2* base = is_reg ? c.reg : is_p ? const_cast<SFp2*>(c.P) : is_jl ? const_cast<SFp2*>(c.jl) : const_cast<SFp2*>(c.F);
  return base + j;
}

struct alignas(16) SopI4 {
  int32_t x, y, z, w;
};
#if !defined(__CUDACC__)
struct alignas(8) int2 {
  int32_t x, y;
};
static inline int2 make_int2(int32_t a, int32_t b) {
  int2 r;
  r.x = a;
  r.y = b;
  return r;
}
#endif
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

// T[i+j] += a[i] * b[j]   (signed IMAD.WIDE; columns wrap modulo 2^64 by design)
BLS_HD void sop_acc(uint64_t* T, const int32_t* a, const int32_t* b) {
#if SOP_KARATSUBA
  // One Karatsuba level over the limbs (7 + 7): 147 multiplies instead of 196.  The identity holds modulo 2^64, so T is
  // bit-identical to the schoolbook columns; |a_i| + |a_{7+i}| < 2^31 because every caller keeps operand limbs < 2^30.
  constexpr int H = NL / 2;
  uint64_t C[2 * H - 1];
#pragma unroll
  for (int part = 0; part < 2; part++) {  // low x low -> columns 0..12, high x high -> 14..26, both leave the middle
#pragma unroll
    for (int i = 0; i < 2 * H - 1; i++) C[i] = 0;
#pragma unroll
    for (int i = 0; i < H; i++) {
#pragma unroll
      for (int j = 0; j < H; j++) C[i + j] += (uint64_t)((int64_t)a[part * H + i] * (int64_t)b[part * H + j]);
    }
#pragma unroll
    for (int i = 0; i < 2 * H - 1; i++) {
      T[2 * part * H + i] += C[i];
      T[H + i] -= C[i];
    }
  }
  int32_t sa[H], sb[H];
#pragma unroll
  for (int i = 0; i < H; i++) {
    sa[i] = alu_add_s32(a[i], a[H + i]);
    sb[i] = alu_add_s32(b[i], b[H + i]);
  }
#pragma unroll
  for (int i = 0; i < H; i++) {
#pragma unroll
    for (int j = 0; j < H; j++) T[H + i + j] += (uint64_t)((int64_t)sa[i] * (int64_t)sb[j]);
  }
#else
#pragma unroll
  for (int i = 0; i < NL; i++) {
#pragma unroll
    for (int j = 0; j < NL; j++) T[i + j] += (uint64_t)((int64_t)a[i] * (int64_t)b[j]);
  }
#endif
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
// promoted to 56 registers, which the pipelined loop below cannot spare).
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

// ---- the unit ---------------------------------------------------------------------------------------------------------
// One branch-free loop body per integer product: the 128-bit loads of the NEXT product's two records are issued first,
// then the 196 IMAD.WIDE of the current product, and the next operands are prepared with ALU instructions only (masks,
// shifts, negations: the multiplier pipe sees nothing but the products) - loads and preparation hide under the
// multiplier stream of the same warp instead of alternating with it.  Code-size discipline (instruction caches: 6 KB per
// SM sub-partition, 32 KB per SM): ONE product body, ONE reduction body, no second variant.
//   fp_mode 0:  three Karatsuba trips  P0 = sum a0 b0, P1 = sum a1 b1, P2 = sum (a0+a1)(b0+b1):
//               real = P0 - P1 ,  imag = P2 - (P0 + P1)
//   fp_mode 1:  every B is an Fp scalar (its c0): two trips, real = sum a0 b0, imag = sum a1 b0
struct SopPrep {  // operand of one integer product from the raw halves: x = +-(((h0 & m0) + (((h1 ^ n1) - n1) & m1)) << sh)
  int32_t m0, m1, n1, ng, sh;
};
BLS_HD SopPrep sop_prep(int half, int xi, int neg, int sh) {  // half: 0 c0, 1 c1, 2 c0 + c1.  Branch-free.
  SopPrep p;
  // plain: (a0), (a1), (a0 + a1) ;  xi: (a0 - a1), (a0 + a1), (2 a0)
  const int p0 = half == 0, p1 = half == 1, p2 = half == 2;
  p.m0 = -((p1 ^ 1) | xi);
  p.m1 = -(((p2 & xi) ^ 1) & ((p0 & (xi ^ 1)) ^ 1));
  p.n1 = -(p0 & xi);
  p.ng = -neg;
  p.sh = sh + (p2 & xi);
  return p;
}
BLS_HD void sop_prep_apply(int32_t* x, const SopI4* r, const SopPrep& p) {
  int32_t t[2 * NL];
#pragma unroll
  for (int i = 0; i < 7; i++) {
    t[4 * i] = r[i].x;
    t[4 * i + 1] = r[i].y;
    t[4 * i + 2] = r[i].z;
    t[4 * i + 3] = r[i].w;
  }
#pragma unroll
  for (int i = 0; i < NL; i++) {
    const int32_t v = (int32_t)((uint32_t)((t[i] & p.m0) + (((t[NL + i] ^ p.n1) - p.n1) & p.m1)) << p.sh);
    x[i] = (v ^ p.ng) - p.ng;
  }
}
BLS_HD void sop_fetch(SopI4* r, const SFp2* p) {
  const SopI4* q = reinterpret_cast<const SopI4*>(p->w);
#pragma unroll
  for (int i = 0; i < 7; i++) r[i] = q[i];
}
BLS_HD int sop_term_xi(uint32_t fl, int lane_k) {
  return (int)((fl & SOP_XI) != 0) | ((int)((fl & SOP_XI_LT2) != 0) & (int)(lane_k < 2)) | ((int)((fl & SOP_XI_LT3) != 0) & (int)(lane_k < 3));
}

// dst may alias any operand: results are written after the last operand read.
// dst_line: the result goes to that line record (expanded form) instead of *dst
BLS_FN void sop2f(SFp2* dst, SLineRec* dst_line, const SopTerm* t, int nt, int fp_mode, const SopSpaces& cx_) {
  const SopSpaces cx = cx_;  // into registers once: the caller's copy lives in local memory
  const int lane_k = cx.k;
  uint64_t T[2 * NL];
  SopKeep keep;
  int32_t res[2 * NL], x[NL], y[NL];
  SopI4 ra[7], rb[7];
#if defined(BLS_TRACK)
  double col = 0, vsum = 0;
  BLS_REQ(nt >= 1 && nt <= SOP_MAX_TERMS, "sop2f term count");
  for (int k = 0; k < nt; k++) {
    const SopTerm m = t[k];
    const SFp2 *pa = sop_rec(cx, m.a), *pb = sop_rec(cx, m.b);
    const double xi = sop_term_xi(m.fl, lane_k) ? 2.0 : 1.0;
    const double la = pa->lb * (double)(1 << m.sha) * xi, lb = pb->lb * (double)(1 << m.shb);
    BLS_REQ(la < 1073741824.0 && lb < 1073741824.0 && m.sha < 4 && m.shb < 4, "sop2f operand limb overflow");
    col += 14.0 * 2.0 * la * lb;
    vsum += 2.0 * pa->vb * (double)(1 << m.sha) * xi * pb->vb * (double)(1 << m.shb);
  }
  {
    // true column values + the reduction's own growth (14 * 2^56 for m*p, < 2^36 of carries) must fit int64
    const double lim = 9223372036854775808.0 - 15.0 * 72057594037927936.0;
    BLS_REQ(col < lim, "sop2f column overflow");
    BLS_REQ(vsum / 2500.0 + 1.0 < 16.0, "sop2f result value bound");
  }
#endif
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
  // prologue: operands of step 0
  {
    const SopTerm m = t[0];
    sop_fetch(ra, sop_rec(cx, m.a));
    sop_fetch(rb, sop_rec(cx, m.b));
    sop_prep_apply(x, ra, sop_prep(0, sop_term_xi(m.fl, lane_k), (m.fl & SOP_NEG) != 0, m.sha));
    sop_prep_apply(y, rb, sop_prep(0, 0, 0, m.shb));
  }
  int pass = 0, k = 0;
  const int npass = fp_mode ? 2 : 3;
  const int nsteps = npass * nt;
#pragma unroll 1
  for (int s = 0; s < nsteps; s++) {
    const int wrap = (k + 1 == nt);
    const int k1 = wrap ? 0 : k + 1, p1 = pass + wrap;
    const int pn = p1 >= npass ? npass - 1 : p1;  // after the last step: a harmless refetch
    const SopTerm m = t[k1];
    sop_fetch(ra, sop_rec(cx, m.a));
    sop_fetch(rb, sop_rec(cx, m.b));
    sop_acc(T, x, y);
    sop_prep_apply(x, ra, sop_prep(pn, sop_term_xi(m.fl, lane_k), (m.fl & SOP_NEG) != 0, m.sha));
    sop_prep_apply(y, rb, sop_prep(fp_mode ? 0 : pn, 0, 0, m.shb));
    if (wrap) {  // a product sum is complete
      if (pass == 0 && !fp_mode) {
#pragma unroll
        for (int i = 0; i < NL; i++) {
          sop_keep_st(keep, i, T[2 * i], T[2 * i + 1]);
          T[2 * i] = T[2 * i + 1] = 0;
        }
      } else {
        // Karatsuba: pass 1: U = P0 - P1, keep <- P0 + P1 ;  pass 2: U = P2 - keep.   Fp scalars: U = T as it is.
        if (!fp_mode) {
#pragma unroll
          for (int i = 0; i < NL; i++) {
            uint64_t u0, u1;
            sop_keep_ld(keep, i, u0, u1);
            const uint64_t v0 = T[2 * i], v1 = T[2 * i + 1];
            if (pass == 1) sop_keep_st(keep, i, u0 + v0, u1 + v1);
            T[2 * i] = pass == 1 ? u0 - v0 : v0 - u0;
            T[2 * i + 1] = pass == 1 ? u1 - v1 : v1 - u1;
          }
        }
        T[2 * NL - 1] = 0;
        int32_t c[NL];
        sop_redc(c, T);
        const bool real_part = fp_mode ? pass == 0 : pass == 1;
#pragma unroll
        for (int i = 0; i < NL; i++) {
          if (real_part) res[i] = c[i]; else res[NL + i] = c[i];
        }
#pragma unroll
        for (int i = 0; i < 2 * NL; i++) T[i] = 0;
      }
    }
    k = k1;
    pass = p1;
  }
  if (dst_line != nullptr) {
    SopI4* q = reinterpret_cast<SopI4*>(dst_line->w);
#pragma unroll
    for (int h = 0; h < 3; h++) {
#pragma unroll
      for (int i = 0; i < 4; i++) {
        int32_t e[4];
#pragma unroll
        for (int z = 0; z < 4; z++) {
          const int j = 4 * i + z;
          e[z] = j >= NL ? 0 : h == 0 ? res[j] : h == 1 ? res[NL + j] : res[j] + res[NL + j];
        }
        SopI4 v;
        v.x = e[0];
        v.y = e[1];
        v.z = e[2];
        v.w = e[3];
        q[4 * h + i] = v;
      }
    }
#if defined(BLS_TRACK)
    STRK(*dst_line, vsum / 2500.0 + 1.0, 134217728.0);
#endif
    return;
  }
  SopI4* q = reinterpret_cast<SopI4*>(dst->w);
#pragma unroll
  for (int i = 0; i < 7; i++) {
    SopI4 v;
    v.x = res[4 * i];
    v.y = res[4 * i + 1];
    v.z = res[4 * i + 2];
    v.w = res[4 * i + 3];
    q[i] = v;
  }
#if defined(BLS_TRACK)
  STRK(*dst, vsum / 2500.0 + 1.0, 134217728.0);
#endif
}

// ---- sop1: ONE coefficient per lane (the lines kernel: two lanes per pair share one record file) -----------------------
// Lane h (0: real, 1: imaginary) computes its coefficient of sum_t A_t B_t with two integer products per term,
//   h = 0:  a0 b0 + (-a1) b1        h = 1:  a0 b1 + a1 b0        (fp_mode: a_h b0, one product per term)
// and ONE reduction; nothing is exchanged between the two lanes and nothing waits in local memory (no Karatsuba: four
// products per term instead of three, but the two lanes of a pair need 7 records of shared memory each instead of 14,
// which is what lets the lines kernel keep two warps per scheduler busy).  Same software pipeline as sop2f.
BLS_HD SopPrep sop1_prep_a(int sel, int xi, int neg, int sh) {  // sel 0: A'0 = a0 - [xi] a1 ; sel 1: A'1 = [xi] a0 + a1
  SopPrep p;
  p.m0 = -((sel ^ 1) | xi);
  p.m1 = -(sel | xi);
  p.n1 = -(sel ^ 1);
  p.ng = -neg;
  p.sh = sh;
  return p;
}
BLS_HD SopPrep sop1_prep_b(int idx, int sh) {
  SopPrep p;
  p.m0 = -(idx ^ 1);
  p.m1 = -idx;
  p.n1 = 0;
  p.ng = 0;
  p.sh = sh;
  return p;
}
// res[14] = lane h's coefficient (balanced limbs).  Returns the value bound of the result under BLS_TRACK (else 0).
// (inlined into its single call site: the result stays in registers; at most three terms: all are read up front so that no
// product waits for a table load)
// operand preparation of a SQUARE term: lane 0 multiplies (a0 + a1) by (a0 - a1), lane 1 multiplies 2 a0 by a1
BLS_HD SopPrep sop1_prep_sqr_x(int h, int neg, int sh) {
  SopPrep p;
  p.m0 = -1;
  p.m1 = -(h ^ 1);
  p.n1 = 0;
  p.ng = -neg;
  p.sh = sh + h;
  return p;
}
BLS_HD SopPrep sop1_prep_sqr_y(int h, int sh) {
  SopPrep p;
  p.m0 = -(h ^ 1);
  p.m1 = -1;
  p.n1 = -(h ^ 1);
  p.ng = 0;
  p.sh = sh;
  return p;
}
BLS_HD double sop1_compute(int32_t* res, const SopTerm* t, int nt, int fp_mode, const SopSpaces& cx_, int h) {
  const SopSpaces cx = cx_;  // into registers once: the caller's copy lives in local memory
  const int lane_k = cx.k;
  const SopTerm tm0 = t[0], tm1 = t[nt > 1 ? 1 : 0], tm2 = t[nt > 2 ? 2 : 0];
  uint64_t T[2 * NL];
  int32_t x[NL], y[NL];
  SopI4 ra[7], rb[7];
  double vout = 0;
#if defined(BLS_TRACK)
  double col = 0, vsum = 0;
  BLS_REQ(nt >= 1 && nt <= 3, "sop1 term count");
  for (int k = 0; k < nt; k++) {
    const SopTerm m = t[k];
    const SFp2 *pa = sop_rec(cx, m.a), *pb = sop_rec(cx, m.b);
    const double xi = sop_term_xi(m.fl, lane_k) ? 2.0 : 1.0;
    const double sq = (m.fl & SOP_SQR) ? 2.0 : 1.0;  // (a0 + a1)(a0 - a1): both factors carry twice the limb / value bound
    const double la = pa->lb * (double)(1 << m.sha) * xi, lb = pb->lb * (double)(1 << m.shb);
    BLS_REQ(la * sq < 1073741824.0 && lb * sq < 1073741824.0 && m.sha < 4 && m.shb < 4, "sop1 operand limb overflow");
    BLS_REQ(!(m.fl & SOP_SQR) || (m.a == m.b && !fp_mode && xi == 1.0), "sop1 square term");
    col += 14.0 * 2.0 * sq * la * lb;
    // value: (a0 + a1)(a0 - a1) = a0^2 - a1^2 and 2 a0 a1 are the parts of the true square: no larger than a general product's
    vsum += 2.0 * pa->vb * (double)(1 << m.sha) * xi * pb->vb * (double)(1 << m.shb);
  }
  {
    const double lim = 9223372036854775808.0 - 15.0 * 72057594037927936.0;
    BLS_REQ(col < lim, "sop1 column overflow");
    BLS_REQ(vsum / 2500.0 + 1.0 < 16.0, "sop1 result value bound");
  }
  vout = vsum / 2500.0 + 1.0;
#endif
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
  const int per = fp_mode ? 1 : 2;
  {
    const SopTerm m = tm0;
    sop_fetch(ra, sop_rec(cx, m.a));
    sop_fetch(rb, sop_rec(cx, m.b));
    if (m.fl & SOP_SQR) {
      sop_prep_apply(x, ra, sop1_prep_sqr_x(h, (m.fl & SOP_NEG) != 0, m.sha));
      sop_prep_apply(y, rb, sop1_prep_sqr_y(h, m.shb));
    } else {
      sop_prep_apply(x, ra, sop1_prep_a(fp_mode ? h : 0, sop_term_xi(m.fl, lane_k), (m.fl & SOP_NEG) != 0, m.sha));
      sop_prep_apply(y, rb, sop1_prep_b(fp_mode ? 0 : h, m.shb));
    }
  }
  int k = 0, q = 0;
  int nsteps = 0;
  nsteps += (tm0.fl & SOP_SQR) ? 1 : per;
  if (nt > 1) nsteps += (tm1.fl & SOP_SQR) ? 1 : per;
  if (nt > 2) nsteps += (tm2.fl & SOP_SQR) ? 1 : per;
#pragma unroll 1
  for (int s = 0; s < nsteps; s++) {
    const SopTerm cur = k == 0 ? tm0 : k == 1 ? tm1 : tm2;
    const int perk = (cur.fl & SOP_SQR) ? 1 : per;
    const int qw = (q + 1 == perk);
    const int q1 = qw ? 0 : q + 1;
    const int k1 = qw ? (k + 1 == nt ? 0 : k + 1) : k;  // after the last step: a harmless refetch of term 0
    const SopTerm m = k1 == 0 ? tm0 : k1 == 1 ? tm1 : tm2;
    sop_fetch(ra, sop_rec(cx, m.a));
    sop_fetch(rb, sop_rec(cx, m.b));
    sop_acc(T, x, y);
    const int sqr = (m.fl & SOP_SQR) != 0;
    const int sel = fp_mode ? h : q1;
    const int neg = (int)((m.fl & SOP_NEG) != 0) ^ ((fp_mode | sqr) ? 0 : (q1 & (h ^ 1)));
    const SopPrep pa = sqr ? sop1_prep_sqr_x(h, neg, m.sha) : sop1_prep_a(sel, sop_term_xi(m.fl, lane_k), neg, m.sha);
    const SopPrep pb = sqr ? sop1_prep_sqr_y(h, m.shb) : sop1_prep_b(fp_mode ? 0 : (q1 ^ h), m.shb);
    sop_prep_apply(x, ra, pa);
    sop_prep_apply(y, rb, pb);
    k = k1;
    q = q1;
  }
  T[2 * NL - 1] = 0;
  sop_redc(res, T);
  return vout;
}
// lane h's half of a compact record (64-bit stores: the imaginary half starts at byte 56)
BLS_HD void sop1_store_rec(SFp2* dst, const int32_t* res, int h, double vb) {
  int2* q = reinterpret_cast<int2*>(dst->w + NL * h);
#pragma unroll
  for (int i = 0; i < 7; i++) q[i] = make_int2(res[2 * i], res[2 * i + 1]);
  (void)vb;
  STRK(*dst, vb, 134217728.0);
}
// lane h's run of an expanded line record; lane 0 also writes the sum run (`other` = the partner lane's coefficient)
BLS_HD void sop1_store_line(SLineRec* dst, const int32_t* res, const int32_t* other, int h, double vb) {
  SopI4* q = reinterpret_cast<SopI4*>(dst->w);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    int32_t e[4], f[4];
#pragma unroll
    for (int z = 0; z < 4; z++) {
      const int j = 4 * i + z;
      e[z] = j >= NL ? 0 : res[j];
      f[z] = j >= NL ? 0 : res[j] + other[j];
    }
    SopI4 v, w;
    v.x = e[0]; v.y = e[1]; v.z = e[2]; v.w = e[3];
    w.x = f[0]; w.y = f[1]; w.z = f[2]; w.w = f[3];
    q[4 * h + i] = v;
    if (h == 0) q[8 + i] = w;
  }
  (void)vb;
  STRK(*dst, vb, 134217728.0);
}
// lane h's half of  sx * [xi] x + sy * y + sz * z  (y, z optional), limbs renormalised
BLS_HD double sfp2_lin_half(int32_t* res, const SFp2* x, int32_t sx, uint32_t flx, const SFp2* y, int32_t sy, const SFp2* z, int32_t sz, int h) {
  int64_t v[NL];
  double vb = 0;
  {
    int32_t t[2 * NL];
    sop_ld28(t, x);
#pragma unroll
    for (int i = 0; i < NL; i++) {
      const int32_t plain = t[NL * h + i];
      const int32_t mixed = h ? t[i] + t[NL + i] : t[i] - t[NL + i];  // limbs < 2^30 (checked): no overflow
      v[i] = (int64_t)sx * (int64_t)((flx & SOP_XI) ? mixed : plain);
    }
#if defined(BLS_TRACK)
    vb = (sx < 0 ? -(double)sx : (double)sx) * x->vb * ((flx & SOP_XI) ? 2.0 : 1.0);
    BLS_REQ(x->lb < 1073741824.0, "sfp2_lin_half limb");
#endif
  }
  if (y != nullptr) {
    int32_t t[2 * NL];
    sop_ld28(t, y);
#pragma unroll
    for (int i = 0; i < NL; i++) v[i] += (int64_t)sy * (int64_t)t[NL * h + i];
#if defined(BLS_TRACK)
    vb += (sy < 0 ? -(double)sy : (double)sy) * y->vb;
#endif
  }
  if (z != nullptr) {
    int32_t t[2 * NL];
    sop_ld28(t, z);
#pragma unroll
    for (int i = 0; i < NL; i++) v[i] += (int64_t)sz * (int64_t)t[NL * h + i];
#if defined(BLS_TRACK)
    vb += (sz < 0 ? -(double)sz : (double)sz) * z->vb;
#endif
  }
  int64_t c = 0;
#pragma unroll
  for (int j = 0; j < NL - 1; j++) {
    c += v[j];
    const int64_t u = c + (1 << 27);
    res[j] = (int32_t)((uint32_t)u & M28) - (1 << 27);
    c = u >> 28;
  }
  res[NL - 1] = (int32_t)(c + v[NL - 1]);
#if defined(BLS_TRACK)
  BLS_REQ(vb < 1000.0, "sfp2_lin_half value bound");
#endif
  return vb;
}

// ---- sopw: the same unit over EXPANDED records (the accumulator kernel) -------------------------------------------------
// A operands: accumulator coefficients (SAccRec), absolute (SOPX_F) or relative to the lane (SOPX_FREL); B operands:
// accumulator coefficients or the incoming line (SOPX_JL, SLineRec).  Terms use sha and the xi flags only.
// Per integer product the loop body is: 8 loads, 14 shifts, 196 IMAD.WIDE.
BLS_HD const int32_t* sopw_run_a(const SAccRec* F, int lane_k, uint32_t i, int pass, int xi) {
  int jr = lane_k - ((int)i - SOPX_FREL);
  jr += jr < 0 ? 6 : 0;
  const int j = i >= SOPX_FREL ? jr : (int)i - SOPX_F;
  const int run = xi ? (pass == 0 ? 3 : pass == 1 ? 2 : 0) : pass;
  return F[j].w + 16 * run;
}
BLS_HD const int32_t* sopw_run_b(const SAccRec* F, const SLineRec* jl, uint32_t i, int pass) {
  const int is_jl = (i >= SOPX_JL) & (i < SOPX_FREL);
  const int32_t* base = is_jl ? jl[is_jl ? (int)i - SOPX_JL : 0].w : F[is_jl ? 0 : (int)i - SOPX_F].w;
  return base + 16 * pass;
}
BLS_HD void sopw_fetch(SopI4* r, const int32_t* p) {
  const SopI4* q = reinterpret_cast<const SopI4*>(p);
#pragma unroll
  for (int i = 0; i < 4; i++) r[i] = q[i];
}
BLS_HD void sopw_take(int32_t* x, const SopI4* r, int sh) {
  int32_t t[16];
#pragma unroll
  for (int i = 0; i < 4; i++) {
    t[4 * i] = r[i].x;
    t[4 * i + 1] = r[i].y;
    t[4 * i + 2] = r[i].z;
    t[4 * i + 3] = r[i].w;
  }
#pragma unroll
  for (int i = 0; i < NL; i++) x[i] = (int32_t)((uint32_t)t[i] << sh);
}
// sync_mask: the lanes that share F (device: a barrier separates the last operand read from the stores, so dst may be the
// lane's own coefficient of F - the accumulator is updated in place, one buffer)
BLS_FN void sopw(SAccRec* dst, const SopTerm* t, int nt, const SAccRec* F, const SLineRec* jl, int lane_k, unsigned sync_mask) {
  uint64_t T[2 * NL];
  SopKeep keep;
  int32_t res[2 * NL], x[NL], y[NL];
  SopI4 ra[4], rb[4];
#if defined(BLS_TRACK)
  double col = 0, vsum = 0;
  BLS_REQ(nt >= 3 && nt <= 4, "sopw term count");
  for (int k = 0; k < nt; k++) {
    const SopTerm m = t[k];
    const double xi = sop_term_xi(m.fl, lane_k) ? 2.0 : 1.0;
    int jr = lane_k - ((int)m.a - SOPX_FREL);
    jr += jr < 0 ? 6 : 0;
    const SAccRec& A = F[m.a >= SOPX_FREL ? jr : (int)m.a - SOPX_F];
    const bool bj = m.b >= SOPX_JL && m.b < SOPX_FREL;
    const double lbb = bj ? jl[m.b - SOPX_JL].lb : F[m.b - SOPX_F].lb, vbb = bj ? jl[m.b - SOPX_JL].vb : F[m.b - SOPX_F].vb;
    const double la = A.lb * (double)(1 << m.sha) * xi;
    BLS_REQ(la < 1073741824.0 && lbb < 1073741824.0 && m.sha < 4 && m.shb == 0 && !(m.fl & SOP_NEG), "sopw operand");
    col += 14.0 * 2.0 * la * lbb;
    vsum += 2.0 * A.vb * (double)(1 << m.sha) * xi * vbb;
  }
  {
    const double lim = 9223372036854775808.0 - 15.0 * 72057594037927936.0;
    BLS_REQ(col < lim, "sopw column overflow");
    BLS_REQ(vsum / 2500.0 + 1.0 < 16.0, "sopw result value bound");
  }
#endif
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) T[i] = 0;
  // No software pipeline here: the operand runs are short (8 loads from shared memory / L1) and the kernel runs three
  // blocks per SM, so other warps' multiplies cover them; the straight form needs no register rotation at all.
#pragma unroll 1
  for (int pass = 0; pass < 3; pass++) {
#pragma unroll 1
    for (int k = 0; k < nt; k++) {
      const SopTerm m = t[k];
      const int xi = sop_term_xi(m.fl, lane_k);
      sopw_fetch(ra, sopw_run_a(F, lane_k, m.a, pass, xi));
      sopw_fetch(rb, sopw_run_b(F, jl, m.b, pass));
      sopw_take(x, ra, m.sha + (xi & (pass == 2)));
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
      // pass 1: U = P0 - P1, keep <- P0 + P1 ;  pass 2: U = P2 - keep
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
#if defined(__CUDA_ARCH__)
  __syncwarp(sync_mask);
#else
  (void)sync_mask;
#endif
  SopI4* q = reinterpret_cast<SopI4*>(dst->w);
#pragma unroll
  for (int h = 0; h < 4; h++) {
#pragma unroll
    for (int i = 0; i < 4; i++) {
      int32_t e[4];
#pragma unroll
      for (int z = 0; z < 4; z++) {
        const int j = 4 * i + z;
        e[z] = j >= NL ? 0 : h == 0 ? res[j] : h == 1 ? res[NL + j] : h == 2 ? res[j] + res[NL + j] : res[j] - res[NL + j];
      }
      SopI4 v;
      v.x = e[0];
      v.y = e[1];
      v.z = e[2];
      v.w = e[3];
      q[4 * h + i] = v;
    }
  }
#if defined(BLS_TRACK)
  STRK(*dst, vsum / 2500.0 + 1.0, 134217728.0);
#endif
#if defined(__CUDA_ARCH__)
  __syncwarp(sync_mask);
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
// expanded accumulator record from a compact one (initial values) and back (the halves are its first two runs)
BLS_HD void sacc_from_sfp2(SAccRec& r, const SFp2& a) {
#pragma unroll
  for (int i = 0; i < 16; i++) {
    const int32_t a0 = i < NL ? a.w[i] : 0, a1 = i < NL ? a.w[NL + i] : 0;
    r.w[i] = a0;
    r.w[16 + i] = a1;
    r.w[32 + i] = a0 + a1;
    r.w[48 + i] = a0 - a1;
  }
  STRK(r, a.vb, a.lb);
}
BLS_HD void sfp2_from_sacc(SFp2& r, const SAccRec& a) {
#pragma unroll
  for (int i = 0; i < NL; i++) {
    r.w[i] = a.w[i];
    r.w[NL + i] = a.w[16 + i];
  }
  STRK(r, a.vb, a.lb);
}
BLS_HD void sfp2_neg(SFp2& r, const SFp2& a) {
#pragma unroll
  for (int i = 0; i < 2 * NL; i++) r.w[i] = -a.w[i];
  STRK(r, a.vb, a.lb);
}

}  // namespace bls
