// Cooperative Miller loop: SIX pairings share one Fp12 accumulator, and the accumulator is spread over SIX lanes.
//
//   lane k of a group owns   (a) its own pair (P_k, Q_k): the running G2 point and the line evaluation, and
//                            (b) coefficient k of  f = sum_k f_k w^k   (Fp12 as Fp2[w]/(w^6 - xi)).
//   per iteration:  f <- f^2 ; every lane evaluates the line of ITS pair ; f <- f * l_0 * l_1 * ... * l_5.
//
// Why this shape on sm_100a (measured, DESIGN.md sections 4-6): the one-thread-per-pairing kernel kept 2.5 KB of Fp12
// temporaries per thread in local memory (47.8 GB of DRAM writes per 65,536 pairings) and spent a third of its
// multiplications squaring an accumulator per pair.  Here the accumulator is one Fp2 per lane, exchanged through
// shared memory (conflict-free 128-bit accesses: 112-byte records), every coefficient of f^2 and f*l is ONE call of the
// fused sum-of-products unit (sfp.cuh: 9-12 integer products, 2 reductions, no temporaries), and the squaring is paid
// once per six pairings - the product of Miller values is all the batch equation needs.
// Replaces `multi_miller_loop` (reference src/helpers.rs:50,62).
#pragma once
#include "pairing.cuh"
#include "sfp.cuh"

namespace bls {

constexpr int M6_GROUP = 6;

// f^2, coefficient k:  sum over unordered {i,j}, i+j = k (mod 6), of  s * a_i * a_j  [* xi when i+j >= 6]
// entry = ia | ib << 3 | xi << 6 | scale << 7   (scale 0: unused slot)
#define M6_E(ia, ib, xi, sc) (uint16_t)((ia) | ((ib) << 3) | ((xi) << 6) | ((sc) << 7))
BLS_CONST uint16_t K_M6_SQR[6][4] = {
    {M6_E(0, 0, 0, 1), M6_E(1, 5, 1, 2), M6_E(2, 4, 1, 2), M6_E(3, 3, 1, 1)},
    {M6_E(0, 1, 0, 2), M6_E(2, 5, 1, 2), M6_E(3, 4, 1, 2), M6_E(0, 0, 0, 0)},
    {M6_E(0, 2, 0, 2), M6_E(1, 1, 0, 1), M6_E(3, 5, 1, 2), M6_E(4, 4, 1, 1)},
    {M6_E(0, 3, 0, 2), M6_E(1, 2, 0, 2), M6_E(4, 5, 1, 2), M6_E(0, 0, 0, 0)},
    {M6_E(0, 4, 0, 2), M6_E(1, 3, 0, 2), M6_E(2, 2, 0, 1), M6_E(5, 5, 1, 1)},
    {M6_E(0, 5, 0, 2), M6_E(1, 4, 0, 2), M6_E(2, 3, 0, 2), M6_E(0, 0, 0, 0)},
};
#undef M6_E

// F: the group's six coefficients (w^0..w^5).  out = coefficient k of f^2.
BLS_HD void m6_sqr_lane(SFp2& out, const SFp2* F, int k) {
  SopT t[4];
#pragma unroll
  for (int s = 0; s < 4; s++) {
    const uint32_t e = K_M6_SQR[k][s];
    t[s] = sop_t(&F[e & 7u], &F[(e >> 3) & 7u], (int32_t)((e >> 7) & 3u), ((e >> 6) & 1u) ? SOP_XI : 0u);
  }
  sop2s(out, t, 4);
}

// out = coefficient k of  f * (l0 + l2 w^2 + l3 w^3)   (line[0..2] = l0, l2, l3: the sparse shape of a Miller line)
BLS_HD void m6_mul_line_lane(SFp2& out, const SFp2* F, const SFp2* line, int k) {
  SopT t[3];
  const int k2 = k >= 2 ? k - 2 : k + 4, k3 = k >= 3 ? k - 3 : k + 3;
  t[0] = sop_t(&F[k], &line[0]);
  t[1] = sop_t(&F[k2], &line[1], 1, k < 2 ? SOP_XI : 0u);
  t[2] = sop_t(&F[k3], &line[2], 1, k < 3 ? SOP_XI : 0u);
  sop2s(out, t, 3);
}

// slot of coefficient k (of w^k) in the tower layout of fp12.cuh: w^0..w^5 = c0.c0, c1.c0, c0.c1, c1.c1, c0.c2, c1.c2
BLS_HD Fp2* fp12_coeff(Fp12& f, int k) {
  Fp6& h = (k & 1) ? f.c1 : f.c0;
  return (k >> 1) == 0 ? &h.c0 : (k >> 1) == 1 ? &h.c1 : &h.c2;
}

// lane k's share of the epilogue: conjugate (x < 0), back to the unsigned form of the tower, reduced
BLS_HD void m6_finish_lane(Fp2& out, const SFp2& fk, int k) {
  SFp2 t = fk;
  if (k & 1) sfp2_neg(t, fk);
  Fp2 u;
  fp2_from_sfp2(u, t);
  fred(out, u);
}

// ---- the line computation of one pair, as a PROGRAM over S-form records ---------------------------------------------
// Every step is one call of sop2s (a sum of at most two Fp2 products, lazily reduced) or of sfp2_lin; the steps are rows
// of a constant table run by a ten-line interpreter, so the hot code of the whole Miller loop is sop2s + sfp2_lin + this
// interpreter (instruction caches: 6 KB / 32 KB per SM sub-partition / SM) instead of 24 KB of
// inlined field glue per step function.
//
// Running point T = (X : Y : Z) in HOMOGENEOUS projective coordinates on the twist y^2 = x^3 + 4 xi; the G1 argument
// enters as three Fp scalars px = Xp Zp, py = Yp, pz = Zp^3 (affine: x, y, 1), so a Jacobian r_i * pk_i needs no inversion.
// Lines are scaled by factors in proper subfields (erased by the final exponentiation):
//   doubling:  B = Y^2, C = Z^2, J = X^2, E = 12 xi C, F = 3E
//              X3 = 2XY (B - F),  Y3 = (B + F)^2 - 12 E^2,  Z3 = 8 B YZ          (4 x the textbook (X3:Y3:Z3))
//              l0 = (B - E) pz,  l2 = -3J px,  l3 = 2YZ py
//   addition:  u = y2 Z - Y, v = x2 Z - X, A = u^2 Z - v^3 - 2 v^2 X
//              X3 = v A,  Y3 = u (v^2 X - A) - v^3 Y,  Z3 = v^3 Z
//              l0 = (u x2 - v y2) pz,  l2 = -u px,  l3 = v py
// (derivation checked against the big-int oracle in tools/proto_lines.py; bounds by the BLS_TRACK build)
enum : uint8_t {
  RX = 0, RY, RZ, RPX, RPY, RPZ, RQX, RQY, RT0, RT1, RT2, RT3, RT4, RT5, RT6,
  M6_NREG,
  RL0 = 32, RL2 = 33, RL3 = 34,  // the lane's line record (shared memory on the device)
  RNONE = 255
};
struct M6Term {
  uint8_t a, a2, b, b2;
  int8_t sa, sa2, sb, sb2;
  uint8_t fl;
};
struct M6Op {
  uint8_t kind;  // 0: sop2s, 1: sfp2_lin (t[0]: sa*[xi]a + sa2*a2 + sb*b)
  uint8_t dst, nt;
  M6Term t[2];
};
#define M6_T1(a, b) {a, RNONE, b, RNONE, 1, 0, 1, 0, 0}
#define M6_NOT {RNONE, RNONE, RNONE, RNONE, 0, 0, 0, 0, 0}
BLS_CONST M6Op K_M6_DBL[] = {
    {0, RT0, 1, {M6_T1(RY, RY), M6_NOT}},                                                         // B
    {0, RT1, 1, {M6_T1(RZ, RZ), M6_NOT}},                                                         // C
    {0, RT2, 1, {M6_T1(RX, RX), M6_NOT}},                                                         // J
    {0, RT3, 1, {M6_T1(RX, RY), M6_NOT}},                                                         // XY
    {0, RT4, 1, {M6_T1(RY, RZ), M6_NOT}},                                                         // YZ
    {1, RT5, 1, {{RT1, RNONE, RNONE, RNONE, 12, 0, 0, 0, SOP_XI}, M6_NOT}},                       // E = 12 xi C
    {1, RT6, 1, {{RT5, RNONE, RNONE, RNONE, 3, 0, 0, 0, 0}, M6_NOT}},                             // F = 3E
    {0, RX, 1, {{RT3, RNONE, RT0, RT6, 2, 0, 1, -1, 0}, M6_NOT}},                                 // X3 = 2XY (B - F)
    {0, RY, 2, {{RT0, RT6, RT0, RT6, 1, 1, 1, 1, 0}, {RT5, RNONE, RT6, RNONE, -4, 0, 1, 0, 0}}},  // Y3 = (B+F)^2 - 4E F
    {0, RZ, 1, {{RT0, RNONE, RT4, RNONE, 4, 0, 2, 0, 0}, M6_NOT}},                                // Z3 = 4B 2YZ
    {0, RL0, 1, {{RT0, RT5, RPZ, RNONE, 1, -1, 1, 0, SOP_BFP}, M6_NOT}},                          // l0 = (B - E) pz
    {0, RL2, 1, {{RT2, RNONE, RPX, RNONE, -3, 0, 1, 0, SOP_BFP}, M6_NOT}},                        // l2 = -3J px
    {0, RL3, 1, {{RT4, RNONE, RPY, RNONE, 2, 0, 1, 0, SOP_BFP}, M6_NOT}},                         // l3 = 2YZ py
};
BLS_CONST M6Op K_M6_ADD[] = {
    {0, RT0, 1, {M6_T1(RQY, RZ), M6_NOT}},                                                        // y2 Z
    {0, RT1, 1, {M6_T1(RQX, RZ), M6_NOT}},                                                        // x2 Z
    {1, RT0, 1, {{RT0, RY, RNONE, RNONE, 1, -1, 0, 0, 0}, M6_NOT}},                               // u
    {1, RT1, 1, {{RT1, RX, RNONE, RNONE, 1, -1, 0, 0, 0}, M6_NOT}},                               // v
    {0, RT2, 1, {M6_T1(RT1, RT1), M6_NOT}},                                                       // vv
    {0, RT3, 1, {M6_T1(RT1, RT2), M6_NOT}},                                                       // vvv
    {0, RT4, 1, {M6_T1(RT2, RX), M6_NOT}},                                                        // Rr = vv X
    {0, RT5, 1, {M6_T1(RT0, RT0), M6_NOT}},                                                       // uu
    {0, RT5, 1, {M6_T1(RT5, RZ), M6_NOT}},                                                        // uu Z
    {1, RT5, 1, {{RT5, RT3, RT4, RNONE, 1, -1, -2, 0, 0}, M6_NOT}},                               // A = uu Z - vvv - 2 Rr
    {0, RT6, 2, {M6_T1(RT0, RQX), {RT1, RNONE, RQY, RNONE, -1, 0, 1, 0, 0}}},                     // u x2 - v y2
    {0, RL0, 1, {{RT6, RNONE, RPZ, RNONE, 1, 0, 1, 0, SOP_BFP}, M6_NOT}},                         // l0
    {0, RL2, 1, {{RT0, RNONE, RPX, RNONE, -1, 0, 1, 0, SOP_BFP}, M6_NOT}},                        // l2 = -u px
    {0, RL3, 1, {{RT1, RNONE, RPY, RNONE, 1, 0, 1, 0, SOP_BFP}, M6_NOT}},                         // l3 = v py
    {0, RX, 1, {M6_T1(RT1, RT5), M6_NOT}},                                                        // X3 = v A
    {0, RY, 2, {{RT0, RNONE, RT4, RT5, 1, 0, 1, -1, 0}, {RT3, RNONE, RY, RNONE, -1, 0, 1, 0, 0}}},  // Y3 = u (Rr - A) - vvv Y
    {0, RZ, 1, {M6_T1(RT3, RZ), M6_NOT}},                                                         // Z3 = vvv Z
};
#undef M6_T1
#undef M6_NOT
constexpr int K_M6_DBL_N = (int)(sizeof(K_M6_DBL) / sizeof(M6Op));
constexpr int K_M6_ADD_N = (int)(sizeof(K_M6_ADD) / sizeof(M6Op));

BLS_HD SFp2* m6_rec(SFp2* reg, SFp2* line, uint32_t i) { return i == RNONE ? nullptr : i >= RL0 ? line + (i - RL0) : reg + i; }

BLS_FN void m6_run(const M6Op* prog, int n, SFp2* reg, SFp2* line) {
#pragma unroll 1
  for (int i = 0; i < n; i++) {
    const M6Op& op = prog[i];
    SFp2* dst = m6_rec(reg, line, op.dst);
    if (op.kind == 0) {
      SopT t[2];
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const M6Term& m = op.t[k];
        t[k] = sop_t2(m6_rec(reg, line, m.a), m.sa, m6_rec(reg, line, m.a2), m.sa2, m6_rec(reg, line, m.b), m.sb,
                      m6_rec(reg, line, m.b2), m.sb2, m.fl);
      }
      sop2s(*dst, t, op.nt);
    } else {
      const M6Term& m = op.t[0];
      sfp2_lin(*dst, m6_rec(reg, line, m.a), m.sa, m.fl, m6_rec(reg, line, m.a2), m.sa2, m6_rec(reg, line, m.b), m.sb);
    }
  }
}

// per-lane private state: the record file of the line programs
struct M6Pair {
  SFp2 reg[M6_NREG];
  const G2Aff* Q;  // stays where it is (HBM): only the 5 addition steps read it
};
// P prepared by miller_prepare (pairing.cuh), Q affine and not the identity
BLS_HD void m6_init_pair(M6Pair& s, const MillerG1& P, const G2Aff* Q) {
  const G2Aff q = *Q;
  s.Q = Q;
  sfp2_from_fp2(s.reg[RX], q.x);
  sfp2_from_fp2(s.reg[RY], q.y);
  sfp2_one(s.reg[RZ]);
  sfp2_from_fp(s.reg[RPX], P.px);
  sfp2_from_fp(s.reg[RPY], P.py);
  sfp2_from_fp(s.reg[RPZ], P.pz);
}
BLS_HD void m6_dbl_line(SFp2* line, M6Pair& s) { m6_run(K_M6_DBL, K_M6_DBL_N, s.reg, line); }
BLS_HD void m6_add_line(SFp2* line, M6Pair& s) {
  const G2Aff q = *s.Q;
  sfp2_from_fp2(s.reg[RQX], q.x);
  sfp2_from_fp2(s.reg[RQY], q.y);
  m6_run(K_M6_ADD, K_M6_ADD_N, s.reg, line);
}

}  // namespace bls
