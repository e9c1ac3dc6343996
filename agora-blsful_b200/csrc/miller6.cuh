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

// per-lane private state of the line computation (the existing Jacobian step functions of pairing.cuh)
struct M6Pair {
  G2Jac R;
  const G2Aff* Q;  // stays where it is (HBM): only the 5 addition steps read it
  MillerG1 P;
};
BLS_HD void m6_line_out(SFp2* line, const Fp2& c0, const Fp2& c2, const Fp2& c3) {
  sfp2_from_fp2(line[0], c0);
  sfp2_from_fp2(line[1], c2);
  sfp2_from_fp2(line[2], c3);
}
BLS_HD void m6_dbl_line(SFp2* line, M6Pair& s) {
  Fp2 c0, c2, c3;
  miller_dbl_step(c0, c2, c3, s.R, s.P);
  m6_line_out(line, c0, c2, c3);
}
BLS_HD void m6_add_line(SFp2* line, M6Pair& s) {
  Fp2 c0, c2, c3;
  const G2Aff q = *s.Q;
  miller_add_step(c0, c2, c3, s.R, q, s.P);
  m6_line_out(line, c0, c2, c3);
}

}  // namespace bls
