// Fp6 = Fp2[v]/(v^3 - xi), Fp12 = Fp6[w]/(w^2 - v), xi = 1+u.  As a degree-6 extension of Fp2 in w:
// (a0 + a1 v + a2 v^2) + (b0 + b1 v + b2 v^2) w  =  a0 + b0 w + a1 w^2 + b1 w^3 + a2 w^4 + b2 w^5.
// Gt arithmetic behind `multi_miller_loop(..).final_exponentiation()` (reference src/helpers.rs:50,62).
#pragma once
#include "fp2.cuh"

namespace bls {

struct Fp6 {
  Fp2 c0, c1, c2;
};
struct Fp12 {
  Fp6 c0, c1;
};

BLS_HD void fp6_add(Fp6& r, const Fp6& a, const Fp6& b) {
  fadd(r.c0, a.c0, b.c0);
  fadd(r.c1, a.c1, b.c1);
  fadd(r.c2, a.c2, b.c2);
}
BLS_HD void fp6_sub(Fp6& r, const Fp6& a, const Fp6& b) {
  fsub(r.c0, a.c0, b.c0);
  fsub(r.c1, a.c1, b.c1);
  fsub(r.c2, a.c2, b.c2);
}
BLS_HD void fp6_neg(Fp6& r, const Fp6& a) {
  fneg(r.c0, a.c0);
  fneg(r.c1, a.c1);
  fneg(r.c2, a.c2);
}
BLS_HD void fp6_mul_by_v(Fp6& r, const Fp6& a) {
  Fp2 t;
  fp2_mul_xi(t, a.c2);
  r.c2 = a.c1;
  r.c1 = a.c0;
  r.c0 = t;
}
BLS_HD bool fp6_is_zero(const Fp6& a) { return fis_zero(a.c0) && fis_zero(a.c1) && fis_zero(a.c2); }

BLS_FN void fp6_mul(Fp6& r, const Fp6& a, const Fp6& b) {
  Fp2 t0, t1, t2, s, u, x;
  fp2_mul(t0, a.c0, b.c0);
  fp2_mul(t1, a.c1, b.c1);
  fp2_mul(t2, a.c2, b.c2);
  // c0 = t0 + xi((a1+a2)(b1+b2) - t1 - t2)
  fadd(s, a.c1, a.c2);
  fadd(u, b.c1, b.c2);
  fp2_mul(x, s, u);
  fsub(x, x, t1);
  fsub(x, x, t2);
  fp2_mul_xi(x, x);
  Fp2 c0;
  fadd(c0, x, t0);
  // c1 = (a0+a1)(b0+b1) - t0 - t1 + xi t2
  fadd(s, a.c0, a.c1);
  fadd(u, b.c0, b.c1);
  fp2_mul(x, s, u);
  fsub(x, x, t0);
  fsub(x, x, t1);
  Fp2 c1;
  fp2_mul_xi(s, t2);
  fadd(c1, x, s);
  // c2 = (a0+a2)(b0+b2) - t0 - t2 + t1
  fadd(s, a.c0, a.c2);
  fadd(u, b.c0, b.c2);
  fp2_mul(x, s, u);
  fsub(x, x, t0);
  fsub(x, x, t2);
  fadd(r.c2, x, t1);
  r.c0 = c0;
  r.c1 = c1;
}

BLS_FN void fp6_sqr(Fp6& r, const Fp6& a) {
  Fp2 s0, s1, s2, s3, s4, t;
  fp2_sqr(s0, a.c0);
  fp2_mul(s1, a.c0, a.c1);
  fdbl(s1, s1);
  fsub(t, a.c0, a.c1);
  fadd(t, t, a.c2);
  fp2_sqr(s2, t);
  fp2_mul(s3, a.c1, a.c2);
  fdbl(s3, s3);
  fp2_sqr(s4, a.c2);
  fp2_mul_xi(t, s3);
  fadd(r.c0, s0, t);
  fp2_mul_xi(t, s4);
  fadd(r.c1, s1, t);
  fadd(t, s1, s2);
  fadd(t, t, s3);
  fsub(t, t, s0);
  fsub(r.c2, t, s4);
}

// a * (b0 + b1 v)
BLS_FN void fp6_mul_by_01(Fp6& r, const Fp6& a, const Fp2& b0, const Fp2& b1) {
  Fp2 t0, t1, s, u, x, y;
  fp2_mul(t0, a.c0, b0);
  fp2_mul(t1, a.c1, b1);
  fadd(s, a.c0, a.c1);
  fadd(u, b0, b1);
  fp2_mul(x, s, u);
  fsub(x, x, t0);
  fsub(x, x, t1);  // c1
  fp2_mul(y, a.c2, b1);
  fp2_mul_xi(y, y);
  fp2_mul(s, a.c2, b0);
  fadd(r.c0, t0, y);
  r.c1 = x;
  fadd(r.c2, t1, s);
}
// a * (b1 v)
BLS_FN void fp6_mul_by_1(Fp6& r, const Fp6& a, const Fp2& b1) {
  Fp2 t0, t1, t2;
  fp2_mul(t2, a.c2, b1);
  fp2_mul(t0, a.c0, b1);
  fp2_mul(t1, a.c1, b1);
  fp2_mul_xi(r.c0, t2);
  r.c1 = t0;
  r.c2 = t1;
}

BLS_FN void fp6_inv(Fp6& r, const Fp6& a) {
  Fp2 t0, t1, t2, x, d;
  fp2_sqr(t0, a.c0);
  fp2_mul(x, a.c1, a.c2);
  fp2_mul_xi(x, x);
  fsub(t0, t0, x);  // a0^2 - xi a1 a2
  fp2_sqr(t1, a.c2);
  fp2_mul_xi(t1, t1);
  fp2_mul(x, a.c0, a.c1);
  fsub(t1, t1, x);  // xi a2^2 - a0 a1
  fp2_sqr(t2, a.c1);
  fp2_mul(x, a.c0, a.c2);
  fsub(t2, t2, x);  // a1^2 - a0 a2
  fp2_mul(d, a.c2, t1);
  fp2_mul(x, a.c1, t2);
  fadd(d, d, x);
  fp2_mul_xi(d, d);
  fp2_mul(x, a.c0, t0);
  fadd(d, d, x);
  fp2_inv(d, d);
  fp2_mul(r.c0, t0, d);
  fp2_mul(r.c1, t1, d);
  fp2_mul(r.c2, t2, d);
}

// ---------------------------------------------------------------------------------------------------------
BLS_HD void fp12_one(Fp12& r) {
  fone(r.c0.c0);
  fzero(r.c0.c1);
  fzero(r.c0.c2);
  fzero(r.c1.c0);
  fzero(r.c1.c1);
  fzero(r.c1.c2);
}
BLS_HD bool fp12_is_one(const Fp12& a) {
  Fp2 one;
  fone(one);
  return feq(a.c0.c0, one) && fis_zero(a.c0.c1) && fis_zero(a.c0.c2) && fp6_is_zero(a.c1);
}
BLS_HD bool fp12_eq(const Fp12& a, const Fp12& b) {
  return feq(a.c0.c0, b.c0.c0) && feq(a.c0.c1, b.c0.c1) && feq(a.c0.c2, b.c0.c2) && feq(a.c1.c0, b.c1.c0) &&
         feq(a.c1.c1, b.c1.c1) && feq(a.c1.c2, b.c1.c2);
}
BLS_HD void fp12_conj(Fp12& r, const Fp12& a) {
  r.c0 = a.c0;
  fp6_neg(r.c1, a.c1);
}

BLS_FN void fp12_mul(Fp12& r, const Fp12& a, const Fp12& b) {
  Fp6 t0, t1, s, u, x;
  fp6_mul(t0, a.c0, b.c0);
  fp6_mul(t1, a.c1, b.c1);
  fp6_add(s, a.c0, a.c1);
  fp6_add(u, b.c0, b.c1);
  fp6_mul(x, s, u);
  fp6_sub(x, x, t0);
  fp6_sub(r.c1, x, t1);
  fp6_mul_by_v(t1, t1);
  fp6_add(r.c0, t0, t1);
}

BLS_FN void fp12_sqr(Fp12& r, const Fp12& a) {
  Fp6 t, s, u, x;
  fp6_mul(t, a.c0, a.c1);
  fp6_add(s, a.c0, a.c1);
  fp6_mul_by_v(u, a.c1);
  fp6_add(u, u, a.c0);
  fp6_mul(x, s, u);
  fp6_sub(x, x, t);
  fp6_mul_by_v(u, t);
  fp6_sub(r.c0, x, u);
  fp6_add(r.c1, t, t);
}

// f * ((c0 + c1 v) + (c4 v) w): the sparse shape of a Miller-loop line (w^0, w^2, w^3 coefficients)
BLS_FN void fp12_mul_by_014(Fp12& f, const Fp2& c0, const Fp2& c1, const Fp2& c4) {
  Fp6 aa, bb, s, x;
  Fp2 c14;
  fp6_mul_by_01(aa, f.c0, c0, c1);
  fp6_mul_by_1(bb, f.c1, c4);
  fadd(c14, c1, c4);
  fp6_add(s, f.c0, f.c1);
  fp6_mul_by_01(x, s, c0, c14);
  fp6_sub(x, x, aa);
  fp6_sub(f.c1, x, bb);
  fp6_mul_by_v(bb, bb);
  fp6_add(f.c0, aa, bb);
}

BLS_FN void fp12_inv(Fp12& r, const Fp12& a) {
  Fp6 t0, t1;
  fp6_sqr(t0, a.c0);
  fp6_sqr(t1, a.c1);
  fp6_mul_by_v(t1, t1);
  fp6_sub(t0, t0, t1);
  fp6_inv(t0, t0);
  fp6_mul(r.c0, a.c0, t0);
  fp6_mul(t1, a.c1, t0);
  fp6_neg(r.c1, t1);
}

// f^p : coefficient of w^i -> conj(c_i) * K_FROB1[i]
BLS_FN void fp12_frob1(Fp12& r, const Fp12& a) {
  Fp2 g, t;
  fp2_conj(r.c0.c0, a.c0.c0);
  fp2_conj(t, a.c1.c0);
  fp2_set(g, K_FROB1[1]);
  fp2_mul(r.c1.c0, t, g);
  fp2_conj(t, a.c0.c1);
  fp2_set(g, K_FROB1[2]);
  fp2_mul(r.c0.c1, t, g);
  fp2_conj(t, a.c1.c1);
  fp2_set(g, K_FROB1[3]);
  fp2_mul(r.c1.c1, t, g);
  fp2_conj(t, a.c0.c2);
  fp2_set(g, K_FROB1[4]);
  fp2_mul(r.c0.c2, t, g);
  fp2_conj(t, a.c1.c2);
  fp2_set(g, K_FROB1[5]);
  fp2_mul(r.c1.c2, t, g);
}
// f^(p^2) : coefficient of w^i -> c_i * K_FROB2[i]  (Fp scalars)
BLS_FN void fp12_frob2(Fp12& r, const Fp12& a) {
  Fp g;
  r.c0.c0 = a.c0.c0;
  fp_set(g, K_FROB2[1]);
  fp2_mul_fp(r.c1.c0, a.c1.c0, g);
  fp_set(g, K_FROB2[2]);
  fp2_mul_fp(r.c0.c1, a.c0.c1, g);
  fp_set(g, K_FROB2[3]);
  fp2_mul_fp(r.c1.c1, a.c1.c1, g);
  fp_set(g, K_FROB2[4]);
  fp2_mul_fp(r.c0.c2, a.c0.c2, g);
  fp_set(g, K_FROB2[5]);
  fp2_mul_fp(r.c1.c2, a.c1.c2, g);
}

// (a + b s)^2 in Fp4 = Fp2[s]/(s^2 - xi): c0 = a^2 + xi b^2, c1 = 2ab
BLS_HD void fp4_sqr(Fp2& c0, Fp2& c1, const Fp2& a, const Fp2& b) {
  Fp2 t0, t1, t2;
  fp2_sqr(t0, a);
  fp2_sqr(t1, b);
  fadd(t2, a, b);
  fp2_sqr(t2, t2);
  fsub(t2, t2, t0);
  fsub(c1, t2, t1);
  fp2_mul_xi(t1, t1);
  fadd(c0, t0, t1);
}

// Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part of the final exponentiation)
BLS_FN void fp12_cyclo_sqr(Fp12& r, const Fp12& a) {
  Fp2 z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2;
  Fp2 t0, t1, t2, t3, x;
  fp4_sqr(t0, t1, z0, z1);
  // z0 = 3 t0 - 2 z0 ; z1 = 3 t1 + 2 z1
  fsub(x, t0, z0);
  fdbl(x, x);
  fadd(z0, x, t0);
  fadd(x, t1, z1);
  fdbl(x, x);
  fadd(z1, x, t1);
  fp4_sqr(t0, t1, z2, z3);
  fp4_sqr(t2, t3, z4, z5);
  // z4 = 3 t0 - 2 z4 ; z5 = 3 t1 + 2 z5
  fsub(x, t0, z4);
  fdbl(x, x);
  fadd(z4, x, t0);
  fadd(x, t1, z5);
  fdbl(x, x);
  fadd(z5, x, t1);
  // z2 = 3 xi t3 + 2 z2 ; z3 = 3 t2 - 2 z3
  fp2_mul_xi(t0, t3);
  fadd(x, t0, z2);
  fdbl(x, x);
  fadd(z2, x, t0);
  fsub(x, t2, z3);
  fdbl(x, x);
  fadd(z3, x, t2);
  r.c0.c0 = z0;
  r.c0.c1 = z4;
  r.c0.c2 = z3;
  r.c1.c0 = z2;
  r.c1.c1 = z1;
  r.c1.c2 = z5;
}

// a^|x| for a in the cyclotomic subgroup, |x| = 0xd201000000010000
BLS_FN void fp12_cyclo_pow_xabs(Fp12& r, const Fp12& a) {
  Fp12 acc = a;
  const uint64_t e = K_X_ABS;
  for (int i = 62; i >= 0; i--) {
    fp12_cyclo_sqr(acc, acc);
    if ((e >> i) & 1) fp12_mul(acc, acc, a);
  }
  r = acc;
}
// a^x with x = -|x| (inverse = conjugate in the cyclotomic subgroup)
BLS_HD void fp12_cyclo_pow_x(Fp12& r, const Fp12& a) {
  Fp12 t;
  fp12_cyclo_pow_xabs(t, a);
  fp12_conj(r, t);
}

// Final exponentiation to the power 3*(p^12-1)/r (the cube of the canonical pairing value; 3 does not divide r, so
// "== 1" is unchanged - and that is all the reference consumes: Gt::is_identity, sig_core.rs:138-140,173).
//   easy: f^((p^6-1)(p^2+1));  hard: 3(p^4-p^2+1)/r = (x-1)^2 (x+p)(x^2+p^2-1) + 3
BLS_FN void final_exponentiation(Fp12& r, const Fp12& f) {
  Fp12 t0, t1, m, a, b, c;
  fp12_conj(t0, f);
  fp12_inv(t1, f);
  fp12_mul(t0, t0, t1);  // f^(p^6-1)
  fp12_frob2(t1, t0);
  fp12_mul(m, t1, t0);  // ^(p^2+1)  -> cyclotomic
  // a = m^((x-1)^2)
  fp12_cyclo_pow_x(t0, m);
  fp12_conj(t1, m);
  fp12_mul(t0, t0, t1);  // m^(x-1)
  fp12_cyclo_pow_x(t1, t0);
  fp12_conj(a, t0);
  fp12_mul(a, a, t1);  // (m^(x-1))^(x-1)
  // b = a^(x+p)
  fp12_cyclo_pow_x(t0, a);
  fp12_frob1(t1, a);
  fp12_mul(b, t0, t1);
  // c = b^(x^2+p^2-1)
  fp12_cyclo_pow_x(t0, b);
  fp12_cyclo_pow_x(t0, t0);
  fp12_frob2(t1, b);
  fp12_mul(t0, t0, t1);
  fp12_conj(t1, b);
  fp12_mul(c, t0, t1);
  // * m^3
  fp12_cyclo_sqr(t0, m);
  fp12_mul(t0, t0, m);
  fp12_mul(r, c, t0);
}

}  // namespace bls
