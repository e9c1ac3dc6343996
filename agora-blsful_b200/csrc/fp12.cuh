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

// Bound contract of the tower (fp.cuh explains vb/lb): "reduced" = output of fred: value < 3p, limbs < 2^28.
//   fp6_mul / fp6_mul_by_01 / fp6_mul_by_1 (the _lazy forms): inputs value <= 11 per Fp2 coefficient part, limbs <= 2^29+64;
//     outputs normalised (limbs <= 2^28+16), value <= 108 - NOT reduced: the Fp12 caller combines them and reduces once.
//   fp12_* : inputs reduced (or value <= 5), outputs reduced.
BLS_HD void fp6_add(Fp6& r, const Fp6& a, const Fp6& b) {
  fadd(r.c0, a.c0, b.c0);
  fadd(r.c1, a.c1, b.c1);
  fadd(r.c2, a.c2, b.c2);
}
template <int K>
BLS_HD void fp6_sub_k(Fp6& r, const Fp6& a, const Fp6& b) {
  fsub_k<K>(r.c0, a.c0, b.c0);
  fsub_k<K>(r.c1, a.c1, b.c1);
  fsub_k<K>(r.c2, a.c2, b.c2);
}
BLS_HD void fp6_norm(Fp6& r, const Fp6& a) {
  fnorm(r.c0, a.c0);
  fnorm(r.c1, a.c1);
  fnorm(r.c2, a.c2);
}
BLS_HD void fp6_red(Fp6& r, const Fp6& a) {
  fred(r.c0, a.c0);
  fred(r.c1, a.c1);
  fred(r.c2, a.c2);
}
// reduced in, reduced out
BLS_HD void fp6_neg(Fp6& r, const Fp6& a) {
  fneg_k<4>(r.c0, a.c0);
  fneg_k<4>(r.c1, a.c1);
  fneg_k<4>(r.c2, a.c2);
  fp6_red(r, r);
}
// a * v, lazily: K bounds the value of a.c2's imaginary part; result c0 not normalised
template <int K>
BLS_HD void fp6_mul_by_v_k(Fp6& r, const Fp6& a) {
  Fp2 t;
  fp2_mul_xi_k<K>(t, a.c2);
  r.c2 = a.c1;
  r.c1 = a.c0;
  r.c0 = t;
}
BLS_HD bool fp6_is_zero(const Fp6& a) { return fis_zero(a.c0) && fis_zero(a.c1) && fis_zero(a.c2); }

BLS_FN void fp6_mul_lazy(Fp6& r, const Fp6& a, const Fp6& b) {
  Fp2 t0, t1, t2, s, u, x, y;
  fp2_mul(t0, a.c0, b.c0);
  fp2_mul(t1, a.c1, b.c1);
  fp2_mul(t2, a.c2, b.c2);
  // c0 = t0 + xi((a1+a2)(b1+b2) - t1 - t2)
  fadd(s, a.c1, a.c2);
  fadd(u, b.c1, b.c2);
  fp2_mul(x, s, u);
  fsub_k<16>(x, x, t1);
  fsub_k<16>(x, x, t2);
  fnorm(x, x);
  fp2_mul_xi_k<64>(y, x);
  Fp2 c0;
  fadd(c0, y, t0);
  // c1 = (a0+a1)(b0+b1) - t0 - t1 + xi t2
  fadd(s, a.c0, a.c1);
  fadd(u, b.c0, b.c1);
  fp2_mul(x, s, u);
  fsub_k<16>(x, x, t0);
  fsub_k<16>(x, x, t1);
  fnorm(x, x);
  fp2_mul_xi_k<16>(y, t2);
  Fp2 c1;
  fadd(c1, x, y);
  // c2 = (a0+a2)(b0+b2) - t0 - t2 + t1
  fadd(s, a.c0, a.c2);
  fadd(u, b.c0, b.c2);
  fp2_mul(x, s, u);
  fsub_k<16>(x, x, t0);
  fsub_k<16>(x, x, t2);
  fadd(x, x, t1);
  fnorm(r.c2, x);
  fnorm(r.c0, c0);
  fnorm(r.c1, c1);
}
BLS_FN void fp6_mul(Fp6& r, const Fp6& a, const Fp6& b) {
  fp6_mul_lazy(r, a, b);
  fp6_red(r, r);
}

// inputs reduced
BLS_FN void fp6_sqr(Fp6& r, const Fp6& a) {
  Fp2 s0, s1, s2, s3, s4, t, y;
  fp2_sqr(s0, a.c0);
  fp2_mul(s1, a.c0, a.c1);
  fdbl(s1, s1);
  fsub_k<4>(t, a.c0, a.c1);
  fadd(t, t, a.c2);
  fnorm(t, t);
  fp2_sqr(s2, t);
  fp2_mul(s3, a.c1, a.c2);
  fdbl(s3, s3);
  fp2_sqr(s4, a.c2);
  fp2_mul_xi_k<32>(y, s3);
  fadd(y, y, s0);
  fred(r.c0, y);
  fp2_mul_xi_k<16>(y, s4);
  fadd(y, y, s1);
  fred(r.c1, y);
  fadd(t, s1, s2);
  fadd(t, t, s3);
  fsub_k<16>(t, t, s0);
  fnorm(t, t);
  fsub_k<16>(t, t, s4);
  fred(r.c2, t);
}

// a * (b0 + b1 v)
BLS_FN void fp6_mul_by_01_lazy(Fp6& r, const Fp6& a, const Fp2& b0, const Fp2& b1) {
  Fp2 t0, t1, s, u, x, y;
  fp2_mul(t0, a.c0, b0);
  fp2_mul(t1, a.c1, b1);
  fadd(s, a.c0, a.c1);
  fadd(u, b0, b1);
  fp2_mul(x, s, u);
  fsub_k<16>(x, x, t0);
  fsub_k<16>(x, x, t1);  // c1
  fp2_mul(y, a.c2, b1);
  fp2_mul_xi_k<16>(y, y);
  fp2_mul(s, a.c2, b0);
  fadd(y, y, t0);
  fnorm(r.c0, y);
  fnorm(r.c1, x);
  fadd(y, t1, s);
  fnorm(r.c2, y);
}
// a * (b1 v)
BLS_FN void fp6_mul_by_1_lazy(Fp6& r, const Fp6& a, const Fp2& b1) {
  Fp2 t0, t1, t2;
  fp2_mul(t2, a.c2, b1);
  fp2_mul(t0, a.c0, b1);
  fp2_mul(t1, a.c1, b1);
  fp2_mul_xi_k<16>(t2, t2);
  fnorm(r.c0, t2);
  r.c1 = t0;
  r.c2 = t1;
}

// input reduced
BLS_FN void fp6_inv(Fp6& r, const Fp6& a) {
  Fp2 t0, t1, t2, x, d;
  fp2_sqr(t0, a.c0);
  fp2_mul(x, a.c1, a.c2);
  fp2_mul_xi_k<16>(x, x);
  fnorm(x, x);
  fsub_k<32>(t0, t0, x);  // a0^2 - xi a1 a2
  fred(t0, t0);
  fp2_sqr(t1, a.c2);
  fp2_mul_xi_k<16>(t1, t1);
  fp2_mul(x, a.c0, a.c1);
  fsub_k<16>(t1, t1, x);  // xi a2^2 - a0 a1
  fred(t1, t1);
  fp2_sqr(t2, a.c1);
  fp2_mul(x, a.c0, a.c2);
  fsub_k<16>(t2, t2, x);  // a1^2 - a0 a2
  fred(t2, t2);
  fp2_mul(d, a.c2, t1);
  fp2_mul(x, a.c1, t2);
  fadd(d, d, x);
  fnorm(d, d);
  fp2_mul_xi_k<32>(d, d);
  fp2_mul(x, a.c0, t0);
  fadd(d, d, x);
  fred(d, d);
  fp2_inv(d, d);
  fp2_mul(x, t0, d);
  fred(r.c0, x);
  fp2_mul(x, t1, d);
  fred(r.c1, x);
  fp2_mul(x, t2, d);
  fred(r.c2, x);
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
  fp6_mul_lazy(t0, a.c0, b.c0);
  fp6_mul_lazy(t1, a.c1, b.c1);
  fp6_add(s, a.c0, a.c1);
  fp6_add(u, b.c0, b.c1);
  fp6_norm(s, s);  // both operands normalised: the fused Fp2 product sums the halves of BOTH operands once more
  fp6_norm(u, u);
  fp6_mul_lazy(x, s, u);
  fp6_sub_k<128>(x, x, t0);
  fp6_sub_k<128>(x, x, t1);
  fp6_red(r.c1, x);
  fp6_mul_by_v_k<128>(t1, t1);
  fp6_add(t0, t0, t1);
  fp6_red(r.c0, t0);
}

BLS_FN void fp12_sqr(Fp12& r, const Fp12& a) {
  Fp6 t, s, u, x;
  fp6_mul_lazy(t, a.c0, a.c1);
  fp6_add(s, a.c0, a.c1);
  fp6_norm(s, s);
  fp6_mul_by_v_k<4>(u, a.c1);
  fp6_add(u, u, a.c0);
  fp6_norm(u, u);
  fp6_mul_lazy(x, s, u);
  fp6_sub_k<128>(x, x, t);
  fp6_mul_by_v_k<128>(u, t);
  fp6_norm(u, u);
  fp6_sub_k<256>(x, x, u);
  fp6_red(r.c0, x);
  fp6_add(t, t, t);
  fp6_red(r.c1, t);
}

// f * ((c0 + c1 v) + (c4 v) w): the sparse shape of a Miller-loop line (w^0, w^2, w^3 coefficients).
// f reduced; line coefficients normalised with value <= 13.
BLS_FN void fp12_mul_by_014(Fp12& f, const Fp2& c0, const Fp2& c1, const Fp2& c4) {
  Fp6 aa, bb, s, x;
  Fp2 c14;
  fp6_mul_by_01_lazy(aa, f.c0, c0, c1);
  fp6_mul_by_1_lazy(bb, f.c1, c4);
  fadd(c14, c1, c4);
  fnorm(c14, c14);
  fp6_add(s, f.c0, f.c1);
  fp6_norm(s, s);
  fp6_mul_by_01_lazy(x, s, c0, c14);
  fp6_sub_k<64>(x, x, aa);
  fp6_sub_k<32>(x, x, bb);
  fp6_red(f.c1, x);
  fp6_mul_by_v_k<16>(bb, bb);
  fp6_add(aa, aa, bb);
  fp6_red(f.c0, aa);
}

BLS_FN void fp12_inv(Fp12& r, const Fp12& a) {
  Fp6 t0, t1;
  fp6_sqr(t0, a.c0);
  fp6_sqr(t1, a.c1);
  fp6_mul_by_v_k<4>(t1, t1);
  fp6_norm(t1, t1);
  fp6_sub_k<16>(t0, t0, t1);
  fp6_red(t0, t0);
  fp6_inv(t0, t0);
  fp6_mul(r.c0, a.c0, t0);
  fp6_mul(t1, a.c1, t0);
  fp6_neg(r.c1, t1);
}

// f^p : coefficient of w^i -> conj(c_i) * K_FROB1[i]   (reduced in, reduced out)
BLS_HD void frob1_coeff(Fp2& r, const Fp2& a, int i) {
  Fp2 g, t;
  fp2_conj_k<4>(t, a);
  fnorm(t, t);
  fp2_set(g, K_FROB1[i]);
  fp2_mul(t, t, g);
  fred(r, t);
}
BLS_FN void fp12_frob1(Fp12& r, const Fp12& a) {
  Fp2 t;
  fp2_conj_k<4>(t, a.c0.c0);
  fred(r.c0.c0, t);
  frob1_coeff(r.c1.c0, a.c1.c0, 1);
  frob1_coeff(r.c0.c1, a.c0.c1, 2);
  frob1_coeff(r.c1.c1, a.c1.c1, 3);
  frob1_coeff(r.c0.c2, a.c0.c2, 4);
  frob1_coeff(r.c1.c2, a.c1.c2, 5);
}
// f^(p^2) : coefficient of w^i -> c_i * K_FROB2[i]  (Fp scalars; products are < 2p with 28-bit limbs: reduced)
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

// (a + b s)^2 in Fp4 = Fp2[s]/(s^2 - xi): c0 = a^2 + xi b^2, c1 = 2ab.  a, b reduced; outputs lazy: c0 <= (12,10), c1 <= 20
BLS_HD void fp4_sqr(Fp2& c0, Fp2& c1, const Fp2& a, const Fp2& b) {
  Fp2 t0, t1, t2;
  fp2_sqr(t0, a);
  fp2_sqr(t1, b);
  fadd(t2, a, b);
  fp2_sqr(t2, t2);
  fsub_k<8>(t2, t2, t0);
  fsub_k<8>(t2, t2, t1);
  fnorm(c1, t2);
  fp2_mul_xi_k<8>(t1, t1);
  fadd(t0, t0, t1);
  fnorm(c0, t0);
}
// 3t - 2z  and  3t + 2z  with t lazy (value <= 48), z reduced; reduced results
BLS_HD void cyc_3t_m2z(Fp2& r, const Fp2& t, const Fp2& z) {
  Fp2 x;
  fsub_k<4>(x, t, z);
  fdbl(x, x);
  fadd(x, x, t);
  fred(r, x);
}
BLS_HD void cyc_3t_p2z(Fp2& r, const Fp2& t, const Fp2& z) {
  Fp2 x;
  fadd(x, t, z);
  fdbl(x, x);
  fadd(x, x, t);
  fred(r, x);
}

// Granger-Scott squaring for elements of the cyclotomic subgroup (after the easy part of the final exponentiation)
BLS_FN void fp12_cyclo_sqr(Fp12& r, const Fp12& a) {
  Fp2 z0 = a.c0.c0, z4 = a.c0.c1, z3 = a.c0.c2, z2 = a.c1.c0, z1 = a.c1.c1, z5 = a.c1.c2;
  Fp2 t0, t1, t2, t3;
  fp4_sqr(t0, t1, z0, z1);
  cyc_3t_m2z(z0, t0, z0);
  cyc_3t_p2z(z1, t1, z1);
  fp4_sqr(t0, t1, z2, z3);
  fp4_sqr(t2, t3, z4, z5);
  cyc_3t_m2z(z4, t0, z4);
  cyc_3t_p2z(z5, t1, z5);
  // z2 = 3 xi t3 + 2 z2 ; z3 = 3 t2 - 2 z3
  fp2_mul_xi_k<32>(t0, t3);
  fnorm(t0, t0);
  cyc_3t_p2z(z2, t0, z2);
  cyc_3t_m2z(z3, t2, z3);
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
