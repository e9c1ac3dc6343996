// Optimal-ate Miller loop on BLS12-381 (|x| = 0xd201000000010000: 63 doubling steps, 5 addition steps), Jacobian running
// point on the twist, sparse line products.  Replaces `multi_miller_loop` (reference src/helpers.rs:50,62).
// The G1 argument may be left in Jacobian form: evaluating a line at (X/Z^2, Y/Z^3) only rescales it by Z^3 in Fp,
// which the final exponentiation erases, so the r_i * pk_i of the batch check never need an inversion.
#pragma once
#include "fp12.cuh"
#include "curve.cuh"

namespace bls {

// G1 argument prepared for line evaluation: px = X Z, py = Y, pz = Z^3 (affine: x, y, 1)
struct MillerG1 {
  Fp px, py, pz;
  uint32_t pz_is_one;
};
BLS_HD void miller_prepare(MillerG1& m, const G1Jac& p) {
  Fp z2;
  fp_mul(m.px, p.X, p.Z);
  m.py = p.Y;
  fp_sqr(z2, p.Z);
  fp_mul(m.pz, z2, p.Z);
  m.pz_is_one = 0;
}
BLS_HD void miller_prepare(MillerG1& m, const G1Aff& p) {
  m.px = p.x;
  m.py = p.y;
  fp_one(m.pz);
  m.pz_is_one = 1;
}

// Bound contract: R's coordinates are reduced on entry and on exit; P's are Fp products (value < 3, limbs < 2^28); the
// line coefficients come out normalised with value <= 13, which is what fp12_mul_by_014 accepts.
//
// tangent line at R evaluated at P, then R = 2R.  (c0, c2, c3) are the w^0, w^2, w^3 coefficients.
BLS_FN void miller_dbl_step(Fp2& c0, Fp2& c2, Fp2& c3, G2Jac& R, const MillerG1& P) {
  Fp2 A, B, C, ZZ, D, E, Fq, t, X3, Y3, Z3;
  fp2_sqr(A, R.X);
  fp2_sqr(B, R.Y);
  fp2_sqr(C, B);
  fp2_sqr(ZZ, R.Z);
  fadd(t, R.X, B);
  fp2_sqr(t, t);
  fsub_k<8>(t, t, A);
  fsub_k<8>(t, t, C);
  fnorm(t, t);
  fdbl(D, t);  // D = 2((X+B)^2 - A - C), value <= 40
  fdbl(E, A);
  fadd(E, E, A);
  fnorm(E, E);  // 3A, value <= (6,12)
  fp2_sqr(Fq, E);
  fdbl(t, D);
  fnorm(t, t);
  fsub_k<128>(X3, Fq, t);
  fred(X3, X3);
  fadd(Z3, R.Y, R.Z);
  fp2_sqr(Z3, Z3);
  fsub_k<8>(Z3, Z3, B);
  fsub_k<8>(Z3, Z3, ZZ);
  fred(Z3, Z3);
  // line
  fp2_mul(c0, E, R.X);
  fsub_k<8>(c0, c0, B);
  fsub_k<8>(c0, c0, B);
  if (!P.pz_is_one) {
    fnorm(c0, c0);
    fp2_mul_fp(c0, c0, P.pz);
  } else {
    fred(c0, c0);
  }
  fp2_mul(t, E, ZZ);
  fp2_mul_fp(t, t, P.px);
  fneg_k<4>(c2, t);
  fnorm(c2, c2);
  fp2_mul(t, Z3, ZZ);
  fp2_mul_fp(c3, t, P.py);
  // finish the doubling
  fsub_k<4>(Y3, D, X3);
  fnorm(Y3, Y3);
  fp2_mul(Y3, E, Y3);
  fmul8(C, C);
  fnorm(C, C);
  fsub_k<64>(Y3, Y3, C);
  fred(R.Y, Y3);
  R.X = X3;
  R.Z = Z3;
}

// chord through R and Q evaluated at P, then R = R + Q (Q affine, reduced coordinates)
BLS_FN void miller_add_step(Fp2& c0, Fp2& c2, Fp2& c3, G2Jac& R, const G2Aff& Q, const MillerG1& P) {
  Fp2 Z1Z1, U2, S2, H, HH, I, J, rr, V, t, X3, Y3, Z3;
  fp2_sqr(Z1Z1, R.Z);
  fp2_mul(U2, Q.x, Z1Z1);
  fp2_mul(S2, Q.y, R.Z);
  fp2_mul(S2, S2, Z1Z1);
  fsub_k<4>(H, U2, R.X);
  fnorm(H, H);
  fsub_k<4>(rr, S2, R.Y);
  fred(rr, rr);
  fdbl(rr, rr);
  fp2_sqr(HH, H);
  fmul4(I, HH);
  fnorm(I, I);
  fp2_mul(J, H, I);
  fp2_mul(V, R.X, I);
  fp2_sqr(X3, rr);
  fdbl(t, V);
  fadd(t, t, J);
  fnorm(t, t);  // J + 2V, value <= (18,30)
  fsub_k<32>(X3, X3, t);
  fred(X3, X3);
  fsub_k<4>(Y3, V, X3);
  fnorm(Y3, Y3);
  fp2_mul(Y3, rr, Y3);
  fp2_mul(t, R.Y, J);
  fdbl(t, t);
  fsub_k<32>(Y3, Y3, t);
  fred(Y3, Y3);
  fadd(Z3, R.Z, H);
  fp2_sqr(Z3, Z3);
  fsub_k<8>(Z3, Z3, Z1Z1);
  fsub_k<8>(Z3, Z3, HH);
  fred(Z3, Z3);
  // line: c0 = rr x2 - Z3 y2 ; c2 = -rr px ; c3 = Z3 py
  fp2_mul(c0, rr, Q.x);
  fp2_mul(t, Z3, Q.y);
  fsub_k<16>(c0, c0, t);
  if (!P.pz_is_one) {
    fnorm(c0, c0);
    fp2_mul_fp(c0, c0, P.pz);
  } else {
    fred(c0, c0);
  }
  fp2_mul_fp(t, rr, P.px);
  fneg_k<4>(c2, t);
  fnorm(c2, c2);
  fp2_mul_fp(c3, Z3, P.py);
  R.X = X3;
  R.Y = Y3;
  R.Z = Z3;
}

// f = f_{|x|,Q}(P) conjugated (x < 0).  Q must be a non-identity affine point of G2, P non-identity.
BLS_FN void miller_loop(Fp12& f, const MillerG1& P, const G2Aff& Q) {
  G2Jac R;
  jac_from_aff(R, Q);
  Fp2 c0, c2, c3;
  fp12_one(f);
  const uint64_t e = K_X_ABS;
  for (int i = 62; i >= 0; i--) {
    if (i != 62) fp12_sqr(f, f);
    miller_dbl_step(c0, c2, c3, R, P);
    fp12_mul_by_014(f, c0, c2, c3);
    if ((e >> i) & 1) {
      miller_add_step(c0, c2, c3, R, Q, P);
      fp12_mul_by_014(f, c0, c2, c3);
    }
  }
  fp12_conj(f, f);
}

}  // namespace bls
