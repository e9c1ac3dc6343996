// G1 = E(Fp): y^2 = x^3 + 4 and G2 = E'(Fp2): y^2 = x^3 + 4(1+u), Jacobian coordinates (x = X/Z^2, y = Y/Z^3, Z = 0 is
// the identity), generic over the field through the f* vocabulary of fp2.cuh.
// Replaces the group operations blsful takes from its curve crate: `+=` / `*` / `is_identity` / neg
// (reference src/traits/sig_core.rs:42-57,126-139; src/secure_aggregation.rs:150-153,201-204), point decoding incl.
// the subgroup check (`from_compressed`, reference src/impls/legacy.rs:107,117,151,161; src/public_key.rs:71) and
// encoding (`to_compressed`, src/impls/legacy.rs:88,132).
#pragma once
#include "fp2.cuh"

namespace bls {

template <class F>
struct Aff {
  F x, y;
  uint32_t inf;
};
template <class F>
struct Jac {
  F X, Y, Z;
};
typedef Aff<Fp> G1Aff;
typedef Aff<Fp2> G2Aff;
typedef Jac<Fp> G1Jac;
typedef Jac<Fp2> G2Jac;

template <class F>
BLS_HD void jac_set_inf(Jac<F>& r) {
  fone(r.X);
  fone(r.Y);
  fzero(r.Z);
}
template <class F>
BLS_HD bool jac_is_inf(const Jac<F>& p) {
  return fis_zero(p.Z);
}
template <class F>
BLS_HD void jac_from_aff(Jac<F>& r, const Aff<F>& p) {
  if (p.inf) {
    jac_set_inf(r);
  } else {
    r.X = p.x;
    r.Y = p.y;
    fone(r.Z);
  }
}
template <class F>
BLS_HD void jac_neg(Jac<F>& r, const Jac<F>& p) {
  r.X = p.X;
  fneg_k<16>(r.Y, p.Y);
  fred(r.Y, r.Y);
  r.Z = p.Z;
}
template <class F>
BLS_HD void aff_neg(Aff<F>& r, const Aff<F>& p) {
  r.x = p.x;
  fneg_k<16>(r.y, p.y);
  fred(r.y, r.y);
  r.inf = p.inf;
}

// Bound contract of this file (see fp.cuh): coordinates of every stored point are normalised with value bound <= 8
// (each point operation ends with fred on X and Y), so the lazy sums below stay far from the limits.

// small multiples (lazy)
template <class F>
BLS_HD void fmul2(F& r, const F& a) { fdbl(r, a); }
template <class F>
BLS_HD void fmul4(F& r, const F& a) {
  fdbl(r, a);
  fdbl(r, r);
}
template <class F>
BLS_HD void fmul8(F& r, const F& a) {
  fdbl(r, a);
  fdbl(r, r);
  fdbl(r, r);
}

// Working-set note (round 2): the compiler gives every named temporary of these functions its own stack slot (stack
// colouring is off, __graft_entry__.py), and with 128 registers per thread all of them live in local memory.  At ~75,000
// resident threads each extra Fp2 temporary is 9.7 MB of L2; the round-1 versions (12 / 13 / 15 temporaries) pushed the
// per-thread hot set of a doubling loop past what the 126 MB L2 holds and the kernels streamed ~400 KB per point through
// HBM.  The versions below do the same arithmetic in 4 / 8 / 9 temporaries.

// doubling (a = 0), dbl-2009-l shape: A = X^2, B = Y^2, C = B^2, D = 2((X+B)^2 - A - C) = 4XB, E = 3A,
// X3 = E^2 - 2D, Y3 = E(D - X3) - 8C, Z3 = 2YZ   (5S + 2M).  r may alias p.
template <class F>
BLS_HD void jac_dbl_impl(Jac<F>& r, const Jac<F>& p) {
  F t0, t1, t2, t3;
  fmul(t0, p.Y, p.Z);  // YZ
  fsqr(t1, p.Y);       // B            value <= 2 (Fp) / (2,4) (Fp2)
  fadd(t2, p.X, t1);
  fnorm(t2, t2);
  fsqr(t2, t2);        // (X+B)^2
  fsqr(t3, p.X);       // A            -- p is not read below this line
  fsqr(t1, t1);        // C
  fsub_k<8>(t2, t2, t3);
  fsub_k<8>(t2, t2, t1);
  fnorm(t2, t2);       // 2XB          value <= 18 (Fp) / 20 (Fp2)
  fdbl(t0, t0);
  fred(r.Z, t0);       // Z3
  fdbl(t0, t3);
  fadd(t0, t0, t3);
  fnorm(t0, t0);       // E = 3A       value <= 6 (Fp) / 12 (Fp2)
  fsqr(t3, t0);        // E^2
  fmul4(r.X, t2);
  fnorm(r.X, r.X);     // 8XB          value <= 80
  fsub_k<128>(t3, t3, r.X);
  fred(r.X, t3);       // X3
  fdbl(t2, t2);        // D = 4XB      value <= 40
  fsub_k<4>(t2, t2, r.X);
  fnorm(t2, t2);
  fmul(t3, t0, t2);    // E (D - X3)
  fmul8(t1, t1);
  fnorm(t1, t1);       // 8C           value <= 16 (Fp) / 32 (Fp2)
  fsub_k<64>(t3, t3, t1);
  fred(r.Y, t3);
}

// madd-2007-bl: 7M + 4S, with the exceptional cases handled (public data, variable time).  r may alias p.
template <class F>
BLS_HD void jac_add_mixed_impl(Jac<F>& r, const Jac<F>& p, const Aff<F>& q) {
  if (q.inf) {
    r = p;
    return;
  }
  if (jac_is_inf(p)) {
    jac_from_aff(r, q);
    return;
  }
  F a, b, c, d, e, f, g, h;
  fsqr(a, p.Z);       // Z1Z1
  fmul(b, q.x, a);    // U2
  fmul(c, q.y, p.Z);
  fmul(c, c, a);      // S2
  fsub_k<16>(b, b, p.X);
  fsub_k<16>(c, c, p.Y);
  fred(b, b);         // H
  fred(c, c);         // r/2
  if (fis_zero(b)) {
    if (fis_zero(c)) {
      jac_dbl_impl(r, p);
    } else {
      jac_set_inf(r);
    }
    return;
  }
  fdbl(c, c);         // rr
  fsqr(d, b);         // HH
  fmul4(e, d);        // I, value <= 8, limbs < 2^30
  fmul(f, b, e);      // J
  fmul(e, p.X, e);    // V
  fsqr(g, c);
  fdbl(h, e);
  fadd(h, h, f);      // J + 2V, value <= 6 (Fp) / 30 (Fp2)
  fsub_k<32>(g, g, h);
  fred(g, g);         // X3
  fsub_k<4>(e, e, g);
  fnorm(e, e);
  fmul(e, c, e);      // rr (V - X3)
  fmul(f, p.Y, f);
  fdbl(f, f);         // 2 Y1 J, value <= 4 (Fp) / 20 (Fp2)
  fsub_k<32>(e, e, f);
  fadd(h, p.Z, b);
  fnorm(h, h);
  fsqr(h, h);         // (Z1 + H)^2
  fadd(a, a, d);
  fsub_k<16>(h, h, a);
  fred(r.Y, e);
  r.X = g;
  fred(r.Z, h);
}

// add-2007-bl: 11M + 5S.  r may alias p or q.
template <class F>
BLS_HD void jac_add_impl(Jac<F>& r, const Jac<F>& p, const Jac<F>& q) {
  if (jac_is_inf(q)) {
    r = p;
    return;
  }
  if (jac_is_inf(p)) {
    r = q;
    return;
  }
  F a, b, c, d, e, f, g, h, i;
  fsqr(a, p.Z);       // Z1Z1
  fsqr(b, q.Z);       // Z2Z2
  fmul(c, p.X, b);    // U1
  fmul(d, q.X, a);    // U2
  fmul(e, p.Y, q.Z);
  fmul(e, e, b);      // S1
  fmul(f, q.Y, p.Z);
  fmul(f, f, a);      // S2
  fsub_k<16>(d, d, c);
  fsub_k<16>(f, f, e);
  fred(d, d);         // H
  fred(f, f);         // r/2
  if (fis_zero(d)) {
    if (fis_zero(f)) {
      jac_dbl_impl(r, p);
    } else {
      jac_set_inf(r);
    }
    return;
  }
  fdbl(f, f);         // rr
  fdbl(g, d);
  fsqr(g, g);         // I
  fmul(h, d, g);      // J
  fmul(c, c, g);      // V
  fsqr(g, f);
  fdbl(i, c);
  fadd(i, i, h);
  fsub_k<32>(g, g, i);
  fred(g, g);         // X3
  fsub_k<4>(c, c, g);
  fnorm(c, c);
  fmul(c, f, c);      // rr (V - X3)
  fmul(e, e, h);
  fdbl(e, e);         // 2 S1 J
  fsub_k<32>(c, c, e);
  fadd(i, p.Z, q.Z);
  fnorm(i, i);
  fsqr(i, i);
  fadd(a, a, b);
  fsub_k<16>(i, i, a);
  fnorm(i, i);
  fmul(i, i, d);
  fred(r.Y, c);
  r.X = g;
  fred(r.Z, i);
}

// out-of-line instances (code size: one copy of each per group)
BLS_FN void jac_dbl(G1Jac& r, const G1Jac& p) { jac_dbl_impl(r, p); }
BLS_FN void jac_dbl(G2Jac& r, const G2Jac& p) { jac_dbl_impl(r, p); }
BLS_FN void jac_add_mixed(G1Jac& r, const G1Jac& p, const G1Aff& q) { jac_add_mixed_impl(r, p, q); }
BLS_FN void jac_add_mixed(G2Jac& r, const G2Jac& p, const G2Aff& q) { jac_add_mixed_impl(r, p, q); }
BLS_FN void jac_add(G1Jac& r, const G1Jac& p, const G1Jac& q) { jac_add_impl(r, p, q); }
BLS_FN void jac_add(G2Jac& r, const G2Jac& p, const G2Jac& q) { jac_add_impl(r, p, q); }

template <class F>
BLS_HD void jac_to_aff(Aff<F>& r, const Jac<F>& p) {
  if (jac_is_inf(p)) {
    fzero(r.x);
    fzero(r.y);
    r.inf = 1;
    return;
  }
  F zi, zi2;
  finv(zi, p.Z);
  fsqr(zi2, zi);
  fmul(r.x, p.X, zi2);
  fmul(zi2, zi2, zi);
  fmul(r.y, p.Y, zi2);
  r.inf = 0;
}

// cnt (<= K) Jacobian points -> affine with ONE field inversion (Montgomery's simultaneous inversion: prefix products
// forward, one inverse, peel backward).  Points at infinity take part with 1 in place of their Z.
template <class F, int K>
BLS_HD void jac_to_aff_batch(Aff<F>* out, const Jac<F>* in, int cnt) {
  F pre[K], acc, z;
  fone(acc);
  for (int j = 0; j < cnt; j++) {
    pre[j] = acc;
    z = in[j].Z;
    if (!fis_zero(z)) fmul(acc, acc, z);
  }
  finv(acc, acc);  // 1 / (product of the non-zero Z)
  for (int j = cnt - 1; j >= 0; j--) {
    Jac<F> p = in[j];
    Aff<F> a;
    if (fis_zero(p.Z)) {
      fzero(a.x);
      fzero(a.y);
      a.inf = 1;
    } else {
      F zi, zi2;
      fmul(zi, acc, pre[j]);  // 1 / Z_j
      fmul(acc, acc, p.Z);    // inverse of the product of the Z before j
      fsqr(zi2, zi);
      fmul(a.x, p.X, zi2);
      fmul(zi2, zi2, zi);
      fmul(a.y, p.Y, zi2);
      a.inf = 0;
    }
    out[j] = a;
  }
}

// projective equality
template <class F>
BLS_HD bool jac_eq(const Jac<F>& a, const Jac<F>& b) {
  bool ia = jac_is_inf(a), ib = jac_is_inf(b);
  if (ia || ib) return ia && ib;
  F za2, zb2, t0, t1;
  fsqr(za2, a.Z);
  fsqr(zb2, b.Z);
  fmul(t0, a.X, zb2);
  fmul(t1, b.X, za2);
  if (!feq(t0, t1)) return false;
  fmul(za2, za2, a.Z);
  fmul(zb2, zb2, b.Z);
  fmul(t0, a.Y, zb2);
  fmul(t1, b.Y, za2);
  return feq(t0, t1);
}

// [k]P for an affine base point, k given as little-endian 32-bit limbs (public scalars: plain double-and-add)
template <class F>
BLS_HD void jac_mul_aff(Jac<F>& r, const Aff<F>& p, const uint32_t* k, int nlimbs) {
  Jac<F> acc;
  jac_set_inf(acc);
  bool started = false;
  for (int i = nlimbs - 1; i >= 0; i--) {
    uint32_t w = k[i];
    for (int b = 31; b >= 0; b--) {
      if (started) jac_dbl(acc, acc);
      if ((w >> b) & 1u) {
        jac_add_mixed(acc, acc, p);
        started = true;
      }
    }
  }
  r = acc;
}
// r = k P for a 64- or 128-bit k: fixed 4-bit windows over a 15-entry Jacobian table.  What matters on a GPU is that all lanes
// of a warp add at the SAME 16 positions: a sparse signed-digit form (NAF) was measured and is twice as SLOW as plain
// double-and-add here, because a warp executes an addition whenever any of its 32 lanes has a non-zero digit - and then
// once more for the other sign.  64-bit k: 60 doublings + 16 full additions + the table (7 doublings, 7 mixed additions).
template <class F>
BLS_HD void jac_mul_aff_w4(Jac<F>& r, const Aff<F>& p, const uint32_t* k, int nwin) {  // k: nwin 4-bit windows, little-endian words
  Jac<F> tbl[16];
  jac_set_inf(tbl[0]);
  jac_from_aff(tbl[1], p);
  for (int i = 2; i < 16; i += 2) {
    jac_dbl(tbl[i], tbl[i / 2]);
    jac_add_mixed(tbl[i + 1], tbl[i], p);
  }
  Jac<F> acc = tbl[(k[(nwin - 1) >> 3] >> (4 * ((nwin - 1) & 7))) & 15u];
  for (int w = nwin - 2; w >= 0; w--) {
    for (int d = 0; d < 4; d++) jac_dbl(acc, acc);
    const uint32_t dg = (k[w >> 3] >> (4 * (w & 7))) & 15u;
    if (dg) jac_add(acc, acc, tbl[dg]);
  }
  r = acc;
}
template <class F>
BLS_HD void jac_mul_aff_w4_64(Jac<F>& r, const Aff<F>& p, uint64_t k) {
  const uint32_t w[2] = {(uint32_t)k, (uint32_t)(k >> 32)};
  jac_mul_aff_w4(r, p, w, 16);
}
// [|x|]P, |x| = 0xd201000000010000 (Hamming weight 6), Jacobian base.  r must not alias p (the running point lives in r:
// no second copy on the stack).
template <class F>
BLS_HD void jac_mul_xabs(Jac<F>& r, const Jac<F>& p) {
  r = p;
  const uint64_t e = K_X_ABS;
  for (int i = 62; i >= 0; i--) {
    jac_dbl(r, r);
    if ((e >> i) & 1) jac_add(r, r, p);
  }
}
template <class F>
BLS_HD void jac_mul_xabs_aff(Jac<F>& r, const Aff<F>& p) {
  jac_from_aff(r, p);
  const uint64_t e = K_X_ABS;
  for (int i = 62; i >= 0; i--) {
    jac_dbl(r, r);
    if ((e >> i) & 1) jac_add_mixed(r, r, p);
  }
}

// ---- endomorphisms and subgroup checks -----------------------------------------------------------------
// G1: phi(x,y) = (beta x, y) acts as [-x^2] on G1; P in G1  <=>  phi(P) + [x^2]P = O   (Scott, eprint 2021/1130)
BLS_FN bool g1_in_subgroup(const G1Aff& p) {
  if (p.inf) return true;
  G1Jac t, t1;
  jac_mul_xabs_aff(t1, p);
  jac_mul_xabs(t, t1);  // [x^2]P  (sign of x cancels)
  G1Aff phi;
  Fp beta;
  fp_set(beta, K_BETA);
  fp_mul(phi.x, p.x, beta);
  phi.y = p.y;
  phi.inf = 0;
  jac_add_mixed(t, t, phi);
  return jac_is_inf(t);
}
// psi on E'(Fp2): (X,Y,Z) -> (conj(X) cx, conj(Y) cy, conj(Z))
BLS_HD void g2_psi(G2Jac& r, const G2Jac& p) {
  Fp2 cx, cy, t;
  fp2_set(cx, K_PSI_CX);
  fp2_set(cy, K_PSI_CY);
  fp2_conj(t, p.X);
  fp2_mul(t, t, cx);
  fred(r.X, t);
  fp2_conj(t, p.Y);
  fp2_mul(t, t, cy);
  fred(r.Y, t);
  fp2_conj(t, p.Z);
  fred(r.Z, t);
}
// psi^2: (x,y) -> (x * K_PSI2_CX, -y)
BLS_HD void g2_psi2(G2Jac& r, const G2Jac& p) {
  Fp c;
  fp_set(c, K_PSI2_CX);
  fp2_mul_fp(r.X, p.X, c);
  fneg(r.Y, p.Y);
  fred(r.Y, r.Y);
  r.Z = p.Z;
}
// G2: P in G2  <=>  psi(P) = [x]P  (Scott), x = -|x|
BLS_FN bool g2_in_subgroup(const G2Aff& p) {
  if (p.inf) return true;
  G2Jac t, pj, ps;
  jac_mul_xabs_aff(t, p);
  jac_neg(t, t);
  jac_from_aff(pj, p);
  g2_psi(ps, pj);
  return jac_eq(t, ps);
}

// ---- compressed encodings (ZCash/IETF "Modern" flags; the Legacy header rewrite is in codec.cuh) --------------
// status codes shared with include/blsgpu.h
enum : uint8_t {
  ST_OK = 0,
  ST_INVALID_SIGNATURE = 1,
  ST_SIG_IDENTITY = 2,
  ST_PK_IDENTITY = 3,
  ST_DESERIALIZE = 4,
  ST_LEGACY_FORMAT = 5,
  ST_INVALID_LENGTH = 6,
  ST_INVALID_COEFFICIENT = 7,
  ST_DUPLICATE_MESSAGES = 8,
  ST_SCHEME = 9,
  ST_MISMATCHED_LENGTHS = 10,
  ST_VSSS = 11,
  ST_INVALID_PROOF = 12,
  ST_COMMITMENT_IDENTITY = 13,
  ST_PROOF_IDENTITY = 14,
  ST_ZERO_CHALLENGE = 15
};

BLS_HD bool bytes_all_zero(const uint8_t* b, int n) {
  uint32_t t = 0;
  for (int i = 0; i < n; i++) t |= b[i];
  return t == 0;
}

// 48 Modern bytes -> G1 affine (Montgomery).  `check_subgroup` false is used by tests only.
BLS_FN uint8_t g1_decompress(G1Aff& r, const uint8_t* in, bool check_subgroup) {
  uint8_t b[48];
  for (int i = 0; i < 48; i++) b[i] = in[i];
  uint8_t h = b[0];
  b[0] &= 0x1f;
  if (!(h & 0x80)) return ST_DESERIALIZE;
  if (h & 0x40) {
    if ((h & 0x20) || !bytes_all_zero(b, 48)) return ST_DESERIALIZE;
    fzero(r.x);
    fzero(r.y);
    r.inf = 1;
    return ST_OK;
  }
  Fp raw, x, y2, y, b4;
  if (!fp_from_be48_raw(raw, b)) return ST_DESERIALIZE;
  fp_to_mont(x, raw);
  fp_sqr(y2, x);
  fp_mul(y2, y2, x);
  fp_set(b4, K_B1);
  fp_add(y2, y2, b4);
  if (!fp_sqrt(y, y2)) return ST_DESERIALIZE;
  if (fp_lex_largest(y) != ((h & 0x20) != 0)) {
    fp_neg(y, y);
    fp_red(y, y);
  }
  r.x = x;
  r.y = y;
  r.inf = 0;
  if (check_subgroup && !g1_in_subgroup(r)) return ST_DESERIALIZE;
  return ST_OK;
}

// 96 Modern bytes (x.c1 || x.c0) -> G2 affine
BLS_FN uint8_t g2_decompress(G2Aff& r, const uint8_t* in, bool check_subgroup) {
  uint8_t b[96];
  for (int i = 0; i < 96; i++) b[i] = in[i];
  uint8_t h = b[0];
  b[0] &= 0x1f;
  if (!(h & 0x80)) return ST_DESERIALIZE;
  if (h & 0x40) {
    if ((h & 0x20) || !bytes_all_zero(b, 96)) return ST_DESERIALIZE;
    fzero(r.x);
    fzero(r.y);
    r.inf = 1;
    return ST_OK;
  }
  Fp raw0, raw1;
  if (!fp_from_be48_raw(raw1, b)) return ST_DESERIALIZE;
  if (!fp_from_be48_raw(raw0, b + 48)) return ST_DESERIALIZE;
  Fp2 x, y2, y, b2, one;
  fp_to_mont(x.c0, raw0);
  fp_to_mont(x.c1, raw1);
  fp2_sqr(y2, x);
  fp2_mul(y2, y2, x);
  fp2_set(b2, K_B2);
  fadd(y2, y2, b2);
  fone(one);
  if (!fp2_sqrt_ratio(y, y2, one)) return ST_DESERIALIZE;
  if (fp2_lex_largest(y) != ((h & 0x20) != 0)) {
    fneg(y, y);
    fred(y, y);
  }
  r.x = x;
  r.y = y;
  r.inf = 0;
  if (check_subgroup && !g2_in_subgroup(r)) return ST_DESERIALIZE;
  return ST_OK;
}

BLS_HD void g1_compress(uint8_t* out, const G1Aff& p) {
  if (p.inf) {
    out[0] = 0xc0;
    for (int i = 1; i < 48; i++) out[i] = 0;
    return;
  }
  Fp raw;
  fp_from_mont(raw, p.x);
  fp_to_be48_raw(out, raw);
  out[0] |= 0x80 | (fp_lex_largest(p.y) ? 0x20 : 0);
}
BLS_HD void g2_compress(uint8_t* out, const G2Aff& p) {
  if (p.inf) {
    out[0] = 0xc0;
    for (int i = 1; i < 96; i++) out[i] = 0;
    return;
  }
  Fp raw;
  fp_from_mont(raw, p.x.c1);
  fp_to_be48_raw(out, raw);
  fp_from_mont(raw, p.x.c0);
  fp_to_be48_raw(out + 48, raw);
  out[0] |= 0x80 | (fp2_lex_largest(p.y) ? 0x20 : 0);
}

// Header handling of the two serialisation formats (reference src/impls/legacy.rs:19-82):
// rewrites byte 0 into the Modern form or reports the reference's error.  format: 0 = Legacy, 1 = Modern.
BLS_HD uint8_t header_to_modern(uint8_t& b0, int format) {
  if (format == 1) {
    if (b0 != 0xc0 && (b0 & 0xc0) != 0x80) return ST_DESERIALIZE;  // validate_modern_format
    return ST_OK;
  }
  if (b0 == 0xc0) return ST_OK;
  uint8_t y_sign = b0 & 0x80;
  uint8_t v = b0 & 0x7f;
  if (v & 0xe0) return ST_LEGACY_FORMAT;
  v |= 0x80;
  if (y_sign) v |= 0x20;
  b0 = v;
  return ST_OK;
}
BLS_HD void header_from_modern(uint8_t& b0, int format) {
  if (format == 1 || b0 == 0xc0) return;
  uint8_t y_sign = b0 & 0x20;
  b0 &= 0x1f;
  if (y_sign) b0 |= 0x80;
}

}  // namespace bls
