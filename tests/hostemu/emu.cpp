// TEST INFRASTRUCTURE ONLY: compiles the engine's device headers for the host (g++) so that the CPU-only test suite can
// check the field/curve/pairing logic against the oracle without a GPU.  Nothing in the product links or loads this;
// the C-ABI library (csrc/blsgpu.cu) has no CPU path and fails loudly without a CUDA device.
#include <cstring>
#include "../../agora-blsful_b200/csrc/pairing.cuh"
#include "../../agora-blsful_b200/csrc/h2c.cuh"
#include "../../agora-blsful_b200/csrc/miller6.cuh"
#include "../../agora-blsful_b200/csrc/finalexp6.cuh"
#include "../../agora-blsful_b200/csrc/fr.cuh"

using namespace bls;

static void fp_in(Fp& r, const uint8_t* b) {
  Fp raw;
  fp_from_be48_raw(raw, b);
  fp_to_mont(r, raw);
}
static void fp_out(uint8_t* b, const Fp& a) {
  Fp raw;
  fp_from_mont(raw, a);
  fp_to_be48_raw(b, raw);
}
static void fp2_in(Fp2& r, const uint8_t* b) { fp_in(r.c0, b); fp_in(r.c1, b + 48); }
static void fp2_out(uint8_t* b, const Fp2& a) { fp_out(b, a.c0); fp_out(b + 48, a.c1); }
static void fp12_out(uint8_t* b, const Fp12& f) {
  // order w^0..w^5 (oracle layout): a0, b0, a1, b1, a2, b2
  fp2_out(b, f.c0.c0); fp2_out(b + 96, f.c1.c0); fp2_out(b + 192, f.c0.c1);
  fp2_out(b + 288, f.c1.c1); fp2_out(b + 384, f.c0.c2); fp2_out(b + 480, f.c1.c2);
}
static void fp12_in(Fp12& f, const uint8_t* b) {
  fp2_in(f.c0.c0, b); fp2_in(f.c1.c0, b + 96); fp2_in(f.c0.c1, b + 192);
  fp2_in(f.c1.c1, b + 288); fp2_in(f.c0.c2, b + 384); fp2_in(f.c1.c2, b + 480);
}

extern "C" {
void emu_fp_mul(const uint8_t* a, const uint8_t* b, uint8_t* out, int which) {
  Fp x, y, z;
  fp_in(x, a); fp_in(y, b);
  if (which == 0) fp_mul_inl(z, x, y); else if (which == 1) fp_mul(z, x, y); else fp_sqr_inl(z, x);
  fp_out(out, z);
}
void emu_fp_addsub(const uint8_t* a, const uint8_t* b, uint8_t* out_add, uint8_t* out_sub, uint8_t* out_neg) {
  Fp x, y, z;
  fp_in(x, a); fp_in(y, b);
  fp_add(z, x, y); fp_out(out_add, z);
  fp_sub(z, x, y); fp_out(out_sub, z);
  fp_neg(z, x); fp_out(out_neg, z);
}
void emu_fp_inv(const uint8_t* a, uint8_t* out) { Fp x, z; fp_in(x, a); fp_inv(z, x); fp_out(out, z); }
int emu_fp_sqrt(const uint8_t* a, uint8_t* out) { Fp x, z; fp_in(x, a); bool ok = fp_sqrt(z, x); fp_out(out, z); return ok; }
int emu_fp2_sqrt_ratio(const uint8_t* n, const uint8_t* d, uint8_t* out) {
  Fp2 x, y, z; fp2_in(x, n); fp2_in(y, d);
  bool ok = fp2_sqrt_ratio(z, x, y); fp2_out(out, z); return ok;
}
void emu_fp2_mul(const uint8_t* a, const uint8_t* b, uint8_t* out_mul, uint8_t* out_sqr, uint8_t* out_inv) {
  Fp2 x, y, z; fp2_in(x, a); fp2_in(y, b);
  fp2_mul(z, x, y); fp2_out(out_mul, z);
  fp2_sqr(z, x); fp2_out(out_sqr, z);
  fp2_inv(z, x); fp2_out(out_inv, z);
}
void emu_fp12_ops(const uint8_t* a, const uint8_t* b, uint8_t* mul, uint8_t* sqr, uint8_t* inv, uint8_t* fr1, uint8_t* fr2) {
  Fp12 x, y, z; fp12_in(x, a); fp12_in(y, b);
  fp12_mul(z, x, y); fp12_out(mul, z);
  fp12_sqr(z, x); fp12_out(sqr, z);
  fp12_inv(z, x); fp12_out(inv, z);
  fp12_frob1(z, x); fp12_out(fr1, z);
  fp12_frob2(z, x); fp12_out(fr2, z);
}
// returns 1 if cyclotomic squaring agrees with the generic squaring on easy_part(a)
int emu_cyclo_check(const uint8_t* a) {
  Fp12 x, t0, t1, m, s0, s1;
  fp12_in(x, a);
  fp12_conj(t0, x); fp12_inv(t1, x); fp12_mul(t0, t0, t1); fp12_frob2(t1, t0); fp12_mul(m, t1, t0);
  fp12_cyclo_sqr(s0, m); fp12_sqr(s1, m);
  return fp12_eq(s0, s1);
}
void emu_final_exp(const uint8_t* a, uint8_t* out) { Fp12 x, z; fp12_in(x, a); final_exponentiation(z, x); fp12_out(out, z); }

int emu_g1_decompress(const uint8_t* in, int format, uint8_t* xy, int* inf) {
  uint8_t b[48]; memcpy(b, in, 48);
  uint8_t st = header_to_modern(b[0], format);
  if (st) return st;
  G1Aff p; st = g1_decompress(p, b, true);
  if (st) return st;
  *inf = p.inf; fp_out(xy, p.x); fp_out(xy + 48, p.y);
  return 0;
}
int emu_g2_decompress(const uint8_t* in, int format, uint8_t* xy, int* inf) {
  uint8_t b[96]; memcpy(b, in, 96);
  uint8_t st = header_to_modern(b[0], format);
  if (st) return st;
  G2Aff p; st = g2_decompress(p, b, true);
  if (st) return st;
  *inf = p.inf; fp2_out(xy, p.x); fp2_out(xy + 96, p.y);
  return 0;
}
// decode (no subgroup check), multiply by a scalar (little-endian 32-bit limbs), re-encode in `format`
int emu_g1_mul(const uint8_t* in, const uint32_t* k, int nl, int format, uint8_t* out) {
  G1Aff p, q; if (g1_decompress(p, in, false)) return -1;
  G1Jac r; jac_mul_aff(r, p, k, nl); jac_to_aff(q, r); g1_compress(out, q); header_from_modern(out[0], format); return 0;
}
int emu_g2_mul(const uint8_t* in, const uint32_t* k, int nl, int format, uint8_t* out) {
  G2Aff p, q; if (g2_decompress(p, in, false)) return -1;
  G2Jac r; jac_mul_aff(r, p, k, nl); jac_to_aff(q, r); g2_compress(out, q); header_from_modern(out[0], format); return 0;
}
int emu_g1_mul_w4(const uint8_t* in, uint64_t k, uint8_t* out) {
  G1Aff p, q; if (g1_decompress(p, in, false)) return -1;
  G1Jac r; jac_mul_aff_w4_64(r, p, k); jac_to_aff(q, r); g1_compress(out, q); return 0;
}
// cnt points (compressed; identity allowed), each scaled by k to give it a non-trivial Z, normalised with one shared inversion
int emu_to_aff_batch(int g2, const uint8_t* in, int cnt, uint64_t k, uint8_t* out) {
  uint32_t kl[2] = {(uint32_t)k, (uint32_t)(k >> 32)};
  if (cnt > 16) return -3;
  if (g2) {
    G2Jac j[16]; G2Aff a[16];
    for (int i = 0; i < cnt; i++) { G2Aff p; if (g2_decompress(p, in + 96 * i, false)) return -1; jac_mul_aff(j[i], p, kl, 2); }
    jac_to_aff_batch<Fp2, 16>(a, j, cnt);
    for (int i = 0; i < cnt; i++) g2_compress(out + 96 * i, a[i]);
  } else {
    G1Jac j[16]; G1Aff a[16];
    for (int i = 0; i < cnt; i++) { G1Aff p; if (g1_decompress(p, in + 48 * i, false)) return -1; jac_mul_aff(j[i], p, kl, 2); }
    jac_to_aff_batch<Fp, 16>(a, j, cnt);
    for (int i = 0; i < cnt; i++) g1_compress(out + 48 * i, a[i]);
  }
  return 0;
}
int emu_g1_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  G1Aff p, q, s; if (g1_decompress(p, a, false) || g1_decompress(q, b, false)) return -1;
  G1Jac x, y, r; jac_from_aff(x, p); jac_from_aff(y, q); jac_add(r, x, y);
  G1Jac r2; jac_add_mixed(r2, x, q); if (!jac_eq(r, r2)) return -2;
  jac_to_aff(s, r); g1_compress(out, s); return 0;
}
int emu_g2_add(const uint8_t* a, const uint8_t* b, uint8_t* out) {
  G2Aff p, q, s; if (g2_decompress(p, a, false) || g2_decompress(q, b, false)) return -1;
  G2Jac x, y, r; jac_from_aff(x, p); jac_from_aff(y, q); jac_add(r, x, y);
  G2Jac r2; jac_add_mixed(r2, x, q); if (!jac_eq(r, r2)) return -2;
  jac_to_aff(s, r); g2_compress(out, s); return 0;
}
int emu_subgroup(const uint8_t* in, int g2) {
  if (g2) { G2Aff p; if (g2_decompress(p, in, false)) return -1; return g2_in_subgroup(p); }
  G1Aff p; if (g1_decompress(p, in, false)) return -1; return g1_in_subgroup(p);
}
void emu_hash_to_g2(const uint8_t* prefix, int plen, const uint8_t* msg, int mlen, const uint8_t* dst, int dlen, uint8_t* out) {
  G2Jac r; hash_to_g2(r, prefix, plen, msg, mlen, dst, dlen);
  G2Aff a; jac_to_aff(a, r); g2_compress(out, a);
}
void emu_hash_to_g1(const uint8_t* prefix, int plen, const uint8_t* msg, int mlen, const uint8_t* dst, int dlen, uint8_t* out) {
  G1Jac r; hash_to_g1(r, prefix, plen, msg, mlen, dst, dlen);
  G1Aff a; jac_to_aff(a, r); g1_compress(out, a);
}
void emu_sha256(const uint8_t* m, int n, uint8_t* out) { Sha256 s; sha256_init(s); sha256_update(s, m, n); sha256_final(s, out); }
// Miller loop of (P in G1 scaled by k to exercise the Jacobian path, Q in G2), raw and after the final exponentiation
int emu_pairing(const uint8_t* p48, const uint8_t* q96, const uint32_t* k, int nl, uint8_t* ml_out, uint8_t* fe_out) {
  G1Aff p; G2Aff q;
  if (g1_decompress(p, p48, false) || g2_decompress(q, q96, false)) return -1;
  MillerG1 mp;
  if (nl) { G1Jac pj; jac_mul_aff(pj, p, k, nl); miller_prepare(mp, pj); } else miller_prepare(mp, p);
  Fp12 f, g; miller_loop(f, mp, q); fp12_out(ml_out, f);
  final_exponentiation(g, f); fp12_out(fe_out, g);
  return 0;
}
// e(p1,q1) e(p2,q2) == 1 ?
int emu_pairing_check2(const uint8_t* p1, const uint8_t* q1, const uint8_t* p2, const uint8_t* q2) {
  G1Aff a, c; G2Aff b, d;
  if (g1_decompress(a, p1, true) || g2_decompress(b, q1, true) || g1_decompress(c, p2, true) || g2_decompress(d, q2, true)) return -1;
  MillerG1 m1, m2; miller_prepare(m1, a); miller_prepare(m2, c);
  Fp12 f, g; miller_loop(f, m1, b); miller_loop(g, m2, d); fp12_mul(f, f, g); final_exponentiation(g, f);
  return fp12_is_one(g);
}

// fused sums of products (plain Fp2 in and out): out0 = a*b + xi*(2c)*d - e*(4f) ; out1 = out0^2 (output aliasing both
// operands) ; out2 = a*k - xi*b*k  (k an Fp scalar: the two-trip mode)
void emu_sop2f(const uint8_t* in, const uint8_t* k48, uint8_t* out) {
  Fp2 v[6], one2; SFp2 s[10];  // records 0..5 inputs, 6: k, 7: one, 8: r
  SopSpaces cx = m6_spaces_line(s, 1, nullptr, nullptr);
  fone(one2); sfp2_from_fp2(s[7], one2);
  for (int i = 0; i < 6; i++) {  // unsigned form -> balanced S-form through one multiplication by 1
    fp2_in(v[i], in + 96 * i); fred(v[i], v[i]); sfp2_from_fp2(s[i], v[i]);
    const SopTerm t = {(uint8_t)i, 7, 0, 0, 0, 0, 0, 0};
    sop2f(&s[i], nullptr, &t, 1, 0, cx);
  }
  Fp kk; fp_in(kk, k48); sfp2_from_fp(s[6], kk);
  const SopTerm t3[3] = {{0, 1, 0, 0, 0, 0, 0, 0}, {2, 3, 1, 0, SOP_XI, 0, 0, 0}, {4, 5, 0, 2, SOP_NEG, 0, 0, 0}};
  sop2f(&s[8], nullptr, t3, 3, 0, cx);
  Fp2 o; fp2_from_sfp2(o, s[8]); fp2_out(out, o);
  const SopTerm sq = {8, 8, 0, 0, 0, 0, 0, 0};
  sop2f(&s[8], nullptr, &sq, 1, 0, cx);
  fp2_from_sfp2(o, s[8]); fp2_out(out + 96, o);
  const SopTerm t2[2] = {{0, 6, 0, 0, 0, 0, 0, 0}, {1, 6, 0, 0, SOP_XI | SOP_NEG, 0, 0, 0}};
  sop2f(&s[9], nullptr, t2, 2, 1, cx);
  fp2_from_sfp2(o, s[9]); fp2_out(out + 192, o);
}
// cooperative Miller loop of miller6.cuh, lanes emulated one after the other: f = prod_j ML(k_j * P_j, Q_j), j < n <= 6.
// Mirrors the two device kernels: every pair's 68 line records first (k_m6_lines), then the shared accumulator (k_m6_accum).
int emu_miller6(int n, const uint8_t* p48, const uint8_t* q96, const uint32_t* k, int nl, int two_lane, uint8_t* ml_out, uint8_t* fe_out) {
  static SLineRec lines[6][M6_STEPS][3];
  const uint64_t e = K_X_ABS;
  for (int j = 0; j < n; j++) {
    G1Aff p; G2Aff q;
    if (g1_decompress(p, p48 + 48 * j, false) || g2_decompress(q, q96 + 96 * j, false)) return -1;
    MillerG1 mp;
    if (nl) { G1Jac pj; jac_mul_aff(pj, p, k + nl * j, nl); miller_prepare(mp, pj); } else miller_prepare(mp, p);
    M6Arg arg; m6_make_arg(arg, mp, q);
    SFp2 reg[2 * M6_NREG];  // exercised with a record stride of 2
    SopSpaces cx = m6_spaces_line(reg, 2, &arg, nullptr);
    m6_init_point(cx, q);
    int step = 0;
    for (int i = 62; i >= 0; i--) {
      for (int pass = 0; pass < 2; pass++) {
        if (pass == 1 && !((e >> i) & 1)) break;
        cx.line = lines[j][step++];
        const M6Op* prog = pass == 0 ? K_M6_DBL : K_M6_ADD;
        const int nops = pass == 0 ? K_M6_DBL_N : K_M6_ADD_N;
        if (two_lane) {  // k_m6_lines: two lanes per pair, emulated one after the other around the store barrier
          for (int o = 0; o < nops; o++) {
            int32_t r0[NL], r1[NL];
            double v0 = m6_op_compute(r0, prog + o, cx, 0), v1 = m6_op_compute(r1, prog + o, cx, 1);
            m6_op_store(prog + o, cx, 0, r0, r1, v0);
            m6_op_store(prog + o, cx, 1, r1, r0, v1);
          }
        } else {
          m6_run(prog, nops, cx);
        }
      }
    }
    if (step != M6_STEPS) return -2;
  }
  SAccRec F[6], T[6];
  { SFp2 o, z; sfp2_one(o); sfp2_zero(z); sacc_from_sfp2(F[0], o); for (int c = 1; c < 6; c++) sacc_from_sfp2(F[c], z); }
  int step = 0;
  for (int i = 62; i >= 0; i--) {
    if (i != 62) {
      for (int c = 0; c < 6; c++) m6_sqr_lane(&T[c], F, c, 0);
      for (int c = 0; c < 6; c++) F[c] = T[c];
    }
    for (int pass = 0; pass < 2; pass++) {
      if (pass == 1 && !((e >> i) & 1)) break;
      for (int j = 0; j < n; j++) {
        for (int c = 0; c < 6; c++) m6_mul_line_lane(&T[c], F, lines[j][step], c, 0);
        for (int c = 0; c < 6; c++) F[c] = T[c];
      }
      step++;
    }
  }
  Fp12 f, g;
  for (int c = 0; c < 6; c++) m6_finish_lane(*fp12_coeff(f, c), F[c], c);
  fp12_out(ml_out, f);
  final_exponentiation(g, f); fp12_out(fe_out, g);
  return 0;
}
// cooperative final exponentiation of finalexp6.cuh (lanes emulated one after the other): out = f^(3 (p^12-1)/r) and the
// lane-wise "== 1" test; mul_out = f * g through the general six-lane product
int emu_final6(const uint8_t* f576, const uint8_t* g576, uint8_t* fe_out, uint8_t* mul_out) {
  static SAccRec R[FE6_NREG * 6];
  Fp12 f, g, scratch, o;
  fp12_in(f, f576); fp12_in(g, g576);
  Fe6 cx;
  cx.R = R; cx.scratch = &scratch; cx.k = 0; cx.mask = 0;
  for (int k = 0; k < 6; k++) { fe6_load_coeff(fe6_reg(cx, 0)[k], *fp12_coeff(f, k)); fe6_load_coeff(fe6_reg(cx, 1)[k], *fp12_coeff(g, k)); }
  fe6_mul(cx, 2, 0, 1);
  for (int k = 0; k < 6; k++) fe6_store_coeff(*fp12_coeff(o, k), fe6_reg(cx, 2)[k]);
  fp12_out(mul_out, o);
  fe6_final_exponentiation(cx);
  for (int k = 0; k < 6; k++) fe6_store_coeff(*fp12_coeff(o, k), fe6_reg(cx, 0)[k]);
  fp12_out(fe_out, o);
  int one = 1;
  for (int k = 0; k < 6; k++) one &= fe6_lane_is_one(cx, k) ? 1 : 0;
  return one;
}
// Fr: c = a*b mod r, d = a^-1 mod r (32-byte big-endian in and out), and the Lagrange coefficient at zero of share i
int emu_fr_ops(const uint8_t* a32, const uint8_t* b32, uint8_t* mul_out, uint8_t* inv_out) {
  uint32_t ra[8], rb[8], o[8];
  fr_raw_from_be32(ra, a32); fr_raw_from_be32(rb, b32);
  if (fr_raw_ge_r(ra) || fr_raw_ge_r(rb)) return -1;
  Fr a, b, c; fr_from_raw(a, ra); fr_from_raw(b, rb);
  fr_mul(c, a, b); fr_to_raw(o, c);
  for (int w = 0; w < 8; w++) for (int k = 0; k < 4; k++) mul_out[31 - 4 * w - k] = (uint8_t)(o[w] >> (8 * k));
  fr_inv(c, a); fr_to_raw(o, c);
  for (int w = 0; w < 8; w++) for (int k = 0; k < 4; k++) inv_out[31 - 4 * w - k] = (uint8_t)(o[w] >> (8 * k));
  return 0;
}
int emu_fr_lagrange(const uint8_t* ids32, int m, int i, uint8_t* out32) {
  uint32_t ids[8 * 64], o[8];
  if (m > 64) return -2;
  for (int j = 0; j < m; j++) fr_raw_from_be32(ids + 8 * j, ids32 + 32 * j);
  if (!fr_lagrange_at_zero(o, ids, (uint32_t)m, (uint32_t)i)) return 1;
  for (int w = 0; w < 8; w++) for (int k = 0; k < 4; k++) out32[31 - 4 * w - k] = (uint8_t)(o[w] >> (8 * k));
  return 0;
}
}
