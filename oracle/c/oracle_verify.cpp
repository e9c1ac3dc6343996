// TEST INFRASTRUCTURE - the timed CPU leg of bench.py (cpu_baseline / --impl reference), nothing else.
//
// The reference's per-signature verification path executed on host cores, call for call:
//   Signature::verify            reference src/signature.rs:130-138  -> BlsSignature{Basic,MessageAugmentation,Pop}::verify
//                                (src/traits/sig_basic.rs:36-38, sig_aug.rs:20-24 prepends pk.to_bytes(), sig_pop.rs:37-39)
//   core_verify                  src/traits/sig_core.rs:120-146: reject identity signature, then identity public key,
//                                H = hash_to_point(msg, dst), accept iff pairing(&[(H, pk), (sig, -g)]) is the Gt identity
//   pairing (G2Impl)             src/impls/g2.rs:36-38 -> src/helpers.rs:41-63: one Miller loop per pair, ONE final
//                                exponentiation per signature (nothing is batched or amortised in the reference)
//   parsing                      src/public_key.rs:55-75, src/signature.rs:120-126: from_compressed = curve + subgroup check
//
// blsful itself cannot be built here (no cargo; blstrs_plus / blst are un-vendored, Cargo.toml:20-28).  This port runs
// the SAME field / curve / hash-to-curve / pairing headers as the engine (agora-blsful_b200/csrc/*.cuh, which compile for
// the host), with g++ -O2 and no assembly: it is a timing restatement, NOT the independent checker - that is
// oracle/bls_oracle.py (big integers, pinned to the reference's golden vectors), against which tests/test_c_oracle.py
// validates this file.  The product never links, loads or calls it.
#include <cstring>
#include <thread>
#include <vector>

#include "../../agora-blsful_b200/csrc/pairing.cuh"
#include "../../agora-blsful_b200/csrc/h2c.cuh"

using namespace bls;

static const char* dst_of(int scheme) {  // src/impls/g2.rs:107-118
  return scheme == 0 ? "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
       : scheme == 1 ? "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_AUG_"
                     : "BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_POP_";
}

extern "C" int oracle_verify_g2impl(int scheme, const uint8_t* pk48, const uint8_t* sig96, const uint8_t* msg, size_t mlen) {
  G1Aff pk;
  G2Aff sig;
  uint8_t st = g1_decompress(pk, pk48, true);
  if (st) return st;
  st = g2_decompress(sig, sig96, true);
  if (st) return st;
  if (sig.inf) return ST_SIG_IDENTITY;  // sig_core.rs:126-130
  if (pk.inf) return ST_PK_IDENTITY;    // sig_core.rs:131-135
  const char* dst = dst_of(scheme);
  G2Jac hj;
  hash_to_g2(hj, pk48, scheme == 1 ? 48 : 0, msg, (uint32_t)mlen, reinterpret_cast<const uint8_t*>(dst), (uint32_t)strlen(dst));
  G2Aff h;
  jac_to_aff(h, hj);
  // pairing(&[(H, pk), (sig, -g)])
  G1Aff ng;
  fp_set(ng.x, K_G1X);
  fp_set(ng.y, K_G1Y);
  ng.inf = 0;
  fp_neg(ng.y, ng.y);
  MillerG1 m1, m2;
  miller_prepare(m1, pk);
  miller_prepare(m2, ng);
  Fp12 f, g, e;
  miller_loop(f, m1, h);
  miller_loop(g, m2, sig);
  fp12_mul(f, f, g);
  final_exponentiation(e, f);
  return fp12_is_one(e) ? ST_OK : ST_INVALID_SIGNATURE;  // sig_core.rs:138-145
}

// n independent verifications over `threads` host threads (one contiguous slice each); messages packed with offsets
extern "C" void oracle_verify_many_g2impl(int scheme, size_t n, const uint8_t* pks, const uint8_t* sigs, const uint8_t* msgs,
                                          const uint64_t* off, uint8_t* status, int threads) {
  if (threads < 1) threads = 1;
  std::vector<std::thread> pool;
  for (int t = 0; t < threads; t++) {
    pool.emplace_back([=]() {
      const size_t lo = n * t / threads, hi = n * (t + 1) / threads;
      for (size_t i = lo; i < hi; i++)
        status[i] = (uint8_t)oracle_verify_g2impl(scheme, pks + 48 * i, sigs + 96 * i, msgs + off[i], (size_t)(off[i + 1] - off[i]));
    });
  }
  for (auto& th : pool) th.join();
}
