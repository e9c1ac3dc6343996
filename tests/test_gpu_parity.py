"""GPU parity tests: the CUDA engine, called through the C ABI, against the CPU oracle and the reference's golden
vectors.  Bit-exact for every byte and status code."""
import hashlib
import json
import os
import random

import numpy as np
import pytest

pytestmark = pytest.mark.gpu

from oracle import bls_oracle as O


@pytest.fixture(scope="module")
def eng():
    import blsful_b200 as B
    e = B.Engine([0])
    yield e
    e.close()


@pytest.fixture(scope="module")
def B():
    import blsful_b200
    return blsful_b200


@pytest.fixture(scope="module")
def cpp(golden_dir):
    return json.load(open(os.path.join(golden_dir, "cpp_integration.json")))


IDENT1 = bytes([0xC0]) + bytes(47)
IDENT2 = bytes([0xC0]) + bytes(95)


def _be(v):
    return v.to_bytes(48, "big")


def test_fp_mul_matches_bigint(eng):
    rnd = random.Random(7)
    n = 20000
    vals_a = [0, 1, O.P - 1, O.P - 2, (1 << 384) % O.P, 2] + [rnd.randrange(O.P) for _ in range(n - 6)]
    vals_b = [O.P - 1, O.P - 1, O.P - 1, 1, (1 << 384) % O.P, 0] + [rnd.randrange(O.P) for _ in range(n - 6)]
    a = np.frombuffer(b"".join(_be(v) for v in vals_a), dtype=np.uint8)
    b = np.frombuffer(b"".join(_be(v) for v in vals_b), dtype=np.uint8)
    want = b"".join(_be(x * y % O.P) for x, y in zip(vals_a, vals_b))
    for variant in (0, 1):
        got = eng.fp_mul_batch(a, b, variant).tobytes()
        assert got == want, f"variant {variant}"


def test_hash_to_curve_vectors(eng, cpp):
    dst2 = b"QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_RO_"
    dst1 = b"QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_RO_"
    msgs = [b"", b"abc", b"hello", bytes(range(200))]
    got2 = eng.hash_to_curve_batch(2, msgs, dst2)
    got1 = eng.hash_to_curve_batch(1, msgs, dst1)
    for m, g2, g1 in zip(msgs, got2, got1):
        assert g2 == O.g2_serialize(O.hash_to_curve_g2(m, dst2))
        assert g1 == O.g1_serialize(O.hash_to_curve_g1(m, dst1))
    kat = eng.hash_to_curve_batch(2, [bytes.fromhex(cpp["message"])], O.sig_dst(O.G2IMPL, O.BASIC))[0]
    assert kat.hex().startswith("8dbf4d3c426badac1e66421c7d65dc01")


def test_recode_points_headers_and_subgroup(eng, B, cpp):
    pk = bytes.fromhex(cpp["signers"][0]["pk"])
    sig = bytes.fromhex(cpp["signers"][0]["sig"])
    for group, enc, deser, ser in ((1, pk, O.g1_deserialize, O.g1_serialize), (2, sig, O.g2_deserialize, O.g2_serialize)):
        L = len(enc)
        pt = deser(enc, O.MODERN)
        cases = [enc, ser(pt, O.LEGACY), (IDENT1 if group == 1 else IDENT2)]
        for b0 in (0x00, 0x20, 0x40, 0x60, 0x80, 0xA0, 0xC0, 0xE0):  # every header-bit combination on a valid x
            cases.append(bytes([(enc[0] & 0x1F) | b0]) + enc[1:])
        bad = bytearray((O.P + 5).to_bytes(48, "big") + bytes(L - 48)); bad[0] |= 0x80
        cases.append(bytes(bad))                       # x >= p
        cases.append(bytes([0xC0]) + bytes(L - 2) + b"\x01")  # infinity with non-zero body
        x = 1
        while group == 1:  # on-curve, outside the subgroup
            y = O.fp_sqrt((x ** 3 + 4) % O.P)
            if y is not None and not O.g1_in_subgroup((x, y)):
                e = bytearray(x.to_bytes(48, "big")); e[0] |= 0x80
                cases.append(bytes(e)); break
            x += 1
        for fin in (O.MODERN, O.LEGACY):
            st, outs = eng.recode_points(group, cases, fin, O.MODERN)
            for c, s, o in zip(cases, st, outs):
                try:
                    want_pt = deser(c, fin); want_st = 0
                except O.BlsError as ex:
                    want_st = ex.code
                assert s == want_st, (group, fin, c.hex()[:8], s, want_st)
                if want_st == 0:
                    assert o == ser(want_pt, O.MODERN)
            st2, outs2 = eng.recode_points(group, [enc], O.MODERN, O.LEGACY)
            assert st2[0] == 0 and outs2[0] == ser(pt, O.LEGACY)


def test_verify_batch_cpp_golden(eng, B, cpp):
    msg = bytes.fromhex(cpp["message"])
    pks = [bytes.fromhex(s["pk"]) for s in cpp["signers"]]
    sigs = [bytes.fromhex(s["sig"]) for s in cpp["signers"]]
    st = eng.verify_batch(B.Bls12381G2Impl, 0, pks, sigs, [msg] * 3)
    assert st.tolist() == [0, 0, 0]
    # rejections and error codes, each compared with the oracle
    pks2 = pks + [pks[0], pks[1], IDENT1, pks[2], IDENT1, pks[0]]
    sigs2 = sigs + [sigs[1], IDENT2, sigs[0], sigs[2], IDENT2, bytes([sigs[0][0] ^ 0x20]) + sigs[0][1:]]
    msgs2 = [msg] * 6 + [b"other", msg, msg]
    st = eng.verify_batch(B.Bls12381G2Impl, 0, pks2, sigs2, msgs2)
    want = [O.verify(O.G2IMPL, O.BASIC, O.MODERN, p, s, m) for p, s, m in zip(pks2, sigs2, msgs2)]
    assert st.tolist() == want
    assert want[:3] == [0, 0, 0] and want[3] == 1 and want[4] == 2 and want[5] == 3 and want[6] == 1
    # batches of one (the tree is a single leaf) and of two
    assert eng.verify_batch(B.Bls12381G2Impl, 0, [pks[0]], [sigs[1]], [msg]).tolist() == [1]
    assert eng.verify_batch(B.Bls12381G2Impl, 0, [pks[0]], [sigs[0]], [msg]).tolist() == [0]
    assert eng.verify_batch(B.Bls12381G2Impl, 0, [pks[0], pks[1]], [sigs[0], sigs[0]], [msg, msg]).tolist() == [0, 1]
    assert eng.verify_batch(B.Bls12381G2Impl, 0, [], [], []).tolist() == []


@pytest.mark.parametrize("impl", [2, 1])
@pytest.mark.parametrize("scheme", [0, 1, 2])
def test_testdata_sign_matches_oracle_and_verifies(eng, impl, scheme):
    rnd = random.Random(100 + impl * 10 + scheme)
    n = 3
    sks = [rnd.randrange(1, O.R) for _ in range(n)]
    msgs = [bytes(rnd.randrange(256) for _ in range(rnd.choice([0, 5, 32, 77]))) for _ in range(n)]
    import blsful_b200 as B
    data, off = B.pack_messages(msgs)
    k = np.frombuffer(b"".join(sk.to_bytes(32, "big") for sk in sks), dtype=np.uint8)
    pks, sigs = eng.testdata_sign(impl, scheme, k, data, off)
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    C = O.IMPLS[impl]
    for i in range(n):
        assert pks[i * pl:(i + 1) * pl].tobytes() == C.pk_ser(O.sk_to_pk(impl, sks[i]))
        assert sigs[i * sl:(i + 1) * sl].tobytes() == C.sig_ser(O.sign(impl, scheme, sks[i], msgs[i]))
    st = eng.verify_batch_packed(impl, scheme, pks, sigs, data, off)
    assert st.tolist() == [0] * n
    # wrong scheme => different DST/framing => invalid
    st = eng.verify_batch_packed(impl, (scheme + 1) % 3, pks, sigs, data, off)
    assert st.tolist() == [1] * n


@pytest.mark.parametrize("impl", [2, 1])
def test_verify_batch_bisection_finds_exact_bad_set(eng, B, impl):
    rnd = random.Random(5 + impl)
    n = 700
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [hashlib.sha256(b"m%d" % i).digest() for i in range(n)]
    data, off = B.pack_messages(msgs)
    pks, sigs = eng.testdata_sign(impl, 0, k, data, off)
    assert eng.verify_batch_packed(impl, 0, pks, sigs, data, off).tolist() == [0] * n
    sl = B.sig_len(impl)
    orig = sigs.copy()
    sigs = sigs.copy()
    bad = sorted(rnd.sample(range(n), 9))
    for i in bad:  # another item's signature: a valid subgroup point, but the wrong one
        src = (i + 1) % n
        sigs[i * sl:(i + 1) * sl] = orig[src * sl:(src + 1) * sl]
    st = eng.verify_batch_packed(impl, 0, pks, sigs, data, off)
    assert [i for i in range(n) if st[i] != 0] == bad
    assert all(st[i] == 1 for i in bad)
    for i in bad[:2]:
        m = msgs[i]
        assert st[i] == O.verify(impl, O.BASIC, O.MODERN, pks[i * B.pk_len(impl):(i + 1) * B.pk_len(impl)].tobytes(),
                                 sigs[i * sl:(i + 1) * sl].tobytes(), m)


@pytest.mark.parametrize("impl,n", [(2, 4500), (1, 4500), (2, 33000)])
def test_verify_batch_large_batch_bucket_msm_and_chunks(eng, B, impl, n):
    """Batches of >= 4096 items take the bucket multi-scalar multiplication for sum r_i sig_i; a failing batch falls back
    to per-group sums for the bisection.  Also covers identity / undecodable items inside cooperative groups of six."""
    rnd = random.Random(50 + impl)  # n = 33000: 11-bit windows with a narrower (9-bit) top window
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [hashlib.sha256(b"big%d" % i).digest() for i in range(n)]
    data, off = B.pack_messages(msgs)
    pks, sigs = eng.testdata_sign(impl, 0, k, data, off)
    assert eng.verify_batch_packed(impl, 0, pks, sigs, data, off).tolist() == [0] * n
    sl, pl = B.sig_len(impl), B.pk_len(impl)
    orig = sigs.copy()
    sigs = sigs.copy()
    pks = pks.copy()
    bad = sorted(rnd.sample(range(n), 5))
    for i in bad:
        src = (i + 7) % n
        sigs[i * sl:(i + 1) * sl] = orig[src * sl:(src + 1) * sl]
    ident = [11, 4002]  # identity signature -> status 2, excluded from the batch equation
    for i in ident:
        sigs[i * sl:(i + 1) * sl] = np.frombuffer(bytes([0xC0]) + bytes(sl - 1), dtype=np.uint8)
    junk = [12]  # undecodable public key -> status 4
    for i in junk:
        pks[i * pl:(i + 1) * pl] = np.frombuffer(bytes([0x80]) + bytes([0xFF]) * (pl - 1), dtype=np.uint8)
    st = eng.verify_batch_packed(impl, 0, pks, sigs, data, off)
    want = [0] * n
    for i in bad:
        want[i] = 1
    for i in ident:
        want[i] = 2
    for i in junk:
        want[i] = 4
    assert st.tolist() == want
    i = junk[0]
    assert O.verify(impl, O.BASIC, O.MODERN, pks[i * pl:(i + 1) * pl].tobytes(), sigs[i * sl:(i + 1) * sl].tobytes(), msgs[i]) == 4


def test_sum_points_golden_and_pairing_product(eng, B, cpp):
    sigs = [bytes.fromhex(s["sig"]) for s in cpp["signers"]]
    pks = [bytes.fromhex(s["pk"]) for s in cpp["signers"]]
    assert eng.sum_points(2, sigs[:2]).hex() == cpp["normal_agg_sig12"]
    agg_pk = eng.sum_public_keys(B.Bls12381G2Impl, pks[:2])
    assert agg_pk == O.g1_serialize(O.g1_add(O.g1_deserialize(pks[0]), O.g1_deserialize(pks[1])))
    # same-message multi-signature verifies with the summed key (cpp_integration_test.rs:171-179, signatures.rs:88-128)
    msg = bytes.fromhex(cpp["message"])
    st = eng.verify_batch(B.Bls12381G2Impl, 0, [agg_pk], [bytes.fromhex(cpp["normal_agg_sig12"])], [msg])
    assert st.tolist() == [0]
    assert eng.sum_points(1, []) == IDENT1 and eng.sum_points(2, [], O.LEGACY) == IDENT2
    with pytest.raises(B.BlsError) as e:
        eng.sum_points(1, [pks[0], bytes(48), pks[1]])
    assert e.value.status == 4
    # bilinearity: e(aP, bQ) * e(-abP, Q) == 1
    a, b = 123456789, 987654321
    g1 = [O.g1_serialize(O.g1_mul(O.G1_GEN, a)), O.g1_serialize(O.g1_neg(O.g1_mul(O.G1_GEN, a * b)))]
    g2 = [O.g2_serialize(O.g2_mul(O.G2_GEN, b)), O.g2_serialize(O.G2_GEN)]
    assert eng.pairing_product_is_one(g1, g2) is True
    assert eng.pairing_product_is_one(g1, [g2[0], g2[0]]) is False


@pytest.mark.parametrize("impl", [2, 1])
def test_aggregate_verify_semantics(eng, B, impl):
    rnd = random.Random(77 + impl)
    n = 40
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [b"msg-%d" % i for i in range(n)]
    data, off = B.pack_messages(msgs)
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    for scheme in (0, 1, 2):
        pks, sigs = eng.testdata_sign(impl, scheme, k, data, off)
        pk_list = [pks[i * pl:(i + 1) * pl].tobytes() for i in range(n)]
        agg = eng.sum_points(1 if impl == 1 else 2, sigs)
        assert eng.aggregate_verify_status(impl, scheme, pk_list, msgs, agg)[0] == 0
        bad_msgs = list(msgs); bad_msgs[7] = b"tampered"
        assert eng.aggregate_verify_status(impl, scheme, pk_list, bad_msgs, agg)[0] == 1
        dup = list(msgs); dup[9] = dup[3]
        st, idx = eng.aggregate_verify_status(impl, scheme, pk_list, dup, agg)
        if scheme == 0:
            assert st == 8 and idx == (3, 9)
        else:
            assert st == 1
        ident_pk = IDENT1 if impl == 2 else IDENT2
        st, idx = eng.aggregate_verify_status(impl, scheme, pk_list[:4] + [ident_pk], msgs[:5], agg)
        assert st == 3 and idx[0] == 5
        assert eng.aggregate_verify_status(impl, scheme, [], [], agg)[0] == 1
        assert eng.aggregate_verify_status(impl, scheme, pk_list, msgs, IDENT2 if impl == 2 else IDENT1)[0] == 2
    # small case against the oracle end to end
    pks, sigs = eng.testdata_sign(impl, 0, k[:3 * 32], *B.pack_messages(msgs[:3]))
    pk_list = [pks[i * pl:(i + 1) * pl].tobytes() for i in range(3)]
    agg = eng.sum_points(1 if impl == 1 else 2, sigs)
    assert O.aggregate_verify(impl, O.BASIC, O.MODERN, pk_list, msgs[:3], agg)[0] == 0


# ---- secure aggregation (reference src/secure_aggregation.rs; tests/secure_aggregation_test.rs, cpp_integration_test.rs) ----
@pytest.fixture(scope="module")
def sec57(golden_dir):
    return json.load(open(os.path.join(golden_dir, "secure_57.json")))


def test_verify_secure_production_57_key_vector(eng, B, sec57):
    # secure_aggregation_test.rs:143-235: 57 production keys, one aggregate signature
    pks = [bytes.fromhex(k) for k in sec57["keys"]]
    sig = bytes.fromhex(sec57["sig"])
    msg = bytes.fromhex(sec57["message"])
    st = eng.verify_secure_batch(B.Bls12381G2Impl, 0, [pks, pks[::-1], pks[:-1], pks], [sig, sig, sig, sig],
                                 [msg, msg, msg, msg + b"x"])
    assert st.tolist() == [0, 0, 1, 1]  # order independent; missing key / wrong message fail (secure_aggregation.rs:562-602)


def test_secure_aggregate_then_verify_cpp_golden(eng, B, cpp):
    # cpp_integration_test.rs:86-192
    msg = bytes.fromhex(cpp["message"])
    pks = [bytes.fromhex(s["pk"]) for s in cpp["signers"]]
    sigs = [bytes.fromhex(s["sig"]) for s in cpp["signers"]]
    sets = [pks[:2], pks[:3], pks[:1]]
    st, aggs = eng.aggregate_secure_batch(B.Bls12381G2Impl, sets, [sigs[:2], sigs[:3], sigs[:1]])
    assert st.tolist() == [0, 0, 0]
    for ks, ss, a in zip(sets, [sigs[:2], sigs[:3], sigs[:1]], aggs):
        assert (0, a) == O.aggregate_secure(O.G2IMPL, O.MODERN, ks, ss)
    normal = bytes.fromhex(cpp["normal_agg_sig12"])
    st = eng.verify_secure_batch(B.Bls12381G2Impl, 0, sets + [pks[:2]], aggs + [normal], [msg] * 4)
    assert st.tolist() == [0, 0, 0, 1]  # the plain (rogue-key-prone) aggregate must fail (:169-192)


@pytest.mark.parametrize("impl,fmt", [(2, 1), (2, 0), (1, 1)])
def test_secure_semantics_against_oracle(eng, B, impl, fmt):
    rnd = random.Random(31 + impl * 7 + fmt)
    C = O.IMPLS[impl]
    msg = b"quorum message"
    sks = [rnd.randrange(1, O.R) for _ in range(5)]
    pk_pts = [O.sk_to_pk(impl, sk) for sk in sks]
    pks = [C.pk_ser(p, fmt) for p in pk_pts]
    ident_sig = bytes([0xC0]) + bytes(B.sig_len(impl) - 1)
    ident_pk = bytes([0xC0]) + bytes(B.pk_len(impl) - 1)
    for scheme in ((0, 1, 2) if (impl, fmt) == (2, 1) else ((impl + fmt) % 3,)):
        # verify_secure hashes the bare message under the scheme's DST (no pk prefix even for MessageAugmentation)
        sigs = [C.sig_ser(C.sig_mul(C.hash(msg, O.sig_dst(impl, scheme)), sk), fmt) for sk in sks]
        dup_keys = pks[:3] + [pks[1]]
        dup_sigs = sigs[:3] + [sigs[0]]  # the duplicate key must reuse the FIRST match's signature (sigs[1]), not this one
        key_sets = [pks, pks[:4], dup_keys, []]
        sig_sets = [sigs, sigs[:4], dup_sigs, []]
        st, aggs = eng.aggregate_secure_batch(impl, key_sets, sig_sets, fmt)
        for ks, ss, s, a in zip(key_sets, sig_sets, st, aggs):
            want_st, want = O.aggregate_secure(impl, fmt, ks, ss)
            assert s == want_st and a == want
        # verify: good, shuffled, wrong message, subset of keys, empty set with identity / non-identity signature,
        # identity signature with keys, one undecodable key, one undecodable signature
        bad_key = bytes(B.pk_len(impl))
        bad_sig = bytes(B.sig_len(impl))
        cases = [
            (pks, aggs[0], msg), (pks[::-1], aggs[0], msg), (pks, aggs[0], msg + b"!"), (pks[:4], aggs[0], msg),
            ([], ident_sig, msg), ([], aggs[0], msg), (pks, ident_sig, msg), (pks[:2] + [bad_key], aggs[0], msg),
            (pks, bad_sig, msg), ([ident_pk], aggs[0], msg), (pks[:4], aggs[1], msg),
        ]
        got = eng.verify_secure_batch(impl, scheme, [c[0] for c in cases], [c[1] for c in cases], [c[2] for c in cases], fmt)
        want = [O.verify_secure(impl, scheme, fmt, c[0], c[1], c[2]) for c in cases]
        assert got.tolist() == want, (impl, fmt, scheme)
        assert want[0] == 0 and want[1] == 0 and want[4] == 0 and want[10] == 0
    if impl == 2:
        # cross-mode must fail: keys/signature of one format checked under the other coefficient derivation (legacy_test.rs:109-171)
        other = 1 - fmt
        st2, agg2 = eng.aggregate_secure_batch(impl, [[C.pk_ser(p, other) for p in pk_pts]],
                                               [[C.sig_ser(C.sig_deser(s, fmt), other) for s in sigs]], other)
        re = C.sig_ser(C.sig_deser(agg2[0], other), fmt)
        assert eng.verify_secure_batch(impl, 2, [pks], [re], [msg], fmt).tolist() == \
            [O.verify_secure(impl, 2, fmt, pks, re, msg)]


def test_combine_shares_matches_oracle_and_golden_signature(eng, B, cpp):
    """blsgpu_combine_shares_batch (Signature::from_shares / PublicKey::from_shares) against the oracle; one set is the
    Shamir sharing of a golden secret key and must combine to the reference's golden signature bytes."""
    rnd = random.Random(91)
    msg = bytes.fromhex(cpp["message"])
    s = cpp["signers"][1]
    sk = int(s["sk"], 16)

    def shamir(secret, t, ids):
        coef = [secret] + [rnd.randrange(O.R) for _ in range(t - 1)]
        return [sum(c * pow(x, k, O.R) for k, c in enumerate(coef)) % O.R for x in ids]

    H = O.hash_to_curve_g2(msg, O.sig_dst(O.G2IMPL, O.BASIC))
    ids = [3, 1, 7, 2 ** 200 + 11, O.R - 1]
    sks = shamir(sk, 4, ids)
    sig_sh = [x.to_bytes(32, "big") + O.g2_serialize(O.g2_mul(H, k)) for x, k in zip(ids, sks)]
    pk_sh = [x.to_bytes(32, "big") + O.g1_serialize(O.g1_mul(O.G1_GEN, k)) for x, k in zip(ids, sks)]
    big_ids = list(range(1, 41))  # more than one chunk of 16
    big = [x.to_bytes(32, "big") + O.g2_serialize(O.g2_mul(O.G2_GEN, rnd.randrange(1, O.R))) for x in big_ids]
    sets2 = [
        sig_sh[:4], sig_sh, sig_sh[1:5], sig_sh[:2], big,
        sig_sh[:1], [],                                                  # fewer than two shares
        [bytes(32) + sig_sh[0][32:], sig_sh[1]],                         # zero identifier
        [sig_sh[0], sig_sh[0][:32] + sig_sh[1][32:], sig_sh[2]],         # duplicate identifier
        [O.R.to_bytes(32, "big") + sig_sh[0][32:], sig_sh[1]],           # identifier >= r
        [sig_sh[0], sig_sh[1][:32] + bytes(96)],                         # undecodable point
        [sig_sh[0], bytes(32) + bytes(96)],                              # zero identifier AND bad point: parse error first
    ]
    st, outs = eng.combine_shares_batch(2, sets2)
    for ss, got_st, got in zip(sets2, st, outs):
        want_st, want = O.combine_shares(2, ss)
        assert got_st == want_st, (len(ss), got_st, want_st)
        assert got == (want if want_st == 0 else bytes(96))
    assert outs[0].hex() == s["sig"] and outs[1].hex() == s["sig"] and outs[2].hex() == s["sig"]
    st, outs = eng.combine_shares_batch(1, [pk_sh[:4], pk_sh[::-1], pk_sh[:1]])
    assert st.tolist() == [0, 0, 11] and outs[0].hex() == s["pk"] and outs[1].hex() == s["pk"]
    # the combined signature verifies under the combined key
    assert eng.verify_batch(2, 0, [outs[0]], [bytes.fromhex(s["sig"])], [msg]).tolist() == [0]


def test_verify_share_batch_against_oracle(eng, B, cpp):
    """blsgpu_verify_share_batch (PublicKeyShare::verify on raw share records): Shamir shares of a golden key verify,
    identifiers are ignored by the check but must be canonical scalars, every other outcome is verify's."""
    rnd = random.Random(17)
    msg = bytes.fromhex(cpp["message"])
    sk = int(cpp["signers"][0]["sk"], 16)
    coef = [sk] + [rnd.randrange(O.R) for _ in range(2)]
    ids = [1, 2 ** 130 + 5, O.R - 1]
    sks = [sum(c * pow(x, k, O.R) for k, c in enumerate(coef)) % O.R for x in ids]
    for impl, scheme in [(2, 0), (1, 1)]:
        C = O.IMPLS[impl]
        pkv = [C.pk_ser(C.pk_mul(C.pk_gen, k), O.MODERN) for k in sks]
        sgv = [C.sig_ser(O.sign(impl, scheme, k, msg), O.MODERN) for k in sks]
        pk_sh = [x.to_bytes(32, "big") + v for x, v in zip(ids, pkv)]
        sg_sh = [x.to_bytes(32, "big") + v for x, v in zip(ids, sgv)]
        cases = list(zip(pk_sh, sg_sh))
        cases.append((pk_sh[0], sg_sh[1]))                                         # another share's signature
        cases.append((pk_sh[0], (7).to_bytes(32, "big") + sgv[0]))                 # identifiers differ: not looked at
        cases.append((O.R.to_bytes(32, "big") + pkv[0], sg_sh[0]))                 # identifier >= r
        cases.append((pk_sh[0], (2 ** 256 - 1).to_bytes(32, "big") + sgv[0]))
        cases.append((pk_sh[0][:32] + bytes(len(pkv[0])), sg_sh[0]))               # undecodable value
        cases.append((pk_sh[0], sg_sh[0][:32] + bytes([0xC0]) + bytes(len(sgv[0]) - 1)))   # identity signature share
        got = eng.verify_share_batch(impl, scheme, [c[0] for c in cases], [c[1] for c in cases], [msg] * len(cases))
        want = [O.verify_share(impl, scheme, a, b, msg) for a, b in cases]
        assert got.tolist() == want, (impl, scheme)
        assert want[:3] == [0] * 3 and want[3] == 1 and want[4] == 0 and want[5:8] == [4, 4, 4] and want[8] == 2


def test_verify_batch_wire_mixed_schemes(eng, B, cpp):
    """blsgpu_verify_batch_wire: serde_bare tagged signatures (signature.rs:112-126), schemes mixed in one call; every
    status equals the oracle's verify under the scheme named by the tag."""
    rnd = random.Random(17)
    for impl in (2, 1):
        C = O.IMPLS[impl]
        items = []
        for i in range(14):
            sk = rnd.randrange(1, O.R)
            scheme = i % 3
            msg = b"wire message %d" % i
            pk = C.pk_ser(O.sk_to_pk(impl, sk), O.MODERN)
            sig = C.sig_ser(O.sign(impl, scheme, sk, msg), O.MODERN)
            tag = scheme
            if i == 5:
                tag = (scheme + 1) % 3      # valid point, wrong scheme tag: verifies under the tag's scheme -> invalid
            if i == 7:
                tag = 3                     # unknown tag
            if i == 9:
                sig = bytes([0xC0]) + bytes(len(sig) - 1)   # identity signature
            items.append((pk, bytes([tag]) + sig, msg, tag))
        st = eng.verify_batch_wire(impl, [x[0] for x in items], [x[1] for x in items], [x[2] for x in items])
        want = [4 if t > 2 else O.verify(impl, t, O.MODERN, pk, ts[1:], m) for pk, ts, m, t in items]
        assert st.tolist() == want
        assert want[5] == 1 and want[7] == 4 and want[9] == 2 and want[0] == 0 and want[1] == 0 and want[2] == 0


def test_pairing_check_batch_against_oracle(eng, B):
    """blsgpu_pairing_check_batch: the 2-pairing public checks of the reference (sign_crypt.rs:69-77,192-207) as sets of
    pairs; compared with the oracle's pairing product, incl. identity points and an undecodable point."""
    rnd = random.Random(23)
    g1, g2 = O.G1_GEN, O.G2_GEN
    a, b, c = (rnd.randrange(1, O.R) for _ in range(3))
    s1, s2 = O.g1_serialize, O.g2_serialize
    neg_g1 = O.g1_neg(g1)
    sets = [
        [(s1(O.g1_mul(g1, a)), s2(O.g2_mul(g2, b))), (s1(neg_g1), s2(O.g2_mul(g2, a * b % O.R)))],        # e(aG,bH) e(-G,abH) = 1
        [(s1(O.g1_mul(g1, a)), s2(O.g2_mul(g2, b))), (s1(neg_g1), s2(O.g2_mul(g2, (a * b + 1) % O.R)))],  # != 1
        [(s1(O.g1_mul(g1, a)), s2(O.g2_mul(g2, b))), (s1(O.g1_mul(g1, c)), s2(g2)),
         (s1(neg_g1), s2(O.g2_mul(g2, (a * b + c) % O.R)))],                                                # three pairs
        [(IDENT1, s2(g2)), (s1(g1), IDENT2)],                                                               # identities only: 1
        [],                                                                                                 # empty product: 1
        [(bytes(48), s2(g2)), (s1(g1), s2(g2))],                                                            # undecodable point
    ]
    ok, st = eng.pairing_check_batch(sets)
    want = []
    for ps in sets[:5]:
        want.append(1 if O.pairing_product_is_one([(_g1(p), _g2(q)) for p, q in ps]) else 0)
    assert ok.tolist() == want + [0]
    assert st.tolist() == [0, 0, 0, 0, 0, 4]
    assert want == [1, 0, 1, 1, 1]


def _g1(b):
    pt, st = O._decode(O.g1_deserialize, b, O.MODERN)
    assert st == 0
    return pt


def _g2(b):
    pt, st = O._decode(O.g2_deserialize, b, O.MODERN)
    assert st == 0
    return pt


def test_verify_batch_in_several_miller_passes(eng, B, monkeypatch):
    """Batches beyond one pass of the Miller kernels (1,048,320 items) go through in chunks with two line buffers on two
    streams; the chunk size is overridden here so that a small batch takes the same path (5 passes, last one partial)."""
    rnd = random.Random(61)
    n = 5000
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [hashlib.sha256(b"chunk%d" % i).digest() for i in range(n)]
    data, off = B.pack_messages(msgs)
    pks, sigs = eng.testdata_sign(2, 0, k, data, off)
    monkeypatch.setenv("BLSGPU_M6_CHUNK", "1200")
    assert eng.verify_batch_packed(2, 0, pks, sigs, data, off).tolist() == [0] * n
    bad = sorted(rnd.sample(range(n), 4)) + [1199, 1200, 4999]
    bad = sorted(set(bad))
    s2 = sigs.copy().reshape(n, 96)
    s2[bad] = s2[[(i + 3) % n for i in bad]]
    st = eng.verify_batch_packed(2, 0, pks, s2.reshape(-1), data, off)
    assert [i for i in range(n) if st[i]] == bad and all(st[i] == 1 for i in bad)


def test_context_on_two_devices_shards_the_batch(eng, B):
    """blsgpu_ctx_create with several devices: contiguous slices, one host thread per device, no collective.
    The status vector must equal the single-device one, with bad items in every slice and ragged message lengths."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 2 * 4096 + 37
    rnd = random.Random(77)
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [hashlib.sha256(b"two%d" % i).digest()[:1 + i % 32] for i in range(n)]
    data, off = B.pack_messages(msgs)
    pks, sigs = eng.testdata_sign(2, 1, k, data, off)       # MessageAugmentation: the pk prefix travels with each slice
    sigs = sigs.copy()
    bad = [0, 5, 4100, 4101, n - 1]
    for i in bad:
        sigs[i * 96:(i + 1) * 96] = sigs[(i + 9) % n * 96:((i + 9) % n + 1) * 96].copy()
    want = eng.verify_batch_packed(2, 1, pks, sigs, data, off).tolist()
    assert [i for i, s in enumerate(want) if s] == bad
    e2 = B.Engine([0, 1])
    try:
        before = e2.launch_count()
        assert e2.verify_batch_packed(2, 1, pks, sigs, data, off).tolist() == want
        assert e2.launch_count() - before > 2 * 20            # both devices launched their own pipeline
        # a batch too small to shard stays on the first device
        before = e2.launch_count()
        assert e2.verify_batch_packed(2, 1, pks[:48 * 64], sigs[:96 * 64], data, off[:65]).tolist() == want[:64]
        assert 0 < e2.launch_count() - before < 2 * 20 + 40
    finally:
        e2.close()


def test_kernel_event_timings_cover_the_hot_kernels(eng, B, cpp):
    """blsgpu_last_kernel_ms: one CUDA-event pair per launch of each hot kernel of the last verify call (what bench.py's
    per-kernel roofline reads)."""
    n = 600
    rnd = random.Random(3)
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    data, off = B.pack_messages([b"t%d" % i for i in range(n)])
    pks, sigs = eng.testdata_sign(2, 0, k, data, off)
    assert eng.verify_batch_packed(2, 0, pks, sigs, data, off).tolist() == [0] * n
    km = eng.last_kernel_ms()
    assert list(km) == B.KERNELS
    for name, (ms, launches) in km.items():
        assert launches == 1 and 0 < ms < 1000, (name, ms, launches)
    stages = eng.last_stage_ms()
    assert km["k_hash"][0] + km["k_clear_cofactor"][0] + km["k_to_affine_batch"][0] <= stages["hash_to_curve"] * 1.05 + 0.05
