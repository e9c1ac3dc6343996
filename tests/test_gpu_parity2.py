"""GPU parity tests, second set (round 2): the rejection paths of the decoders on the device, the Legacy format through
every batch entry point, proofs of possession, the random-linear-combination modes.  Every expectation comes from the
big-int oracle (oracle/bls_oracle.py) or from the reference's golden vectors; everything goes through the C ABI."""
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


def ident(length):
    return bytes([0xC0]) + bytes(length - 1)


# ---- adversarial encodings (module-level so that several tests share the search) ------------------------------------
def _g1_outside_subgroup():
    x = 1
    while True:
        y = O.fp_sqrt((x ** 3 + 4) % O.P)
        if y is not None and not O.g1_in_subgroup((x, y)):
            return (x, y)
        x += 1


def _g2_outside_subgroup():
    k = 1
    while True:
        x = (k, 1)
        y = O.f2_sqrt(O.f2_add(O.f2_mul(O.f2_sqr(x), x), O.B2))
        if y is not None and not O.g2_in_subgroup((x, y)):
            return (x, y)
        k += 1


def _g1_off_curve_x():
    x = 1
    while O.fp_sqrt((x ** 3 + 4) % O.P) is not None:
        x += 1
    return x


def _g2_off_curve_x():
    k = 1
    while O.f2_sqrt(O.f2_add(O.f2_mul(O.f2_sqr((k, 2)), (k, 2)), O.B2)) is not None:
        k += 1
    return (k, 2)


@pytest.fixture(scope="module")
def bad_points():
    p1, p2 = _g1_outside_subgroup(), _g2_outside_subgroup()
    assert O.g1_on_curve(p1) and O.g2_on_curve(p2)
    x1, x2 = _g1_off_curve_x(), _g2_off_curve_x()
    e1 = bytearray(x1.to_bytes(48, "big")); e1[0] |= 0x80
    e2 = bytearray(x2[1].to_bytes(48, "big") + x2[0].to_bytes(48, "big")); e2[0] |= 0x80
    return {
        1: {"outside": O.g1_compress_modern(p1), "off_curve": bytes(e1)},
        2: {"outside": O.g2_compress_modern(p2), "off_curve": bytes(e2)},
    }


def test_hash_to_curve_hello_full_known_answer(eng, cpp):
    """The reference's golden triples pin H("hello") completely: sig = sk * H, so H = sk^-1 * sig for every signer
    (cpp_integration_test.rs:19-82).  All 96 bytes are compared (round 1 compared a 16-byte prefix)."""
    msg = bytes.fromhex(cpp["message"])
    got = eng.hash_to_curve_batch(2, [msg], O.sig_dst(O.G2IMPL, O.BASIC))[0]
    for s in cpp["signers"]:
        sk = int(s["sk"], 16)
        h = O.g2_mul(O.g2_deserialize(bytes.fromhex(s["sig"])), pow(sk, -1, O.R))
        assert got == O.g2_serialize(h)
    assert got.hex() == ("8dbf4d3c426badac1e66421c7d65dc017c05fb7631833f3c9a72f531bedf7995f2309d2fd6831018c83de0c27b6a10c8"
                         "10946937ad15674b2f3976d10f50ae5a66a07f5da23a4f177870702d0dbf8463225a493a8c221032e15d445afeac748a")


@pytest.mark.parametrize("group", [1, 2])
def test_decoder_rejects_points_outside_the_subgroup_and_off_the_curve(eng, B, bad_points, group):
    """k_subgroup_check's REJECT path and the square-root failure of k_decode, for both groups and both formats
    (from_compressed: legacy.rs:107,117,151,161 -> DeserializationError)."""
    deser, ser = (O.g1_deserialize, O.g1_serialize) if group == 1 else (O.g2_deserialize, O.g2_serialize)
    good = ser(O.g1_mul(O.G1_GEN, 5) if group == 1 else O.g2_mul(O.G2_GEN, 5))
    enc = [bad_points[group]["outside"], bad_points[group]["off_curve"], good]
    enc += [bytes([e[0] ^ 0x20]) + e[1:] for e in enc]           # the other y
    for fin in (O.MODERN, O.LEGACY):
        cases = enc if fin == O.MODERN else [O.modern_to_legacy(e) for e in enc]
        st, outs = eng.recode_points(group, cases, fin, O.MODERN)
        want = []
        for c in cases:
            try:
                deser(c, fin); want.append(0)
            except O.BlsError as ex:
                want.append(ex.code)
        assert st.tolist() == want
        assert want == [4, 4, 0, 4, 4, 0]
        # many copies in one launch: every lane of whole warps takes the reject path, mixed with accepting lanes
        many = (cases * 700)[:4099]
        st, _ = eng.recode_points(group, many, fin, O.MODERN)
        assert st.tolist() == (want * 700)[:4099]


@pytest.mark.parametrize("impl,fmt", [(2, 1), (2, 0), (1, 1), (1, 0)])
def test_verify_batch_with_adversarial_encodings_inside_a_large_batch(eng, B, bad_points, impl, fmt):
    """>= 4096 items (bucket path, cooperative groups of six) with out-of-subgroup / off-curve public keys and signatures,
    in the given serialization format: statuses equal the oracle's, neighbours are unaffected."""
    rnd = random.Random(900 + impl * 10 + fmt)
    n = 4200
    C = O.IMPLS[impl]
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    pkg, sgg = (1, 2) if impl == 2 else (2, 1)
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [hashlib.sha256(b"adv%d" % i).digest() for i in range(n)]
    data, off = B.pack_messages(msgs)
    pks, sigs = eng.testdata_sign(impl, 0, k, data, off)
    if fmt == O.LEGACY:
        st, p = eng.recode_points(pkg, pks, O.MODERN, O.LEGACY); assert int(st.max()) == 0
        pks = np.frombuffer(b"".join(p), dtype=np.uint8)
        st, s = eng.recode_points(sgg, sigs, O.MODERN, O.LEGACY); assert int(st.max()) == 0
        sigs = np.frombuffer(b"".join(s), dtype=np.uint8)
    pks, sigs = pks.copy(), sigs.copy()
    conv = (lambda e: e) if fmt == O.MODERN else O.modern_to_legacy
    plant = {
        5: ("pk", conv(bad_points[pkg]["outside"])), 6: ("sig", conv(bad_points[sgg]["outside"])),
        7: ("pk", conv(bad_points[pkg]["off_curve"])), 8: ("sig", conv(bad_points[sgg]["off_curve"])),
        4100: ("sig", conv(bad_points[sgg]["outside"])), 4101: ("pk", conv(bad_points[pkg]["outside"])),
        2000: ("sig", ident(sl)), 2001: ("pk", ident(pl)),
        2002: ("sig", bytes([sigs[2002 * sl] ^ (0x20 if fmt == O.MODERN else 0x80)]) + sigs[2002 * sl + 1:2003 * sl].tobytes()),  # -sig
    }
    for i, (which, enc) in plant.items():
        if which == "pk":
            pks[i * pl:(i + 1) * pl] = np.frombuffer(enc, dtype=np.uint8)
        else:
            sigs[i * sl:(i + 1) * sl] = np.frombuffer(enc, dtype=np.uint8)
    st = eng.verify_batch_packed(impl, 0, pks, sigs, data, off, fmt)
    want = [0] * n
    for i in plant:
        want[i] = O.verify(impl, O.BASIC, fmt, pks[i * pl:(i + 1) * pl].tobytes(), sigs[i * sl:(i + 1) * sl].tobytes(), msgs[i])
    assert st.tolist() == want
    assert [want[i] for i in (5, 6, 7, 8, 4100, 4101, 2000, 2001, 2002)] == [4, 4, 4, 4, 4, 4, 2, 3, 1]
    for i in (0, 4, 9, 4099, n - 1):   # untouched neighbours, checked one by one
        assert O.verify(impl, O.BASIC, fmt, pks[i * pl:(i + 1) * pl].tobytes(), sigs[i * sl:(i + 1) * sl].tobytes(), msgs[i]) == 0


@pytest.mark.parametrize("impl", [2, 1])
def test_legacy_format_through_every_batch_entry_point(eng, B, impl):
    """format = Legacy through verify_batch, aggregate_verify and sum_points (legacy_test.rs:70-171,
    legacy_comprehensive_test.rs): results equal the oracle's under the same format, including the cross-format traps
    of legacy.rs:39-67 (a Modern encoding with y-flag 0 decodes as -P under Legacy; one with y-flag 1 is a
    LegacyFormatError; a Legacy encoding with sign 0 fails the Modern decoder)."""
    rnd = random.Random(300 + impl)
    C = O.IMPLS[impl]
    n = 6
    sks = [rnd.randrange(1, O.R) for _ in range(n)]
    msgs = [b"legacy message %d" % i for i in range(n)]
    pk_pts = [O.sk_to_pk(impl, sk) for sk in sks]
    for scheme in (0, 1, 2):
        m = n if scheme == 0 else 2
        sig_pts = [O.sign(impl, scheme, sk, mm) for sk, mm in zip(sks[:m], msgs[:m])]
        pk_m, sg_m = [C.pk_ser(p, O.MODERN) for p in pk_pts[:m]], [C.sig_ser(s, O.MODERN) for s in sig_pts]
        pk_l, sg_l = [C.pk_ser(p, O.LEGACY) for p in pk_pts[:m]], [C.sig_ser(s, O.LEGACY) for s in sig_pts]
        assert eng.verify_batch(impl, scheme, pk_l, sg_l, msgs[:m], O.LEGACY).tolist() == [0] * m
        assert O.verify(impl, scheme, O.LEGACY, pk_l[0], sg_l[0], msgs[0]) == 0
        if scheme != 0:
            continue
        # Modern bytes handed to the Legacy decoder and the reverse: whatever the oracle says, item by item
        seen = set()
        for fmt, pk_x, sg_x in ((O.LEGACY, pk_m, sg_m), (O.MODERN, pk_l, sg_l), (O.LEGACY, pk_l, sg_m), (O.LEGACY, pk_m, sg_l)):
            got = eng.verify_batch(impl, scheme, pk_x, sg_x, msgs, fmt).tolist()
            want = [O.verify(impl, scheme, fmt, p, s, mm) for p, s, mm in zip(pk_x, sg_x, msgs)]
            assert got == want, (impl, scheme, fmt)
            seen |= set(want)
        assert seen >= {1, 4, 5} or seen >= {4, 5}
    # aggregate_verify and sum_points in Legacy
    sig_pts = [O.sign(impl, 0, sk, m) for sk, m in zip(sks[:6], msgs[:6])]
    sg_l = [C.sig_ser(s, O.LEGACY) for s in sig_pts]
    pk_l = [C.pk_ser(p, O.LEGACY) for p in pk_pts[:6]]
    grp_sig, grp_pk = (2, 1) if impl == 2 else (1, 2)
    agg = eng.sum_points(grp_sig, sg_l, O.LEGACY)
    want_st, want_agg, _ = O.sum_points(grp_sig, O.LEGACY, sg_l)
    assert want_st == 0 and agg == want_agg
    assert eng.sum_points(grp_pk, pk_l, O.LEGACY) == O.sum_points(grp_pk, O.LEGACY, pk_l)[1]
    assert eng.aggregate_verify_status(impl, 0, pk_l, msgs[:6], agg, O.LEGACY)[0] == \
        O.aggregate_verify(impl, O.BASIC, O.LEGACY, pk_l, msgs[:6], agg)[0] == 0
    # the Modern bytes of the same aggregate under Legacy: the oracle decides (sign flip or format error)
    agg_m = C.sig_ser(C.sig_deser(agg, O.LEGACY), O.MODERN)
    assert eng.aggregate_verify_status(impl, 0, pk_l, msgs[:6], agg_m, O.LEGACY)[0] == \
        O.aggregate_verify(impl, O.BASIC, O.LEGACY, pk_l, msgs[:6], agg_m)[0] != 0
    with pytest.raises(B.BlsError) as e:
        eng.sum_points(grp_sig, [C.sig_ser(s, O.MODERN) for s in sig_pts], O.LEGACY)
    assert e.value.status == O.sum_points(grp_sig, O.LEGACY, [C.sig_ser(s, O.MODERN) for s in sig_pts])[0]


@pytest.mark.parametrize("impl,fmt", [(2, 1), (2, 0), (1, 1), (1, 0)])
def test_pop_verify_batch_against_oracle(eng, B, bad_points, impl, fmt):
    """blsgpu_pop_verify_batch = n x ProofOfPossession::verify (proof_of_possession.rs:77-81 -> pop_verify,
    sig_pop.rs:61-70: core_verify(pk, proof, pk.to_bytes(), POP_DST)); the message is ALWAYS the Modern key bytes."""
    rnd = random.Random(40 + impl * 10 + fmt)
    C = O.IMPLS[impl]
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    pkg, sgg = (1, 2) if impl == 2 else (2, 1)
    sks = [rnd.randrange(1, O.R) for _ in range(9)]
    pk_pts = [O.sk_to_pk(impl, sk) for sk in sks]
    pops = [C.sig_mul(C.hash(C.pk_ser(p, O.MODERN), O.pop_dst(impl)), sk) for p, sk in zip(pk_pts, sks)]
    pk_b = [C.pk_ser(p, fmt) for p in pk_pts]
    pop_b = [C.sig_ser(s, fmt) for s in pops]
    conv = (lambda e: e) if fmt == O.MODERN else O.modern_to_legacy
    cases = list(zip(pk_b, pop_b))
    cases += [
        (pk_b[0], pop_b[1]),                                              # another key's proof
        (pk_b[1], C.sig_ser(O.sign(impl, O.POP, sks[1], C.pk_ser(pk_pts[1], O.MODERN)), fmt)),  # a SIGNATURE over the key bytes: wrong DST
        (ident(pl), pop_b[2]), (pk_b[2], ident(sl)), (ident(pl), ident(sl)),  # identity key / proof / both (proof checked first)
        (bytes(pl), pop_b[3]), (pk_b[3], bytes(sl)),                      # undecodable
        (conv(bad_points[pkg]["outside"]), pop_b[4]), (pk_b[4], conv(bad_points[sgg]["outside"])),
        (conv(bad_points[pkg]["off_curve"]), pop_b[5]),
        (C.pk_ser(pk_pts[6], 1 - fmt), pop_b[6]),                         # the key in the OTHER format
    ]
    got = eng.pop_verify_batch(impl, [c[0] for c in cases], [c[1] for c in cases], fmt)
    want = [O.pop_verify(impl, fmt, a, b) for a, b in cases]
    assert got.tolist() == want, (impl, fmt)
    assert want[:9] == [0] * 9 and want[9] == 1 and want[10] == 1 and want[11:14] == [3, 2, 2] and want[14:19] == [4] * 5
    # a batch large enough for the bucket path: proofs from the synthetic-data helper (scheme 3 = proof of possession),
    # pinned to the oracle on a sample, then planted failures
    n = 4200
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    none, off0 = np.zeros(0, dtype=np.uint8), np.zeros(n + 1, dtype=np.uint64)
    pks, proofs = eng.testdata_sign(impl, 3, k, none, off0)
    for i in (0, 1, n - 1):
        ki = int.from_bytes(k[i * 32:(i + 1) * 32].tobytes(), "big")
        pk_pt = C.pk_mul(C.pk_gen, ki)
        assert pks[i * pl:(i + 1) * pl].tobytes() == C.pk_ser(pk_pt, O.MODERN)
        assert proofs[i * sl:(i + 1) * sl].tobytes() == C.sig_ser(C.sig_mul(C.hash(C.pk_ser(pk_pt, O.MODERN), O.pop_dst(impl)), ki), O.MODERN)
    if fmt == O.LEGACY:
        st, p = eng.recode_points(pkg, pks, O.MODERN, O.LEGACY); assert int(st.max()) == 0
        pks = np.frombuffer(b"".join(p), dtype=np.uint8)
        st, p = eng.recode_points(sgg, proofs, O.MODERN, O.LEGACY); assert int(st.max()) == 0
        proofs = np.frombuffer(b"".join(p), dtype=np.uint8)
    assert eng.pop_verify_batch(impl, pks, proofs, fmt).tolist() == [0] * n
    pr2 = proofs.copy().reshape(n, sl)
    bad = [17, 2048, 4199]
    pr2[bad] = pr2[[b - 1 for b in bad]]
    st = eng.pop_verify_batch(impl, pks, pr2.reshape(-1), fmt)
    assert [i for i in range(n) if st[i]] == bad and all(st[i] == 1 for i in bad)
    assert O.pop_verify(impl, fmt, pks[17 * pl:18 * pl].tobytes(), pr2[17].tobytes()) == 1


def test_rlc_salt_is_random_by_default_and_statuses_do_not_depend_on_it(eng, B):
    """Default: a fresh OS-random salt per call.  Pinned salt: reproducible.  128-bit scalars: same statuses.
    (The accept/reject vector is a property of the inputs; only the scalars of the batch equation change.)"""
    rnd = random.Random(1234)
    n = 4500
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    data, off = B.pack_messages([hashlib.sha256(b"salt%d" % i).digest() for i in range(n)])
    pks, sigs = eng.testdata_sign(2, 0, k, data, off)
    bad = [3, 700, 4499]
    s2 = sigs.copy().reshape(n, 96)
    s2[bad] = s2[[b + 1 if b + 1 < n else 0 for b in bad]]
    want = [1 if i in bad else 0 for i in range(n)]
    e2 = B.Engine([0])
    try:
        for bits in (64, 128):
            e2.set_rlc_bits(bits)
            for rep in range(2):      # unpinned: two calls draw two salts
                assert e2.verify_batch_packed(2, 0, pks, s2.reshape(-1), data, off).tolist() == want
            assert e2.verify_batch_packed(2, 0, pks, sigs, data, off).tolist() == [0] * n
            small = e2.verify_batch_packed(2, 0, pks[:48 * 100], s2.reshape(-1)[:96 * 100], data, off[:101])   # per-item scaling path
            assert small.tolist() == want[:100]
        e2.set_rlc_salt(bytes(range(32)))
        assert e2.verify_batch_packed(2, 0, pks, s2.reshape(-1), data, off).tolist() == want
        e2.set_rlc_bits(64)
        with pytest.raises(B.EngineError):
            e2.set_rlc_bits(96)
        for impl in (1,):
            pk1, sg1 = e2.testdata_sign(impl, 0, k[:32 * 300], data, off[:301])
            e2.set_rlc_bits(128)
            assert e2.verify_batch_packed(impl, 0, pk1, sg1, data, off[:301]).tolist() == [0] * 300
    finally:
        e2.close()


def test_bad_offsets_and_chunk_overrides_are_refused_or_clamped(eng, B, monkeypatch):
    n = 300
    rnd = random.Random(2)
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    data, off = B.pack_messages([b"o%d" % i for i in range(n)])
    pks, sigs = eng.testdata_sign(2, 0, k, data, off)
    bad_off = off.copy(); bad_off[10], bad_off[11] = off[11], off[10] - 1   # decreasing
    with pytest.raises(B.EngineError):
        eng.verify_batch_packed(2, 0, pks, sigs, data, bad_off)
    for chunk in ("0", "7", "junk", "119"):   # used to loop forever (chunk 0) - now clamped to one group-multiple pass size
        monkeypatch.setenv("BLSGPU_M6_CHUNK", chunk)
        assert eng.verify_batch_packed(2, 0, pks, sigs, data, off).tolist() == [0] * n
    monkeypatch.delenv("BLSGPU_M6_CHUNK")


# ---- the other public 2-pairing checks, pairs assembled on the device -------------------------------------
@pytest.mark.parametrize("impl", [2, 1])
def test_signcrypt_valid_and_share_checks_against_oracle(eng, B, impl):
    """blsgpu_signcrypt_valid_batch / blsgpu_signcrypt_verify_share_batch against the oracle's restatement of
    BlsSignCrypt::valid / verify_share (sign_crypt.rs:69-77,192-207); ciphertexts built as `seal` builds them (:36-61)."""
    rnd = random.Random(70 + impl)
    C = O.IMPLS[impl]
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    items = []
    for i in range(4):
        scheme = i % 3
        r = rnd.randrange(1, O.R)
        u = C.pk_mul(C.pk_gen, r)
        v = bytes(rnd.randrange(256) for _ in range(32 + 5 * i))
        w = C.sig_mul(C.hash(C.pk_ser(u, O.MODERN) + v, O.sig_dst(impl, scheme)), r)
        items.append((scheme, C.pk_ser(u, O.MODERN), v, C.sig_ser(w, O.MODERN), r))
    for scheme in (0, 1, 2):
        sel = [it for it in items if it[0] == scheme] + [items[0]]          # the last one is valid only under scheme 0
        cases = [(it[1], it[2], it[3]) for it in sel]
        cases += [(sel[0][1], sel[0][2] + b"!", sel[0][3]), (ident(pl), sel[0][2], sel[0][3]), (sel[0][1], sel[0][2], ident(sl)),
                  (bytes(pl), sel[0][2], sel[0][3])]
        ok, st = eng.signcrypt_valid_batch(impl, scheme, [c[0] for c in cases], [c[2] for c in cases], [c[1] for c in cases])
        want = [O.signcrypt_valid(impl, scheme, u, v, w) for u, v, w in cases]
        assert list(zip(st.tolist(), [bool(x) for x in ok])) == want, (impl, scheme)
    # decryption shares: share = U * sk_i, pk_share = g * sk_i   (create_decryption_share, sign_crypt.rs:166-184)
    scheme, ub, v, wb, r = items[0]
    u = C.pk_deser(ub, O.MODERN)
    sks = [rnd.randrange(1, O.R) for _ in range(2)]
    sh = [C.pk_ser(C.pk_mul(u, sk), O.MODERN) for sk in sks]
    pk = [C.pk_ser(C.pk_mul(C.pk_gen, sk), O.MODERN) for sk in sks]
    cases = [(sh[0], pk[0]), (sh[1], pk[1]), (sh[0], pk[1]), (ident(pl), pk[0]), (sh[0], ident(pl)), (bytes(pl), pk[0])]
    ok, st = eng.signcrypt_verify_share_batch(impl, 0, [c[0] for c in cases], [c[1] for c in cases], [ub] * 6, [wb] * 6, [v] * 6)
    want = [O.signcrypt_verify_share(impl, 0, a, b, ub, v, wb) for a, b in cases]
    assert list(zip(st.tolist(), [bool(x) for x in ok])) == want
    assert [w[1] for w in want] == [True, True, False, False, False, False] and want[5][0] == 4
    ok, st = eng.signcrypt_verify_share_batch(impl, 0, [sh[0]], [pk[0]], [ub], [ident(sl)], [v])
    assert (int(st[0]), bool(ok[0])) == O.signcrypt_verify_share(impl, 0, sh[0], pk[0], ub, v, ident(sl)) == (0, False)


@pytest.mark.parametrize("impl", [2, 1])
def test_pok_verify_batch_against_oracle(eng, B, impl):
    """blsgpu_pok_verify_batch against the oracle's restatement of BlsSignatureProof::verify (sig_proof.rs:102-142); proofs
    built as generate_proof builds them (:49-74): U = a*x, V = -(sig*(x+y))."""
    rnd = random.Random(80 + impl)
    C = O.IMPLS[impl]
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    cases = []
    for i in range(3):
        scheme = i % 3
        sk, x, y = (rnd.randrange(1, O.R) for _ in range(3))
        msg = b"pok message %d" % i
        a = C.hash(msg, O.sig_dst(impl, scheme))
        U = C.sig_ser(C.sig_mul(a, x), O.MODERN)
        V = C.sig_ser(C.sig_mul(a, (-(sk * (x + y))) % O.R), O.MODERN)
        cases.append((scheme, U, V, C.pk_ser(C.pk_mul(C.pk_gen, sk), O.MODERN), y.to_bytes(32, "big"), msg))
    s0 = cases[0]
    extra = [
        (0, s0[1], s0[2], s0[3], (int.from_bytes(s0[4], "big") + 1).to_bytes(32, "big"), s0[5]),   # wrong challenge
        (0, s0[1], s0[2], s0[3], s0[4], s0[5] + b"?"),                                               # wrong message
        (0, ident(sl), s0[2], s0[3], s0[4], s0[5]), (0, s0[1], ident(sl), s0[3], s0[4], s0[5]), (0, s0[1], s0[2], ident(pl), s0[4], s0[5]),
        (0, s0[1], s0[2], s0[3], bytes(32), s0[5]), (0, s0[1], s0[2], s0[3], O.R.to_bytes(32, "big"), s0[5]),
        (0, ident(sl), ident(sl), ident(pl), bytes(32), s0[5]),                                      # first check wins
        (0, bytes(sl), s0[2], s0[3], s0[4], s0[5]),
    ]
    for scheme in (0, 1, 2):
        sel = [c for c in cases if c[0] == scheme] + (extra if scheme == 0 else [cases[0]])
        got = eng.pok_verify_batch(impl, scheme, [c[1] for c in sel], [c[2] for c in sel], [c[3] for c in sel], [c[4] for c in sel],
                                   [c[5] for c in sel])
        want = [O.pok_verify(impl, scheme, c[1], c[2], c[3], c[4], c[5]) for c in sel]
        assert got.tolist() == want, (impl, scheme)
        if scheme == 0:
            assert want == [0, 12, 12, 13, 14, 3, 15, 4, 13, 4]


@pytest.mark.parametrize("impl", [2, 1])
def test_verify_batch_records_lengths_tags_and_formats(eng, B, impl):
    """blsgpu_verify_batch_records: ragged network records.  Length rules (public_key.rs:159-164, signature.rs:236-241),
    serde tags (signature.rs:112-126) and the Legacy / Modern header rules (legacy.rs:39-82), each against the oracle."""
    rnd = random.Random(90 + impl)
    C = O.IMPLS[impl]
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    sks = [rnd.randrange(1, O.R) for _ in range(3)]
    msgs = [b"rec-%d" % i for i in range(3)]
    pk_pts = [O.sk_to_pk(impl, sk) for sk in sks]
    for fmt in (O.MODERN, O.LEGACY):
        pk = [C.pk_ser(p, fmt) for p in pk_pts]
        sg = [C.sig_ser(O.sign(impl, 0, sk, m), fmt) for sk, m in zip(sks, msgs)]
        other_pk = C.pk_ser(pk_pts[1], 1 - fmt)
        recs = [(pk[0], sg[0], msgs[0]), (pk[1], sg[1], msgs[1]), (pk[2][:-1], sg[2], msgs[2]), (pk[2], sg[2] + b"\0", msgs[2]),
                (b"", sg[0], msgs[0]), (pk[0], b"", msgs[0]), (other_pk, sg[1], msgs[1]), (pk[2], sg[2], b"")]
        got = eng.verify_batch_records(impl, fmt, 0, [r[0] for r in recs], [r[1] for r in recs], [r[2] for r in recs])
        want = [6 if len(p) != pl or len(s) != sl else O.verify(impl, 0, fmt, p, s, m) for p, s, m in recs]
        assert got.tolist() == want, (impl, fmt)
        assert want[:6] == [0, 0, 6, 6, 6, 6] and want[6] != 0 and want[7] == 1
    # the serde_bare form: tag + IETF point; schemes mixed; short / long / unknown tag are serde errors
    pk = [C.pk_ser(p, O.MODERN) for p in pk_pts]
    tagged = [bytes([sch]) + C.sig_ser(O.sign(impl, sch, sk, m), O.MODERN) for sch, sk, m in zip((0, 1, 2), sks, msgs)]
    recs = list(zip(pk, tagged, msgs)) + [(pk[0], tagged[0][:-1], msgs[0]), (pk[0], bytes([3]) + tagged[0][1:], msgs[0]),
                                          (pk[0], bytes([1]) + tagged[0][1:], msgs[0]), (pk[0][:5], tagged[0], msgs[0])]
    got = eng.verify_batch_records(impl, O.MODERN, -1, [r[0] for r in recs], [r[1] for r in recs], [r[2] for r in recs])
    assert got.tolist() == [0, 0, 0, 4, 4, O.verify(impl, 1, O.MODERN, pk[0], tagged[0][1:], msgs[0]), 6]
    assert got[5] == 1


# ---- one batch cut over several slices, folded into a single Miller loop + final exponentiation ----------------
def test_partial_results_fold_and_finish(B, cpp):
    """blsgpu_miller_partial / blsgpu_final_exp_is_one / blsgpu_partial_finish with three contexts on one GPU standing in
    for three devices: statuses equal blsgpu_verify_batch's; the partial results satisfy the batch equation under the
    ORACLE's pairing as well; a tampered partial result makes the fold fail."""
    rnd = random.Random(555)
    n = 3 * 4200 + 5
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [hashlib.sha256(b"fold%d" % i).digest()[: 1 + i % 32] for i in range(n)]
    data, off = B.pack_messages(msgs)
    engs = [B.Engine([0]) for _ in range(3)]
    try:
        pks, sigs = engs[0].testdata_sign(2, 1, k, data, off)       # MessageAugmentation: the pk prefix travels with each slice
        sigs = sigs.copy()
        for trial, bad in enumerate(([], [4200 + 7, 4200 + 8, 2 * 4200 - 1])):   # all valid; three bad items, all in slice 1
            s2 = sigs.reshape(n, 96).copy()
            for i in bad:
                s2[i] = sigs.reshape(n, 96)[(i + 11) % n]
            s2[3] = 0; s2[3, 0] = 0xC0                               # an identity signature in slice 0: status 2, no failure
            flat = s2.reshape(-1)
            want = engs[0].verify_batch_packed(2, 1, pks, flat, data, off)
            assert [i for i in range(n) if want[i] == 1] == bad and want[3] == 2
            parts, ranges = [], []
            for r, e in enumerate(engs):
                lo, hi = B.shard_range(n, r, 3)
                ranges.append((lo, hi))
                parts.append(e.miller_partial(2, 1, pks[48 * lo:48 * hi], flat[96 * lo:96 * hi], data, off[lo:hi + 1]))
            ok = engs[2].final_exp_is_one(2, [p[0] for p in parts], [p[1] for p in parts])
            assert ok == (not bad)
            got = np.concatenate([e.partial_finish(hi - lo, ok) for e, (lo, hi) in zip(engs, ranges)])
            assert np.array_equal(got, want)
            if trial == 0:
                # the oracle agrees that the folded equation holds:  prod F_j * e(-g, sum S_j) == 1
                f = O.F12_ONE
                s = None
                for gt, sm in parts:
                    f = O.f12_mul(f, tuple((int.from_bytes(gt[96 * i:96 * i + 48], "big"), int.from_bytes(gt[96 * i + 48:96 * i + 96], "big")) for i in range(6)))
                    s = O.g2_add(s, O.g2_deserialize(sm))
                f = O.f12_mul(f, O.miller_loop(O.g1_neg(O.G1_GEN), s))
                assert O.final_exponentiation(f) == O.F12_ONE
                # and that a tampered partial result breaks it
                assert not engs[1].final_exp_is_one(2, [parts[0][0], parts[1][0], parts[1][0]], [p[1] for p in parts])
                assert not engs[1].final_exp_is_one(2, [p[0] for p in parts], [parts[0][1], parts[0][1], parts[2][1]])
                assert engs[1].final_exp_is_one(2, [p[0] for p in parts[::-1]], [p[1] for p in parts[::-1]])   # order-free
        # an empty slice is (1, O)
        gt, sm = engs[0].miller_partial(2, 0, np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(0, np.uint8), np.zeros(1, np.uint64))
        assert gt == (1).to_bytes(48, "big") + bytes(528) and sm == ident(96)
        assert engs[0].final_exp_is_one(2, [gt], [sm]) and engs[0].partial_finish(0, True).size == 0
        with pytest.raises(B.EngineError):
            engs[0].final_exp_is_one(2, [bytes([0xff]) * 576], [sm])
    finally:
        for e in engs:
            e.close()


def test_small_batches_and_probe_paths(eng, B):
    """The cooperative probes (six-lane Miller loop + final exponentiation) on every small shape: 1, 2, 5, 6, 7, 13, 97 items,
    each all-valid and with every second item invalid (all leaves of the tree probed), both impls."""
    rnd = random.Random(777)
    for impl in (2, 1):
        sl = B.sig_len(impl)
        nmax = 97
        k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(nmax)), dtype=np.uint8)
        data, off = B.pack_messages([b"small-%d" % i for i in range(nmax)])
        pks, sigs = eng.testdata_sign(impl, 0, k, data, off)
        pl = B.pk_len(impl)
        for n in (1, 2, 5, 6, 7, 13, 97):
            assert eng.verify_batch_packed(impl, 0, pks[:pl * n], sigs[:sl * n], data, off[:n + 1]).tolist() == [0] * n
            s2 = sigs[:sl * n].copy().reshape(n, sl)
            bad = list(range(0, n, 2))
            s2[bad] = sigs.reshape(nmax, sl)[[(b + 1) % nmax for b in bad]]
            st = eng.verify_batch_packed(impl, 0, pks[:pl * n], s2.reshape(-1), data, off[:n + 1])
            assert st.tolist() == [1 if i % 2 == 0 else 0 for i in range(n)], (impl, n)


def test_context_on_two_devices_shards_sums_and_quorums(eng, B):
    """A context on several devices cuts blsgpu_sum_points (cfg 3: per-device partial sums, then the sum of the partial results)
    and blsgpu_verify_secure_batch / blsgpu_aggregate_secure_batch (cfg 5: quorums are independent) over its devices
    (SURVEY 8e).  Results must equal the single-device ones, including the index of the first bad element and the statuses of
    bad quorums on either side of the cut."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    rnd = random.Random(91)
    e2 = B.Engine([0, 1])
    try:
        # point sums, both groups
        n = 2 * 65536 + 11
        k = np.zeros((n, 32), dtype=np.uint8)
        k[:, 24:] = np.frombuffer(rnd.randbytes(8 * n), dtype=np.uint8).reshape(n, 8)
        k[:, 31] |= 1
        msg = np.frombuffer(b"same message", dtype=np.uint8)
        off = np.arange(n + 1, dtype=np.uint64) * msg.size
        pks, sigs = eng.testdata_sign(2, 2, k.reshape(-1), np.tile(msg, n), off)
        for group, pts, L in ((1, pks, 48), (2, sigs, 96)):
            before = e2.launch_count()
            assert e2.sum_points(group, pts) == eng.sum_points(group, pts)
            assert e2.launch_count() - before >= 2 * 4                      # both devices decoded and summed
            bad = pts.copy()
            bad[(n - 5) * L:(n - 4) * L] = 0xFF                              # undecodable, in the second slice
            bad[(70000) * L + 1] ^= 0x55                                     # earlier one (second slice as well)
            for e in (eng, e2):
                with pytest.raises(B.BlsError) as err:
                    e.sum_points(group, bad)
                assert err.value.status == B.ST_DESERIALIZE and "element 70000" in str(err.value)
        # quorums: 40 key sets of ragged sizes, bad aggregates on both sides of the cut, one Legacy pass
        q = 40
        sizes = [150 + 23 * (j % 7) for j in range(q)]
        tot = sum(sizes)
        kk = np.zeros((tot, 32), dtype=np.uint8)
        kk[:, 20:] = np.frombuffer(rnd.randbytes(12 * tot), dtype=np.uint8).reshape(tot, 12)
        kk[:, 31] |= 1
        qmsgs = [(b"quorum %d" % j) * (1 + j % 3) for j in range(q)]
        m5, o5 = B.pack_messages([qmsgs[j] for j in range(q) for _ in range(sizes[j])])
        pk5, sg5 = eng.testdata_sign(2, 0, kk.reshape(-1), m5, o5)
        koff = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64)
        qm, qoff = B.pack_messages(qmsgs)
        for fmt in (1, 0):
            if fmt == 0:
                pk5, sg5 = eng.recode_points_packed(1, pk5, 1, 0), eng.recode_points_packed(2, sg5, 1, 0)
            st1, agg1 = eng.aggregate_secure_batch_packed(2, koff, pk5, sg5, fmt)
            before = e2.launch_count()
            st2, agg2 = e2.aggregate_secure_batch_packed(2, koff, pk5, sg5, fmt)
            assert e2.launch_count() - before >= 2 * 4
            assert st1.tolist() == st2.tolist() == [0] * q and np.array_equal(agg1, agg2)
            wrong = agg1.copy().reshape(q, 96)
            wrong[[2, 3, q - 2]] = wrong[[3, 2, 5]]
            want = eng.verify_secure_batch_packed(2, 0, koff, pk5, wrong.reshape(-1), qm, qoff, fmt)
            assert np.nonzero(want)[0].tolist() == [2, 3, q - 2]
            assert e2.verify_secure_batch_packed(2, 0, koff, pk5, wrong.reshape(-1), qm, qoff, fmt).tolist() == want.tolist()
    finally:
        e2.close()


@pytest.mark.parametrize("impl,scheme", [(2, 0), (2, 1), (1, 2)])
def test_context_on_two_devices_shards_aggregate_verify(eng, B, impl, scheme):
    """cfg 4 on several devices: the pairs of AggregateSignature::verify are cut over the devices, the partial products of Miller
    values are folded on the first device.  Status and reported indices must equal the single-device ones for a valid aggregate,
    a wrong one, and every rule the reference applies before the pairing (first decode error, duplicate messages, identity key)
    when the offending element lies in the second slice."""
    import torch
    if torch.cuda.device_count() < 2:
        pytest.skip("needs two GPUs")
    n = 2 * 4096 + 5
    rnd = random.Random(17 + scheme)
    pl, sl = B.pk_len(impl), B.sig_len(impl)
    k = np.frombuffer(b"".join(rnd.randrange(1, O.R).to_bytes(32, "big") for _ in range(n)), dtype=np.uint8)
    msgs = [hashlib.sha256(b"agg%d" % i).digest()[:1 + i % 32] + bytes([i & 255, i >> 8 & 255]) for i in range(n)]
    data, off = B.pack_messages(msgs)
    pks, sigs = eng.testdata_sign(impl, scheme, k, data, off)
    agg = eng.sum_points(2 if impl == 2 else 1, sigs)
    pk_rows = pks.reshape(n, pl)
    ident = bytes([0xC0]) + bytes(pl - 1)
    e2 = B.Engine([0, 1])
    try:
        def both(pk_arr, ms, sg):
            a = eng.aggregate_verify_status(impl, scheme, pk_arr, ms, sg)
            before = e2.launch_count()
            b = e2.aggregate_verify_status(impl, scheme, pk_arr, ms, sg)
            assert a == b, (a, b)
            return a, e2.launch_count() - before
        (st, _), launched = both(pks, msgs, agg)
        assert st == 0 and launched > 2 * 12                       # both devices ran their own pipeline
        wrong = eng.sum_points(2 if impl == 2 else 1, sigs[sl:])
        assert both(pks, msgs, wrong)[0][0] == B.ST_INVALID_SIGNATURE
        swapped = list(msgs)
        swapped[6000], swapped[6001] = swapped[6001], swapped[6000]   # same multiset of pairs? no: keys stay - must fail
        assert both(pks, swapped, agg)[0][0] == B.ST_INVALID_SIGNATURE
        bad = pk_rows.copy()
        bad[7000] = 0xFF                                           # undecodable key in the second slice
        bad[8000] = np.frombuffer(ident, dtype=np.uint8)           # identity key later on
        assert both(bad.reshape(-1), msgs, agg)[0] == (B.ST_DESERIALIZE, (7000, -1))
        bad[7000] = pk_rows[7000]
        assert both(bad.reshape(-1), msgs, agg)[0] == (B.ST_PK_IDENTITY, (8001, -1))
        if scheme == 0:
            dup = list(msgs)
            dup[8100] = dup[3]
            assert both(pks, dup, agg)[0] == (B.ST_DUPLICATE_MESSAGES, (3, 8100))
        assert both(pks, msgs, bytes([0xC0]) + bytes(sl - 1))[0][0] == B.ST_SIG_IDENTITY
        # a batch too small to shard stays on the first device
        small = eng.sum_points(2 if impl == 2 else 1, sigs[:64 * sl])
        (st, _), launched = both(pks[:64 * pl], msgs[:64], small)
        assert st == 0 and launched < 2 * 12 + 40
    finally:
        e2.close()
