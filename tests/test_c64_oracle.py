"""The 64-bit-limb C oracle (oracle/c64/bls64.c) against the big-int oracle and the reference's golden vectors.
CPU only.  bls64.c shares no code with the CUDA engine; it is the fast checker of the large GPU parity tests and the timed
CPU leg of bench.py, so it is pinned here: field products on 10^4 random + edge operands, every decoder status on an
adversarial sweep, hash_to_curve, the pairing value, and the verification entry points."""
import hashlib
import json
import os
import random

import pytest

from oracle import bls_oracle as O
from oracle import c64_oracle as C


@pytest.fixture(scope="module", autouse=True)
def built():
    C.build()
    assert C.available()


@pytest.fixture(scope="module")
def cpp(golden_dir):
    return json.load(open(os.path.join(golden_dir, "cpp_integration.json")))


def test_field_products_on_random_and_edge_operands():
    rnd = random.Random(11)
    edge = [0, 1, 2, O.P - 1, O.P - 2, (1 << 380), (1 << 384) % O.P, (O.P - 1) // 2, (O.P + 1) // 2]
    pairs = [(a, b) for a in edge for b in edge] + [(rnd.randrange(O.P), rnd.randrange(O.P)) for _ in range(10000)]
    for a, b in pairs:
        assert C.fp_mul(a, b) == a * b % O.P
    with pytest.raises(ValueError):
        C.fp_mul(O.P, 1)


def test_generators_scalar_multiplication_and_sums():
    rnd = random.Random(12)
    assert C.generator(1) == O.g1_serialize(O.G1_GEN) and C.generator(2) == O.g2_serialize(O.G2_GEN)
    pts1, pts2 = [], []
    for k in [1, O.R - 1, O.R + 1, 2 ** 255 - 19] + [rnd.randrange(O.R) for _ in range(2)]:
        a, b = C.point_mul(1, C.generator(1), k % 2 ** 256), C.point_mul(2, C.generator(2), k % 2 ** 256)
        assert a == O.g1_serialize(O.g1_mul(O.G1_GEN, k)) and b == O.g2_serialize(O.g2_mul(O.G2_GEN, k))
        pts1.append(a); pts2.append(b)
    for fmt in (O.MODERN, O.LEGACY):
        e1 = [O.g1_serialize(O.g1_deserialize(p), fmt) for p in pts1]
        e2 = [O.g2_serialize(O.g2_deserialize(p), fmt) for p in pts2]
        assert C.sum_points(1, fmt, e1) == O.sum_points(1, fmt, e1)
        assert C.sum_points(2, fmt, e2) == O.sum_points(2, fmt, e2)
    assert C.sum_points(1, 1, [pts1[0], bytes(48), pts1[1]])[::2] == (4, 1)
    assert C.sum_points(2, 0, []) == (0, bytes([0xC0]) + bytes(95), -1)


def _status(deser, enc, fmt):
    try:
        deser(enc, fmt)
        return 0
    except O.BlsError as e:
        return e.code


def test_decoder_statuses_on_an_adversarial_sweep():
    """Every header-bit combination on valid and invalid x, both groups, both formats: status and re-encoded bytes."""
    rnd = random.Random(13)
    n_cases = 0
    for group, L, deser, ser, gen_mul in ((1, 48, O.g1_deserialize, O.g1_serialize, lambda k: O.g1_mul(O.G1_GEN, k)),
                                          (2, 96, O.g2_deserialize, O.g2_serialize, lambda k: O.g2_mul(O.G2_GEN, k))):
        bodies = [ser(gen_mul(rnd.randrange(1, O.R))) for _ in range(3)]
        # x on the curve but outside the subgroup / off the curve / not canonical
        x = 1
        found = {"outside": None, "off": None}
        while None in found.values():
            if group == 1:
                y = O.fp_sqrt((x ** 3 + 4) % O.P)
                enc = bytearray(x.to_bytes(48, "big"))
                on, sub = y is not None, y is not None and O.g1_in_subgroup((x, y))
            else:
                xx = (x, 1)
                y = O.f2_sqrt(O.f2_add(O.f2_mul(O.f2_sqr(xx), xx), O.B2))
                enc = bytearray((1).to_bytes(48, "big") + x.to_bytes(48, "big"))
                on, sub = y is not None, y is not None and O.g2_in_subgroup((xx, y))
            enc[0] |= 0x80
            if on and not sub and found["outside"] is None:
                found["outside"] = bytes(enc)
            if not on and found["off"] is None:
                found["off"] = bytes(enc)
            x += 1
        big = bytearray((O.P + 3).to_bytes(48, "big") + bytes(L - 48)); big[0] |= 0x80
        bodies += [found["outside"], found["off"], bytes(big), bytes([0xC0]) + bytes(L - 2) + b"\x01", bytes(L),
                   bytes([0xC0]) + bytes(L - 1)]
        for body in bodies:
            for b0 in range(0, 256, 32):
                enc = bytes([(body[0] & 0x1F) | b0]) + body[1:]
                for fin in (O.MODERN, O.LEGACY):
                    want = _status(deser, enc, fin)
                    st, out = C.recode(group, fin, O.MODERN, enc)
                    assert st == want, (group, fin, enc.hex()[:12])
                    if want == 0:
                        assert out == ser(deser(enc, fin), O.MODERN)
                        assert C.recode(group, fin, O.LEGACY, enc)[1] == ser(deser(enc, fin), O.LEGACY)
                    n_cases += 1
    assert n_cases >= 250


def test_endomorphism_checks_and_psi_cofactor_clearing_agree_with_the_definitions():
    """Scott's subgroup tests == [r]P = O, psi-based clear_cofactor == h_eff scalar multiplication, on curve points that are
    NOT in the subgroup (small x) and on subgroup points."""
    seen = 0
    for group in (1, 2):
        x = 1
        while seen < (40 if group == 1 else 70):
            enc = x.to_bytes(48, "big") if group == 1 else (3).to_bytes(48, "big") + x.to_bytes(48, "big")
            r = C.selfcheck_point(group, enc, x % 2 == 0)
            if r >= 0:
                assert r == 3, (group, x, r)
                seen += 1
            x += 1
        g = O.G1_GEN if group == 1 else O.G2_GEN
        ser = (lambda p: p[0].to_bytes(48, "big")) if group == 1 else (lambda p: p[0][1].to_bytes(48, "big") + p[0][0].to_bytes(48, "big"))
        mul = O.g1_mul if group == 1 else O.g2_mul
        for k in (1, 7, 2 ** 100 + 1):
            p = mul(g, k)
            assert C.selfcheck_point(group, ser(p), False) == 3


def test_hash_to_curve_both_suites():
    rnd = random.Random(14)
    msgs = [b"", b"abc", b"abcdef0123456789", b"a" * 133] + [bytes(rnd.randrange(256) for _ in range(rnd.randrange(1, 90))) for _ in range(4)]
    for g, f, ser in ((1, O.hash_to_curve_g1, O.g1_serialize), (2, O.hash_to_curve_g2, O.g2_serialize)):
        for dst in (b"QUUX-V01-CS02-with-BLS12381G%d_XMD:SHA-256_SSWU_RO_" % g, O.sig_dst(g, O.BASIC), O.pop_dst(g)):
            for m in msgs[:4] if dst.startswith(b"QUUX") else msgs[4:6]:
                assert C.hash_to_curve(g, m, dst) == ser(f(m, dst))
    # RFC 9380 J.9.1 / J.10.1, msg = "" 
    assert C.hash_to_curve(1, b"", b"QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_RO_")[1:].hex() == \
        "052926add2207b76ca4fa57a8734416c8dc95e24501772c814278700eed6d1e4e8cf62d9c09db0fac349612b759e79a1"[2:]


def test_pairing_value_equals_the_big_int_pairing_cubed():
    a, b = 0x1234567, 0x7654321
    P1, Q2 = O.g1_mul(O.G1_GEN, a), O.g2_mul(O.G2_GEN, b)
    want = O.f12_pow(O.pairing(P1, Q2), 3)
    wb = b"".join(c[0].to_bytes(48, "big") + c[1].to_bytes(48, "big") for c in want)
    assert C.pairing_cubed(O.g1_serialize(P1), O.g2_serialize(Q2)) == wb
    s1, s2 = O.g1_serialize, O.g2_serialize
    assert C.pairing_product_is_one([s1(P1), s1(O.g1_neg(O.g1_mul(O.G1_GEN, a * b % O.R)))], [s2(Q2), s2(O.G2_GEN)])
    assert not C.pairing_product_is_one([s1(P1), s1(O.g1_neg(O.g1_mul(O.G1_GEN, a * b % O.R + 1)))], [s2(Q2), s2(O.G2_GEN)])
    assert C.pairing_product_is_one([s1(None), s1(P1)], [s2(Q2), s2(None)]) and C.pairing_product_is_one([], [])


def test_reference_golden_vectors_directly(cpp, golden_dir):
    """cpp_integration_test.rs:19-192 and secure_aggregation_test.rs:143-235 without the Python oracle in between."""
    msg = bytes.fromhex(cpp["message"])
    pks = [bytes.fromhex(s["pk"]) for s in cpp["signers"]]
    sigs = [bytes.fromhex(s["sig"]) for s in cpp["signers"]]
    for s, pk, sig in zip(cpp["signers"], pks, sigs):
        assert C.verify(2, 0, 1, pk, sig, msg) == 0
        assert C.point_mul(1, C.generator(1), int(s["sk"], 16)) == pk
        assert C.point_mul(2, C.hash_to_curve(2, msg, O.sig_dst(2, 0)), int(s["sk"], 16)) == sig
    assert C.verify(2, 0, 1, pks[0], sigs[1], msg) == 1 and C.verify(2, 1, 1, pks[0], sigs[0], msg) == 1
    assert C.sum_points(2, 1, sigs[:2])[1].hex() == cpp["normal_agg_sig12"]
    for k in (2, 3):
        st, agg = C.aggregate_secure(2, 1, pks[:k], sigs[:k])
        assert st == 0 and C.verify_secure(2, 0, 1, pks[:k], agg, msg) == 0
        assert C.verify_secure(2, 0, 1, pks[:k][::-1], agg, msg) == 0
    assert C.verify_secure(2, 0, 1, pks[:2], bytes.fromhex(cpp["normal_agg_sig12"]), msg) == 1
    sec = json.load(open(os.path.join(golden_dir, "secure_57.json")))
    keys = [bytes.fromhex(k) for k in sec["keys"]]
    assert C.verify_secure(2, 0, 1, keys, bytes.fromhex(sec["sig"]), bytes.fromhex(sec["message"])) == 0
    assert C.verify_secure(2, 0, 1, keys[:-1], bytes.fromhex(sec["sig"]), bytes.fromhex(sec["message"])) == 1


@pytest.mark.parametrize("impl", [2, 1])
def test_verification_entry_points_against_the_big_int_oracle(impl):
    rnd = random.Random(15 + impl)
    Ci = O.IMPLS[impl]
    sks = [rnd.randrange(1, O.R) for _ in range(3)]
    msgs = [b"c64-%d" % i for i in range(3)]
    ident_pk, ident_sig = bytes([0xC0]) + bytes(Ci.pk_len - 1), bytes([0xC0]) + bytes(Ci.sig_len - 1)
    for scheme, fmt in (((0, 1), (1, 0)) if impl == 2 else ((2, 1),)):
        pks = [Ci.pk_ser(O.sk_to_pk(impl, sk), fmt) for sk in sks]
        sigs = [Ci.sig_ser(O.sign(impl, scheme, sk, m), fmt) for sk, m in zip(sks, msgs)]
        cases = [(pks[0], sigs[0], msgs[0]), (pks[1], sigs[0], msgs[0]), (pks[0], ident_sig, msgs[0]), (ident_pk, sigs[0], msgs[0]),
                 (bytes(Ci.pk_len), sigs[0], msgs[0])]
        for pk, sig, m in cases:
            assert C.verify(impl, scheme, fmt, pk, sig, m) == O.verify(impl, scheme, fmt, pk, sig, m)
        # the C oracle alone, multi-threaded, on the whole little batch
        import numpy as np
        off = np.cumsum([0] + [len(m) for m in msgs]).astype(np.uint64)
        st = C.verify_many(impl, scheme, fmt, np.frombuffer(b"".join(pks), dtype=np.uint8), np.frombuffer(b"".join(sigs), dtype=np.uint8),
                           np.frombuffer(b"".join(msgs), dtype=np.uint8), off, threads=2)
        assert st.tolist() == [0, 0, 0]
    fmt = 1
    pk_pts = [O.sk_to_pk(impl, sk) for sk in sks]
    pks = [Ci.pk_ser(p, fmt) for p in pk_pts]
    # proof of possession
    pop = Ci.sig_ser(Ci.sig_mul(Ci.hash(pks[0], O.pop_dst(impl)), sks[0]), fmt)
    for pk, pr in ((pks[0], pop), (pks[1], pop), (ident_pk, pop), (pks[0], ident_sig)):
        assert C.pop_verify(impl, fmt, pk, pr) == O.pop_verify(impl, fmt, pk, pr)
    # aggregate verify: accept, duplicate messages, identity key, tampered
    sigs = [O.sign(impl, 0, sk, m) for sk, m in zip(sks, msgs)]
    agg = Ci.sig_ser(Ci.sig_add(Ci.sig_add(sigs[0], sigs[1]), sigs[2]), fmt)
    for ks, ms in ((pks, msgs), (pks, [msgs[0], msgs[1], msgs[0]]), (pks[:2] + [ident_pk], msgs)) + (((pks, [msgs[0], msgs[1], b"x"]),) if impl == 2 else ()):
        got = C.aggregate_verify(impl, 0, fmt, ks, ms, agg)
        want = O.aggregate_verify(impl, 0, fmt, ks, ms, agg)
        assert got[0] == want[0] and tuple(i for i in got[1] if i >= 0) == tuple(want[1])
    # secure aggregation incl. a duplicate key (first-match rule) and the legacy coefficient derivation
    for f2 in ((1, 0) if impl == 2 else (1,)):
        kb = [Ci.pk_ser(p, f2) for p in pk_pts] + [Ci.pk_ser(pk_pts[1], f2)]
        sb = [Ci.sig_ser(Ci.sig_mul(Ci.hash(b"q", O.sig_dst(impl, 0)), sk), f2) for sk in sks + [sks[0]]]
        want = O.aggregate_secure(impl, f2, kb, sb)
        assert C.aggregate_secure(impl, f2, kb, sb) == want
        assert C.verify_secure(impl, 0, f2, kb, want[1], b"q") == O.verify_secure(impl, 0, f2, kb, want[1], b"q")
        assert C.verify_secure(impl, 0, f2, kb[:3], want[1], b"q") == 1
        assert C.verify_secure(impl, 0, f2, [], ident_sig, b"q") == 0 and C.verify_secure(impl, 0, f2, [], want[1], b"q") == 1
