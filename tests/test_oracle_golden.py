"""Pins the CPU oracle against every golden vector the reference's tests hold for the hot path
 and the public RFC 9380 vectors. CPU only."""
import hashlib
import json
import os

import pytest

from oracle import bls_oracle as O


@pytest.fixture(scope="module")
def cpp(golden_dir):
    return json.load(open(os.path.join(golden_dir, "cpp_integration.json")))


@pytest.fixture(scope="module")
def sec57(golden_dir):
    return json.load(open(os.path.join(golden_dir, "secure_57.json")))


def test_cpp_keys_and_signatures_bytes(cpp):
    # reference tests/cpp_integration_test.rs:99-104,141-148: pk == sk*G1, sig == sk*H(m), verify ok
    msg = bytes.fromhex(cpp["message"])
    for s in cpp["signers"]:
        sk = int(s["sk"], 16)
        assert O.g1_serialize(O.sk_to_pk(O.G2IMPL, sk)).hex() == s["pk"]
        assert O.g2_serialize(O.sign(O.G2IMPL, O.BASIC, sk, msg)).hex() == s["sig"]


def test_cpp_hash_to_curve_kat(cpp):
    # H = sk^-1 * sig, identical for all three keys
    h = O.hash_to_curve_g2(bytes.fromhex(cpp["message"]), O.sig_dst(O.G2IMPL, O.BASIC))
    assert O.g2_serialize(h).hex() == (
        "8dbf4d3c426badac1e66421c7d65dc017c05fb7631833f3c9a72f531bedf7995f2309d2fd6831018c83de0c27b6a10c8"
        "10946937ad15674b2f3976d10f50ae5a66a07f5da23a4f177870702d0dbf8463225a493a8c221032e15d445afeac748a")


def test_cpp_verify_accept_and_reject(cpp):
    msg = bytes.fromhex(cpp["message"])
    s0, s1 = cpp["signers"][0], cpp["signers"][1]
    assert O.verify(O.G2IMPL, O.BASIC, O.MODERN, bytes.fromhex(s0["pk"]), bytes.fromhex(s0["sig"]), msg) == O.OK
    assert O.verify(O.G2IMPL, O.BASIC, O.MODERN, bytes.fromhex(s0["pk"]), bytes.fromhex(s1["sig"]), msg) \
        == O.ERR_INVALID_SIGNATURE


def test_cpp_normal_aggregate_bytes(cpp):
    # cpp_integration_test.rs:171-179 == compress(sig1 + sig2)
    st, out, bad = O.sum_points(2, O.MODERN, [bytes.fromhex(cpp["signers"][i]["sig"]) for i in (0, 1)])
    assert st == O.OK and bad == -1
    assert out.hex() == cpp["normal_agg_sig12"]


def test_cpp_secure_coefficients(cpp):
    # sorted order pk2, pk1; t[0] needs the mod-r reduction
    pks = sorted(bytes.fromhex(cpp["signers"][i]["pk"]) for i in (0, 1))
    assert pks[0].hex() == cpp["signers"][1]["pk"]
    assert hashlib.sha256(b"".join(pks)).hexdigest() == \
        "6040b788e954eb9df1a0d581cf020f7b1946d0ed0dd48de1d03bab45e3b3a29a"
    t = O.secure_coefficients(pks)
    assert "%064x" % t[0] == "584ccd89aaf51f8b06067b165b36a9096ae4abc23189c97ca1d34accb015244a"
    assert "%064x" % t[1] == "350f133013a3e8f028ab28c14c710b88cc15bfaa853497887728fe591c20a17d"


@pytest.mark.parametrize("n", [2, 3])
def test_cpp_secure_aggregate_then_verify(cpp, n):
    # cpp_integration_test.rs:86-165
    msg = bytes.fromhex(cpp["message"])
    pks = [bytes.fromhex(s["pk"]) for s in cpp["signers"][:n]]
    sigs = [bytes.fromhex(s["sig"]) for s in cpp["signers"][:n]]
    st, agg = O.aggregate_secure(O.G2IMPL, O.MODERN, pks, sigs)
    assert st == O.OK
    assert O.verify_secure(O.G2IMPL, O.BASIC, O.MODERN, pks, agg, msg) == O.OK
    # order independence (secure_aggregation_test.rs:13-140)
    assert O.verify_secure(O.G2IMPL, O.BASIC, O.MODERN, pks[::-1], agg, msg) == O.OK


def test_cpp_normal_aggregate_fails_verify_secure(cpp):
    # cpp_integration_test.rs:169-192
    msg = bytes.fromhex(cpp["message"])
    pks = [bytes.fromhex(s["pk"]) for s in cpp["signers"][:2]]
    assert O.verify_secure(O.G2IMPL, O.BASIC, O.MODERN, pks, bytes.fromhex(cpp["normal_agg_sig12"]), msg) \
        == O.ERR_INVALID_SIGNATURE


def test_production_57_key_vector(sec57):
    # secure_aggregation_test.rs:143-235
    pks = [bytes.fromhex(k) for k in sec57["keys"]]
    sig = bytes.fromhex(sec57["sig"])
    assert O.g2_serialize(O.g2_deserialize(sig, O.MODERN), O.MODERN) == sig  # :224-227
    assert O.verify_secure(O.G2IMPL, O.BASIC, O.MODERN, pks, sig, bytes.fromhex(sec57["message"])) == O.OK


def test_rfc9380_g1_vectors():
    dst = b"QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_RO_"
    x, y = O.hash_to_curve_g1(b"", dst)
    assert x == 0x052926ADD2207B76CA4FA57A8734416C8DC95E24501772C814278700EED6D1E4E8CF62D9C09DB0FAC349612B759E79A1
    assert y == 0x08BA738453BFED09CB546DBB0783DBB3A5F1F566ED67BB6BE0E8C67E2E81A4CC68EE29813BB7994998F3EAE0C9C6A265
    x, y = O.hash_to_curve_g1(b"abc", dst)
    assert x == 0x03567BC5EF9C690C2AB2ECDF6A96EF1C139CC0B2F284DCA0A9A7943388A49A3AEE664BA5379A7655D3C68900BE2F6903
    assert y == 0x0B9C15F3FE6E5CF4211F346271D7B01C8F3B28BE689C8429C85B67AF215533311F0B8DFAAA154FA6B88176C229F2885D


def test_rfc9380_g2_vector():
    dst = b"QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_RO_"
    x, y = O.hash_to_curve_g2(b"", dst)
    assert x == (0x0141EBFBDCA40EB85B87142E130AB689C673CF60F1A3E98D69335266F30D9B8D4AC44C1038E9DCDD5393FAF5C41FB78A,
                 0x05CB8437535E20ECFFAEF7752BADDF98034139C38452458BAEEFAB379BA13DFF5BF5DD71B72418717047F5B0F37DA03D)
    assert y == (0x0503921D7F6A12805E72940B963C0CF3471C7B2A524950CA195D11062EE75EC076DAF2D4BC358C4B190C0C98064FDD92,
                 0x12424AC32561493F3FE3C260708A12B7C620E7BE00099A974E259DDC7D1F6395C3C811CDD19F1E8DBF3E9ECFDCBAB8D6)


def test_expand_message_xmd_rfc_vector():
    # RFC 9380 K.1, DST = QUUX-V01-CS02-with-expander-SHA256-128, msg = "", len 0x20
    out = O.expand_message_xmd(b"", b"QUUX-V01-CS02-with-expander-SHA256-128", 32)
    assert out.hex() == "68a985b87eb6b46952128911f2a4412bbc302a9d759667f87f7a21d803f07235"


def test_pairing_bilinearity_and_g1impl_roundtrip():
    # G1Impl has no absolute vectors in the reference: pinned by RFC J.9.1 above + this round trip
    a, b = 0x1234567, 0x89ABCDE
    lhs = O.pairing(O.g1_mul(O.G1_GEN, a), O.g2_mul(O.G2_GEN, b))
    rhs = O.f12_pow(O.pairing(O.G1_GEN, O.G2_GEN), a * b)
    assert lhs == rhs and lhs != O.F12_ONE
    sk = 0x2A06164DAE4751E566EE2854F2865F782F28E2420EC2ADE059ED434919B67B5D
    msg = b"g1impl message"
    for scheme in (O.BASIC, O.AUG, O.POP):
        pk = O.g2_serialize(O.sk_to_pk(O.G1IMPL, sk))
        sig = O.g1_serialize(O.sign(O.G1IMPL, scheme, sk, msg))
        assert O.verify(O.G1IMPL, scheme, O.MODERN, pk, sig, msg) == O.OK
        assert O.verify(O.G1IMPL, scheme, O.MODERN, pk, sig, msg + b"x") == O.ERR_INVALID_SIGNATURE


def test_identity_and_header_rules():
    # legacy_test.rs:26-35,70-106 ; legacy_comprehensive_test.rs:211-240,379-402 ; sig_core.rs:126-135
    ident1 = bytes([0xC0]) + bytes(47)
    ident2 = bytes([0xC0]) + bytes(95)
    assert O.g1_deserialize(ident1, O.MODERN) is None and O.g1_deserialize(ident1, O.LEGACY) is None
    assert O.g1_serialize(None, O.LEGACY) == ident1 and O.g2_serialize(None, O.LEGACY) == ident2
    pk = O.g1_serialize(O.G1_GEN)
    sig = O.g2_serialize(O.G2_GEN)
    assert O.verify(O.G2IMPL, O.BASIC, O.MODERN, pk, ident2, b"m") == O.ERR_SIG_IDENTITY
    assert O.verify(O.G2IMPL, O.BASIC, O.MODERN, ident1, sig, b"m") == O.ERR_PK_IDENTITY
    assert O.verify(O.G2IMPL, O.BASIC, O.MODERN, ident1, ident2, b"m") == O.ERR_SIG_IDENTITY
    # Modern bytes with y-flag 0 decode under Legacy as -P; with y-flag 1 they are a LegacyFormatError
    p = O.g1_mul(O.G1_GEN, 5)
    for q in (p, O.g1_neg(p)):
        m = O.g1_serialize(q, O.MODERN)
        if m[0] & 0x20:
            with pytest.raises(O.BlsError) as e:
                O.g1_deserialize(m, O.LEGACY)
            assert e.value.code == O.ERR_LEGACY_FORMAT
        else:
            assert O.g1_deserialize(m, O.LEGACY) == O.g1_neg(q)
        leg = O.g1_serialize(q, O.LEGACY)
        assert O.g1_deserialize(leg, O.LEGACY) == q
        if leg[0] < 0x80:  # legacy sign 0 always fails the Modern decoder (legacy_test.rs:85-105)
            with pytest.raises(O.BlsError) as e:
                O.g1_deserialize(leg, O.MODERN)
            assert e.value.code == O.ERR_DESERIALIZE
    with pytest.raises(O.BlsError) as e:
        O.g1_deserialize(bytes(47), O.MODERN)
    assert e.value.code == O.ERR_INVALID_LENGTH
    # x >= p, off-curve, on-curve-but-outside-subgroup
    bad = bytearray((O.P + 1).to_bytes(48, "big")); bad[0] |= 0x80
    with pytest.raises(O.BlsError):
        O.g1_deserialize(bytes(bad), O.MODERN)
    x = 1
    while True:
        y = O.fp_sqrt((x ** 3 + 4) % O.P)
        if y is not None and not O.g1_in_subgroup((x, y)):
            break
        x += 1
    enc = bytearray(x.to_bytes(48, "big")); enc[0] |= 0x80
    with pytest.raises(O.BlsError) as e:
        O.g1_deserialize(bytes(enc), O.MODERN)
    assert e.value.code == O.ERR_DESERIALIZE


def test_aggregate_verify_semantics():
    # tests/signatures.rs:130-173 ; sig_basic.rs:46-58 ; sig_core.rs:149-178
    sks = [11, 22, 33]
    msgs = [b"m0", b"m1", b"m2"]
    for scheme in (O.BASIC, O.AUG):
        pks = [O.g1_serialize(O.sk_to_pk(O.G2IMPL, sk)) for sk in sks]
        agg = None
        for sk, m in zip(sks, msgs):
            agg = O.g2_add(agg, O.sign(O.G2IMPL, scheme, sk, m))
        sig = O.g2_serialize(agg)
        assert O.aggregate_verify(O.G2IMPL, scheme, O.MODERN, pks, msgs, sig)[0] == O.OK
        assert O.aggregate_verify(O.G2IMPL, scheme, O.MODERN, pks, [b"m0", b"mX", b"m2"], sig)[0] \
            == O.ERR_INVALID_SIGNATURE
    pks = [O.g1_serialize(O.sk_to_pk(O.G2IMPL, sk)) for sk in sks]
    st, idx = O.aggregate_verify(O.G2IMPL, O.BASIC, O.MODERN, pks, [b"m0", b"m1", b"m0"], sig)
    assert st == O.ERR_DUPLICATE_MESSAGES and idx == (0, 2)
    assert O.aggregate_verify(O.G2IMPL, O.BASIC, O.MODERN, [], [], sig)[0] == O.ERR_INVALID_SIGNATURE
    st, idx = O.aggregate_verify(O.G2IMPL, O.POP, O.MODERN, [pks[0], bytes([0xC0]) + bytes(47)], msgs[:2], sig)
    assert st == O.ERR_PK_IDENTITY and idx == (2,)


def test_verify_secure_edge_semantics():
    # secure_aggregation.rs:542-560 (empty list), :125-129 (length mismatch)
    ident2 = bytes([0xC0]) + bytes(95)
    assert O.verify_secure(O.G2IMPL, O.BASIC, O.MODERN, [], ident2, b"m") == O.OK
    assert O.verify_secure(O.G2IMPL, O.BASIC, O.MODERN, [], O.g2_serialize(O.G2_GEN), b"m") \
        == O.ERR_INVALID_SIGNATURE
    assert O.aggregate_secure(O.G2IMPL, O.MODERN, [O.g1_serialize(O.G1_GEN)], [])[0] == O.ERR_MISMATCHED_LENGTHS
    assert O.aggregate_secure(O.G2IMPL, O.MODERN, [], []) == (O.OK, ident2)


def _shamir_shares(secret, threshold, ids, seed):
    import random
    rnd = random.Random(seed)
    coef = [secret] + [rnd.randrange(O.R) for _ in range(threshold - 1)]
    return [sum(c * pow(x, k, O.R) for k, c in enumerate(coef)) % O.R for x in ids]


def test_combine_shares_reproduces_the_golden_signature(cpp):
    """Signature::from_shares / PublicKey::from_shares (signature.rs:151-165): Shamir shares of a golden secret key give
    signature / public-key shares whose combination must be the reference's own golden sig / pk bytes
    (tests/cpp_integration_test.rs:19-82) - an absolute pin for the Lagrange combination."""
    msg = bytes.fromhex(cpp["message"])
    s = cpp["signers"][0]
    sk = int(s["sk"], 16)
    ids = [1, 2, 5, 2 ** 130 + 3]
    sks = _shamir_shares(sk, 3, ids, seed=4)
    H = O.hash_to_curve_g2(msg, O.sig_dst(O.G2IMPL, O.BASIC))
    sig_shares = [x.to_bytes(32, "big") + O.g2_serialize(O.g2_mul(H, k)) for x, k in zip(ids, sks)]
    pk_shares = [x.to_bytes(32, "big") + O.g1_serialize(O.g1_mul(O.G1_GEN, k)) for x, k in zip(ids, sks)]
    for subset in ([0, 1, 2], [3, 1, 0], [0, 1, 2, 3]):
        st, sig = O.combine_shares(2, [sig_shares[i] for i in subset])
        assert st == O.OK and sig.hex() == s["sig"]
        st, pk = O.combine_shares(1, [pk_shares[i] for i in subset])
        assert st == O.OK and pk.hex() == s["pk"]
    # two shares of a degree-2 polynomial combine to something else (no error: the threshold is not known to combine)
    st, sig = O.combine_shares(2, sig_shares[:2])
    assert st == O.OK and sig.hex() != s["sig"]
    # vsss errors: fewer than two shares, zero identifier, duplicate identifier; parse errors: identifier >= r, bad point
    assert O.combine_shares(2, sig_shares[:1])[0] == O.ERR_VSSS
    assert O.combine_shares(2, [bytes(32) + sig_shares[0][32:], sig_shares[1]])[0] == O.ERR_VSSS
    assert O.combine_shares(2, [sig_shares[0], sig_shares[0][:32] + sig_shares[1][32:]])[0] == O.ERR_VSSS
    assert O.combine_shares(2, [O.R.to_bytes(32, "big") + sig_shares[0][32:], sig_shares[1]])[0] == O.ERR_DESERIALIZE
    assert O.combine_shares(2, [sig_shares[0], sig_shares[1][:32] + bytes(96)])[0] == O.ERR_DESERIALIZE


def test_verify_share_on_golden_triple(golden_dir):
    """oracle verify_share (PublicKeyShare::verify on raw share records): the share VALUES decide, identifiers only have
    to be canonical scalars (lib.rs:126-133).  Pinned on the reference's golden pk/sig pair (cpp_integration_test.rs)."""
    import json, os
    g = json.load(open(os.path.join(golden_dir, "cpp_integration.json")))
    msg = bytes.fromhex(g["message"])
    s0, s1 = g["signers"][0], g["signers"][1]
    pk = (5).to_bytes(32, "big") + bytes.fromhex(s0["pk"])
    sg = (9).to_bytes(32, "big") + bytes.fromhex(s0["sig"])
    assert O.verify_share(2, 0, pk, sg, msg) == 0
    assert O.verify_share(2, 0, pk, (9).to_bytes(32, "big") + bytes.fromhex(s1["sig"]), msg) == 1
    assert O.verify_share(2, 0, O.R.to_bytes(32, "big") + pk[32:], sg, msg) == 4
    assert O.verify_share(2, 0, pk, sg[:-1], msg) == 4
