"""CPU-only checks of the engine's device headers (csrc/*.cuh) compiled for the host by tests/hostemu/emu.cpp, against
the big-int oracle.  Two builds are exercised: the plain one, and one with -DBLS_TRACK, in which every Fp carries the
worst-case value/limb bounds of the lazy-reduction discipline (csrc/fp.cuh) and every operation aborts if a bound could
be exceeded for ANY input - so one pass over each code path proves the absence of limb/column overflow on the GPU too.
TEST INFRASTRUCTURE ONLY: nothing here is on the product path (the C-ABI library has no CPU fallback)."""
import ctypes
import hashlib
import json
import os
import random
import subprocess

import pytest

from oracle import bls_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "hostemu", "emu.cpp")
CSRC = os.path.join(ROOT, "agora-blsful_b200", "csrc")
P = O.P


def _build(out, flags):
    deps = [SRC] + [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    if os.path.exists(out) and all(os.path.getmtime(d) <= os.path.getmtime(out) for d in deps):
        return out
    subprocess.run(["g++", "-shared", "-fPIC", "-o", out, SRC] + flags, check=True)
    return out


@pytest.fixture(scope="module", params=["plain", "track"])
def L(request):
    if request.param == "plain":
        so = _build(os.path.join(ROOT, "tests", "_hostemu.so"), ["-O2"])
    else:
        so = _build(os.path.join(ROOT, "tests", "_hostemu_track.so"), ["-O1", "-g", "-DBLS_TRACK", "-rdynamic"])
    return ctypes.CDLL(so)


def be(v):
    return v.to_bytes(48, "big")


def buf(n):
    return ctypes.create_string_buffer(n)


def limbs(k, n):
    return (ctypes.c_uint32 * n)(*[(k >> (32 * i)) & 0xFFFFFFFF for i in range(n)])


def f2b(x):
    return be(x[0]) + be(x[1])


def bf2(b):
    return (int.from_bytes(b[:48], "big"), int.from_bytes(b[48:96], "big"))


def f12b(f):
    return b"".join(be(c[0]) + be(c[1]) for c in f)


def bf12(b):
    return tuple(bf2(b[96 * i:96 * i + 96]) for i in range(6))


def test_fp_ops(L):
    rnd = random.Random(1)
    for i in range(600):
        a, b = rnd.randrange(P), rnd.randrange(P)
        if i < 4:
            a = [0, 1, P - 1, P - 2][i]
        o, o2, o3 = buf(48), buf(48), buf(48)
        L.emu_fp_mul(be(a), be(b), o, 0)
        L.emu_fp_mul(be(a), be(b), o2, 1)
        L.emu_fp_mul(be(a), be(a), o3, 2)
        assert int.from_bytes(o.raw, "big") == a * b % P
        assert int.from_bytes(o2.raw, "big") == a * b % P
        assert int.from_bytes(o3.raw, "big") == a * a % P
        s, d, n = buf(48), buf(48), buf(48)
        L.emu_fp_addsub(be(a), be(b), s, d, n)
        assert int.from_bytes(s.raw, "big") == (a + b) % P
        assert int.from_bytes(d.raw, "big") == (a - b) % P
        assert int.from_bytes(n.raw, "big") == (-a) % P
    for i in range(12):
        a = rnd.randrange(1, P)
        o = buf(48)
        L.emu_fp_inv(be(a), o)
        assert int.from_bytes(o.raw, "big") * a % P == 1
        ok = L.emu_fp_sqrt(be(a), o)
        r = int.from_bytes(o.raw, "big")
        assert bool(ok) == (pow(a, (P - 1) // 2, P) == 1)
        if ok:
            assert r * r % P == a


def test_fp2_sqrt_ratio(L):
    rnd = random.Random(11)
    nsq = 0
    for i in range(40):
        n = (rnd.randrange(P), rnd.randrange(P))
        d = (rnd.randrange(P), rnd.randrange(1, P))
        if i % 5 == 0:
            n = (n[0], 0)
        if i % 7 == 0:
            d = (1, 0)
        if i == 3:
            n = (0, n[1])
        o = buf(96)
        ok = L.emu_fp2_sqrt_ratio(f2b(n), f2b(d), o)
        r = bf2(o.raw)
        ratio = O.f2_mul(n, O.f2_inv(d))
        if ok:
            assert O.f2_sqr(r) == ratio
        else:
            nsq += 1
            assert O.f2_sqr(r) == O.f2_mul(O.SSWU_G2_Z, ratio)
        assert bool(ok) == (O.f2_sqrt(ratio) is not None)
    assert nsq > 5
    a, b = (rnd.randrange(P), rnd.randrange(P)), (rnd.randrange(P), rnd.randrange(P))
    m, s, v = buf(96), buf(96), buf(96)
    L.emu_fp2_mul(f2b(a), f2b(b), m, s, v)
    assert bf2(m.raw) == O.f2_mul(a, b) and bf2(s.raw) == O.f2_sqr(a) and O.f2_mul(bf2(v.raw), a) == (1, 0)


def test_fp12_tower_and_final_exp(L):
    rnd = random.Random(3)
    a = tuple((rnd.randrange(P), rnd.randrange(P)) for _ in range(6))
    b = tuple((rnd.randrange(P), rnd.randrange(P)) for _ in range(6))
    o = [buf(576) for _ in range(5)]
    L.emu_fp12_ops(f12b(a), f12b(b), *o)
    assert bf12(o[0].raw) == O.f12_mul(a, b)
    assert bf12(o[1].raw) == O.f12_mul(a, a)
    assert O.f12_mul(bf12(o[2].raw), a) == O.F12_ONE
    assert bf12(o[3].raw) == O.f12_pow(a, P)
    assert bf12(o[4].raw) == O.f12_pow(a, P * P)
    assert L.emu_cyclo_check(f12b(a)) == 1
    fe = buf(576)
    L.emu_final_exp(f12b(a), fe)
    ref = O.final_exponentiation(a)
    assert bf12(fe.raw) == O.f12_mul(O.f12_mul(ref, ref), ref)  # the engine computes the cube (fp12.cuh)


def test_pairing_and_golden_verify(L, golden_dir):
    rnd = random.Random(5)
    pa, qb = O.g1_mul(O.G1_GEN, 7), O.g2_mul(O.G2_GEN, 11)
    ml, fe = buf(576), buf(576)
    assert L.emu_pairing(O.g1_serialize(pa), O.g2_serialize(qb), limbs(0, 1), 0, ml, fe) == 0
    e = O.pairing(pa, qb)
    assert bf12(fe.raw) == O.f12_pow(e, 3)
    assert O.final_exponentiation(bf12(ml.raw)) == e
    k = rnd.randrange(2 ** 64)
    assert L.emu_pairing(O.g1_serialize(pa), O.g2_serialize(qb), limbs(k, 2), 2, ml, fe) == 0
    assert bf12(fe.raw) == O.f12_pow(e, 3 * k)  # Jacobian (scaled) G1 argument
    g = json.load(open(os.path.join(golden_dir, "cpp_integration.json")))
    s = g["signers"][0]
    H = O.hash_to_curve_g2(bytes.fromhex(g["message"]), O.sig_dst(2, 0))
    ng = O.g1_serialize(O.g1_neg(O.G1_GEN))
    assert L.emu_pairing_check2(bytes.fromhex(s["pk"]), O.g2_serialize(H), ng, bytes.fromhex(s["sig"])) == 1
    assert L.emu_pairing_check2(bytes.fromhex(s["pk"]), O.g2_serialize(H), ng,
                                bytes.fromhex(g["signers"][1]["sig"])) == 0


def test_curve_ops_and_codecs(L):
    rnd = random.Random(2)
    g1, g2 = O.g1_serialize(O.G1_GEN), O.g2_serialize(O.G2_GEN)
    for _ in range(2):
        k, k2 = rnd.randrange(O.R), rnd.randrange(O.R)
        o = buf(48)
        assert L.emu_g1_mul(g1, limbs(k, 8), 8, 1, o) == 0 and o.raw == O.g1_serialize(O.g1_mul(O.G1_GEN, k))
        o2 = buf(96)
        assert L.emu_g2_mul(g2, limbs(k, 8), 8, 1, o2) == 0 and o2.raw == O.g2_serialize(O.g2_mul(O.G2_GEN, k))
        a, b = O.g1_mul(O.G1_GEN, k), O.g1_mul(O.G1_GEN, k2)
        for x, y in ((a, b), (a, a), (a, O.g1_neg(a))):
            assert L.emu_g1_add(O.g1_serialize(x), O.g1_serialize(y), o) == 0
            assert o.raw == O.g1_serialize(O.g1_add(x, y))
        a, b = O.g2_mul(O.G2_GEN, k), O.g2_mul(O.G2_GEN, k2)
        for x, y in ((a, b), (a, a)):
            assert L.emu_g2_add(O.g2_serialize(x), O.g2_serialize(y), o2) == 0
            assert o2.raw == O.g2_serialize(O.g2_add(x, y))
    xy, inf = buf(192), ctypes.c_int(0)
    for fmt in (0, 1):
        for k in (1, 2, 12345, O.R - 1):
            pt = O.g1_mul(O.G1_GEN, k)
            assert L.emu_g1_decompress(O.g1_serialize(pt, fmt), fmt, xy, ctypes.byref(inf)) == 0
            assert (int.from_bytes(xy.raw[:48], "big"), int.from_bytes(xy.raw[48:96], "big")) == pt
            pt = O.g2_mul(O.G2_GEN, k)
            assert L.emu_g2_decompress(O.g2_serialize(pt, fmt), fmt, xy, ctypes.byref(inf)) == 0
            assert (bf2(xy.raw[:96]), bf2(xy.raw[96:192])) == pt
    cnt, x = 0, 5
    while cnt < 3:  # on-curve points outside G1 / x without a y
        x += 1
        y = O.fp_sqrt((x ** 3 + 4) % P)
        enc = bytearray(be(x))
        enc[0] |= 0x80
        assert L.emu_g1_decompress(bytes(enc), 1, xy, ctypes.byref(inf)) == 4
        if y is not None:
            assert L.emu_subgroup(bytes(enc), 0) == int(O.g1_in_subgroup((x, y)))
            cnt += 1
    cnt, x0 = 0, 7
    while cnt < 2:
        x0 += 1
        xx = (x0, 3)
        y = O.f2_sqrt(O.f2_add(O.f2_mul(O.f2_sqr(xx), xx), O.B2))
        enc = bytearray(be(xx[1]) + be(xx[0]))
        enc[0] |= 0x80
        if y is None:
            assert L.emu_g2_decompress(bytes(enc), 1, xy, ctypes.byref(inf)) == 4
            continue
        assert L.emu_subgroup(bytes(enc), 1) == int(O.g2_in_subgroup((xx, y))) == 0
        cnt += 1


def test_sha256_and_hash_to_curve(L):
    rnd = random.Random(9)
    for n in (0, 1, 55, 56, 63, 64, 65, 200):
        m = bytes(rnd.randrange(256) for _ in range(n))
        o = buf(32)
        L.emu_sha256(m, n, o)
        assert o.raw == hashlib.sha256(m).digest()
    q2 = b"QUUX-V01-CS02-with-BLS12381G2_XMD:SHA-256_SSWU_RO_"
    q1 = b"QUUX-V01-CS02-with-BLS12381G1_XMD:SHA-256_SSWU_RO_"
    for msg, dst in ((b"", q2), (b"hello", O.sig_dst(2, 0)), (b"x" * 100, O.sig_dst(2, 1))):
        o = buf(96)
        L.emu_hash_to_g2(b"", 0, msg, len(msg), dst, len(dst), o)
        assert o.raw == O.g2_serialize(O.hash_to_curve_g2(msg, dst))
    for msg, dst in ((b"", q1), (b"abc", q1), (b"hello", O.sig_dst(1, 0))):
        o = buf(48)
        L.emu_hash_to_g1(b"", 0, msg, len(msg), dst, len(dst), o)
        assert o.raw == O.g1_serialize(O.hash_to_curve_g1(msg, dst))
    pk = O.g1_serialize(O.g1_mul(O.G1_GEN, 99))
    o = buf(96)
    L.emu_hash_to_g2(pk, 48, b"aug", 3, O.sig_dst(2, 1), 43, o)  # MessageAugmentation framing: pk || msg
    assert o.raw == O.g2_serialize(O.hash_to_curve_g2(pk + b"aug", O.sig_dst(2, 1)))


def test_sop2f_fused_unit(L):
    """csrc/sfp.cuh: signed-limb sum of Fp2 products with ONE reduction per coefficient (shifts, xi, negation, the
    Fp-scalar mode, output aliasing)."""
    rnd = random.Random(21)
    XI = (1, 1)
    mul, add, sub = O.f2_mul, O.f2_add, O.f2_sub
    sc = lambda s, x: ((s * x[0]) % P, (s * x[1]) % P)
    for it in range(40):
        v = [(rnd.randrange(P), rnd.randrange(P)) for _ in range(6)]
        if it == 0:
            v = [(P - 1, P - 1)] * 6
        if it == 1:
            v = [(0, 0), (P - 1, 0), (0, P - 1), (1, 0), (0, 1), (P - 1, 1)]
        a, b, c, d, e, f = v
        k = rnd.randrange(P)
        out = buf(288)
        L.emu_sop2f(b"".join(f2b(x) for x in v), be(k), out)
        want = sub(add(mul(a, b), mul(mul(XI, sc(2, c)), d)), mul(e, sc(4, f)))
        assert bf2(out.raw[:96]) == want
        assert bf2(out.raw[96:192]) == mul(want, want)
        assert bf2(out.raw[192:]) == sub(sc(k, a), sc(k, mul(XI, b)))


def test_miller6_cooperative(L):
    """csrc/miller6.cuh: six pairings, one shared accumulator spread over six lanes == the product of the pairings."""
    rnd = random.Random(77)
    for n, nl, two_lane in ((6, 2, 1), (1, 0, 0), (3, 2, 0), (2, 0, 1)):
        ps = [O.g1_mul(O.G1_GEN, rnd.randrange(1, O.R)) for _ in range(n)]
        qs = [O.g2_mul(O.G2_GEN, rnd.randrange(1, O.R)) for _ in range(n)]
        ks = [rnd.randrange(1, 2 ** 64) for _ in range(n)]
        karr = (ctypes.c_uint32 * (2 * n))(*[w for k in ks for w in (k & 0xFFFFFFFF, k >> 32)])
        ml, fe = buf(576), buf(576)
        assert L.emu_miller6(n, b"".join(O.g1_serialize(p) for p in ps), b"".join(O.g2_serialize(q) for q in qs),
                             karr, nl, two_lane, ml, fe) == 0
        want = O.F12_ONE
        for p, q, k in zip(ps, qs, ks):
            want = O.f12_mul(want, O.f12_pow(O.pairing(p, q), 3 * (k if nl else 1)))
        assert bf12(fe.raw) == want


def test_final6_cooperative_final_exponentiation(L):
    """finalexp6.cuh: the six-lane final exponentiation equals the one-thread tower version (fp12.cuh), which equals the
    big-int oracle's plain power cubed; the six-lane general product equals the oracle's Fp12 product."""
    rnd = random.Random(66)
    def rand12():
        return tuple((rnd.randrange(P), rnd.randrange(P)) for _ in range(6))
    for trial in range(2):
        f, g = rand12(), rand12()
        fe, mul = buf(576), buf(576)
        one = L.emu_final6(f12b(f), f12b(g), fe, mul)
        assert mul.raw == f12b(O.f12_mul(f, g))
        want = buf(576)
        L.emu_final_exp(f12b(f), want)
        assert fe.raw == want.raw and one == 0
        if trial == 0:
            assert fe.raw == f12b(O.f12_pow(O.final_exponentiation(f), 3))
    # a value whose exponentiation IS one: e(aP, Q) * e(-P, aQ)
    a = 0xabcdef
    m = O.f12_mul(O.miller_loop(O.g1_mul(O.G1_GEN, a), O.G2_GEN), O.miller_loop(O.g1_neg(O.G1_GEN), O.g2_mul(O.G2_GEN, a)))
    fe, mul = buf(576), buf(576)
    assert L.emu_final6(f12b(m), f12b(O.F12_ONE), fe, mul) == 1
    assert fe.raw == f12b(O.F12_ONE) and mul.raw == f12b(m)


def test_windowed_scalar_multiplication(L):
    """csrc/curve.cuh jac_mul_aff_w4_64 (the r_i * pk_i of the batch equation): zero digits, zero top window, all-ones."""
    rnd = random.Random(9)
    L.emu_g1_mul_w4.argtypes = [ctypes.c_char_p, ctypes.c_uint64, ctypes.c_char_p]
    pt = O.g1_mul(O.G1_GEN, 0xABCDEF)
    ser = O.g1_serialize(pt)
    ks = [1, 2, 3, 15, 16, 17, 2 ** 63, 2 ** 64 - 1, 2 ** 64 - 2, 0xF000000000000000, 0x000000000000F000, 0x1010101010101010]
    ks += [rnd.randrange(1, 2 ** 64) for _ in range(12)]
    for k in ks:
        o = buf(48)
        assert L.emu_g1_mul_w4(ser, k, o) == 0
        assert o.raw == O.g1_serialize(O.g1_mul(pt, k)), hex(k)


def test_shared_inversion_normalisation(L):
    """csrc/curve.cuh jac_to_aff_batch (hash_to_curve outputs, 16 per inversion): ragged counts, identities anywhere."""
    rnd = random.Random(5)
    L.emu_to_aff_batch.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_int, ctypes.c_uint64, ctypes.c_char_p]
    for g2 in (0, 1):
        gen, mul, ser, ln = (O.G2_GEN, O.g2_mul, O.g2_serialize, 96) if g2 else (O.G1_GEN, O.g1_mul, O.g1_serialize, 48)
        for cnt, inf_at in [(1, ()), (1, (0,)), (2, (1,)), (5, (0, 4)), (16, ()), (16, (0, 7, 15)), (3, (0, 1, 2))]:
            pts = [None if i in inf_at else mul(gen, rnd.randrange(1, O.R)) for i in range(cnt)]
            k = rnd.randrange(2, 2 ** 64)
            o = buf(ln * cnt)
            assert L.emu_to_aff_batch(g2, b"".join(ser(p) for p in pts), cnt, k, o) == 0
            want = b"".join(ser(None if p is None else mul(p, k)) for p in pts)
            assert o.raw == want, (g2, cnt, inf_at)


def test_fr_arithmetic_and_lagrange(L):
    """csrc/fr.cuh: scalar-field Montgomery arithmetic and the Lagrange basis at zero used by share combination."""
    rnd = random.Random(33)
    R = O.R
    b32 = lambda v: v.to_bytes(32, "big")
    for it in range(40):
        a, b = rnd.randrange(1, R), rnd.randrange(R)
        if it == 0:
            a, b = R - 1, R - 1
        if it == 1:
            a, b = 1, 0
        m, inv = buf(32), buf(32)
        assert L.emu_fr_ops(b32(a), b32(b), m, inv) == 0
        assert int.from_bytes(m.raw, "big") == a * b % R
        assert int.from_bytes(inv.raw, "big") == pow(a, -1, R)
    assert L.emu_fr_ops(b32(R), b32(1), buf(32), buf(32)) == -1
    for ids in ([1, 2, 3], [5, 1, 9, 2 ** 200 + 7, R - 1], list(range(1, 41))):
        for i in range(len(ids)):
            want = 1
            for j, x in enumerate(ids):
                if j != i:
                    want = want * x % R * pow((x - ids[i]) % R, -1, R) % R
            o = buf(32)
            assert L.emu_fr_lagrange(b"".join(b32(x) for x in ids), len(ids), i, o) == 0
            assert int.from_bytes(o.raw, "big") == want
    assert L.emu_fr_lagrange(b32(4) + b32(7) + b32(4), 3, 0, buf(32)) == 1  # duplicate identifier
