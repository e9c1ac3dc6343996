"""CPU ORACLE (test infrastructure, never shipped, never on the product path).

Big-int restatement of the verification hot path of dashpay/agora-blsful (blsful 3.0.0-pre8).
The reference holds NO arithmetic: field/curve/pairing/hash-to-curve code lives in the un-vendored
crates `blstrs_plus 0.8` (-> blst) / `bls12_381_plus 0.8` (reference Cargo.toml:20-28, no lockfile).
This file therefore restates the *published* algorithms those crates implement
(RFC 9380 hash_to_curve suites BLS12381G1/G2_XMD:SHA-256_SSWU_RO_, the ZCash/IETF compressed
encoding, the optimal-ate pairing) and follows the reference's own call order, DSTs, byte formats
and error semantics, each function citing the reference file:line it mirrors.

Parity pinning: tests/test_oracle_golden.py checks this file against every golden vector the
reference's tests hold for the path (tests/cpp_integration_test.rs:19-192,
tests/secure_aggregation_test.rs:143-235) and the RFC 9380 J.9.1 / J.10.1 vectors.
Unpinned (no absolute vectors in the reference): G1Impl signatures (pinned by RFC 9380 J.9.1
instead), Legacy-mode verify_secure bytes, Gt serialisation.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline leg may import this module.

Deliberately textbook: affine curve arithmetic with inversions, Fp12 as Fp2[w]/(w^6 - (1+u)) with
schoolbook multiplication, final exponentiation as one plain power by (p^12-1)/r, cofactor clearing
by the h_eff scalars. The CUDA engine uses different algorithms (projective lines, sparse products,
psi endomorphism, cyclotomic squarings), so agreement is meaningful.
"""
from __future__ import annotations

import hashlib
from typing import List, Optional, Sequence, Tuple

try:  # allow both `import oracle.bls_oracle` and path-based import
    from . import constants as K
except ImportError:  # pragma: no cover
    import constants as K

# ----------------------------------------------------------------------------------------------
# parameters
# ----------------------------------------------------------------------------------------------
P = 0x1A0111EA397FE69A4B1BA7B6434BACD764774B84F38512BF6730D2A0F6B0F6241EABFFFEB153FFFFB9FEFFFFFFFFAAAB
R = 0x73EDA753299D7D483339D80809A1D80553BDA402FFFE5BFEFFFFFFFF00000001
X_ABS = 0xD201000000010000  # the curve parameter is -X_ABS
H_EFF_G1 = 0xD201000000010001
H_EFF_G2 = K.H_EFF_G2

# status codes of the C ABI (include/blsgpu.h) == the reference's BlsError outcomes
OK = 0
ERR_INVALID_SIGNATURE = 1      # sig_core.rs:144,176
ERR_SIG_IDENTITY = 2           # sig_core.rs:126-130,155-159
ERR_PK_IDENTITY = 3            # sig_core.rs:131-135,162-167
ERR_DESERIALIZE = 4            # legacy.rs:110,121,154,165 ; public_key.rs:73
ERR_LEGACY_FORMAT = 5          # legacy.rs:53-58
ERR_INVALID_LENGTH = 6         # public_key.rs:159-164 ; signature.rs:236-241
ERR_INVALID_COEFFICIENT = 7    # secure_aggregation.rs:98-100
ERR_DUPLICATE_MESSAGES = 8     # sig_basic.rs:51-56
ERR_SCHEME = 9                 # aggregate_signature.rs:127-133 ; multi_signature.rs:84-96
ERR_MISMATCHED_LENGTHS = 10    # secure_aggregation.rs:125-129

G1IMPL, G2IMPL = 1, 2          # impls.rs:102-109
BASIC, AUG, POP = 0, 1, 2      # sig_types.rs:8-12
LEGACY, MODERN = 0, 1          # serialization.rs:10-17


class BlsError(Exception):
    def __init__(self, code: int, msg: str = "", index: Optional[Tuple[int, ...]] = None):
        super().__init__(f"status {code}: {msg}")
        self.code = code
        self.index = index


def sig_dst(impl: int, scheme: int) -> bytes:
    """impls/g2.rs:107-118, impls/g1.rs:109-120."""
    g = b"G2" if impl == G2IMPL else b"G1"
    tag = {BASIC: b"NUL_", AUG: b"AUG_", POP: b"POP_"}[scheme]
    return b"BLS_SIG_BLS12381" + g + b"_XMD:SHA-256_SSWU_RO_" + tag


def pop_dst(impl: int) -> bytes:
    g = b"G2" if impl == G2IMPL else b"G1"
    return b"BLS_POP_BLS12381" + g + b"_XMD:SHA-256_SSWU_RO_POP_"


# ----------------------------------------------------------------------------------------------
# Fp, Fp2
# ----------------------------------------------------------------------------------------------
def fp_inv(a: int) -> int:
    return pow(a, P - 2, P)


def fp_sqrt(a: int) -> Optional[int]:
    s = pow(a, (P + 1) // 4, P)
    return s if s * s % P == a % P else None


Fp2 = Tuple[int, int]
F2_ZERO: Fp2 = (0, 0)
F2_ONE: Fp2 = (1, 0)


def f2_add(a, b): return ((a[0] + b[0]) % P, (a[1] + b[1]) % P)
def f2_sub(a, b): return ((a[0] - b[0]) % P, (a[1] - b[1]) % P)
def f2_neg(a): return (-a[0] % P, -a[1] % P)
def f2_mul(a, b): return ((a[0] * b[0] - a[1] * b[1]) % P, (a[0] * b[1] + a[1] * b[0]) % P)
def f2_sqr(a): return f2_mul(a, a)
def f2_muls(a, k: int): return (a[0] * k % P, a[1] * k % P)
def f2_conj(a): return (a[0], -a[1] % P)


def f2_inv(a):
    n = fp_inv((a[0] * a[0] + a[1] * a[1]) % P)
    return (a[0] * n % P, -a[1] * n % P)


def f2_pow(a, e: int):
    r = F2_ONE
    while e:
        if e & 1:
            r = f2_mul(r, a)
        a = f2_sqr(a)
        e >>= 1
    return r


def f2_sqrt(a) -> Optional[Fp2]:
    """Square root in Fp2 = Fp[u]/(u^2+1) through the norm (complex method)."""
    a0, a1 = a[0] % P, a[1] % P
    if a1 == 0:
        s = fp_sqrt(a0)
        if s is not None:
            return (s, 0)
        s = fp_sqrt(-a0 % P)
        return (0, s) if s is not None else None
    n = fp_sqrt((a0 * a0 + a1 * a1) % P)
    if n is None:
        return None
    half = fp_inv(2)
    for cand in ((a0 + n) * half % P, (a0 - n) * half % P):
        x = fp_sqrt(cand)
        if x is not None and x != 0:
            y = a1 * fp_inv(2 * x % P) % P
            if f2_sqr((x, y)) == (a0, a1):
                return (x, y)
    return None


def f2_sgn0(a) -> int:
    """RFC 9380 4.1 sgn0 for m=2."""
    s0, z0 = a[0] & 1, a[0] == 0
    return s0 | (z0 and (a[1] & 1))


def fp_lex_largest(y: int) -> bool:
    return y > (P - 1) // 2


def f2_lex_largest(y) -> bool:
    """compare c1 first, then c0."""
    if y[1] != 0:
        return y[1] > (P - 1) // 2
    return y[0] > (P - 1) // 2


# ----------------------------------------------------------------------------------------------
# generic short-Weierstrass affine arithmetic over a field given by an ops record
# point = None (identity) or (x, y)
# ----------------------------------------------------------------------------------------------
class _Ops:
    def __init__(self, add, sub, mul, inv, neg, zero, one, muls):
        self.add, self.sub, self.mul, self.inv, self.neg = add, sub, mul, inv, neg
        self.zero, self.one, self.muls = zero, one, muls


OPS1 = _Ops(lambda a, b: (a + b) % P, lambda a, b: (a - b) % P, lambda a, b: a * b % P, fp_inv,
            lambda a: -a % P, 0, 1, lambda a, k: a * k % P)
OPS2 = _Ops(f2_add, f2_sub, f2_mul, f2_inv, f2_neg, F2_ZERO, F2_ONE, f2_muls)


def ec_neg(F: _Ops, p):
    return None if p is None else (p[0], F.neg(p[1]))


def ec_add(F: _Ops, p, q, a_coeff=None):
    """Affine addition on y^2 = x^3 + a x + b (a_coeff None means a = 0)."""
    if p is None:
        return q
    if q is None:
        return p
    x1, y1 = p
    x2, y2 = q
    if x1 == x2:
        if y1 != y2 or y1 == F.zero:
            return None
        num = F.muls(F.mul(x1, x1), 3)
        if a_coeff is not None:
            num = F.add(num, a_coeff)
        lam = F.mul(num, F.inv(F.muls(y1, 2)))
    else:
        lam = F.mul(F.sub(y2, y1), F.inv(F.sub(x2, x1)))
    x3 = F.sub(F.sub(F.mul(lam, lam), x1), x2)
    y3 = F.sub(F.mul(lam, F.sub(x1, x3)), y1)
    return (x3, y3)


def ec_mul(F: _Ops, p, k: int, a_coeff=None):
    if k < 0:
        return ec_mul(F, ec_neg(F, p), -k, a_coeff)
    r = None
    while k:
        if k & 1:
            r = ec_add(F, r, p, a_coeff)
        p = ec_add(F, p, p, a_coeff)
        k >>= 1
    return r


B1 = 4
B2: Fp2 = (4, 4)
G1_GEN = (0x17F1D3A73197D7942695638C4FA9AC0FC3688C4F9774B905A14E3A3F171BAC586C55E83FF97A1AEFFB3AF00ADB22C6BB,
          0x08B3F481E3AAA0F1A09E30ED741D8AE4FCF5E095D5D00AF600DB18CB2C04B3EDD03CC744A2888AE40CAA232946C5E7E1)
G2_GEN = ((0x024AA2B2F08F0A91260805272DC51051C6E47AD4FA403B02B4510B647AE3D1770BAC0326A805BBEFD48056C8C121BDB8,
           0x13E02B6052719F607DACD3A088274F65596BD0D09920B61AB5DA61BBDC7F5049334CF11213945D57E5AC7D055D042B7E),
          (0x0CE5D527727D6E118CC9CDC6DA2E351AADFD9BAA8CBDD3A76D429A695160D12C923AC9CC3BACA289E193548608B82801,
           0x0606C4A02EA734CC32ACD2B02BC28B99CB3E287E85A763AF267492AB572E99AB3F370D275CEC1DA1AAA9075FF05F79BE))


def g1_add(p, q): return ec_add(OPS1, p, q)
def g2_add(p, q): return ec_add(OPS2, p, q)
def g1_mul(p, k): return ec_mul(OPS1, p, k)
def g2_mul(p, k): return ec_mul(OPS2, p, k)
def g1_neg(p): return ec_neg(OPS1, p)
def g2_neg(p): return ec_neg(OPS2, p)


def g1_on_curve(p) -> bool:
    return p is None or (p[1] * p[1] - p[0] ** 3 - B1) % P == 0


def g2_on_curve(p) -> bool:
    if p is None:
        return True
    return f2_sub(f2_sqr(p[1]), f2_add(f2_mul(f2_sqr(p[0]), p[0]), B2)) == F2_ZERO


def g1_in_subgroup(p) -> bool:
    return g1_mul(p, R) is None


def g2_in_subgroup(p) -> bool:
    return g2_mul(p, R) is None


# ----------------------------------------------------------------------------------------------
# compressed encodings, Modern (ZCash/IETF) and Legacy (Dash/relic header)  -- legacy.rs:19-170
# ----------------------------------------------------------------------------------------------
def _modern_decode_fields(b: bytes):
    """Common flag handling of `G{1,2}Affine::from_compressed` (external crate; published ZCash format)."""
    c_flag, i_flag, s_flag = b[0] >> 7 & 1, b[0] >> 6 & 1, b[0] >> 5 & 1
    body = bytes([b[0] & 0x1F]) + b[1:]
    return c_flag, i_flag, s_flag, body


def g1_decompress_modern(b: bytes):
    if len(b) != 48:
        raise BlsError(ERR_INVALID_LENGTH, "G1 expects 48 bytes")
    c, i, s, body = _modern_decode_fields(b)
    if not c:
        raise BlsError(ERR_DESERIALIZE, "compression flag clear")
    x = int.from_bytes(body, "big")
    if i:
        if s or x != 0:
            raise BlsError(ERR_DESERIALIZE, "bad infinity encoding")
        return None
    if x >= P:
        raise BlsError(ERR_DESERIALIZE, "x not canonical")
    y = fp_sqrt((x * x * x + B1) % P)
    if y is None:
        raise BlsError(ERR_DESERIALIZE, "not on curve")
    if fp_lex_largest(y) != bool(s):
        y = -y % P
    pt = (x, y)
    if not g1_in_subgroup(pt):
        raise BlsError(ERR_DESERIALIZE, "not in subgroup")
    return pt


def g2_decompress_modern(b: bytes):
    if len(b) != 96:
        raise BlsError(ERR_INVALID_LENGTH, "G2 expects 96 bytes")
    c, i, s, body = _modern_decode_fields(b)
    if not c:
        raise BlsError(ERR_DESERIALIZE, "compression flag clear")
    x1 = int.from_bytes(body[:48], "big")
    x0 = int.from_bytes(body[48:], "big")
    if i:
        if s or x0 != 0 or x1 != 0:
            raise BlsError(ERR_DESERIALIZE, "bad infinity encoding")
        return None
    if x0 >= P or x1 >= P:
        raise BlsError(ERR_DESERIALIZE, "x not canonical")
    x = (x0, x1)
    y = f2_sqrt(f2_add(f2_mul(f2_sqr(x), x), B2))
    if y is None:
        raise BlsError(ERR_DESERIALIZE, "not on curve")
    if f2_lex_largest(y) != bool(s):
        y = f2_neg(y)
    pt = (x, y)
    if not g2_in_subgroup(pt):
        raise BlsError(ERR_DESERIALIZE, "not in subgroup")
    return pt


def g1_compress_modern(p) -> bytes:
    if p is None:
        return bytes([0xC0]) + bytes(47)
    b = bytearray(p[0].to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if fp_lex_largest(p[1]) else 0)
    return bytes(b)


def g2_compress_modern(p) -> bytes:
    if p is None:
        return bytes([0xC0]) + bytes(95)
    b = bytearray(p[0][1].to_bytes(48, "big") + p[0][0].to_bytes(48, "big"))
    b[0] |= 0x80 | (0x20 if f2_lex_largest(p[1]) else 0)
    return bytes(b)


def modern_to_legacy(b: bytes) -> bytes:
    """legacy.rs:19-35."""
    if b[0] == 0xC0:
        return b
    y_sign = b[0] & 0x20
    b0 = b[0] & 0x1F
    if y_sign:
        b0 |= 0x80
    return bytes([b0]) + b[1:]


def legacy_to_modern(b: bytes) -> bytes:
    """legacy.rs:39-67."""
    if b[0] == 0xC0:
        return b
    y_sign = b[0] & 0x80
    b0 = b[0] & 0x7F
    if b0 & 0xE0:
        raise BlsError(ERR_LEGACY_FORMAT, "unexpected bits in byte[0]")
    b0 |= 0x80
    if y_sign:
        b0 |= 0x20
    return bytes([b0]) + b[1:]


def validate_modern(b0: int):
    """legacy.rs:71-82."""
    if b0 != 0xC0 and (b0 & 0xC0) != 0x80:
        raise BlsError(ERR_DESERIALIZE, "expected bit pattern 10xxxxxx")


def g1_deserialize(b: bytes, fmt: int = MODERN):
    """legacy.rs:100-125 (LegacyG1Point::deserialize_g1); length check public_key.rs:159-164."""
    if len(b) != 48:
        raise BlsError(ERR_INVALID_LENGTH, "expected 48")
    if fmt == MODERN:
        validate_modern(b[0])
        return g1_decompress_modern(b)
    return g1_decompress_modern(legacy_to_modern(b))


def g2_deserialize(b: bytes, fmt: int = MODERN):
    """legacy.rs:144-169; length check signature.rs:236-241."""
    if len(b) != 96:
        raise BlsError(ERR_INVALID_LENGTH, "expected 96")
    if fmt == MODERN:
        validate_modern(b[0])
        return g2_decompress_modern(b)
    return g2_decompress_modern(legacy_to_modern(b))


def g1_serialize(p, fmt: int = MODERN) -> bytes:
    """legacy.rs:86-98."""
    b = g1_compress_modern(p)
    return b if fmt == MODERN else modern_to_legacy(b)


def g2_serialize(p, fmt: int = MODERN) -> bytes:
    """legacy.rs:130-142."""
    b = g2_compress_modern(p)
    return b if fmt == MODERN else modern_to_legacy(b)


# ----------------------------------------------------------------------------------------------
# RFC 9380 hash_to_curve  -- reference call sites impls/g1.rs:17-19, impls/g2.rs:15-17
# ----------------------------------------------------------------------------------------------
def expand_message_xmd(msg: bytes, dst: bytes, n: int) -> bytes:
    """RFC 9380 5.3.1 with SHA-256 (b=32, s=64)."""
    if len(dst) > 255:
        dst = hashlib.sha256(b"H2C-OVERSIZE-DST-" + dst).digest()
    ell = (n + 31) // 32
    assert ell <= 255 and n <= 65535
    dst_prime = dst + bytes([len(dst)])
    b0 = hashlib.sha256(bytes(64) + msg + n.to_bytes(2, "big") + b"\x00" + dst_prime).digest()
    bi = hashlib.sha256(b0 + b"\x01" + dst_prime).digest()
    out = bi
    for i in range(2, ell + 1):
        bi = hashlib.sha256(bytes(a ^ b for a, b in zip(b0, bi)) + bytes([i]) + dst_prime).digest()
        out += bi
    return out[:n]


def hash_to_field_fp(msg: bytes, dst: bytes, count: int) -> List[int]:
    u = expand_message_xmd(msg, dst, count * 64)
    return [int.from_bytes(u[64 * i:64 * i + 64], "big") % P for i in range(count)]


def hash_to_field_fp2(msg: bytes, dst: bytes, count: int) -> List[Fp2]:
    u = expand_message_xmd(msg, dst, count * 128)
    out = []
    for i in range(count):
        e0 = int.from_bytes(u[128 * i:128 * i + 64], "big") % P
        e1 = int.from_bytes(u[128 * i + 64:128 * i + 128], "big") % P
        out.append((e0, e1))
    return out


def _sswu(F: _Ops, u, A, Bc, Z, sqrt, sgn0):
    """RFC 9380 6.6.2 map_to_curve_simple_swu (the plain, non-optimised statement)."""
    u2 = F.mul(u, u)
    zu2 = F.mul(Z, u2)
    tv1 = F.add(F.mul(zu2, zu2), zu2)
    if tv1 == F.zero:
        x1 = F.mul(Bc, F.inv(F.mul(Z, A)))
    else:
        x1 = F.mul(F.mul(F.neg(Bc), F.inv(A)), F.add(F.one, F.inv(tv1)))
    gx1 = F.add(F.add(F.mul(F.mul(x1, x1), x1), F.mul(A, x1)), Bc)
    x2 = F.mul(zu2, x1)
    gx2 = F.add(F.add(F.mul(F.mul(x2, x2), x2), F.mul(A, x2)), Bc)
    y1 = sqrt(gx1)
    if y1 is not None:
        x, y = x1, y1
    else:
        x, y = x2, sqrt(gx2)
        assert y is not None
    if sgn0(u) != sgn0(y):
        y = F.neg(y)
    return (x, y)


def _poly(F: _Ops, coeffs, x):
    acc = F.zero
    for c in reversed(coeffs):
        acc = F.add(F.mul(acc, x), c)
    return acc


def iso11(pt):
    """RFC 9380 appendix E.2."""
    x, y = pt
    xn, xd = _poly(OPS1, K.ISO11_K1, x), _poly(OPS1, K.ISO11_K2, x)
    yn, yd = _poly(OPS1, K.ISO11_K3, x), _poly(OPS1, K.ISO11_K4, x)
    if xd == 0 or yd == 0:
        return None
    return (xn * fp_inv(xd) % P, y * yn % P * fp_inv(yd) % P)


def iso3(pt):
    """RFC 9380 appendix E.3."""
    x, y = pt
    xn, xd = _poly(OPS2, K.ISO3_K1, x), _poly(OPS2, K.ISO3_K2, x)
    yn, yd = _poly(OPS2, K.ISO3_K3, x), _poly(OPS2, K.ISO3_K4, x)
    if xd == F2_ZERO or yd == F2_ZERO:
        return None
    return (f2_mul(xn, f2_inv(xd)), f2_mul(f2_mul(y, yn), f2_inv(yd)))


SSWU_G2_A: Fp2 = (0, 240)
SSWU_G2_B: Fp2 = (1012, 1012)
SSWU_G2_Z: Fp2 = (-2 % P, -1 % P)


def map_to_curve_g1(u: int):
    return iso11(_sswu(OPS1, u, K.SSWU_G1_A, K.SSWU_G1_B, 11, fp_sqrt, lambda a: a & 1))


def map_to_curve_g2(u: Fp2):
    return iso3(_sswu(OPS2, u, SSWU_G2_A, SSWU_G2_B, SSWU_G2_Z, f2_sqrt, f2_sgn0))


def hash_to_curve_g1(msg: bytes, dst: bytes):
    u = hash_to_field_fp(msg, dst, 2)
    q = g1_add(map_to_curve_g1(u[0]), map_to_curve_g1(u[1]))
    return g1_mul(q, H_EFF_G1)


def hash_to_curve_g2(msg: bytes, dst: bytes):
    u = hash_to_field_fp2(msg, dst, 2)
    q = g2_add(map_to_curve_g2(u[0]), map_to_curve_g2(u[1]))
    return g2_mul(q, H_EFF_G2)


# ----------------------------------------------------------------------------------------------
# Fp12 = Fp2[w]/(w^6 - xi), xi = 1+u ; elements are 6-tuples of Fp2
# ----------------------------------------------------------------------------------------------
XI: Fp2 = (1, 1)
F12_ONE = (F2_ONE,) + (F2_ZERO,) * 5


def f12_mul(a, b):
    t = [F2_ZERO] * 11
    for i in range(6):
        ai = a[i]
        if ai == F2_ZERO:
            continue
        for j in range(6):
            bj = b[j]
            if bj == F2_ZERO:
                continue
            t[i + j] = f2_add(t[i + j], f2_mul(ai, bj))
    return tuple(f2_add(t[i], f2_mul(t[i + 6], XI)) if i < 5 else t[5] for i in range(6))


def f12_conj6(a):
    """a^(p^6): w -> -w."""
    return tuple(a[i] if i % 2 == 0 else f2_neg(a[i]) for i in range(6))


def f12_pow(a, e: int):
    r = F12_ONE
    for bit in bin(e)[2:]:
        r = f12_mul(r, r)
        if bit == "1":
            r = f12_mul(r, a)
    return r


FINAL_EXP = (P ** 12 - 1) // R


def final_exponentiation(f):
    return f12_pow(f, FINAL_EXP)


def _line(T, Q, Pt):
    """Line through T and Q (twist coordinates over Fp2) evaluated at Pt in E(Fp), scaled by w^3
    (an Fp4 element, erased by the final exponentiation). Returns (line, T+Q)."""
    xT, yT = T
    xQ, yQ = Q
    if T == Q:
        lam = f2_mul(f2_muls(f2_sqr(xT), 3), f2_inv(f2_muls(yT, 2)))
    else:
        lam = f2_mul(f2_sub(yQ, yT), f2_inv(f2_sub(xQ, xT)))
    x3 = f2_sub(f2_sub(f2_sqr(lam), xT), xQ)
    y3 = f2_sub(f2_mul(lam, f2_sub(xT, x3)), yT)
    xp, yp = Pt
    c0 = f2_sub(f2_mul(lam, xT), yT)
    c2 = f2_neg(f2_muls(lam, xp))
    c3 = (yp % P, 0)
    return (c0, F2_ZERO, c2, c3, F2_ZERO, F2_ZERO), (x3, y3)


def miller_loop(Pt, Q):
    """Optimal-ate Miller loop f_{|x|,Q}(P), conjugated because x < 0. Pt in G1 affine, Q in G2 affine."""
    if Pt is None or Q is None:
        return F12_ONE
    f = F12_ONE
    T = Q
    for bit in bin(X_ABS)[3:]:
        l, T2 = _line(T, T, Pt)
        f = f12_mul(f12_mul(f, f), l)
        T = T2
        if bit == "1":
            l, T2 = _line(T, Q, Pt)
            f = f12_mul(f, l)
            T = T2
    return f12_conj6(f)


def pairing_product_is_one(pairs: Sequence[Tuple[object, object]]) -> bool:
    """helpers.rs:41-63: multi_miller_loop(..).final_exponentiation() then Gt::is_identity."""
    f = F12_ONE
    for p1, q2 in pairs:
        f = f12_mul(f, miller_loop(p1, q2))
    return final_exponentiation(f) == F12_ONE


def pairing(p1, q2):
    return final_exponentiation(miller_loop(p1, q2))


# ----------------------------------------------------------------------------------------------
# impl-generic helpers: G2Impl = (pk in G1, sig in G2); G1Impl = (pk in G2, sig in G1)
# ----------------------------------------------------------------------------------------------
class Impl:
    def __init__(self, impl_id: int):
        self.id = impl_id
        if impl_id == G2IMPL:
            self.pk_len, self.sig_len = 48, 96
            self.pk_deser, self.sig_deser = g1_deserialize, g2_deserialize
            self.pk_ser, self.sig_ser = g1_serialize, g2_serialize
            self.pk_add, self.sig_add = g1_add, g2_add
            self.pk_mul, self.sig_mul = g1_mul, g2_mul
            self.pk_gen = G1_GEN
            self.pk_neg = g1_neg
            self.hash = hash_to_curve_g2
        else:
            self.pk_len, self.sig_len = 96, 48
            self.pk_deser, self.sig_deser = g2_deserialize, g1_deserialize
            self.pk_ser, self.sig_ser = g2_serialize, g1_serialize
            self.pk_add, self.sig_add = g2_add, g1_add
            self.pk_mul, self.sig_mul = g2_mul, g1_mul
            self.pk_gen = G2_GEN
            self.pk_neg = g2_neg
            self.hash = hash_to_curve_g1

    def pair(self, sig_side, pk_side):
        """helpers.rs:41-63: a (Signature-group, PublicKey-group) pair as (G1, G2)."""
        return (pk_side, sig_side) if self.id == G2IMPL else (sig_side, pk_side)


IMPLS = {G1IMPL: Impl(G1IMPL), G2IMPL: Impl(G2IMPL)}


# ----------------------------------------------------------------------------------------------
# the hot path
# ----------------------------------------------------------------------------------------------
def core_verify(impl: int, pk, sig, msg: bytes, dst: bytes) -> int:
    """sig_core.rs:120-146 on decoded points. Returns a status code."""
    C = IMPLS[impl]
    if sig is None:
        return ERR_SIG_IDENTITY
    if pk is None:
        return ERR_PK_IDENTITY
    a = C.hash(msg, dst)
    neg_g = C.pk_neg(C.pk_gen)
    ok = pairing_product_is_one([C.pair(a, pk), C.pair(sig, neg_g)])
    return OK if ok else ERR_INVALID_SIGNATURE


def _decode(fn, b, fmt):
    try:
        return fn(bytes(b), fmt), OK
    except BlsError as e:
        return None, e.code


def scheme_message(impl: int, scheme: int, pk, msg: bytes) -> bytes:
    """sig_aug.rs:20-24,41-47: MessageAugmentation prepends pk.to_bytes() (always Modern bytes)."""
    if scheme == AUG:
        return IMPLS[impl].pk_ser(pk, MODERN) + msg
    return msg


def verify(impl: int, scheme: int, fmt: int, pk_bytes: bytes, sig_bytes: bytes, msg: bytes) -> int:
    """Signature::verify (signature.rs:130-138) fed from bytes: decode pk, decode sig, then core_verify.
    Decode order pk-then-sig is this engine's batch convention (the reference decodes at parse time)."""
    C = IMPLS[impl]
    pk, st = _decode(C.pk_deser, pk_bytes, fmt)
    if st != OK:
        return st
    sig, st = _decode(C.sig_deser, sig_bytes, fmt)
    if st != OK:
        return st
    if sig is None:
        return ERR_SIG_IDENTITY
    if pk is None:
        return ERR_PK_IDENTITY
    return core_verify(impl, pk, sig, scheme_message(impl, scheme, pk, msg), sig_dst(impl, scheme))


def verify_share(impl: int, scheme: int, pk_share: bytes, sig_share: bytes, msg: bytes) -> int:
    """PublicKeyShare::verify / SignatureShare::verify (public_key_share.rs:55-71, signature_share.rs:98-101) fed from the
    raw share records of lib.rs:117-157 / 219-259 (32-byte big-endian identifier || compressed point): both records must
    parse (identifier a canonical scalar < r, lib.rs:126-133; the point through from_compressed), then the scheme's
    verify runs on the share VALUES - the identifiers play no part in it."""
    C = IMPLS[impl]
    pl, sl = (48, 96) if impl == G2IMPL else (96, 48)
    if len(pk_share) != 32 + pl or len(sig_share) != 32 + sl:
        return ERR_DESERIALIZE
    if int.from_bytes(pk_share[:32], "big") >= R or int.from_bytes(sig_share[:32], "big") >= R:
        return ERR_DESERIALIZE
    return verify(impl, scheme, MODERN, pk_share[32:], sig_share[32:], msg)


def pop_verify(impl: int, fmt: int, pk_bytes: bytes, sig_bytes: bytes) -> int:
    """sig_pop.rs:61-70 / proof_of_possession.rs:77-81."""
    C = IMPLS[impl]
    pk, st = _decode(C.pk_deser, pk_bytes, fmt)
    if st != OK:
        return st
    sig, st = _decode(C.sig_deser, sig_bytes, fmt)
    if st != OK:
        return st
    return core_verify(impl, pk, sig, C.pk_ser(pk, MODERN), pop_dst(impl))


def aggregate_verify(impl: int, scheme: int, fmt: int, pks: Sequence[bytes], msgs: Sequence[bytes],
                     sig_bytes: bytes) -> Tuple[int, Tuple[int, ...]]:
    """AggregateSignature::verify (aggregate_signature.rs:230-239) -> aggregate_verify
    (sig_basic.rs:41-64 / sig_aug.rs:27-38 / sig_pop.rs:52-58) -> core_aggregate_verify (sig_core.rs:149-178).
    Returns (status, indices): indices = (old, new) for duplicates, (i+1,) for identity pk, (i,) for a bad pk encoding."""
    C = IMPLS[impl]
    dec = []
    for i, b in enumerate(pks):
        pk, st = _decode(C.pk_deser, b, fmt)
        if st != OK:
            return st, (i,)
        dec.append(pk)
    sig, st = _decode(C.sig_deser, sig_bytes, fmt)
    if st != OK:
        return st, ()
    if scheme == BASIC:
        seen = {}
        for i, m in enumerate(msgs):
            m = bytes(m)
            if m in seen:
                return ERR_DUPLICATE_MESSAGES, (seen[m], i)
            seen[m] = i
    if sig is None:
        return ERR_SIG_IDENTITY, ()
    dst = sig_dst(impl, scheme)
    pairs = []
    for i, (pk, m) in enumerate(zip(dec, msgs)):
        if pk is None:
            return ERR_PK_IDENTITY, (i + 1,)
        pairs.append(C.pair(C.hash(scheme_message(impl, scheme, pk, bytes(m)), dst), pk))
    pairs.append(C.pair(sig, C.pk_neg(C.pk_gen)))
    return (OK if pairing_product_is_one(pairs) else ERR_INVALID_SIGNATURE), ()


def sum_points(group: int, fmt: int, encs: Sequence[bytes]) -> Tuple[int, bytes, int]:
    """aggregate_signatures/aggregate_public_keys (sig_core.rs:38-59), sig_multi.rs:7-13, pk_multi.rs:7-13.
    group 1 = G1 points (48 B), 2 = G2 points (96 B). Returns (status, compressed sum, first bad index)."""
    deser, ser, add = ((g1_deserialize, g1_serialize, g1_add) if group == 1
                       else (g2_deserialize, g2_serialize, g2_add))
    acc = None
    for i, b in enumerate(encs):
        pt, st = _decode(deser, b, fmt)
        if st != OK:
            return st, b"", i
        acc = add(acc, pt)
    return OK, ser(acc, fmt), -1


def secure_coefficients(sorted_pk_bytes: Sequence[bytes]) -> List[int]:
    """secure_aggregation.rs:44-103 / 284-330: base = SHA256(pk_0 || ...); t_i = BE(SHA256(be32(i)||base)) mod r.
    Reduce-mod-r semantics (SURVEY header fact 3); zero => InvalidCoefficient."""
    base = hashlib.sha256(b"".join(sorted_pk_bytes)).digest()
    out = []
    for i in range(len(sorted_pk_bytes)):
        t = int.from_bytes(hashlib.sha256(i.to_bytes(4, "big") + base).digest(), "big") % R
        if t == 0:
            raise BlsError(ERR_INVALID_COEFFICIENT, "zero coefficient")
        out.append(t)
    return out


def _sorted_keys(impl: int, fmt: int, pks: Sequence[bytes]):
    """Decode keys, re-serialise in `fmt` (canonical bytes) and sort ascending (secure_aggregation.rs:42, 278-283)."""
    C = IMPLS[impl]
    dec = []
    for i, b in enumerate(pks):
        pk, st = _decode(C.pk_deser, b, fmt)
        if st != OK:
            raise BlsError(st, "bad public key", (i,))
        dec.append((C.pk_ser(pk, fmt), pk, i))
    order = sorted(range(len(dec)), key=lambda j: dec[j][0])  # stable
    return dec, order


def verify_secure(impl: int, scheme: int, fmt: int, pks: Sequence[bytes], sig_bytes: bytes, msg: bytes) -> int:
    """Signature::verify_secure[_with_mode] (signature.rs:177-197, 256-276) ->
    verify_secure_with_dst_internal (secure_aggregation.rs:173-208). AUG uses its DST but no pk prefix (:236-247)."""
    C = IMPLS[impl]
    try:
        dec, order = _sorted_keys(impl, fmt, pks)
    except BlsError as e:
        return e.code
    sig, st = _decode(C.sig_deser, sig_bytes, fmt)
    if st != OK:
        return st
    if len(pks) == 0:
        return OK if sig is None else ERR_INVALID_SIGNATURE
    try:
        coeffs = secure_coefficients([dec[j][0] for j in order])
    except BlsError as e:
        return e.code
    agg = None
    for t, j in zip(coeffs, order):
        agg = C.pk_add(agg, C.pk_mul(dec[j][1], t))
    return core_verify(impl, agg, sig, bytes(msg), sig_dst(impl, scheme))


def aggregate_secure(impl: int, fmt: int, pks: Sequence[bytes], sigs: Sequence[bytes]) -> Tuple[int, bytes]:
    """aggregate_secure[_with_mode] (secure_aggregation.rs:110-169, 338-352). Duplicate keys reuse the FIRST
    matching original index (`position`, :140-147)."""
    C = IMPLS[impl]
    if len(pks) != len(sigs):
        return ERR_MISMATCHED_LENGTHS, b""
    if len(pks) == 0:
        return OK, C.sig_ser(None, fmt)
    try:
        dec, order = _sorted_keys(impl, fmt, pks)
    except BlsError as e:
        return e.code, b""
    dsig = []
    for b in sigs:
        s, st = _decode(C.sig_deser, b, fmt)
        if st != OK:
            return st, b""
        dsig.append(s)
    try:
        coeffs = secure_coefficients([dec[j][0] for j in order])
    except BlsError as e:
        return e.code, b""
    first = {}
    for idx, (ser, _, _) in enumerate(dec):
        first.setdefault(ser, idx)
    agg = None
    for t, j in zip(coeffs, order):
        agg = C.sig_add(agg, C.sig_mul(dsig[first[dec[j][0]]], t))
    return OK, C.sig_ser(agg, fmt)


# ----------------------------------------------------------------------------------------------
# test-data helpers (CPU only; signing is NOT part of the GPU path)
# ----------------------------------------------------------------------------------------------
def sk_to_pk(impl: int, sk: int):
    C = IMPLS[impl]
    return C.pk_mul(C.pk_gen, sk)


def sign(impl: int, scheme: int, sk: int, msg: bytes):
    """core_sign (sig_core.rs:108-118) incl. scheme framing."""
    C = IMPLS[impl]
    pk = sk_to_pk(impl, sk)
    return C.sig_mul(C.hash(scheme_message(impl, scheme, pk, msg), sig_dst(impl, scheme)), sk)


def rlc_scalars(seed: bytes, n: int) -> List[int]:
    """Deterministic non-zero 64-bit random-linear-combination scalars used by the engine's batch verify
    (engine convention, stated in DESIGN.md): r_i = LE64(SHA256(seed || LE64(i))[0:8]) | 1."""
    out = []
    for i in range(n):
        d = hashlib.sha256(seed + i.to_bytes(8, "little")).digest()
        out.append(int.from_bytes(d[:8], "little") | 1)
    return out


# ----------------------------------------------------------------------------------------------
# threshold shares on public data
# ----------------------------------------------------------------------------------------------
ERR_VSSS = 11  # BlsError::VsssError: every vsss_rs::Error maps to it (error.rs:24-26,60-64)


def combine_shares(group: int, shares: Sequence[bytes]) -> Tuple[int, bytes]:
    """Signature::from_shares / PublicKey::from_shares (signature.rs:151-165; sig_core.rs:92-105) = vsss-rs 5.0.0-rc2
    `combine` (dependency absent from /root/reference, Cargo.toml:43; published algorithm: at least two shares, no zero
    identifier, no duplicate identifier, then Lagrange interpolation at zero: sum_i value_i * prod_{j!=i} x_j/(x_j - x_i)).
    A share is the raw form of InnerPointShareG1/G2 (lib.rs:117-157): 32-byte big-endian identifier (must be < r:
    Scalar::from_be_bytes) followed by the IETF compressed point (from_compressed: curve + subgroup check)."""
    deser, ser, add, mul = ((g1_deserialize, g1_serialize, g1_add, g1_mul) if group == 1
                            else (g2_deserialize, g2_serialize, g2_add, g2_mul))
    length = 48 if group == 1 else 96
    ids, vals = [], []
    for sh in shares:
        if len(sh) != 32 + length:
            return ERR_DESERIALIZE, b""
        x = int.from_bytes(sh[:32], "big")
        if x >= R:
            return ERR_DESERIALIZE, b""
        pt, st = _decode(deser, sh[32:], MODERN)
        if st != OK:
            return ERR_DESERIALIZE, b""
        ids.append(x)
        vals.append(pt)
    if len(ids) < 2 or any(x == 0 for x in ids) or len(set(ids)) != len(ids):
        return ERR_VSSS, b""
    acc = None
    for i, (xi, v) in enumerate(zip(ids, vals)):
        lam = 1
        for j, xj in enumerate(ids):
            if j != i:
                lam = lam * xj % R * pow((xj - xi) % R, -1, R) % R
        acc = add(acc, mul(v, lam))
    return OK, ser(acc, MODERN)


# ----------------------------------------------------------------------------------------------
# the other public 2-pairing checks
# ----------------------------------------------------------------------------------------------
ERR_INVALID_PROOF = 12          # BlsError::InvalidProof (sig_proof.rs:140) / InvalidDecryptionShare (sign_decryption_share.rs:58)
ERR_COMMITMENT_IDENTITY = 13    # InvalidInputs("commitment is the identity point")   sig_proof.rs:110-114
ERR_PROOF_IDENTITY = 14         # InvalidInputs("proof is the identity point")        sig_proof.rs:115-119
ERR_ZERO_CHALLENGE = 15         # InvalidInputs("y is the zero")                      sig_proof.rs:125-127


def signcrypt_valid(impl: int, scheme: int, u_bytes: bytes, v: bytes, w_bytes: bytes) -> Tuple[int, bool]:
    """SignCryptCiphertext::is_valid (sign_crypt_ciphertext.rs:86-101) -> BlsSignCrypt::valid (sign_crypt.rs:69-77):
    W' = hash_to_point(U.to_bytes() || V, DST(scheme)) (compute_w, :151-158); valid iff
    pairing([(W, -g), (W', U)]) is the identity and neither U nor W is.  U lives in the public-key group, W in the signature
    group.  Returns (parse status, valid)."""
    C = IMPLS[impl]
    u, st = _decode(C.pk_deser, u_bytes, MODERN)
    if st != OK:
        return st, False
    w, st = _decode(C.sig_deser, w_bytes, MODERN)
    if st != OK:
        return st, False
    if u is None or w is None:
        return OK, False
    w_tick = C.hash(C.pk_ser(u, MODERN) + bytes(v), sig_dst(impl, scheme))
    return OK, pairing_product_is_one([C.pair(w, C.pk_neg(C.pk_gen)), C.pair(w_tick, u)])


def signcrypt_verify_share(impl: int, scheme: int, share_bytes: bytes, pk_bytes: bytes, u_bytes: bytes, v: bytes,
                           w_bytes: bytes) -> Tuple[int, bool]:
    """BlsSignCrypt::verify_share (sign_crypt.rs:192-207) behind SignDecryptionShare::verify (sign_decryption_share.rs:45-61,
    which always passes the Basic DST): hash = -compute_w(u, v, dst); ok iff share, pk, w are not the identity and
    pairing([(hash, share), (w, pk)]) is the identity.  share, pk, u: public-key group; w: signature group."""
    C = IMPLS[impl]
    pts = []
    for fn, b in ((C.pk_deser, share_bytes), (C.pk_deser, pk_bytes), (C.pk_deser, u_bytes), (C.sig_deser, w_bytes)):
        p, st = _decode(fn, b, MODERN)
        if st != OK:
            return st, False
        pts.append(p)
    share, pk, u, w = pts
    if share is None or pk is None or w is None:
        return OK, False
    h = C.hash(C.pk_ser(u, MODERN) + bytes(v), sig_dst(impl, scheme))
    sig_neg = g2_neg if impl == G2IMPL else g1_neg
    return OK, pairing_product_is_one([C.pair(sig_neg(h), share), C.pair(w, pk)])


def pok_verify(impl: int, scheme: int, commitment_bytes: bytes, proof_bytes: bytes, pk_bytes: bytes, y_be32: bytes,
               msg: bytes) -> int:
    """ProofOfKnowledge::verify (proof_of_knowledge.rs:132-165) -> BlsSignatureProof::verify (sig_proof.rs:102-142):
    reject identity commitment / proof / pk and y = 0 (in that order), a = hash_to_point(msg, DST(scheme)), accept iff
    pairing([(proof, g), (commitment + a*y, pk)]) is the identity.  commitment, proof: signature group; y: a scalar < r."""
    C = IMPLS[impl]
    cm, st = _decode(C.sig_deser, commitment_bytes, MODERN)
    if st != OK:
        return st
    pr, st = _decode(C.sig_deser, proof_bytes, MODERN)
    if st != OK:
        return st
    pk, st = _decode(C.pk_deser, pk_bytes, MODERN)
    if st != OK:
        return st
    y = int.from_bytes(y_be32, "big")
    if len(y_be32) != 32 or y >= R:
        return ERR_DESERIALIZE
    if cm is None:
        return ERR_COMMITMENT_IDENTITY
    if pr is None:
        return ERR_PROOF_IDENTITY
    if pk is None:
        return ERR_PK_IDENTITY
    if y == 0:
        return ERR_ZERO_CHALLENGE
    a = C.hash(bytes(msg), sig_dst(impl, scheme))
    t = C.sig_add(cm, C.sig_mul(a, y))
    ok = pairing_product_is_one([C.pair(pr, C.pk_gen), C.pair(t, pk)])
    return OK if ok else ERR_INVALID_PROOF
