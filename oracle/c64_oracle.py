"""ctypes front of the 64-bit-limb C oracle (oracle/c64/bls64.c -> oracle/_build/libbls64.so).
TEST INFRASTRUCTURE: the fast checker of the GPU parity tests and the timed CPU legs of bench.py; validated against the
big-int oracle by tests/test_c64_oracle.py.  Nothing in the product loads it."""
import ctypes
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(HERE, "_build", "libbls64.so")
_lib = None


def build(force=False):
    src = os.path.join(HERE, "c64")
    subprocess.run(["make", "-s", "-C", src] + (["-B"] if force else []), check=True)


def _load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(_PATH):
        return None
    lib = ctypes.CDLL(_PATH)
    c = ctypes
    vp, sz, i64p = c.c_void_p, c.c_size_t, c.POINTER(c.c_int64)
    sigs = {
        "bls64_verify": (c.c_int, [c.c_int, c.c_int, c.c_int, c.c_char_p, c.c_char_p, c.c_char_p, sz]),
        "bls64_pop_verify": (c.c_int, [c.c_int, c.c_int, c.c_char_p, c.c_char_p]),
        "bls64_verify_many": (None, [c.c_int, c.c_int, c.c_int, c.c_int, sz, vp, vp, vp, vp, vp, c.c_int]),
        "bls64_hash_to_curve": (None, [c.c_int, c.c_char_p, sz, c.c_char_p, sz, c.c_char_p]),
        "bls64_recode": (c.c_int, [c.c_int, c.c_int, c.c_int, c.c_char_p, c.c_char_p]),
        "bls64_sum_points": (c.c_int, [c.c_int, c.c_int, sz, c.c_char_p, c.c_char_p, i64p]),
        "bls64_point_mul": (c.c_int, [c.c_int, c.c_char_p, c.c_char_p, c.c_char_p]),
        "bls64_generator": (None, [c.c_int, c.c_char_p]),
        "bls64_pairing_product_is_one": (c.c_int, [sz, c.c_char_p, c.c_char_p]),
        "bls64_pairing_cubed": (c.c_int, [c.c_char_p, c.c_char_p, c.c_char_p]),
        "bls64_aggregate_verify": (c.c_int, [c.c_int, c.c_int, c.c_int, sz, c.c_char_p, c.c_char_p, vp, c.c_char_p, i64p]),
        "bls64_verify_secure": (c.c_int, [c.c_int, c.c_int, c.c_int, sz, c.c_char_p, c.c_char_p, c.c_char_p, sz]),
        "bls64_aggregate_secure": (c.c_int, [c.c_int, c.c_int, sz, c.c_char_p, c.c_char_p, c.c_char_p]),
        "bls64_selfcheck_point": (c.c_int, [c.c_int, c.c_char_p, c.c_int]),
        "bls64_fp_mul": (c.c_int, [c.c_char_p, c.c_char_p, c.c_char_p]),
    }
    for name, (res, args) in sigs.items():
        fn = getattr(lib, name)
        fn.restype, fn.argtypes = res, args
    _lib = lib
    return lib


def available() -> bool:
    return _load() is not None


def _plen(impl):
    return (48, 96) if impl == 2 else (96, 48)


def verify(impl, scheme, fmt, pk: bytes, sig: bytes, msg: bytes) -> int:
    """Signature::verify from bytes -> status code (same convention as bls_oracle.verify)."""
    pl, sl = _plen(impl)
    if len(pk) != pl or len(sig) != sl:
        return 6
    return int(_load().bls64_verify(impl, scheme, fmt, bytes(pk), bytes(sig), bytes(msg), len(msg)))


def pop_verify(impl, fmt, pk: bytes, proof: bytes) -> int:
    return int(_load().bls64_pop_verify(impl, fmt, bytes(pk), bytes(proof)))


def verify_many(impl, scheme, fmt, pks, sigs, msgs, off, threads=0, pop=False) -> np.ndarray:
    n = off.size - 1
    st = np.empty(n, dtype=np.uint8)
    pks, sigs, msgs = (np.ascontiguousarray(a, dtype=np.uint8) for a in (pks, sigs, msgs))
    off = np.ascontiguousarray(off, dtype=np.uint64)
    _load().bls64_verify_many(impl, scheme, fmt, 1 if pop else 0, n, pks.ctypes.data, sigs.ctypes.data, msgs.ctypes.data if msgs.size else None,
                              off.ctypes.data, st.ctypes.data, threads or (os.cpu_count() or 1))
    return st


def hash_to_curve(group, msg: bytes, dst: bytes) -> bytes:
    out = ctypes.create_string_buffer(48 if group == 1 else 96)
    _load().bls64_hash_to_curve(group, bytes(msg), len(msg), bytes(dst), len(dst), out)
    return out.raw


def recode(group, fmt_in, fmt_out, enc: bytes):
    out = ctypes.create_string_buffer(48 if group == 1 else 96)
    st = _load().bls64_recode(group, fmt_in, fmt_out, bytes(enc), out)
    return int(st), (out.raw if st == 0 else b"")


def sum_points(group, fmt, encs):
    L = 48 if group == 1 else 96
    out = ctypes.create_string_buffer(L)
    bad = ctypes.c_int64(-1)
    data = b"".join(bytes(e) for e in encs)
    st = _load().bls64_sum_points(group, fmt, len(data) // L, data, out, ctypes.byref(bad))
    return int(st), (out.raw if st == 0 else b""), int(bad.value)


def point_mul(group, enc: bytes, k: int) -> bytes:
    out = ctypes.create_string_buffer(48 if group == 1 else 96)
    st = _load().bls64_point_mul(group, bytes(enc), int(k).to_bytes(32, "big"), out)
    if st:
        raise ValueError(f"undecodable point (status {st})")
    return out.raw


def generator(group) -> bytes:
    out = ctypes.create_string_buffer(48 if group == 1 else 96)
    _load().bls64_generator(group, out)
    return out.raw


def pairing_product_is_one(g1s, g2s) -> bool:
    r = _load().bls64_pairing_product_is_one(len(g1s), b"".join(g1s), b"".join(g2s))
    if r < 0:
        raise ValueError(f"undecodable point (status {-r})")
    return bool(r)


def pairing_cubed(g1: bytes, g2: bytes) -> bytes:
    out = ctypes.create_string_buffer(576)
    st = _load().bls64_pairing_cubed(bytes(g1), bytes(g2), out)
    if st:
        raise ValueError(f"undecodable point (status {st})")
    return out.raw


def aggregate_verify(impl, scheme, fmt, pks, msgs, sig):
    off = np.zeros(len(msgs) + 1, dtype=np.uint64)
    if msgs:
        off[1:] = np.cumsum([len(m) for m in msgs])
    idx = (ctypes.c_int64 * 2)(-1, -1)
    st = _load().bls64_aggregate_verify(impl, scheme, fmt, len(pks), b"".join(pks), b"".join(msgs), off.ctypes.data, bytes(sig), idx)
    return int(st), (int(idx[0]), int(idx[1]))


def verify_secure(impl, scheme, fmt, pks, sig: bytes, msg: bytes) -> int:
    return int(_load().bls64_verify_secure(impl, scheme, fmt, len(pks), b"".join(pks), bytes(sig), bytes(msg), len(msg)))


def aggregate_secure(impl, fmt, pks, sigs):
    if len(pks) != len(sigs):
        return 10, b""
    out = ctypes.create_string_buffer(96 if impl == 2 else 48)
    st = _load().bls64_aggregate_secure(impl, fmt, len(pks), b"".join(pks), b"".join(sigs), out)
    return int(st), (out.raw if st == 0 else b"")


def selfcheck_point(group, x_be: bytes, want_largest: bool) -> int:
    return int(_load().bls64_selfcheck_point(group, bytes(x_be), 1 if want_largest else 0))


def fp_mul(a: int, b: int) -> int:
    out = ctypes.create_string_buffer(48)
    if _load().bls64_fp_mul(a.to_bytes(48, "big"), b.to_bytes(48, "big"), out):
        raise ValueError("operand >= p")
    return int.from_bytes(out.raw, "big")
