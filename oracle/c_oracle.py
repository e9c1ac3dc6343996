"""ctypes front of the C restatement used by the timed CPU legs (oracle/c/oracle_verify.cpp -> oracle/_build/liboracle.so).
TEST INFRASTRUCTURE: bench.py's cpu_baseline / --impl reference legs and tests/test_c_oracle.py only."""
import ctypes
import os

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
_PATH = os.path.join(HERE, "_build", "liboracle.so")
_lib = None


def _load():
    global _lib
    if _lib is None and os.path.exists(_PATH):
        lib = ctypes.CDLL(_PATH)
        lib.oracle_verify_g2impl.restype = ctypes.c_int
        lib.oracle_verify_g2impl.argtypes = [ctypes.c_int, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_char_p, ctypes.c_size_t]
        lib.oracle_verify_many_g2impl.restype = None
        lib.oracle_verify_many_g2impl.argtypes = [ctypes.c_int, ctypes.c_size_t, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                                                  ctypes.c_void_p, ctypes.c_void_p, ctypes.c_int]
        _lib = lib
    return _lib


def available() -> bool:
    return _load() is not None


def verify(impl: int, scheme: int, fmt: int, pk: bytes, sig: bytes, msg: bytes) -> int:
    """Signature::verify for Bls12381G2Impl, Modern encoding (the benchmark's configuration)."""
    if impl != 2 or fmt != 1:
        raise NotImplementedError("the C restatement covers Bls12381G2Impl with the Modern encoding")
    return int(_load().oracle_verify_g2impl(scheme, pk, sig, msg, len(msg)))


def verify_many(scheme: int, pks: np.ndarray, sigs: np.ndarray, msgs: np.ndarray, off: np.ndarray, threads: int) -> np.ndarray:
    n = off.size - 1
    st = np.empty(n, dtype=np.uint8)
    pks, sigs, msgs, off = (np.ascontiguousarray(a) for a in (pks, sigs, msgs, off.astype(np.uint64)))
    _load().oracle_verify_many_g2impl(scheme, n, pks.ctypes.data, sigs.ctypes.data, msgs.ctypes.data, off.ctypes.data, st.ctypes.data,
                                      threads)
    return st
