"""CPU-only checks of the drop-in boundary: the C-ABI library loads and exports every symbol include/blsgpu.h
declares, and refuses to run without a CUDA device (no CPU fallback)."""
import os
import re

import pytest


@pytest.fixture(scope="module", autouse=True)
def built_library():
    """nvcc cross-compiles without a GPU: (re)build the library when it is missing or older than its sources."""
    import __graft_entry__ as g
    g.build_engine()


def test_library_exports_every_declared_symbol():
    import blsful_b200 as B
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    header = open(os.path.join(root, "include", "blsgpu.h")).read()
    declared = set(re.findall(r"\b(blsgpu_[a-z0-9_]+)\s*\(", header))
    assert declared == set(B.EXPORTED_SYMBOLS)
    lib = B.load_library()
    for name in declared:
        assert hasattr(lib, name), name


def test_no_cpu_fallback():
    import torch
    import blsful_b200 as B
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    with pytest.raises(B.EngineError):
        B.Engine([0])


def test_length_rules_are_enforced_host_side():
    import blsful_b200 as B
    with pytest.raises(B.BlsError) as e:
        B._pack_points([bytes(47)], 48, "public key")
    assert e.value.status == B.ST_INVALID_LENGTH


def test_msm_window_plan_never_has_a_narrow_top_window():
    """blsgpu_plan_msm (host-only): a narrow top window concentrates n / 2^bits signatures in each of its few buckets -
    one thread each.  (n = 500,000 once took 1.5 s with a 4-bit top window.)"""
    import ctypes
    import blsful_b200 as B
    lib = B.load_library()
    for bits, n in [(b, n) for b in (64, 128) for n in [4096, 4097, 5000, 33000, 65536, 100000, 262144, 500000, 524288, 999999,
                                                           1000000, 1048576, 4000000, 10 ** 8]]:
        c, w, top = ctypes.c_int(), ctypes.c_int(), ctypes.c_int()
        assert lib.blsgpu_plan_msm(n, bits, ctypes.byref(c), ctypes.byref(w), ctypes.byref(top)) == 0
        c, w, top = c.value, w.value, top.value
        assert (w - 1) * c + top == bits and 0 < top <= c
        assert n >> c >= 4 or c == 4          # enough signatures per bucket to amortise the bucket reduction
        assert (n >> top) <= 16 * max(1, n >> c) or n < (1 << 16), (n, c, top)   # top-window buckets at most 16x larger


def test_rust_sys_crate_declares_the_whole_header():
    """rust/blsful-gpu-sys/src/lib.rs is source only (no cargo here): at least keep it in step with include/blsgpu.h -
    the same functions with the same number of arguments, and the same status constants."""
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    hdr = open(os.path.join(root, "include", "blsgpu.h")).read()
    h = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    h = re.sub(r"//[^\n]*", "", h)
    c_fns = {m.group(1): len([a for a in m.group(2).split(",") if a.strip() and a.strip() != "void"])
             for m in re.finditer(r"\b(blsgpu_[a-z0-9_]+)\s*\(([^;{]*?)\)\s*;", h, flags=re.S)}
    rs = open(os.path.join(root, "rust", "blsful-gpu-sys", "src", "lib.rs")).read()
    r = re.sub(r"/\*.*?\*/", "", rs, flags=re.S)
    r = re.sub(r"//[^\n]*", "", r)
    rs_fns = {m.group(1): len([a for a in m.group(2).split(",") if a.strip()])
              for m in re.finditer(r"pub fn (blsgpu_[a-z0-9_]+)\(([^)]*)\)", r, flags=re.S)}
    assert c_fns and set(c_fns) == set(rs_fns), (set(c_fns) ^ set(rs_fns))
    assert c_fns == rs_fns, {k: (c_fns[k], rs_fns[k]) for k in c_fns if c_fns[k] != rs_fns[k]}
    c_consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"#define (BLSGPU_(?:ST|E)_[A-Z_]+|BLSGPU_OK|BLSGPU_STAGE_COUNT|BLSGPU_KERNEL_COUNT) \(?(-?\d+)\)?", hdr)}
    rs_consts = {m.group(1): int(m.group(2)) for m in re.finditer(r"pub const (BLSGPU_[A-Z_]+): \w+ = (-?\d+);", rs)}
    assert c_consts == rs_consts, set(c_consts.items()) ^ set(rs_consts.items())


def test_shard_plan_of_a_multi_device_context():
    """blsgpu_plan_shards (host-only): contiguous runs of key sets per device, balanced by member count, covering every set
    exactly once, for ragged, empty and very uneven inputs (what blsgpu_verify_secure_batch does on a context with several
    devices)."""
    import numpy as np
    import blsful_b200 as B
    lib = B.load_library()
    rng = np.random.default_rng(11)
    cases = [[400] * 10000, [1], [], [0, 0, 0], [5, 0, 0, 0, 1000000, 3], [1] * 7, list(rng.integers(0, 900, size=997))]
    for sizes in cases:
        off = np.concatenate([[0], np.cumsum(sizes)]).astype(np.uint64) + np.uint64(17)   # offsets need not start at zero
        for ndev in (1, 2, 3, 8):
            cut = np.zeros(ndev + 1, dtype=np.uint64)
            assert lib.blsgpu_plan_shards(len(sizes), off.ctypes.data, ndev, cut.ctypes.data) == 0
            cut = cut.tolist()
            assert cut[0] == 0 and cut[-1] == len(sizes) and all(a <= b for a, b in zip(cut, cut[1:]))
            total = int(off[-1] - off[0])
            loads = [int(off[cut[d + 1]] - off[cut[d]]) for d in range(ndev)]
            assert sum(loads) == total
            if sizes:   # no run exceeds its fair share by more than one set
                assert max(loads) <= total / ndev + max(sizes) + 1e-9, (sizes[:8], ndev, loads)
    bad = np.array([0, 5, 3], dtype=np.uint64)
    cut = np.zeros(3, dtype=np.uint64)
    assert lib.blsgpu_plan_shards(2, bad.ctypes.data, 2, cut.ctypes.data) == -1   # BLSGPU_E_ARG: offsets must not decrease
