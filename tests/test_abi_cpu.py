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
