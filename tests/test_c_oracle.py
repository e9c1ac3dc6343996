"""The C++ timing restatement used by bench.py's CPU legs (oracle/c) against the big-int oracle: every status it returns
must be the oracle's.  CPU only.  (It shares the engine's headers: this validates the timed baseline, not the engine.)"""
import json
import os
import random
import subprocess

import numpy as np
import pytest

from oracle import bls_oracle as O

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def C():
    subprocess.run(["make", "-s", "-C", os.path.join(ROOT, "oracle", "c")], check=True)
    from oracle import c_oracle
    assert c_oracle.available()
    return c_oracle


def test_c_restatement_matches_the_oracle(C, golden_dir):
    g = json.load(open(os.path.join(golden_dir, "cpp_integration.json")))
    msg = bytes.fromhex(g["message"])
    pks = [bytes.fromhex(s["pk"]) for s in g["signers"]]
    sigs = [bytes.fromhex(s["sig"]) for s in g["signers"]]
    ident1, ident2 = bytes([0xC0]) + bytes(47), bytes([0xC0]) + bytes(95)
    cases = [(0, pks[i], sigs[i], msg) for i in range(3)]
    cases += [(0, pks[0], sigs[1], msg), (0, pks[0], sigs[0], msg + b"!"), (0, pks[0], ident2, msg), (0, ident1, sigs[0], msg),
              (0, bytes(48), sigs[0], msg), (0, pks[0], bytes(96), msg), (1, pks[0], sigs[0], msg)]
    rnd = random.Random(3)
    for scheme in (0, 1, 2):
        sk = rnd.randrange(1, O.R)
        m = b"scheme %d" % scheme
        cases.append((scheme, O.g1_serialize(O.sk_to_pk(2, sk)), O.g2_serialize(O.sign(2, scheme, sk, m)), m))
    for scheme, pk, sig, m in cases:
        assert C.verify(2, scheme, 1, pk, sig, m) == O.verify(2, scheme, O.MODERN, pk, sig, m)
    # the threaded batch form used for timing
    pk_a = np.frombuffer(b"".join(c[1] for c in cases[:7]), dtype=np.uint8)
    sg_a = np.frombuffer(b"".join(c[2] for c in cases[:7]), dtype=np.uint8)
    ms = [c[3] for c in cases[:7]]
    off = np.zeros(8, dtype=np.uint64)
    off[1:] = np.cumsum([len(x) for x in ms])
    st = C.verify_many(0, pk_a, sg_a, np.frombuffer(b"".join(ms), dtype=np.uint8), off, 3)
    assert st.tolist() == [O.verify(2, 0, O.MODERN, c[1], c[2], c[3]) for c in cases[:7]]
