"""CPU timing legs for bench.py (test infrastructure): the reference's per-signature verification path
(core_verify, reference src/traits/sig_core.rs:120-146: hash_to_curve, two Miller loops, one final exponentiation per
signature, no batching) restated by the oracle and run on the host cores.

Uses the C++ restatement (oracle/c -> oracle/_build/liboracle.so: the reference's call sequence over the engine's
field / curve / pairing headers compiled for the host, validated against the big-int oracle by tests/test_c_oracle.py)
when it has been built, otherwise the big-int Python oracle.  blsful itself cannot be built here (no cargo, un-vendored
blstrs_plus/blst), so kind is always "port"."""
import json
import multiprocessing as mp
import os
import time

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)


def _golden_triples():
    g = json.load(open(os.path.join(ROOT, "tests", "golden", "cpp_integration.json")))
    msg = bytes.fromhex(g["message"])
    return [(bytes.fromhex(s["pk"]), bytes.fromhex(s["sig"]), msg) for s in g["signers"]]


def _c_oracle():
    try:
        from oracle import c_oracle
        return c_oracle if c_oracle.available() else None
    except Exception:
        return None


def _verify_chunk(args):
    count, impl = args
    tr = _golden_triples()
    co = _c_oracle()
    ok = 0
    for i in range(count):
        pk, sig, msg = tr[i % len(tr)]
        if co is not None:
            st = co.verify(impl, 0, 1, pk, sig, msg)
        else:
            from oracle import bls_oracle as O
            st = O.verify(impl, O.BASIC, O.MODERN, pk, sig, msg)
        ok += st == 0
    return ok


def _run(total, impl, procs):
    per = [total // procs + (1 if i < total % procs else 0) for i in range(procs)]
    per = [p for p in per if p]
    t0 = time.perf_counter()
    if len(per) == 1:
        ok = _verify_chunk((per[0], impl))
    else:
        with mp.get_context("fork").Pool(len(per)) as pool:
            ok = sum(pool.map(_verify_chunk, [(p, impl) for p in per]))
    dt = time.perf_counter() - t0
    assert ok == total, "reference vectors must verify"
    return dt


def _auto_sample(procs):
    # ~10-30 s of CPU work: C++ restatement ~11 ms/verify, Python big-int ~1 s/verify
    return (1000 if _c_oracle() is not None else 4) * procs


def time_verify_sample(sample=0, impl=2):
    procs = os.cpu_count() or 1
    kind = ("C++ restatement (oracle/c: per-signature core_verify over the engine's host-compiled field/curve headers, "
            "g++ -O3, no asm)") if _c_oracle() is not None else "Python big-int oracle"
    n1 = max(1, (sample or _auto_sample(1)) // (1 if sample else 4))
    dt1 = _run(n1, impl, 1)
    nall = sample or _auto_sample(procs)
    dtall = _run(nall, impl, procs)
    return {"value": nall / dtall, "unit": "sigs/s", "cores": procs, "kind": "port",
            "sample": f"{nall} Signature::verify calls over the reference's 3 golden triples on {procs} processes; {kind}",
            "single_core_sigs_per_s": n1 / dt1}


def reference_arm(n_per_step=0, steps=3, warmup=1):
    procs = os.cpu_count() or 1
    n = n_per_step or max(procs, _auto_sample(procs) // max(1, steps + warmup))
    for _ in range(warmup):
        _run(n, 2, procs)
    t0 = time.perf_counter()
    for _ in range(steps):
        _run(n, 2, procs)
    dt = time.perf_counter() - t0
    kind = "C++ restatement (oracle/c, g++ -O3, no asm)" if _c_oracle() is not None else "Python big-int oracle"
    return {"value": n * steps / dt, "ms_per_step": dt / steps * 1e3, "n_per_step": n, "cores": procs, "kind": "port",
            "sample": f"{n} per-signature verifies per step on {procs} processes; {kind}; blsful+blst cannot be built here"}
