"""CPU timing legs for bench.py (test infrastructure): the reference's per-signature verification path
(`Signature::verify` -> core_verify, reference src/signature.rs:130-138, src/traits/sig_core.rs:120-146: parse both points
with curve + subgroup checks, hash_to_curve, two Miller loops, ONE final exponentiation per signature, nothing batched)
executed by the 64-bit-limb C oracle (oracle/c64/bls64.c -> oracle/_build/libbls64.so: portable C, `unsigned __int128`,
no assembly, its own source - not the engine's headers) on the host cores, one pthread per core.

blsful itself cannot be built here (no cargo; blstrs_plus / blst are un-vendored, Cargo.toml:20-28), so `kind` is always
"port".  Expect blst's hand-written assembly to be a few times faster per core than this portable C."""
import os
import time

import numpy as np

KIND_TEXT = ("C oracle oracle/c64/bls64.c: the reference's per-signature Signature::verify call sequence (decode + subgroup checks, "
             "hash_to_curve, 2 Miller loops, 1 final exponentiation), 6x64-bit limbs, gcc -O3, no asm; blsful+blst cannot be built here")


def _c64():
    from oracle import c64_oracle
    if not c64_oracle.available():
        c64_oracle.build()
    return c64_oracle


def synth_triples(n, seed=2024):
    """n distinct (pk 48 B, 32-byte message, sig 96 B) Bls12381G2Impl/Basic triples made by the C oracle itself
    (pk = [sk]G, sig = [sk]H(msg)); signing is outside every timed region."""
    C = _c64()
    rng = np.random.default_rng(seed)
    dst = b"BLS_SIG_BLS12381G2_XMD:SHA-256_SSWU_RO_NUL_"
    g1 = C.generator(1)
    pks, sigs, msgs = [], [], []
    for i in range(n):
        sk = int.from_bytes(rng.bytes(31), "big") | 1
        m = i.to_bytes(8, "little") + rng.bytes(24)
        pks.append(C.point_mul(1, g1, sk))
        sigs.append(C.point_mul(2, C.hash_to_curve(2, m, dst), sk))
        msgs.append(m)
    off = np.arange(n + 1, dtype=np.uint64) * 32
    return (np.frombuffer(b"".join(pks), dtype=np.uint8), np.frombuffer(b"".join(sigs), dtype=np.uint8),
            np.frombuffer(b"".join(msgs), dtype=np.uint8), off)


def _time_verify(batch, threads):
    C = _c64()
    pks, sigs, msgs, off = batch
    t0 = time.perf_counter()
    st = C.verify_many(2, 0, 1, pks, sigs, msgs, off, threads=threads)
    dt = time.perf_counter() - t0
    assert int(st.max()) == 0, "the sample must verify"
    return dt


def time_verify_sample(batch, threads=0):
    """cpu_baseline leg: `batch` = (pks, sigs, msgs, off) is a bounded sample of the benchmark's own workload."""
    procs = threads or (os.cpu_count() or 1)
    n = batch[3].size - 1
    n1 = max(1, min(n, 64))
    dt1 = _time_verify((batch[0][:48 * n1], batch[1][:96 * n1], batch[2][:int(batch[3][n1])], batch[3][:n1 + 1]), 1)
    dt = _time_verify(batch, procs)
    return {"value": n / dt, "unit": "sigs/s", "cores": procs, "kind": "port",
            "sample": f"the first {n} signatures of the benchmark batch, one Signature::verify each, {procs} threads; {KIND_TEXT}",
            "single_core_sigs_per_s": n1 / dt1}


def reference_arm(n_per_step=0, steps=3, warmup=1):
    """`bench.py --impl reference`: per step, n_per_step distinct triples through the per-signature path on every core."""
    procs = os.cpu_count() or 1
    n = n_per_step or 1024   # configs[0]: Signature::verify over 1,024 distinct (pk, 32-byte msg, sig) triples
    batch = synth_triples(n)
    for _ in range(warmup):
        _time_verify(batch, procs)
    dt = 0.0
    for _ in range(steps):
        dt += _time_verify(batch, procs)
    return {"value": n * steps / dt, "ms_per_step": dt / steps * 1e3, "n_per_step": n, "cores": procs, "kind": "port",
            "sample": f"{n} distinct triples per step, one Signature::verify each, {procs} threads; {KIND_TEXT}"}
