"""Failure path against the accept path at 1M: wall clock of the host-buffer call with 0 / 1 / 16 / 1,000 corrupted signatures."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np, blsful_b200 as B, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
eng = B.Engine([0])
pks, sigs, msgs, off = bench.synth_batch(eng, n, seed=5)
eng.verify_batch_packed(2, 0, pks, sigs, msgs, off)
rng = np.random.default_rng(1)
t0 = None
for nbad in (0, 1, 16, 1000):
    bad = np.sort(rng.choice(n, nbad, replace=False))
    s2 = sigs.copy().reshape(n, 96)
    s2[bad] = s2[(bad + 1) % n]
    best = 1e9
    for _ in range(2):
        t = time.perf_counter(); st = eng.verify_batch_packed(2, 0, pks, s2.reshape(-1), msgs, off); best = min(best, (time.perf_counter() - t) * 1e3)
    assert np.array_equal(np.nonzero(st)[0], bad)
    t0 = t0 or best
    s = eng.last_stage_ms()
    print(f"{nbad:5d} bad: {best:7.1f} ms  x{best / t0:.3f}  final={s['final']:.1f} bisect={s['bisect']:.1f}", flush=True)
