"""One failing batch (argv[2] corrupted signatures among argv[1]) through blsgpu_verify_batch: for ncu launch lists of the failure path."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np, blsful_b200 as B, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 245760
nbad = int(sys.argv[2]) if len(sys.argv) > 2 else 1
eng = B.Engine([0])
pks, sigs, msgs, off = bench.synth_batch(eng, n, seed=5)
rng = np.random.default_rng(1)
bad = np.sort(rng.choice(n, nbad, replace=False))
s2 = sigs.copy().reshape(n, 96); s2[bad] = s2[(bad + 1) % n]
st = eng.verify_batch_packed(2, 0, pks, s2.reshape(-1), msgs, off)
assert np.array_equal(np.nonzero(st)[0], bad)
print("ok", eng.last_stage_ms())
