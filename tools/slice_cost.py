"""Fixed costs of a slice: stage times and wall clock of one rank's share of a 1M batch cut over 1/2/4/8/16 ranks
(verify_batch_folded with world = 1: partial results, fold, finish - everything but the exchange)."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np, torch
import blsful_b200 as B, bench
eng = B.Engine([0])
N = 1000000
pks, sigs, msgs, off = bench.synth_batch(eng, N, seed=7)
pin = lambda a: torch.from_numpy(a).pin_memory().numpy()
for world in (1, 2, 4, 8, 16):
    n = N // world
    a = [pin(pks[:48 * n]), pin(sigs[:96 * n]), pin(msgs[:32 * n]), pin(off[:n + 1].view(np.int64)).view(np.uint64)]
    for name, fn in (("one call ", lambda: eng.verify_batch_packed(2, 0, *a)),
                     ("fold path", lambda: B.verify_batch_folded(eng, 2, 0, a[0], a[1], a[2], a[3], 0, 1, lambda p: [p]))):
        fn()
        best = 1e9
        for _ in range(3):
            t = time.perf_counter(); fn(); best = min(best, (time.perf_counter() - t) * 1e3)
        s = eng.last_stage_ms()
        print(f"n={n:8d} {name}: wall {best:7.1f} ms  ideal {807.0 / world:6.1f}  stages {sum(s.values()):7.1f} | " + " ".join(f"{k}={v:.1f}" for k, v in s.items()), flush=True)
