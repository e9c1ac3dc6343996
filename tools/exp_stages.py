"""Experiment helper: per-stage device times of one verify_batch_dev pass for a given library build (BLSGPU_LIB)."""
import os, sys, json
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np, torch
import blsful_b200 as B
import bench

n = int(sys.argv[1]) if len(sys.argv) > 1 else 262144
impl = int(sys.argv[2]) if len(sys.argv) > 2 else 2
eng = B.Engine([0])
cache = f"/tmp/exp_data_{n}_{impl}.npz"
if os.path.exists(cache):
    z = np.load(cache); pks, sigs, msgs, off = z["pks"], z["sigs"], z["msgs"], z["off"]
else:
    pks, sigs, msgs, off = bench.synth_batch(eng, n, seed=7, impl=impl)
    np.savez(cache, pks=pks, sigs=sigs, msgs=msgs, off=off)
dev = torch.device("cuda", 0)
d = [torch.from_numpy(a).to(dev) for a in (pks, sigs, msgs, off.view(np.int64))]
st = torch.empty(n, dtype=torch.uint8, device=dev)
best = None
for it in range(int(os.environ.get('ITERS', 3))):
    eng.verify_batch_dev(impl, 0, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), st.data_ptr())
    s = eng.last_stage_ms()
    tot = sum(s.values())
    if best is None or tot < best[0]:
        best = (tot, s)
assert int(st.max().item()) == 0
tot, s = best
print(os.environ.get("BLSGPU_LIB", "default"), f"n={n} total={tot:.1f}ms  {n/tot/1e3:.3f} Msig/s |",
      " ".join(f"{k}={v:.1f}" for k, v in s.items()), flush=True)
