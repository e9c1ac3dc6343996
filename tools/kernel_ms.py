"""Per-kernel device times of one 1M (or argv[1]) G2Impl verify for the library named by BLSGPU_LIB (experiments)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np, torch, blsful_b200 as B, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
eng = B.Engine([0]); dev = torch.device('cuda', 0)
cache = f"/tmp/exp_data_{n}_2.npz"
if os.path.exists(cache):
    z = np.load(cache); pks, sigs, msgs, off = z["pks"], z["sigs"], z["msgs"], z["off"]
else:
    pks, sigs, msgs, off = bench.synth_batch(eng, n, seed=7); np.savez(cache, pks=pks, sigs=sigs, msgs=msgs, off=off)
d = [torch.from_numpy(a).to(dev) for a in (pks, sigs, msgs, off.view(np.int64))]
st = torch.empty(n, dtype=torch.uint8, device=dev)
best = None
for _ in range(3):
    eng.verify_batch_dev(2, 0, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), st.data_ptr())
    s = eng.last_stage_ms(); tot = sum(s.values())
    if best is None or tot < best[0]: best = (tot, s, eng.last_kernel_ms())
assert int(st.max().item()) == 0
print(os.path.basename(os.environ.get("BLSGPU_LIB", "default")), f"total={best[0]:.1f}", {k: round(v, 1) for k, v in best[1].items()}, {k: round(v[0], 1) for k, v in best[2].items()}, flush=True)
