"""Where the host-buffer call spends its time beyond the device stages: wall clock vs stage sum, host vs device inputs."""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np, torch
import blsful_b200 as B, bench
n = int(sys.argv[1]) if len(sys.argv) > 1 else 1000000
eng = B.Engine([0])
if os.environ.get("TORCH_STREAM"):
    stream = torch.cuda.Stream(device=torch.device("cuda", 0))
    eng.set_stream(stream.cuda_stream)
pks, sigs, msgs, off = bench.synth_batch(eng, n, seed=7)
pin = lambda a: torch.from_numpy(a).pin_memory()
h = [pin(pks), pin(sigs), pin(msgs), pin(off.view(np.int64))]
npk, nsg, nms, nof = h[0].numpy(), h[1].numpy(), h[2].numpy(), h[3].numpy().view(np.uint64)
dev = torch.device("cuda", 0)
d = [t.to(dev) for t in h]
st = torch.empty(n, dtype=torch.uint8, device=dev)
for name, fn in (("dev ", lambda: eng.verify_batch_dev(2, 0, n, d[0].data_ptr(), d[1].data_ptr(), d[2].data_ptr(), d[3].data_ptr(), st.data_ptr())),
                 ("host pinned", lambda: eng.verify_batch_packed(2, 0, npk, nsg, nms, nof)),
                 ("host pageable", lambda: eng.verify_batch_packed(2, 0, pks, sigs, msgs, off))):
    fn(); torch.cuda.synchronize()
    best = 1e9
    for _ in range(3):
        t = time.perf_counter(); fn(); torch.cuda.synchronize(); dt = (time.perf_counter() - t) * 1e3
        s = eng.last_stage_ms()
        best = min(best, dt)
        print(f"{name}: wall {dt:7.1f} ms  stages {sum(s.values()):7.1f} ms  {' '.join(f'{k}={v:.1f}' for k, v in s.items())}", flush=True)
