"""One process, one context on G devices: end-to-end time of blsgpu_verify_batch on host buffers (the in-context sharding
of include/blsgpu.h blsgpu_ctx_create).  usage: python tools/multi_dev_e2e.py G [n_per_device]"""
import json, os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "agora-blsful_b200"))
import numpy as np
import blsful_b200 as B
import bench

G = int(sys.argv[1]) if len(sys.argv) > 1 else 2
per = int(sys.argv[2]) if len(sys.argv) > 2 else 1000000
one = B.Engine([0])
pks, sigs, msgs, off = bench.synth_batch(one, per, seed=7, impl=2)
n = per * G
pks, sigs, msgs = np.tile(pks, G), np.tile(sigs, G), np.tile(msgs, G)
off = np.concatenate([off[:-1].astype(np.uint64) + np.uint64(d * int(off[-1])) for d in range(G)] + [np.array([G * int(off[-1])], dtype=np.uint64)])
res = {}
for name, cnt in (("1 device", per), (f"{G} devices", n)):
    eng = one if cnt == per else B.Engine(list(range(G)))   # one arena per device at a time: 1M items take ~80 GB
    best = None
    for it in range(3):
        t0 = time.perf_counter()
        st = eng.verify_batch_packed(2, 0, pks[:48 * cnt], sigs[:96 * cnt], msgs[:int(off[cnt])], off[:cnt + 1])
        dt = time.perf_counter() - t0
        assert int(st.max()) == 0
        best = dt if best is None else min(best, dt)
    res[name] = {"n": cnt, "seconds": round(best, 4), "sigs_per_s": round(cnt / best)}
    eng.close()
print(json.dumps(res))
