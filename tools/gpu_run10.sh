mkdir -p gpurun_out
python -m pytest tests/test_gpu_parity.py tests/test_gpu_scale.py -m gpu -x -q -k "bisection or large_batch or several_miller or cfg2" > gpurun_out/r2j_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2j_pytest.log
python - > gpurun_out/r2j_fail.log 2>&1 <<'PY'
import sys, time; sys.path.insert(0,'.'); sys.path.insert(0,'agora-blsful_b200')
import numpy as np, blsful_b200 as B, bench
eng=B.Engine([0]); n=1000000
pks,sigs,msgs,off=bench.synth_batch(eng,n,seed=5)
eng.verify_batch_packed(2,0,pks,sigs,msgs,off)
t=time.time(); st=eng.verify_batch_packed(2,0,pks,sigs,msgs,off); t0=(time.time()-t)*1e3; print("all valid", round(t0,1))
rng=np.random.default_rng(1)
for nbad in (1,1000,20000):
    bad=np.sort(rng.choice(n,nbad,replace=False)); s2=sigs.copy().reshape(n,96); s2[bad]=s2[(bad+1)%n]
    t=time.time(); st=eng.verify_batch_packed(2,0,pks,s2.reshape(-1),msgs,off); dt=(time.time()-t)*1e3
    assert np.array_equal(np.nonzero(st)[0],bad)
    print(nbad,"bad:",round(dt,1),"ms ratio",round(dt/t0,3), {k:round(v,1) for k,v in eng.last_stage_ms().items() if k in ('final','bisect')})
PY
