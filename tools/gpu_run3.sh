set -x
mkdir -p gpurun_out
python tools/debug_fold.py > gpurun_out/r2c_debug.log 2>&1
python -m pytest tests -m gpu -x -q -k "secure or shares or share_batch or fold" > gpurun_out/r2c_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2c_pytest.log
tail -5 gpurun_out/r2c_pytest.log
python - > gpurun_out/r2c_cfg5.log 2>&1 <<'PY'
import sys, os, json
sys.path.insert(0, '.'); sys.path.insert(0, 'agora-blsful_b200')
import numpy as np, bench, blsful_b200 as B
eng = B.Engine([0])
batch = bench.synth_batch(eng, 20000, seed=3)
out = bench.other_configs(eng, B, batch, eng.imad_peak(), 10000)
print(json.dumps({k: v for k, v in out.items() if k.startswith('cfg5') or k.startswith('cfg1')}, indent=1))
PY
tail -30 gpurun_out/r2c_cfg5.log
