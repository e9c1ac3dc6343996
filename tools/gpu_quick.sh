mkdir -p gpurun_out
T=${TAG:-r2m}
python -m pytest tests/test_gpu_parity.py -m gpu -x -q -k "golden or bisection or large_batch or aggregate_verify or pairing or hash or recode or sum_points or testdata or secure or shares" > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
python tools/exp_stages.py 1000000 > gpurun_out/${T}_stages.log 2>&1
python - >> gpurun_out/${T}_stages.log 2>&1 <<'PY'
import sys; sys.path.insert(0,'.'); sys.path.insert(0,'agora-blsful_b200')
import numpy as np, torch, blsful_b200 as B
z=np.load('/tmp/exp_data_1000000_2.npz'); eng=B.Engine([0]); dev=torch.device('cuda',0)
d=[torch.from_numpy(z[k]).to(dev) for k in ('pks','sigs','msgs')]; off=torch.from_numpy(z['off'].view(np.int64)).to(dev); st=torch.empty(1000000,dtype=torch.uint8,device=dev)
for _ in range(2): eng.verify_batch_dev(2,0,1000000,d[0].data_ptr(),d[1].data_ptr(),d[2].data_ptr(),off.data_ptr(),st.data_ptr())
print({k:(round(v[0],2),v[1]) for k,v in eng.last_kernel_ms().items()})
PY
