# usage: LIBS="build/a.so build/b.so" TAG=r2x bash tools/gpu_variants.sh   -> stage + kernel times of each build at 1M
mkdir -p gpurun_out
T=${TAG:-r2v}
: > gpurun_out/${T}_variants.log
python tools/exp_stages.py 1000000 > /dev/null 2>&1   # creates the cached batch
for L in $LIBS; do
BLSGPU_LIB=$PWD/$L python - >> gpurun_out/${T}_variants.log 2>&1 <<'PY'
import os, sys; sys.path.insert(0,'.'); sys.path.insert(0,'agora-blsful_b200')
import numpy as np, torch, blsful_b200 as B
z=np.load('/tmp/exp_data_1000000_2.npz'); eng=B.Engine([0]); dev=torch.device('cuda',0)
d=[torch.from_numpy(z[k]).to(dev) for k in ('pks','sigs','msgs')]; off=torch.from_numpy(z['off'].view(np.int64)).to(dev); st=torch.empty(1000000,dtype=torch.uint8,device=dev)
best=None
for _ in range(3):
    eng.verify_batch_dev(2,0,1000000,d[0].data_ptr(),d[1].data_ptr(),d[2].data_ptr(),off.data_ptr(),st.data_ptr())
    s=eng.last_stage_ms(); tot=sum(s.values())
    if best is None or tot<best[0]: best=(tot,s,eng.last_kernel_ms())
assert int(st.max().item())==0
print(os.environ['BLSGPU_LIB'].split('/')[-1], 'total=%.1f'%best[0], {k:round(v,1) for k,v in best[1].items()})
print('   ', {k:round(v[0],2) for k,v in best[2].items()})
PY
done
