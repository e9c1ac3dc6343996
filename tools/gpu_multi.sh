# Multi-GPU evidence on one box with N GPUs:  N=2 TAG=r2 bash tools/gpu_multi.sh
mkdir -p gpurun_out
T=${TAG:-r2}
N=${N:-2}
python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29517 bench.py --gpus $N --steps 3 --warmup 3 \
    > gpurun_out/${T}_bench_${N}gpu.json 2> gpurun_out/${T}_bench_${N}gpu.err; echo "rc=$?" >> gpurun_out/${T}_bench_${N}gpu.err
if [ "$N" = "2" ]; then
  python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q -x -k "two_devices or fold" > gpurun_out/${T}_pytest_2gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest_2gpu.log
fi
python tools/multi_dev_e2e.py $N > gpurun_out/${T}_multi_device_ctx_${N}gpu.json 2> gpurun_out/${T}_multi_device_ctx_${N}gpu.err
