set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=8 > gpurun_out/r2f_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest.log
tail -4 gpurun_out/r2f_pytest.log
BLSGPU_LIB=$PWD/build/libblsgpu_color.so python -m pytest tests/test_gpu_parity.py tests/test_gpu_parity2.py -m gpu -q > gpurun_out/r2f_pytest_stack_colouring_on.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2f_pytest_stack_colouring_on.log
tail -4 gpurun_out/r2f_pytest_stack_colouring_on.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2f_bench.json 2> gpurun_out/r2f_bench.err; echo "bench rc=$?"
