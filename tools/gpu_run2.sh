set -x
mkdir -p gpurun_out
python __graft_entry__.py smoke > gpurun_out/r2b_smoke.log 2>&1; echo "smoke rc=$?"
python -m pytest tests -m gpu -x -q --durations=12 > gpurun_out/r2b_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2b_pytest.log
tail -3 gpurun_out/r2b_pytest.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2b_bench.json 2> gpurun_out/r2b_bench.err; echo "bench rc=$?"
tail -5 gpurun_out/r2b_bench.err
