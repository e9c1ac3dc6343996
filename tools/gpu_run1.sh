set -x
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q --durations=15 > gpurun_out/r2a_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/r2a_pytest.log
python bench.py --steps 2 --warmup 3 > gpurun_out/r2a_bench.json 2> gpurun_out/r2a_bench.err; echo "bench rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2a_ref.json 2> gpurun_out/r2a_ref.err
for tool in memcheck initcheck racecheck synccheck; do
  timeout 900 compute-sanitizer --tool $tool --log-file gpurun_out/r2a_san_$tool.log python tools/sanitize_all.py 200 > gpurun_out/r2a_san_$tool.out 2>&1; echo "$tool rc=$?" >> gpurun_out/r2a_san_$tool.out
done
nproc > gpurun_out/r2a_nproc.txt; grep -m1 "model name" /proc/cpuinfo >> gpurun_out/r2a_nproc.txt
