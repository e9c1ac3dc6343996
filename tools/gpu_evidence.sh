# Round evidence on one B200: full GPU test suite, smoke(), both bench arms, the ncu launch list of a bench run and one
# `ncu --set full` capture of every hot kernel.  Everything lands in gpurun_out/<TAG>_*; tools/collect_evidence.py reduces it
# into profiles/.   usage: TAG=r2 bash tools/gpu_evidence.sh [tests|bench|ncu ...]
mkdir -p gpurun_out
T=${TAG:-r2}
WHAT="${@:-tests bench ncu}"
for w in $WHAT; do
case $w in
tests)
  python -m pytest tests -m gpu -q -x > gpurun_out/${T}_pytest.log 2>&1; echo "pytest rc=$?" >> gpurun_out/${T}_pytest.log
  python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/${T}_smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/${T}_smoke.log
  ;;
bench)
  python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/${T}_bench_reference.json 2> gpurun_out/${T}_bench_reference.err
  python bench.py > gpurun_out/${T}_bench.json 2> gpurun_out/${T}_bench.err; echo "bench rc=$?" >> gpurun_out/${T}_bench.err
  ;;
ncu)
  # (1) launch list of the same command, default batch (per-launch times are cold-cache and serialised: shares only)
  ncu --metrics gpu__time_duration.sum --clock-control none -c 600 --csv --log-file gpurun_out/${T}_launches.csv \
      python bench.py --steps 1 --warmup 1 --no-configs --no-cpu-baseline > gpurun_out/${T}_launches_run.log 2>&1
  # (2) one full capture of each hot kernel at n = 245,760 (the self-test launches are skipped)
  ITERS=1 BLSGPU_SKIP_SELFTEST=1 ncu --set full --clock-control none --import-source on \
      -k "regex:k_decode|k_subgroup_check|k_hash|k_clear_cofactor|k_to_affine_batch|k_m6_prep|k_m6_lines|k_m6_accum|k_msm_bucket|k_msm_chunk" \
      -c 24 -f -o gpurun_out/${T}_hot python tools/exp_stages.py 245760 > gpurun_out/${T}_ncu.log 2>&1
  echo "ncu rc=$?" >> gpurun_out/${T}_ncu.log
  # reduce on the box: the report with source pages is larger than what gpurun brings back
  PROFILES_OUT=gpurun_out/${T}_profiles python tools/collect_evidence.py ${T} > gpurun_out/${T}_collect.log 2>&1
  rm -f gpurun_out/${T}_src_*.csv
  if [ $(stat -c %s gpurun_out/${T}_hot.ncu-rep) -gt 45000000 ]; then rm -f gpurun_out/${T}_hot.ncu-rep; fi
  ;;
esac
done
