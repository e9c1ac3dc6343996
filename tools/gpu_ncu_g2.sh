mkdir -p gpurun_out
T=${TAG:-r2n}
python tools/exp_stages.py 245760 > gpurun_out/${T}_plain.log 2>&1 || exit 1
BLSGPU_SKIP_SELFTEST=1 ncu --set full --clock-control none --import-source on -k "regex:${KREGEX:-k_clear_cofactor|k_subgroup_check}" -c ${KCOUNT:-3} -f -o gpurun_out/${T}_g2 python tools/exp_stages.py 245760 > gpurun_out/${T}_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/${T}_ncu.log
