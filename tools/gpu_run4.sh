mkdir -p gpurun_out
python tools/debug_fold.py > gpurun_out/r2d_debug.log 2>&1
