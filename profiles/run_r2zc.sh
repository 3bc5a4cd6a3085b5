# fix-up pass of folded boxes dealt (row, vector) items round-robin over the lanes + vector copy for mirrored rows
set -x
mkdir -p gpurun_out/r2zc
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2zc/pytest_pipe.log 2>&1; tail -3 gpurun_out/r2zc/pytest_pipe.log
for b in 3 0 1 2 4 6; do
timeout 120 python profiles/time_ops.py --ops pull,grad --bound $b > gpurun_out/r2zc/time_ops_o3_bound$b.txt 2>&1
done
grep -H Mvox gpurun_out/r2zc/time_ops_*.txt
