# channel pairs in the tile pull / grad kernels: parity tests + cfg 3 timings
set -x
mkdir -p gpurun_out/r2q
timeout 1200 python -m pytest tests/test_gpu_tile_parity.py tests/test_gpu_ops.py tests/test_gpu_pipe.py -x -q > gpurun_out/r2q/pytest.log 2>&1; tail -4 gpurun_out/r2q/pytest.log
timeout 120 python profiles/time_ops.py --channels 4 --ops pull,grad > gpurun_out/r2q/time_ops_256_o3_c4.txt 2>&1
timeout 120 python profiles/time_ops.py --channels 2 --ops pull,grad > gpurun_out/r2q/time_ops_256_o3_c2.txt 2>&1
timeout 120 python profiles/time_ops.py --channels 3 --ops pull,grad --order 1 > gpurun_out/r2q/time_ops_256_o1_c3.txt 2>&1
grep -h Mvox gpurun_out/r2q/time_ops_*.txt
