# channel-interleaved boxes (C % 4 == 0): parity + cfg 3 timings
set -x
mkdir -p gpurun_out/r2s
timeout 1200 python -m pytest tests/test_gpu_tile_parity.py -x -q -k "interleaved or cfg3" > gpurun_out/r2s/pytest.log 2>&1; tail -6 gpurun_out/r2s/pytest.log
timeout 120 python profiles/time_ops.py --channels 4 --ops pull,grad > gpurun_out/r2s/time_ops_256_o3_c4.txt 2>&1
timeout 120 python profiles/time_ops.py --channels 4 --ops pull,grad --order 1 > gpurun_out/r2s/time_ops_256_o1_c4.txt 2>&1
timeout 120 python profiles/time_ops.py --channels 8 --ops pull,grad --size 192 > gpurun_out/r2s/time_ops_192_o3_c8.txt 2>&1
grep -h Mvox gpurun_out/r2s/time_ops_*.txt
