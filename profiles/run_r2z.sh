# dynamic tile claims in the persistent pull (global counter, two tiles ahead) vs the static round-robin
set -x
mkdir -p gpurun_out/r2z
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py -x -q > gpurun_out/r2z/pytest_pipe.log 2>&1; tail -3 gpurun_out/r2z/pytest_pipe.log
timeout 120 python profiles/time_ops.py --ops pull,grad,bwd_grid > gpurun_out/r2z/time_ops_dyn.txt 2>&1
IB200_STATIC_TILES=1 timeout 120 python profiles/time_ops.py --ops pull,grad,bwd_grid > gpurun_out/r2z/time_ops_static.txt 2>&1
IB200_NCW=22 timeout 120 python profiles/time_ops.py --ops pull > gpurun_out/r2z/time_ops_dyn_ncw22.txt 2>&1
IB200_NCW=18 timeout 120 python profiles/time_ops.py --ops pull > gpurun_out/r2z/time_ops_dyn_ncw18.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull,grad --order 1 > gpurun_out/r2z/time_ops_dyn_o1.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull,grad --order 2 > gpurun_out/r2z/time_ops_dyn_o2.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull --bound 0 > gpurun_out/r2z/time_ops_dyn_zero.txt 2>&1
grep -H Mvox gpurun_out/r2z/time_ops_*.txt
