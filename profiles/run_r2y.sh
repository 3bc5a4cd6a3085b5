# scout warp (look-ahead box off the consumers' path) + consumer-warp count sweep for the persistent pull
set -x
mkdir -p gpurun_out/r2y
timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q > gpurun_out/r2y/pytest_pipe.log 2>&1; tail -3 gpurun_out/r2y/pytest_pipe.log
for ncw in 14 18 22; do
  IB200_NCW=$ncw timeout 120 python profiles/time_ops.py --ops pull,grad,bwd_grid > gpurun_out/r2y/time_ops_ncw$ncw.txt 2>&1
  grep -h Mvox gpurun_out/r2y/time_ops_ncw$ncw.txt
done
IB200_NCW=22 timeout 120 python profiles/time_ops.py --ops pull --order 3 --bound 0 > gpurun_out/r2y/time_ops_ncw22_zero.txt 2>&1; grep -h Mvox gpurun_out/r2y/time_ops_ncw22_zero.txt
