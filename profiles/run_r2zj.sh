# coordinate prefetch for the next row; 16 consumer warps (4 rows each per tile)
set -x
mkdir -p gpurun_out/r2zj
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2zj/pytest_pipe.log 2>&1; tail -3 gpurun_out/r2zj/pytest_pipe.log
timeout 120 python profiles/time_pull_b2b.py > gpurun_out/r2zj/b2b_prefetch.txt 2>&1
for v in rotated ncw16; do
  IB200_LIB=$PWD/profiles/lab_so/lib_$v.so timeout 120 python profiles/time_pull_b2b.py > gpurun_out/r2zj/b2b_$v.txt 2>&1
done
IB200_LIB=$PWD/profiles/lab_so/lib_ncw16.so timeout 600 python -m pytest tests/test_gpu_pipe.py -x -q > gpurun_out/r2zj/pytest_pipe_ncw16.log 2>&1; tail -2 gpurun_out/r2zj/pytest_pipe_ncw16.log
cat gpurun_out/r2zj/b2b_*.txt
