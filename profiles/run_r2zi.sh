# z-rows dealt statically with a per-box rotation vs claimed from a shared counter vs plain static
set -x
mkdir -p gpurun_out/r2zi
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2zi/pytest_pipe.log 2>&1; tail -3 gpurun_out/r2zi/pytest_pipe.log
timeout 120 python profiles/time_pull_b2b.py > gpurun_out/r2zi/b2b_rotated.txt 2>&1
for v in dynrows static; do
  IB200_LIB=$PWD/profiles/lab_so/lib_$v.so timeout 120 python profiles/time_pull_b2b.py > gpurun_out/r2zi/b2b_$v.txt 2>&1
done
cat gpurun_out/r2zi/b2b_*.txt
