# persistent pull: output strides / coordinate-slot address / output pointer pinned in registers (no S2R, fewer LDCU per row)
set -x
mkdir -p gpurun_out/r2zr
timeout 600 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2zr/pytest.log 2>&1; tail -2 gpurun_out/r2zr/pytest.log
timeout 100 python profiles/time_pull_b2b.py > gpurun_out/r2zr/b2b_pinned.txt 2>&1
IB200_LIB=$PWD/profiles/lab_so/lib_before.so timeout 100 python profiles/time_pull_b2b.py > gpurun_out/r2zr/b2b_before.txt 2>&1
cat gpurun_out/r2zr/b2b_*.txt
