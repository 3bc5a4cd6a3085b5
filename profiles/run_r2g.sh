# k-split for stretched rows in the pull pipe: pipe tests + timings
set -x
mkdir -p gpurun_out/r2g
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py -x -q > gpurun_out/r2g/pytest_pipe.log 2>&1; tail -5 gpurun_out/r2g/pytest_pipe.log
timeout 120 python profiles/time_ops.py --ops pull,grad,push > gpurun_out/r2g/time_ops_256_o3.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull --order 1 > gpurun_out/r2g/time_ops_256_o1.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull --order 2 > gpurun_out/r2g/time_ops_256_o2.txt 2>&1
grep -h Mvox gpurun_out/r2g/time_ops_*.txt
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2g/smoke.log 2>&1; tail -3 gpurun_out/r2g/smoke.log
