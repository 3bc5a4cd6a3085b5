# pull pipe trimming: pipe tests + timings (+ ncu source counters of the pull kernel)
set -x
mkdir -p gpurun_out/r2i
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py -x -q > gpurun_out/r2i/pytest_pipe.log 2>&1; tail -3 gpurun_out/r2i/pytest_pipe.log
timeout 120 python profiles/time_ops.py --ops pull,grad > gpurun_out/r2i/time_ops_256_o3.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull,grad --order 1 > gpurun_out/r2i/time_ops_256_o1.txt 2>&1
grep -h Mvox gpurun_out/r2i/time_ops_*.txt
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pull_pipe3d -s 3 -c 1 -f -o gpurun_out/r2i/prof_pull_pipe python profiles/time_ops.py --ops pull > gpurun_out/r2i/ncu_pull.log 2>&1
