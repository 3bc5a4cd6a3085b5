# adaptive number of tiles along z (probe kernel) + dynamic tile claims: tests and timings
set -x
mkdir -p gpurun_out/r2za
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2za/pytest_pipe.log 2>&1; tail -5 gpurun_out/r2za/pytest_pipe.log
timeout 120 python profiles/time_ops.py --ops pull,grad,bwd_grid > gpurun_out/r2za/time_ops_adapt.txt 2>&1
IB200_NO_ZADAPT=1 timeout 120 python profiles/time_ops.py --ops pull,grad,bwd_grid > gpurun_out/r2za/time_ops_noadapt.txt 2>&1
IB200_NCW=18 timeout 120 python profiles/time_ops.py --ops pull > gpurun_out/r2za/time_ops_adapt_ncw18.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull,grad --order 1 > gpurun_out/r2za/time_ops_adapt_o1.txt 2>&1
timeout 120 python profiles/time_ops.py --ops pull,grad --order 2 > gpurun_out/r2za/time_ops_adapt_o2.txt 2>&1
grep -H Mvox gpurun_out/r2za/time_ops_*.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2za/bench_headline.json 2> gpurun_out/r2za/bench_headline.err; cat gpurun_out/r2za/bench_headline.json
