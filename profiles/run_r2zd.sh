# fix-up pass: vector copies from the volume for rows / planes whose source lies outside the box, whole-vector dft wrap
set -x
mkdir -p gpurun_out/r2zd
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2zd/pytest_pipe.log 2>&1; tail -3 gpurun_out/r2zd/pytest_pipe.log
for b in 3 4 6; do
timeout 120 python profiles/time_ops.py --ops pull,grad --bound $b > gpurun_out/r2zd/time_ops_o3_bound$b.txt 2>&1
done
grep -H Mvox gpurun_out/r2zd/time_ops_*.txt
timeout 600 compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/memcheck_pipe.py > gpurun_out/r2zd/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2zd/memcheck.log; tail -4 gpurun_out/r2zd/memcheck.log
timeout 300 python bench.py --config cfg5 --steps 5 --warmup 3 > gpurun_out/r2zd/bench_cfg5.json 2> gpurun_out/r2zd/bench_cfg5.err; python -c "
import json; d=json.load(open('gpurun_out/r2zd/bench_cfg5.json')); print(d['value'], d['ms_per_step'], d['roofline']['ops'], d.get('parity'))"
