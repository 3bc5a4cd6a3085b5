# where do the 120 clk of a conflict-free row go?  A/B builds of pull_pipe.cu: static row assignment, no validity checks
set -x
mkdir -p gpurun_out/r2zh
timeout 120 python profiles/time_pull_b2b.py > gpurun_out/r2zh/b2b_default.txt 2>&1
for v in static nochk both; do
  IB200_LIB=$PWD/profiles/lab_so/lib_$v.so timeout 120 python profiles/time_pull_b2b.py > gpurun_out/r2zh/b2b_$v.txt 2>&1
done
cat gpurun_out/r2zh/b2b_*.txt
timeout 300 python bench.py --config backward --steps 10 --warmup 3 > gpurun_out/r2zh/bench_backward.json 2> gpurun_out/r2zh/bench_backward.err; python -c "
import json; d=json.load(open('gpurun_out/r2zh/bench_backward.json')); print(d['value'], d['ms_per_step'], d.get('parity_rel'), d.get('parity'))"
