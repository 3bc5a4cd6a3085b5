set -x
mkdir -p gpurun_out/r2zs
timeout 120 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zs/smoke.log 2>&1; tail -2 gpurun_out/r2zs/smoke.log
timeout 120 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2zs/bench_headline.json 2> gpurun_out/r2zs/bench_headline.err
python -c "
import json; d=json.load(open('gpurun_out/r2zs/bench_headline.json')); print(d['value'], d['ms_per_step'], {k: round(v['ms'],4) for k,v in d['roofline']['ops'].items()}, d['roofline']['frac'])"
