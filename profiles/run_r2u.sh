# channel-interleaved tile kernels: full suite, fuzz with C = 4, smoke, cfg 3 bench line
set -x
mkdir -p gpurun_out/r2u
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2u/pytest_gpu.log 2>&1; tail -4 gpurun_out/r2u/pytest_gpu.log
for seed in 21 22 23; do timeout 300 python profiles/fuzz_fast_vs_generic.py 400 $seed > gpurun_out/r2u/fuzz_seed$seed.txt 2>&1; tail -1 gpurun_out/r2u/fuzz_seed$seed.txt | cut -c1-200; grep -c MISMATCH gpurun_out/r2u/fuzz_seed$seed.txt; done
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2u/smoke.log 2>&1; tail -2 gpurun_out/r2u/smoke.log
timeout 900 python bench.py --config cfg3 --steps 20 --warmup 3 > gpurun_out/r2u/bench_cfg3.json 2> gpurun_out/r2u/bench_cfg3.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2u/bench_cfg3.json').read().strip().splitlines()[-1])
print(d['value'], d['ms_per_step'], {k: round(v['ms'], 3) for k, v in d['roofline']['ops'].items()}, d.get('parity_rel'), d['e2e']['ms_per_step'])
PY
