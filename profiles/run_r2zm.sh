# Final commit of round 2: full GPU suite, smoke(), headline bench (both arms), cfg 5 on one GPU.
set -x
mkdir -p gpurun_out/r2zm
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2zm/pytest_gpu.log 2>&1; tail -3 gpurun_out/r2zm/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zm/smoke.log 2>&1; tail -2 gpurun_out/r2zm/smoke.log
timeout 600 python bench.py > gpurun_out/r2zm/bench_headline.json 2> gpurun_out/r2zm/bench_headline.err
timeout 600 python bench.py --config backward > gpurun_out/r2zm/bench_backward.json 2> gpurun_out/r2zm/bench_backward.err
timeout 600 python bench.py --config cfg5 --steps 5 --warmup 3 --no-cpu-baseline > gpurun_out/r2zm/bench_cfg5.json 2> gpurun_out/r2zm/bench_cfg5.err
timeout 600 python bench.py --config cfg2 --no-cpu-baseline > gpurun_out/r2zm/bench_cfg2.json 2> gpurun_out/r2zm/bench_cfg2.err
for f in gpurun_out/r2zm/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline', {})
    print(d['metric'], '| value %.0f' % d['value'], '| ms %.3f' % d['ms_per_step'], '| e2e', d.get('e2e', {}).get('value'), d.get('e2e', {}).get('ms_per_step'),
          '| frac', r.get('frac'), r.get('kernel'), '| cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'),
          '| parity', d.get('parity_rel'), d.get('parity'), '| ops', {k: round(v['ms'], 3) for k, v in r.get('ops', {}).items()}, '| launches', d.get('gpu_launches'))
except Exception as e:
    print('unreadable', e)
PY
done
