# End of round 2: full GPU suite, smoke(), every BASELINE configuration through bench.py at the final commit.
set -x
mkdir -p gpurun_out/r2zg
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2zg/pytest_gpu.log 2>&1; tail -3 gpurun_out/r2zg/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zg/smoke.log 2>&1; tail -2 gpurun_out/r2zg/smoke.log
timeout 600 python bench.py > gpurun_out/r2zg/bench_headline.json 2> gpurun_out/r2zg/bench_headline.err
for cfg in cfg3 backward; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 3 > gpurun_out/r2zg/bench_$cfg.json 2> gpurun_out/r2zg/bench_$cfg.err
done
for cfg in cfg1 cfg2 cfg4; do
  timeout 600 python bench.py --config $cfg --steps 20 --warmup 3 --no-cpu-baseline > gpurun_out/r2zg/bench_$cfg.json 2> gpurun_out/r2zg/bench_$cfg.err
done
for f in gpurun_out/r2zg/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline', {})
    print(d['metric'], '| value %.0f' % d['value'], '| ms %.3f' % d['ms_per_step'], '| e2e', d.get('e2e', {}).get('value'), d.get('e2e', {}).get('ms_per_step'),
          '| frac', r.get('frac'), r.get('kernel'), '| cpu', d.get('cpu_baseline', {}).get('value'), d.get('cpu_baseline', {}).get('kind'),
          '| parity', d.get('parity_rel'), '| ops', {k: round(v['ms'], 3) for k, v in r.get('ops', {}).items()})
except Exception as e:
    print('unreadable', e)
PY
done
