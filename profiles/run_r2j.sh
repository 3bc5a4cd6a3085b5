# Final round-2 evidence, one GPU: every BASELINE configuration through bench.py (both arms for the headline).
set -x
mkdir -p gpurun_out/r2j
for cfg in headline cfg1 cfg2 cfg3 cfg4 cfg4i backward; do
  timeout 900 python bench.py --config $cfg --steps 20 --warmup 3 > gpurun_out/r2j/bench_$cfg.json 2> gpurun_out/r2j/bench_$cfg.err
done
timeout 900 python bench.py --config cfg5 --steps 5 --warmup 3 > gpurun_out/r2j/bench_cfg5_n1.json 2> gpurun_out/r2j/bench_cfg5_n1.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2j/bench_ref.json 2> gpurun_out/r2j/bench_ref.err
for f in gpurun_out/r2j/bench_*.json; do echo "== $f"; python - "$f" <<'PY'
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
