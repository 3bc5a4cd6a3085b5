# Round-2 (session 2) first call: full GPU suite + the two configurations whose bench parity_rel was > 1 in r2c.
set -x
mkdir -p gpurun_out/r2e
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2e/pytest_gpu.log 2>&1; tail -5 gpurun_out/r2e/pytest_gpu.log
for cfg in cfg5 cfg4i; do
  timeout 900 python bench.py --config $cfg --steps 5 --warmup 3 > gpurun_out/r2e/bench_$cfg.json 2> gpurun_out/r2e/bench_$cfg.err
  tail -c 600 gpurun_out/r2e/bench_$cfg.err
  python - gpurun_out/r2e/bench_$cfg.json <<'PY'
import json, sys
d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d['metric'], d['value'], d.get('parity_rel'), d.get('parity'))
PY
done
