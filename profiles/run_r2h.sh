# streamed host-tensor pull: test + e2e of the headline configuration
set -x
mkdir -p gpurun_out/r2h
timeout 600 python -m pytest tests/test_gpu_ops.py -x -q -k "host_tensors or edge or api" > gpurun_out/r2h/pytest.log 2>&1; tail -15 gpurun_out/r2h/pytest.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/r2h/bench_headline.json 2> gpurun_out/r2h/bench_headline.err; tail -3 gpurun_out/r2h/bench_headline.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2h/bench_headline.json').read().strip().splitlines()[-1])
print('value', d['value'], 'ms', d['ms_per_step'], 'e2e', d['e2e'])
PY
