# N = 8 headline at the final commit (gpurun --gpus 8)
set -x
mkdir -p gpurun_out/r2zn
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29508 bench.py --gpus 8 --steps 20 --warmup 3 --no-e2e > gpurun_out/r2zn/headline_n8.json 2> gpurun_out/r2zn/headline_n8.err
python - <<'PY'
import json
d = json.loads(open('gpurun_out/r2zn/headline_n8.json').read().strip().splitlines()[-1])
r = d['roofline']
print('n', d['n_gpus'], '| value %.0f' % d['value'], '| ms %.3f' % d['ms_per_step'], '| ops', {k: round(v['ms'], 3) for k, v in r['ops'].items()}, d.get('clocks'))
PY
