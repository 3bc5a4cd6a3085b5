# N = 2 sanity at the final commit (gpurun --gpus 2): headline weak scaling and cfg 5 sharded with the collectives
set -x
mkdir -p gpurun_out/r2zk
run() { n=$1; shift; timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
run 2 --steps 20 --warmup 3 > gpurun_out/r2zk/headline_n2.json 2> gpurun_out/r2zk/headline_n2.err
run 2 --config cfg5 --steps 5 --warmup 3 --collectives --no-e2e > gpurun_out/r2zk/cfg5_n2.json 2> gpurun_out/r2zk/cfg5_n2.err
for f in gpurun_out/r2zk/*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline', {})
    print('n', d['n_gpus'], '| value %.0f' % d['value'], '| ms %.3f' % d['ms_per_step'], '| e2e', d.get('e2e', {}).get('value'), d.get('e2e', {}).get('ms_per_step'),
          '| frac', r.get('frac'), r.get('frac_of_n_gpus_peak'), '| coll', d.get('collectives'),
          '| ops', {k: round(v['ms'], 3) for k, v in r.get('ops', {}).items()})
except Exception as e:
    print('unreadable', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
done
