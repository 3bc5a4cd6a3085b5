# Round-2 multi-GPU evidence (gpurun --gpus 8): headline weak scaling at N = 2 / 4 / 8 (device and e2e), BASELINE
# config 5 (batch 64 x 192^3, mixed bounds) sharded over 8 / 4 / 2 ranks with the optional collectives timed apart.
set -x
mkdir -p gpurun_out/r2d
nvidia-smi topo -m > gpurun_out/r2d/topo.txt 2>&1
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 3 > gpurun_out/r2d/headline_n8.json 2> gpurun_out/r2d/headline_n8.err
run 4 --steps 20 --warmup 3 > gpurun_out/r2d/headline_n4.json 2> gpurun_out/r2d/headline_n4.err
run 2 --steps 20 --warmup 3 > gpurun_out/r2d/headline_n2.json 2> gpurun_out/r2d/headline_n2.err
NCCL_DEBUG=INFO run 8 --config cfg5 --steps 10 --warmup 3 --collectives > gpurun_out/r2d/cfg5_n8.json 2> gpurun_out/r2d/cfg5_n8.err
run 4 --config cfg5 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2d/cfg5_n4.json 2> gpurun_out/r2d/cfg5_n4.err
run 2 --config cfg5 --steps 10 --warmup 3 --no-e2e > gpurun_out/r2d/cfg5_n2.json 2> gpurun_out/r2d/cfg5_n2.err
grep -E "NVLS|Connected|via P2P|nRanks" gpurun_out/r2d/cfg5_n8.err | head -12 > gpurun_out/r2d/nccl_info.txt
for f in gpurun_out/r2d/*.json; do echo "== $f"; python - "$f" <<'PY'
import json, sys
try:
    d = json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
    r = d.get('roofline', {})
    print('n', d['n_gpus'], '| value %.0f' % d['value'], '| ms %.3f' % d['ms_per_step'], '| e2e', d.get('e2e', {}).get('value'), d.get('e2e', {}).get('ms_per_step'),
          '| frac', r.get('frac'), r.get('frac_of_n_gpus_peak'), '| coll', d.get('collectives'), '| cpus', d.get('host_cpus_bound'),
          '| ops', {k: round(v['ms'], 3) for k, v in r.get('ops', {}).items()})
except Exception as e:
    print('unreadable', e); print(open(sys.argv[1].replace('.json', '.err')).read()[-1500:])
PY
done
