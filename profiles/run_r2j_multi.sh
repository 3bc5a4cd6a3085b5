# Final round-2 multi-GPU evidence (gpurun --gpus 8): headline weak scaling at N = 8 (device and e2e, slab-streamed host path),
# BASELINE config 5 sharded over 8 ranks with the two optional collectives timed apart (warm, mean of 3).
set -x
mkdir -p gpurun_out/r2j_multi
nvidia-smi topo -m > gpurun_out/r2j_multi/topo.txt 2>&1
run() { n=$1; shift; timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $n --master-addr 127.0.0.1 --master-port $((29500 + n)) bench.py --gpus $n "$@"; }
run 8 --steps 20 --warmup 3 > gpurun_out/r2j_multi/headline_n8.json 2> gpurun_out/r2j_multi/headline_n8.err
NCCL_DEBUG=INFO NCCL_DEBUG_FILE=gpurun_out/r2j_multi/nccl_%h_%p.log run 8 --config cfg5 --steps 10 --warmup 3 --collectives > gpurun_out/r2j_multi/cfg5_n8.json 2> gpurun_out/r2j_multi/cfg5_n8.err


cat gpurun_out/r2j_multi/nccl_*.log | grep -E "NVLS|Connected|via P2P|nRanks" | head -12 > gpurun_out/r2j_multi/nccl_info.txt
for f in gpurun_out/r2j_multi/*.json; do echo "== $f"; python - "$f" <<'PY'
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
