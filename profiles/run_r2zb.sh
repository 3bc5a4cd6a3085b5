# persistent pull: scout warp + dynamic tile claims (claim issued early, coordinates requested after the boxes)
set -x
mkdir -p gpurun_out/r2zb
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2zb/pytest_pipe.log 2>&1; tail -5 gpurun_out/r2zb/pytest_pipe.log
for o in 3 2 1; do
timeout 120 python profiles/time_ops.py --ops pull,grad,bwd_grid --order $o > gpurun_out/r2zb/time_ops_o$o.txt 2>&1
done
IB200_STATIC_TILES=1 timeout 120 python profiles/time_ops.py --ops pull,grad > gpurun_out/r2zb/time_ops_o3_static.txt 2>&1
grep -H Mvox gpurun_out/r2zb/time_ops_*.txt
timeout 300 python bench.py --steps 20 --warmup 5 > gpurun_out/r2zb/bench_headline.json 2> gpurun_out/r2zb/bench_headline.err; python -c "
import json; d=json.load(open('gpurun_out/r2zb/bench_headline.json')); print(d['value'], d['ms_per_step'], d['roofline']['ops'], d['e2e']['ms_per_step'], d['parity'])"
