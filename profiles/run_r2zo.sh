# boxed push: geometry words of the flush loop in registers instead of local memory (7 LDL per vector)
set -x
mkdir -p gpurun_out/r2zo
timeout 900 python -m pytest tests/test_gpu_pipe.py tests/test_gpu_tile_parity.py tests/test_gpu_fuzz.py -x -q > gpurun_out/r2zo/pytest.log 2>&1; tail -3 gpurun_out/r2zo/pytest.log
timeout 300 python bench.py --no-cpu-baseline --no-e2e > gpurun_out/r2zo/bench_headline.json 2> gpurun_out/r2zo/bench_headline.err
python -c "
import json; d=json.load(open('gpurun_out/r2zo/bench_headline.json')); print(d['value'], d['ms_per_step'], {k: round(v['ms'],4) for k,v in d['roofline']['ops'].items()})"
timeout 120 python profiles/time_ops.py --ops push,count > gpurun_out/r2zo/time_ops.txt 2>&1; grep Mvox gpurun_out/r2zo/time_ops.txt
timeout 120 python profiles/time_ops.py --ops push,count --order 5 --dtype f16 --bound 6 > gpurun_out/r2zo/time_ops_cfg4.txt 2>&1; grep Mvox gpurun_out/r2zo/time_ops_cfg4.txt
