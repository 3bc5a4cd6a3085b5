# Round-1 evidence run: parity tests, bench (both arms), per-op timings, ncu launch list + full captures.
set -x
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/smi.txt
(timeout 900 python -m pytest tests -m gpu -x -q 2>&1 | tail -8) > gpurun_out/pytest_gpu.log
timeout 600 python bench.py --steps 20 --warmup 3 > gpurun_out/bench_ours.json 2> gpurun_out/bench_ours.err
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err
timeout 120 python profiles/time_ops.py > gpurun_out/time_ops_256_o3.txt 2>&1
timeout 120 python profiles/time_ops.py --channels 4 --ops pull,grad,push,coeff > gpurun_out/time_ops_256_o3_c4.txt 2>&1
timeout 120 python profiles/time_ops.py --dtype f16 --order 5 --bound 6 --ops push,count > gpurun_out/time_ops_256_o5_f16.txt 2>&1
timeout 120 python profiles/time_ops.py --size 128 --ops pull,push > gpurun_out/time_ops_128.txt 2>&1
timeout 120 python profiles/time_ops.py --order 1 --ops pull,push,count,grad > gpurun_out/time_ops_256_o1.txt 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/bench_under_ncu.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:pull_pipe3d -s 3 -c 1 -f -o gpurun_out/prof_pull python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_pull.log 2>&1
timeout 400 ncu --set full --clock-control none --import-source on -k regex:push_tile3d -s 3 -c 1 -f -o gpurun_out/prof_push python bench.py --steps 2 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_push.log 2>&1
cat gpurun_out/pytest_gpu.log gpurun_out/bench_ours.json gpurun_out/bench_ref.json; grep -h Mvox gpurun_out/time_ops_*.txt
