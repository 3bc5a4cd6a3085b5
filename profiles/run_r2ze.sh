# full ncu capture of the persistent pull at the current commit (scout warp, dynamic claims, new fix-up) and the
# launch list of the default bench command
set -x
mkdir -p gpurun_out/r2ze
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pull_pipe3d -s 3 -c 1 -f -o gpurun_out/r2ze/prof_pull_pipe python profiles/time_ops.py --ops pull > gpurun_out/r2ze/ncu_pull.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2ze/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2ze/bench_under_ncu.log 2>&1
tail -3 gpurun_out/r2ze/ncu_pull.log; grep -c . gpurun_out/r2ze/launches.csv
