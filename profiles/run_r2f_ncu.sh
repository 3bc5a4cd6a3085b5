# Round 2, session 2: full ncu captures of the two headline kernels at the current commit (pull pipe, boxed push)
# and the launch list of the default bench command.
set -x
mkdir -p gpurun_out/r2f
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pull_pipe3d -s 3 -c 1 -f -o gpurun_out/r2f/prof_pull_pipe python profiles/time_ops.py --ops pull > gpurun_out/r2f/ncu_pull.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_box3d -s 3 -c 1 -f -o gpurun_out/r2f/prof_push_box python profiles/time_ops.py --ops push > gpurun_out/r2f/ncu_push.log 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2f/launches.csv python bench.py --steps 2 --warmup 1 --no-cpu-baseline > gpurun_out/r2f/bench_under_ncu.log 2>&1
timeout 120 python profiles/time_ops.py > gpurun_out/r2f/time_ops_256_o3.txt 2>&1
timeout 120 python profiles/time_ops.py --channels 4 --ops pull,grad,push,coeff > gpurun_out/r2f/time_ops_256_o3_c4.txt 2>&1
timeout 120 python profiles/time_coeff.py > gpurun_out/r2f/time_coeff.txt 2>&1
tail -3 gpurun_out/r2f/ncu_pull.log gpurun_out/r2f/ncu_push.log; grep -h Mvox gpurun_out/r2f/time_ops_*.txt; tail -20 gpurun_out/r2f/time_coeff.txt
