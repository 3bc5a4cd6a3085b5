# ncu --set full of both headline kernels at the last commit (traffic.json provenance)
set -x
mkdir -p gpurun_out/r2zq
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pull_pipe3d -s 3 -c 1 -f -o gpurun_out/r2zq/prof_pull_pipe python profiles/time_ops.py --ops pull > gpurun_out/r2zq/ncu_pull.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:push_box3d -s 3 -c 1 -f -o gpurun_out/r2zq/prof_push_box python profiles/time_ops.py --ops push > gpurun_out/r2zq/ncu_push.log 2>&1
ls -la gpurun_out/r2zq
