mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pull_pipe3d -s 3 -c 1 -o gpurun_out/prof_pull_pipe -f python profiles/time_ops.py --ops pull > gpurun_out/ncu_pull_pipe.log 2>&1
tail -3 gpurun_out/ncu_pull_pipe.log
