set -x
mkdir -p gpurun_out/r2t
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pull_tile3d -s 3 -c 1 -f -o gpurun_out/r2t/prof_pull_c4il python profiles/time_ops.py --channels 4 --ops pull > gpurun_out/r2t/ncu.log 2>&1
python profiles/ncu_phases.py gpurun_out/r2t/prof_pull_c4il.ncu-rep > gpurun_out/r2t/phases.txt 2>&1
python profiles/ncu_summary.py gpurun_out/r2t/prof_pull_c4il.ncu-rep > gpurun_out/r2t/raw.txt 2>&1
cut -c1-250 gpurun_out/r2t/phases.txt; grep -E "gpu__time|lsu_wavefronts.sum.pct|registers_per|warps_active|issue_active|occupancy_limit" gpurun_out/r2t/raw.txt
