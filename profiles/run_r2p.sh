# cfg 3 (C = 4) on the tile kernels: where does the time go
set -x
mkdir -p gpurun_out/r2p
timeout 600 ncu --set full --clock-control none --import-source on -k regex:pull_tile3d -s 3 -c 1 -f -o gpurun_out/r2p/prof_pull_tile_c4 python profiles/time_ops.py --channels 4 --ops pull > gpurun_out/r2p/ncu_pull_c4.log 2>&1
python profiles/ncu_phases.py gpurun_out/r2p/prof_pull_tile_c4.ncu-rep > gpurun_out/r2p/ncu_pull_tile_c4_phases.txt 2>&1
python profiles/ncu_summary.py gpurun_out/r2p/prof_pull_tile_c4.ncu-rep > gpurun_out/r2p/ncu_pull_tile_c4_raw.txt 2>&1
cut -c1-260 gpurun_out/r2p/ncu_pull_tile_c4_phases.txt
