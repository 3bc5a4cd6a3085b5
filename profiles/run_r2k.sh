set -x
mkdir -p gpurun_out/r2k
for n in 15 19 23; do IB200_PULL_NCW=$n timeout 120 python profiles/time_ops.py --ops pull > gpurun_out/r2k/time_pull_ncw$n.txt 2>&1; grep -h Mvox gpurun_out/r2k/time_pull_ncw$n.txt; done
