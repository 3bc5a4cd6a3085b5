# Round-2 push evidence: full GPU test suite with the boxed push kernel as default, timings, ncu full capture +
# phase attribution.
set -x
mkdir -p gpurun_out/r2b
(timeout 1500 python -m pytest tests -m gpu -q 2>&1 | tail -15) > gpurun_out/r2b/pytest_gpu.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:push_box3d -s 3 -c 1 -f -o gpurun_out/r2b/prof_push_box python profiles/time_ops.py --ops push > gpurun_out/r2b/ncu_push.log 2>&1
python profiles/ncu_phases.py gpurun_out/r2b/prof_push_box.ncu-rep > gpurun_out/r2b/ncu_push_box_phases.txt 2>&1
python profiles/ncu_summary.py gpurun_out/r2b/prof_push_box.ncu-rep > gpurun_out/r2b/ncu_push_box_raw.txt 2>&1
timeout 120 python profiles/time_ops.py > gpurun_out/r2b/time_ops_256_o3.txt 2>&1
timeout 120 python profiles/time_ops.py --channels 4 --ops pull,grad,push,coeff > gpurun_out/r2b/time_ops_256_o3_c4.txt 2>&1
timeout 120 python profiles/time_ops.py --dtype f16 --order 5 --bound 6 --ops push,count > gpurun_out/r2b/time_ops_256_o5_f16.txt 2>&1
timeout 120 python profiles/time_ops.py --dtype f16 --order 5 --bound 6 --ops push,count --incoherent > gpurun_out/r2b/time_ops_256_o5_f16_incoherent.txt 2>&1
timeout 120 python profiles/time_ops.py --size 128 --ops pull,push > gpurun_out/r2b/time_ops_128.txt 2>&1
cat gpurun_out/r2b/pytest_gpu.log; cut -c1-230 gpurun_out/r2b/ncu_push_box_phases.txt; grep -h Mvox gpurun_out/r2b/time_ops_*.txt
