set -x
mkdir -p gpurun_out/r2zf
timeout 300 python profiles/time_stretch.py > gpurun_out/r2zf/time_stretch.txt 2>&1; cat gpurun_out/r2zf/time_stretch.txt
