set -x
mkdir -p gpurun_out/r2zl
timeout 300 python profiles/time_pipe_threshold.py > gpurun_out/r2zl/pipe_threshold.txt 2>&1; cat gpurun_out/r2zl/pipe_threshold.txt
