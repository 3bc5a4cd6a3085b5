set -x
mkdir -p gpurun_out/r2w
timeout 300 python profiles/time_api_latency.py > gpurun_out/r2w/api_latency.txt 2>&1; cat gpurun_out/r2w/api_latency.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2w/pytest_gpu.log 2>&1; tail -3 gpurun_out/r2w/pytest_gpu.log
