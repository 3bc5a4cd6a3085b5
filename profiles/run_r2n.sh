# native label storage types: label tests + timing of the 256^3 int64 label pull
set -x
mkdir -p gpurun_out/r2n
timeout 600 python -m pytest tests/test_gpu_ops.py tests/test_gpu_golden.py -x -q -k "label" > gpurun_out/r2n/pytest_labels.log 2>&1; tail -5 gpurun_out/r2n/pytest_labels.log
timeout 300 python profiles/time_labels.py > gpurun_out/r2n/time_labels.txt 2>&1; tail -5 gpurun_out/r2n/time_labels.txt
