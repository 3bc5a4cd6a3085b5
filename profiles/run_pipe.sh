mkdir -p gpurun_out
(timeout 400 python -m pytest tests/test_gpu_pipe.py -x -q 2>&1 | tail -15) > gpurun_out/pytest_pipe.log
timeout 60 python profiles/time_ops.py --ops pull,grad,push,count > gpurun_out/time_pipe_256.txt 2>&1
timeout 60 python profiles/time_ops.py --ops pull,grad,push --channels 4 > gpurun_out/time_pipe_256_c4.txt 2>&1
timeout 60 python profiles/time_ops.py --ops pull,push --size 128 > gpurun_out/time_pipe_128.txt 2>&1
timeout 60 python profiles/time_ops.py --ops pull,push --order 1 > gpurun_out/time_pipe_256_o1.txt 2>&1
cat gpurun_out/pytest_pipe.log; grep -h "Mvox" gpurun_out/time_pipe_*.txt
