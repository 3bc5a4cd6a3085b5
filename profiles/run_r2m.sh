# compute-sanitizer on the persistent pull / grad and the boxed push / count kernels
set -x
mkdir -p gpurun_out/r2m
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 1 python profiles/memcheck_pipe.py > gpurun_out/r2m/memcheck.log 2>&1; echo "memcheck rc=$?" >> gpurun_out/r2m/memcheck.log; tail -5 gpurun_out/r2m/memcheck.log
timeout 900 compute-sanitizer --tool initcheck --error-exitcode 1 python profiles/memcheck_pipe.py > gpurun_out/r2m/initcheck.log 2>&1; echo "initcheck rc=$?" >> gpurun_out/r2m/initcheck.log; tail -5 gpurun_out/r2m/initcheck.log
