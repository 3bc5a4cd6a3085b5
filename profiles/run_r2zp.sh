# last commit of round 2: full GPU suite + smoke()
set -x
mkdir -p gpurun_out/r2zp
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/r2zp/pytest_gpu.log 2>&1; tail -2 gpurun_out/r2zp/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2zp/smoke.log 2>&1; tail -2 gpurun_out/r2zp/smoke.log
