# full GPU suite + smoke after the host-layer refactor
set -x
mkdir -p gpurun_out/r2l
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/r2l/pytest_gpu.log 2>&1; tail -15 gpurun_out/r2l/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/r2l/smoke.log 2>&1; tail -3 gpurun_out/r2l/smoke.log
