#!/bin/bash
# Final 1-GPU validation: what the driver runs at round end (pytest -m gpu, smoke, bench N=1).
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -x -q -m gpu -p no:cacheprovider > gpurun_out/final_pytest_gpu.log 2>&1
echo "pytest rc=$?" >> gpurun_out/final_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/final_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/final_smoke.log
timeout 600 python bench.py > gpurun_out/final_bench1.log 2>&1
echo "bench rc=$?" >> gpurun_out/final_bench1.log
tail -4 gpurun_out/final_pytest_gpu.log; tail -3 gpurun_out/final_smoke.log; tail -c 1500 gpurun_out/final_bench1.log
