#!/bin/bash
# 1 GPU: new kernels of this batch (exchange, relay broadcast, hybrid build) + the whole round-2 file + smoke
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm --format=csv > gpurun_out/f_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_cuda_round2.py tests/test_cuda_collectives.py -x -q -m gpu 2>&1 | tail -15 > gpurun_out/f_tests.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/f_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/f_smoke.log
tail -5 gpurun_out/f_tests.log; tail -12 gpurun_out/f_smoke.log
