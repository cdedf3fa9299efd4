#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_cuda_round2.py tests/test_cuda_collectives.py tests/test_cuda_allreduce.py -q --timeout 300 -p no:cacheprovider -x > gpurun_out/d_tests.log 2>&1
echo "tests rc=$?" >> gpurun_out/d_tests.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/d_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/d_smoke.log
tail -n 5 gpurun_out/d_tests.log gpurun_out/d_smoke.log
