#!/bin/bash
# GPU call A (1 GPU): new tests + whole gpu suite + smoke (plain and under ncu) + N=1 bench
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/a_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_cuda_round2.py -q --timeout 240 -p no:cacheprovider > gpurun_out/a_round2.log 2>&1
echo "round2 rc=$?" >> gpurun_out/a_round2.log
timeout 1500 python -m pytest tests -m gpu -q --timeout 240 -p no:cacheprovider --deselect tests/test_cuda_round2.py > gpurun_out/a_gpu_suite.log 2>&1
echo "suite rc=$?" >> gpurun_out/a_gpu_suite.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/a_smoke.log 2>&1
echo "smoke rc=$?" >> gpurun_out/a_smoke.log
timeout 600 python bench.py --gpus 1 > gpurun_out/a_bench1.log 2>&1
echo "bench1 rc=$?" >> gpurun_out/a_bench1.log
timeout 600 python bench.py --gpus 1 --impl reference > gpurun_out/a_bench1_ref.log 2>&1
echo "bench1 ref rc=$?" >> gpurun_out/a_bench1_ref.log
for f in a_round2 a_gpu_suite a_smoke a_bench1 a_bench1_ref; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
