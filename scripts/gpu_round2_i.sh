#!/bin/bash
# 1 GPU: ncu --set full of the hot kernels (flagship N=1 step at 400 MB, loop-back collectives at 64 MB),
# compute-sanitizer over the self-tests, p2p tests with the wider mailbox.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_round2.py -q -m gpu -k "p2p or exchange or relay" -p no:cacheprovider 2>&1 | tail -4 > gpurun_out/i_p2p.log
export PYTHONPATH=$PWD
timeout 900 ncu --set full --clock-control none --import-source on \
  -k regex:'localAllreduceMany|twoShotAllreduceKernel|pipelinedAllreduce|exchangeKernel|peerBulkCopy|gatherBulk|llAllreduce|hybridAllreduce|broadcastKernel' \
  -c 14 -f -o gpurun_out/prof_r2 python scripts/ncu_target.py > gpurun_out/i_ncu.log 2>&1
echo "ncu rc=$?" >> gpurun_out/i_ncu.log
ncu -i gpurun_out/prof_r2.ncu-rep --page raw --csv > gpurun_out/prof_r2_raw.csv 2>> gpurun_out/i_ncu.log
ls -la gpurun_out/prof_r2* >> gpurun_out/i_ncu.log
bash scripts/compute_sanitizer_run.sh gpurun_out > gpurun_out/i_sanitizer.log 2>&1
cat gpurun_out/i_p2p.log; tail -5 gpurun_out/i_ncu.log; cat gpurun_out/sanitizer_summary.txt
