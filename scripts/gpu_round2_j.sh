#!/bin/bash
# GPU call J (8 GPUs): new kernels at 8 ranks, focused re-tune of the multicast family (few-CTA NVLS, NVLS + P2P
# hybrid), broadcast variants, long-context rows with the zero-copy exchange, NVLink counters, bench N=8 / N=4.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 500 python -m pytest tests/test_cuda_multiproc.py -q -k vmm --timeout 400 -p no:cacheprovider > gpurun_out/j_multiproc8.log 2>&1
echo "multiproc8 rc=$?" >> gpurun_out/j_multiproc8.log
timeout 500 $TR --nproc-per-node 8 --master-port 29571 -m gloo_b200.tune --focus nvls --kinds sym --collectives allreduce --min-bytes 1562500 --max-bytes 400000000 --quick --no-nccl --out gpurun_out/tune_P8_nvls > gpurun_out/j_tune8.log 2>&1
echo "tune8 rc=$?" >> gpurun_out/j_tune8.log
cp gloo_b200/tuning/b200.tune gpurun_out/j_b200.tune
[ -s gpurun_out/tune_P8_nvls.tune ] && python -m gloo_b200.tune --merge gpurun_out/j_b200.tune gpurun_out/tune_P8_nvls.tune gpurun_out/j_b200.tune --merge-min-bytes 1100000
timeout 300 $TR --nproc-per-node 8 --master-port 29572 scripts/bench_broadcast.py --out gpurun_out/broadcast_P8.json > gpurun_out/j_bcast8.log 2>&1
echo "bcast8 rc=$?" >> gpurun_out/j_bcast8.log
timeout 400 $TR --nproc-per-node 8 --master-port 29573 scripts/bench_longcontext.py --out gpurun_out/longcontext_P8.json > gpurun_out/j_lc8.log 2>&1
echo "lc8 rc=$?" >> gpurun_out/j_lc8.log
timeout 300 $TR --nproc-per-node 8 --master-port 29574 scripts/nvlink_evidence.py --steps 20 --out gpurun_out/nvlink_P8.json > gpurun_out/j_nvlink8.log 2>&1
echo "nvlink8 rc=$?" >> gpurun_out/j_nvlink8.log
# P=4 focus tune on GPUs 0-3 while GPUs 4-7 run the broadcast variants at P=4
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 400 $TR --nproc-per-node 4 --master-port 29575 -m gloo_b200.tune --focus nvls --kinds sym --collectives allreduce --min-bytes 1562500 --max-bytes 400000000 --quick --no-nccl --out gpurun_out/tune_P4_nvls > gpurun_out/j_tune4.log 2>&1; echo "tune4 rc=$?" >> gpurun_out/j_tune4.log ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 300 $TR --nproc-per-node 4 --master-port 29576 scripts/bench_broadcast.py --out gpurun_out/broadcast_P4.json > gpurun_out/j_bcast4.log 2>&1; echo "bcast4 rc=$?" >> gpurun_out/j_bcast4.log ) &
wait
[ -s gpurun_out/tune_P4_nvls.tune ] && python -m gloo_b200.tune --merge gpurun_out/j_b200.tune gpurun_out/tune_P4_nvls.tune gpurun_out/j_b200.tune --merge-min-bytes 1100000
export GLB_TUNE_FILE=$PWD/gpurun_out/j_b200.tune
timeout 900 $TR --nproc-per-node 8 --master-port 29577 bench.py --gpus 8 > gpurun_out/j_bench8.log 2>&1
echo "bench8 rc=$?" >> gpurun_out/j_bench8.log
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29578 bench.py --gpus 4 > gpurun_out/j_bench4.log 2>&1; echo "bench4 rc=$?" >> gpurun_out/j_bench4.log ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 400 $TR --nproc-per-node 4 --master-port 29579 scripts/bench_longcontext.py --out gpurun_out/longcontext_P4.json > gpurun_out/j_lc4.log 2>&1; echo "lc4 rc=$?" >> gpurun_out/j_lc4.log ) &
wait
for f in j_multiproc8 j_tune8 j_bcast8 j_lc8 j_nvlink8 j_tune4 j_bcast4 j_lc4; do echo "== $f"; tail -n 5 gpurun_out/$f.log | cut -c 1-600; done
for f in j_bench8 j_bench4; do echo "== $f"; tail -c 500 gpurun_out/$f.log; done
