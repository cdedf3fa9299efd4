#!/bin/bash
# GPU call C (8 GPUs): correctness at 8 ranks (NVLS), tuning tables for P=2 (user kind), 4 and 8,
# bench at N=8 / N=4 with the fresh table, NVLink byte counters, long-context micro-benchmarks.
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/c_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/c_gpus.txt 2>&1
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
# -- correctness at 8 ranks (cuMem + NVLS multicast path) --------------------------------------
timeout 600 python -m pytest tests/test_cuda_multiproc.py -q -k vmm --timeout 500 -p no:cacheprovider > gpurun_out/c_multiproc8.log 2>&1
echo "multiproc8 rc=$?" >> gpurun_out/c_multiproc8.log
# -- tuning: P=4 on GPUs 0-3, at the same time P=2 (plain-pointer kinds) on GPUs 4-5 and the
#    long-context benchmark on GPUs 6-7 (disjoint GPU sets: every GPU has its own NVLink ports)
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29541 -m gloo_b200.tune --out gpurun_out/tune_P4 > gpurun_out/c_tune4.log 2>&1; echo "tune4 rc=$?" >> gpurun_out/c_tune4.log ) &
( CUDA_VISIBLE_DEVICES=4,5 timeout 600 $TR --nproc-per-node 2 --master-port 29542 -m gloo_b200.tune --out gpurun_out/tune_P2user --kinds user --collectives allreduce > gpurun_out/c_tune2u.log 2>&1; echo "tune2u rc=$?" >> gpurun_out/c_tune2u.log ) &
( CUDA_VISIBLE_DEVICES=6,7 timeout 600 $TR --nproc-per-node 2 --master-port 29543 scripts/bench_longcontext.py --out gpurun_out/longcontext_P2.json > gpurun_out/c_lc2.log 2>&1; echo "lc2 rc=$?" >> gpurun_out/c_lc2.log ) &
wait
timeout 1200 $TR --nproc-per-node 8 --master-port 29544 -m gloo_b200.tune --out gpurun_out/tune_P8 > gpurun_out/c_tune8.log 2>&1
echo "tune8 rc=$?" >> gpurun_out/c_tune8.log
# -- merged table ------------------------------------------------------------------------------
{ cat gloo_b200/tuning/b200.tune; grep "buf=user" gpurun_out/tune_P2user.tune; cat gpurun_out/tune_P4.tune gpurun_out/tune_P8.tune; } > gpurun_out/b200.tune 2>/dev/null
export GLB_TUNE_FILE=$PWD/gpurun_out/b200.tune
timeout 1200 $TR --nproc-per-node 8 --master-port 29545 bench.py --gpus 8 > gpurun_out/c_bench8.log 2>&1
echo "bench8 rc=$?" >> gpurun_out/c_bench8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29546 scripts/nvlink_evidence.py --out gpurun_out/nvlink_P8.json > gpurun_out/c_nvlink8.log 2>&1
echo "nvlink8 rc=$?" >> gpurun_out/c_nvlink8.log
timeout 600 $TR --nproc-per-node 8 --master-port 29547 scripts/bench_longcontext.py --out gpurun_out/longcontext_P8.json > gpurun_out/c_lc8.log 2>&1
echo "lc8 rc=$?" >> gpurun_out/c_lc8.log
( CUDA_VISIBLE_DEVICES=0,1,2,3 timeout 900 $TR --nproc-per-node 4 --master-port 29548 bench.py --gpus 4 > gpurun_out/c_bench4.log 2>&1; echo "bench4 rc=$?" >> gpurun_out/c_bench4.log ) &
( CUDA_VISIBLE_DEVICES=4,5,6,7 timeout 600 $TR --nproc-per-node 4 --master-port 29549 scripts/nvlink_evidence.py --out gpurun_out/nvlink_P4.json > gpurun_out/c_nvlink4.log 2>&1; echo "nvlink4 rc=$?" >> gpurun_out/c_nvlink4.log ) &
wait
for f in c_multiproc8 c_tune4 c_tune2u c_lc2 c_tune8 c_nvlink8 c_lc8 c_nvlink4; do echo "== $f"; tail -n 4 gpurun_out/$f.log; done
for f in c_bench8 c_bench4; do echo "== $f"; tail -c 600 gpurun_out/$f.log; done
