#!/bin/bash
# GPU call H (2 GPUs): new kernels across two real GPUs (hybrid at P=2, zero-copy exchange), fault tests,
# long-context rows with the exchange column (+ a wider mailbox ring), reference arm at N=2, bench N=2.
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --master-addr 127.0.0.1"
timeout 700 python -m pytest tests/test_cuda_multiproc.py tests/test_cuda_faults.py -q -m gpu --timeout 600 -p no:cacheprovider > gpurun_out/h_multiproc2.log 2>&1
echo "multiproc2 rc=$?" >> gpurun_out/h_multiproc2.log
timeout 300 $TR --nproc-per-node 2 --master-port 29561 scripts/bench_longcontext.py --seq 32768,131072,524288 --out gpurun_out/h_lc2.json > gpurun_out/h_lc2.log 2>&1
echo "lc2 rc=$?" >> gpurun_out/h_lc2.log
GLB_CUDA_P2P_LANES=30 GLB_CUDA_P2P_SLOT_KB=2048 timeout 300 $TR --nproc-per-node 2 --master-port 29562 scripts/bench_longcontext.py --seq 131072,524288 --out gpurun_out/h_lc2_wide.json > gpurun_out/h_lc2_wide.log 2>&1
echo "lc2 wide rc=$?" >> gpurun_out/h_lc2_wide.log
GLB_CUDA_EXCHANGE_BLOCKS=64 timeout 300 $TR --nproc-per-node 2 --master-port 29563 scripts/bench_longcontext.py --seq 131072,524288 --out gpurun_out/h_lc2_x64.json > gpurun_out/h_lc2_x64.log 2>&1
echo "lc2 x64 rc=$?" >> gpurun_out/h_lc2_x64.log
timeout 400 $TR --nproc-per-node 2 --master-port 29564 -m gloo_b200.tune --focus nvls --kinds sym --collectives allreduce --min-bytes 4194304 --max-bytes 268435456 --quick --no-nccl --out gpurun_out/h_tune2_nvls > gpurun_out/h_tune2_nvls.log 2>&1
echo "tune2 nvls rc=$?" >> gpurun_out/h_tune2_nvls.log
timeout 600 $TR --nproc-per-node 2 --master-port 29565 bench.py --impl reference --gpus 2 --steps 5 --warmup 3 --quick --no-sweep > gpurun_out/h_bench2_ref.log 2>&1
echo "bench2 ref rc=$?" >> gpurun_out/h_bench2_ref.log
timeout 900 $TR --nproc-per-node 2 --master-port 29566 bench.py --gpus 2 --quick > gpurun_out/h_bench2.log 2>&1
echo "bench2 rc=$?" >> gpurun_out/h_bench2.log
for f in h_multiproc2 h_lc2 h_lc2_wide h_lc2_x64 h_tune2_nvls; do echo "== $f"; tail -n 6 gpurun_out/$f.log | cut -c 1-400; done
for f in h_bench2_ref h_bench2; do echo "== $f"; tail -c 700 gpurun_out/$f.log; done
