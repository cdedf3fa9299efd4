#!/bin/bash
# GPU call B (2 GPUs): multi-process correctness (cuMem/cudaIpc, p2p, graphs, faults, torch backend),
# the previously failing single-GPU tests, the tuner at P=2, bench at N=2 (both arms).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/b_gpus.txt 2>&1
nvidia-smi topo -m >> gpurun_out/b_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_cuda_multiproc.py tests/test_cuda_faults.py -q --timeout 400 -p no:cacheprovider > gpurun_out/b_multiproc.log 2>&1
echo "multiproc rc=$?" >> gpurun_out/b_multiproc.log
timeout 600 python -m pytest tests/test_process_group.py -q -m gpu --timeout 400 -p no:cacheprovider > gpurun_out/b_pg.log 2>&1
echo "pg rc=$?" >> gpurun_out/b_pg.log
timeout 900 python -m pytest tests/test_cuda_collectives.py tests/test_cuda_round2.py -q --timeout 300 -p no:cacheprovider -k "broadcast or alltoall or nvl or local_op or skewed or small_allgather" > gpurun_out/b_fixed.log 2>&1
echo "fixed rc=$?" >> gpurun_out/b_fixed.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29533 -m gloo_b200.tune --out gpurun_out/tune_P2 > gpurun_out/b_tune2.log 2>&1
echo "tune rc=$?" >> gpurun_out/b_tune2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29534 bench.py --gpus 2 > gpurun_out/b_bench2.log 2>&1
echo "bench2 rc=$?" >> gpurun_out/b_bench2.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29535 bench.py --gpus 2 --impl reference > gpurun_out/b_bench2_ref.log 2>&1
echo "bench2 ref rc=$?" >> gpurun_out/b_bench2_ref.log
for f in b_multiproc b_pg b_fixed b_tune2; do echo "== $f"; tail -n 6 gpurun_out/$f.log; done
for f in b_bench2 b_bench2_ref; do echo "== $f"; tail -c 1500 gpurun_out/$f.log; done
