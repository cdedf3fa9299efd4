#!/bin/bash
# GPU call E (4 GPUs): first contact of the NVLS (multicast) code paths that need P > 2.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_cuda_multiproc.py -q -k vmm --timeout 500 -p no:cacheprovider > gpurun_out/e_multiproc4.log 2>&1
echo "multiproc4 rc=$?" >> gpurun_out/e_multiproc4.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 4 --master-addr 127.0.0.1 --master-port 29561 bench.py --gpus 4 --quick --steps 10 > gpurun_out/e_bench4q.log 2>&1
echo "bench4q rc=$?" >> gpurun_out/e_bench4q.log
tail -n 8 gpurun_out/e_multiproc4.log; tail -c 1800 gpurun_out/e_bench4q.log
