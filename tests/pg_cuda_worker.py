"""One rank of the "glb" backend on CUDA tensors: usage pg_cuda_worker.py INIT_FILE RANK SIZE.
Rank r uses cuda:(r % device_count)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
os.environ.setdefault("CUDA_MODULE_LOADING", "EAGER")

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gloo_b200.parallel.process_group  # noqa: E402,F401


def main():
    path, rank, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dev = torch.device("cuda", rank % torch.cuda.device_count())
    torch.cuda.set_device(dev)
    dist.init_process_group("glb", init_method=f"file://{path}", rank=rank, world_size=size)
    tri = size * (size + 1) // 2

    for n in (5, 4099, 1 << 20):
        t = torch.full((n,), float(rank + 1), device=dev)
        dist.all_reduce(t)
        assert float(t[0]) == tri and float(t[-1]) == tri, (n, float(t[0]))
    t = torch.full((257,), float(rank + 1), device=dev, dtype=torch.bfloat16)
    dist.all_reduce(t, op=dist.ReduceOp.AVG)
    assert abs(float(t[0]) - tri / size) < 0.05
    b = torch.arange(1000, device=dev) if rank == size - 1 else torch.zeros(1000, dtype=torch.int64, device=dev)
    dist.broadcast(b, src=size - 1)
    assert torch.equal(b.cpu(), torch.arange(1000))
    flat = torch.zeros(size * 3, device=dev)
    dist.all_gather_into_tensor(flat, torch.full((3,), float(rank), device=dev))
    assert flat.cpu().tolist() == [float(r) for r in range(size) for _ in range(3)]
    outs = [torch.zeros(2, device=dev) for _ in range(size)]
    dist.all_gather(outs, torch.full((2,), float(rank), device=dev))
    assert [float(o[0]) for o in outs] == [float(r) for r in range(size)]
    out = torch.zeros(4, device=dev)
    dist.reduce_scatter_tensor(out, torch.arange(4 * size, dtype=torch.float32, device=dev))
    assert out.cpu().tolist() == [float(size * (4 * rank + k)) for k in range(4)]
    inp = torch.arange(size * 2, dtype=torch.float32, device=dev) + 100 * rank
    out = torch.zeros(size * 2, device=dev)
    dist.all_to_all_single(out, inp)
    assert out.cpu().tolist() == [100.0 * r + 2 * rank + k for r in range(size) for k in range(2)]
    t = torch.full((3,), float(rank + 1), device=dev)
    dist.reduce(t, dst=0)
    if rank == 0:
        assert float(t[0]) == tri
    if size > 1:
        got = torch.zeros(4, device=dev)
        right, left = (rank + 1) % size, (rank - 1) % size
        if rank % 2 == 0:
            dist.send(torch.full((4,), float(rank), device=dev), dst=right)
            dist.recv(got, src=left)
        else:
            dist.recv(got, src=left)
            dist.send(torch.full((4,), float(rank), device=dev), dst=right)
        assert float(got[0]) == left
        # NVLink p2p: everybody posts isend then irecv, payload larger than the mailbox ring
        big = torch.full((3_000_000,), float(rank), device=dev)
        inbox = torch.zeros(3_000_000, device=dev)
        reqs = [dist.isend(big, dst=right), dist.irecv(inbox, src=left)]
        [r.wait() for r in reqs]
        torch.cuda.synchronize()
        assert float(inbox[0]) == left and float(inbox[-1]) == left
        ops = [dist.P2POp(dist.isend, big, right), dist.P2POp(dist.irecv, inbox, left)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        torch.cuda.synchronize()
        assert float(inbox[12345]) == left
    avg = torch.full((1 << 20,), float(rank + 1), device=dev)
    dist.all_reduce(avg, op=dist.ReduceOp.AVG)   # fused scale epilogue, pipelined plain-pointer kernel
    assert abs(float(avg[-1]) - tri / size) < 1e-5
    dist.barrier()
    torch.manual_seed(0)
    model = torch.nn.parallel.DistributedDataParallel(torch.nn.Linear(8, 4).to(dev), device_ids=[dev.index])
    model(torch.full((2, 8), float(rank + 1), device=dev)).sum().backward()
    gw = model.module.weight.grad
    assert abs(float(gw[0, 0]) - 2.0 * tri / size) < 1e-4, float(gw[0, 0])
    torch.cuda.synchronize()
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
