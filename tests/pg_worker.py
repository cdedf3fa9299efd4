"""One rank of the torch.distributed "glb" backend test: usage pg_worker.py INIT_FILE RANK SIZE."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402
import torch.distributed as dist  # noqa: E402

import gloo_b200.parallel.process_group  # noqa: E402,F401  (registers "glb")


def main():
    path, rank, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dist.init_process_group("glb", init_method=f"file://{path}", rank=rank, world_size=size)
    assert dist.get_backend() == "glb"
    tri = size * (size + 1) // 2

    t = torch.full((1000,), float(rank + 1))
    dist.all_reduce(t)
    assert t[0] == tri and t[-1] == tri
    t = torch.full((7,), float(rank + 1))
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    assert t[0] == size
    t = torch.full((5,), float(rank + 1), dtype=torch.bfloat16)
    dist.all_reduce(t, op=dist.ReduceOp.AVG)
    assert abs(float(t[0]) - tri / size) < 0.05
    nc = torch.arange(12, dtype=torch.float32).reshape(3, 4).t()  # non-contiguous
    assert not nc.is_contiguous()
    dist.all_reduce(nc)
    assert torch.equal(nc, (torch.arange(12, dtype=torch.float32).reshape(3, 4).t() * size))

    b = torch.arange(10) if rank == 1 % size else torch.zeros(10, dtype=torch.int64)
    dist.broadcast(b, src=1 % size)
    assert torch.equal(b, torch.arange(10))

    outs = [torch.zeros(3, dtype=torch.int32) for _ in range(size)]
    dist.all_gather(outs, torch.full((3,), rank, dtype=torch.int32))
    assert [int(o[0]) for o in outs] == list(range(size))
    flat = torch.zeros(size * 2)
    dist.all_gather_into_tensor(flat, torch.full((2,), float(rank)))
    assert flat.tolist() == [float(r) for r in range(size) for _ in range(2)]

    out = torch.zeros(4)
    dist.reduce_scatter(out, [torch.full((4,), float(rank + r)) for r in range(size)])
    assert out[0] == sum(q + rank for q in range(size))
    out = torch.zeros(2)
    dist.reduce_scatter_tensor(out, torch.arange(2 * size, dtype=torch.float32))
    assert out.tolist() == [size * (2 * rank), size * (2 * rank + 1)]

    inp = torch.arange(size * 2, dtype=torch.float32) + 100 * rank
    out = torch.zeros(size * 2)
    dist.all_to_all_single(out, inp)
    assert out.tolist() == [100.0 * r + 2 * rank + k for r in range(size) for k in range(2)]
    # uneven splits: rank r sends (q + 1) rows to rank q
    in_splits = [q + 1 for q in range(size)]
    out_splits = [rank + 1] * size
    inp = torch.full((sum(in_splits), 2), float(rank))
    out = torch.zeros(sum(out_splits), 2)
    dist.all_to_all_single(out, inp, out_splits, in_splits)
    assert out[:, 0].tolist() == [float(r) for r in range(size) for _ in range(rank + 1)]
    outs = [torch.zeros(2) for _ in range(size)]
    dist.all_to_all(outs, [torch.full((2,), float(rank * 10 + q)) for q in range(size)])
    assert [float(o[0]) for o in outs] == [float(r * 10 + rank) for r in range(size)]

    t = torch.full((3,), float(rank + 1))
    dist.reduce(t, dst=0)
    if rank == 0:
        assert t[0] == tri
    g = [torch.zeros(2) for _ in range(size)] if rank == 0 else None
    dist.gather(torch.full((2,), float(rank)), g, dst=0)
    if rank == 0:
        assert [float(x[0]) for x in g] == [float(r) for r in range(size)]
    s = torch.zeros(2)
    dist.scatter(s, [torch.full((2,), float(r * 3)) for r in range(size)] if rank == 0 else None, src=0)
    assert s[0] == rank * 3

    if size > 1:
        right, left = (rank + 1) % size, (rank - 1) % size
        got = torch.zeros(4)
        if rank % 2 == 0:
            dist.send(torch.full((4,), float(rank)), dst=right)
            dist.recv(got, src=left)
        else:
            dist.recv(got, src=left)
            dist.send(torch.full((4,), float(rank)), dst=right)
        assert got[0] == left
        reqs = [dist.isend(torch.full((2,), float(rank)), dst=right, tag=5), dist.irecv(got[:2], src=left, tag=5)]
        [r.wait() for r in reqs]
        assert got[0] == left
        # Every rank posts isend first, then irecv, with payloads above the single-copy
        # threshold (256 KiB): a synchronous isend would wait for the peer's recv forever.
        big = torch.full((200_000,), float(rank))  # 800 KB
        inbox = torch.zeros(200_000)
        reqs = [dist.isend(big, dst=right, tag=6), dist.irecv(inbox, src=left, tag=6)]
        [r.wait() for r in reqs]
        assert inbox[0] == left and inbox[-1] == left
        ops = [dist.P2POp(dist.isend, big, right, tag=7), dist.P2POp(dist.irecv, inbox, left, tag=7)]
        for r in dist.batch_isend_irecv(ops):
            r.wait()
        assert inbox[12345] == left
        strided = torch.zeros(4, 50_000).t()  # non-contiguous receive buffer
        reqs = [dist.isend(big, dst=right, tag=8), dist.irecv(strided, src=left, tag=8)]
        [r.wait() for r in reqs]
        assert strided[0, 0] == left and strided[-1, -1] == left

    if size > 1:  # recv without a source: whoever sends first
        if rank == 0:
            seen = set()
            for _ in range(size - 1):
                buf = torch.zeros(2)
                src = dist.recv(buf, tag=9)
                assert int(buf[0]) == src
                seen.add(src)
            assert seen == set(range(1, size))
        else:
            dist.send(torch.full((2,), float(rank)), dst=0, tag=9)

    # object collectives ride on byte / long tensors
    objs = [None] * size
    dist.all_gather_object(objs, {"rank": rank, "name": "r" * (rank + 1)})
    assert [o["rank"] for o in objs] == list(range(size)) and objs[-1]["name"] == "r" * size
    box = [{"cfg": 42}] if rank == 0 else [None]
    dist.broadcast_object_list(box, src=0)
    assert box[0] == {"cfg": 42}
    dist.barrier()

    # a sub-group gets its own context through a prefixed store
    if size >= 3:
        sub = dist.new_group(ranks=[0, 2])
        if rank in (0, 2):
            t = torch.full((4,), float(rank + 1))
            dist.all_reduce(t, group=sub)
            assert t[0] == 4
    dist.barrier()
    # DistributedDataParallel on CPU: gradients averaged through the backend
    torch.manual_seed(0)
    model = torch.nn.parallel.DistributedDataParallel(torch.nn.Linear(8, 4))
    x = torch.full((2, 8), float(rank + 1))
    model(x).sum().backward()
    gw = model.module.weight.grad.clone()
    ref = [torch.zeros_like(gw) for _ in range(size)]
    dist.all_gather(ref, gw)
    assert all(torch.allclose(ref[0], r) for r in ref)  # identical on every rank
    assert abs(float(gw[0, 0]) - 2.0 * tri / size) < 1e-5  # mean over ranks of d/dW sum(Wx+b) = sum_batch x
    dist.destroy_process_group()
    print(f"rank {rank} ok")


if __name__ == "__main__":
    main()
