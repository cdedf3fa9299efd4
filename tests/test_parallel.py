"""parallel/ strategies on the host path (CPU tensors over TCP)."""
import pytest
import torch

import gloo_b200 as gb
from gloo_b200.models import DDPMLP, train_step
from gloo_b200.parallel import (DataParallel, GradientBucketer, MoEDispatcher, RingExchange, TensorParallel, UlyssesAttention,
                                ZeroShard)


def test_ddp_matches_single_process():
    size = 4
    torch.manual_seed(0)
    ref = DDPMLP()
    x = torch.randn(size * 8, 64)
    y = torch.randn(size * 8, 8)
    # single process, full batch
    single = DDPMLP()
    single.load_state_dict(ref.state_dict())
    loss = torch.nn.functional.mse_loss(single(x), y)
    loss.backward()
    want = [p.grad.clone() for p in single.parameters()]

    def fn(ctx):
        m = DDPMLP()
        if ctx.rank == 0:
            m.load_state_dict(ref.state_dict())
        dp = DataParallel(ctx, bucket_bytes=16 << 10)
        dp.broadcast_parameters(m.parameters())
        xs, ys = x[ctx.rank * 8:(ctx.rank + 1) * 8], y[ctx.rank * 8:(ctx.rank + 1) * 8]
        train_step(m, dp, xs, ys, lr=0.0)
        return [p.grad.clone() for p in m.parameters()]

    for grads in gb.spawn_threads(size, fn):
        for g, w in zip(grads, want):
            torch.testing.assert_close(g, w, rtol=1e-4, atol=1e-5)


def test_gradient_bucketer_views_and_overlap_order():
    """Gradients are views into the flat buckets, buckets are reduced in order while backward runs, two
    steps (zero_grad keeps the views; a rebound .grad is adopted), a parameter without gradient."""
    size = 3
    torch.manual_seed(1)
    ref = DDPMLP()
    x = torch.randn(size * 4, 64)
    y = torch.randn(size * 4, 8)
    single = DDPMLP()
    single.load_state_dict(ref.state_dict())
    torch.nn.functional.mse_loss(single(x), y).backward()
    want = [p.grad.clone() for p in single.parameters()]

    def fn(ctx):
        m = DDPMLP()
        m.load_state_dict(ref.state_dict())
        unused = torch.nn.Parameter(torch.ones(5))          # never part of the loss: no hook fires for it
        params = list(m.parameters()) + [unused]
        gbk = GradientBucketer(ctx, None, params, bucket_bytes=4 << 10)
        assert len(gbk.buckets) > 2
        flat0 = gbk.buckets[0]["flat"]
        assert all(p.grad is not None and p.grad.untyped_storage().data_ptr() in
                   {b["flat"].untyped_storage().data_ptr() for b in gbk.buckets} for p in params)
        xs, ys = x[ctx.rank * 4:(ctx.rank + 1) * 4], y[ctx.rank * 4:(ctx.rank + 1) * 4]
        out = []
        for step in range(2):
            loss = torch.nn.functional.mse_loss(m(xs), ys)
            if step == 1:
                params[0].grad = None                          # what zero_grad(set_to_none=True) does
            loss.backward()
            gbk.finish()
            out.append([p.grad.clone() for p in m.parameters()])
            assert float(unused.grad.abs().sum()) == 0.0
            assert gbk.buckets[0]["flat"] is flat0 and params[0].grad.untyped_storage().data_ptr() in \
                {b["flat"].untyped_storage().data_ptr() for b in gbk.buckets}
            gbk.zero_grad()
            assert all(float(p.grad.abs().sum()) == 0.0 for p in params)
        gbk.remove()
        return out

    for steps in gb.spawn_threads(size, fn):
        for grads in steps:
            for g, w in zip(grads, want):
                # mean over ranks of per-rank mean losses == full-batch mean loss (equal shards)
                torch.testing.assert_close(g, w, rtol=1e-4, atol=1e-5)


def test_zero_shard_roundtrip():
    size = 3

    def fn(ctx):
        z = ZeroShard(ctx)
        n = 1000
        start, cnt, counts = z.shard_range(n)
        g = torch.arange(n, dtype=torch.float32) + ctx.rank
        shard = torch.zeros(cnt)
        z.reduce_scatter_gradients(g, shard, average=False)
        exp = (torch.arange(n, dtype=torch.float32) * size + sum(range(size)))[start:start + cnt]
        torch.testing.assert_close(shard, exp)
        full = torch.zeros(n)
        z.allgather_parameters(full, shard)
        torch.testing.assert_close(full, torch.arange(n, dtype=torch.float32) * size + sum(range(size)))
        return True

    assert all(gb.spawn_threads(size, fn))


def test_moe_dispatch_combine():
    size = 4

    def fn(ctx):
        moe = MoEDispatcher(ctx)
        width = 8
        send = [(ctx.rank + j) % 3 + 1 for j in range(size)]
        rows = sum(send)
        tokens = torch.cat([torch.full((send[j], width), float(ctx.rank * 10 + j)) for j in range(size)])
        recv = moe.exchange_counts(send)
        assert recv == [(j + ctx.rank) % 3 + 1 for j in range(size)]
        out = torch.zeros(sum(recv), width)
        moe.dispatch(tokens, send, out, recv)
        exp = torch.cat([torch.full((recv[j], width), float(j * 10 + ctx.rank)) for j in range(size)])
        torch.testing.assert_close(out, exp)
        back = torch.zeros(rows, width)
        moe.combine(out, recv, back, send)
        torch.testing.assert_close(back, tokens)
        return True

    assert all(gb.spawn_threads(size, fn))


def test_tp_ulysses_ring():
    size = 4

    def fn(ctx):
        tp = TensorParallel(ctx)
        part = torch.full((16,), float(ctx.rank + 1))
        tp.row_parallel_output(part)
        assert float(part[0]) == 10.0
        out = torch.zeros(size * 4)
        tp.column_parallel_gather(torch.full((4,), float(ctx.rank)), out)
        torch.testing.assert_close(out, torch.arange(size).repeat_interleave(4).float())
        u = UlyssesAttention(ctx)
        x = torch.arange(size * 6, dtype=torch.float32) + 100 * ctx.rank
        y = torch.zeros_like(x)
        u.seq_to_heads(x, y)
        exp = torch.cat([torch.arange(ctx.rank * 6, ctx.rank * 6 + 6, dtype=torch.float32) + 100 * j for j in range(size)])
        torch.testing.assert_close(y, exp)
        ring = RingExchange(ctx)
        kv = torch.full((32,), float(ctx.rank))
        nxt = torch.zeros(32)
        for step in range(1, size):
            ring.rotate(kv, nxt, step)
            assert float(nxt[0]) == float((ctx.rank - step) % size)
        return True

    assert all(gb.spawn_threads(size, fn))


def test_affinity_helpers_are_best_effort():
    """CPU placement never raises: without a GPU (or on a single-node box) it leaves the
    affinity untouched and reports what is in effect."""
    import os

    from gloo_b200.utils.affinity import _parse_cpulist, bind_to_gpu

    assert _parse_cpulist("0-3,8,10-11\n") == [0, 1, 2, 3, 8, 10, 11]
    assert _parse_cpulist("") == []
    before = os.sched_getaffinity(0)
    got = bind_to_gpu(0)
    assert isinstance(got, list)
    assert os.sched_getaffinity(0) == before or set(got) <= before


def test_training_example_runs_on_two_processes():
    """examples/example_training.py end to end (CPU tensors here): rendezvous, parameter broadcast,
    GradientBucketer, optimizer steps; the loss must go down."""
    import os
    import re
    import subprocess
    import sys

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    env = dict(os.environ, STEPS="12", CUDA_VISIBLE_DEVICES="")
    out = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2",
                          "--master-addr", "127.0.0.1", "--master-port", "29613",
                          os.path.join(root, "examples", "example_training.py")],
                         capture_output=True, text=True, timeout=240, env=env)
    assert out.returncode == 0, out.stdout[-2000:] + out.stderr[-2000:]
    m = re.search(r"loss ([0-9.]+) -> ([0-9.]+) over 12 steps on 2 rank", out.stdout + out.stderr)
    assert m and float(m.group(2)) < float(m.group(1)), out.stdout[-500:]


def test_zero_optimizer_matches_replicated_adam():
    """ZeRO-1: reduce_scatter of the flat gradient, AdamW on the local shard, allgather of the parameters ==
    AdamW on the full batch in one process; optimizer state is 1/P of the replicated one."""
    from gloo_b200.parallel import ZeroOptimizer

    size = 3
    torch.manual_seed(7)
    ref = DDPMLP()
    xs_all = [torch.randn(size * 4, 64) for _ in range(3)]
    ys_all = [torch.randn(size * 4, 8) for _ in range(3)]
    single = DDPMLP()
    single.load_state_dict(ref.state_dict())
    opt = torch.optim.AdamW(single.parameters(), lr=1e-2)
    for x, y in zip(xs_all, ys_all):
        opt.zero_grad()
        torch.nn.functional.mse_loss(single(x), y).backward()
        opt.step()
    want = [p.detach().clone() for p in single.parameters()]
    full_state = sum(v.numel() * v.element_size() for st in opt.state.values() for v in st.values() if isinstance(v, torch.Tensor) and v.numel() > 1)

    def fn(ctx):
        m = DDPMLP()
        m.load_state_dict(ref.state_dict())
        zo = ZeroOptimizer(ctx, None, m.parameters(), torch.optim.AdamW, lr=1e-2)
        for step, (x, y) in enumerate(zip(xs_all, ys_all)):
            loss = torch.nn.functional.mse_loss(m(x[ctx.rank * 4:(ctx.rank + 1) * 4]), y[ctx.rank * 4:(ctx.rank + 1) * 4])
            if step == 1:
                next(m.parameters()).grad = None      # zero_grad(set_to_none=True) on the model
            loss.backward()
            zo.step()
            zo.zero_grad()
        return [p.detach().clone() for p in m.parameters()], zo.state_bytes()

    for params, state in gb.spawn_threads(size, fn):
        for p, w in zip(params, want):
            torch.testing.assert_close(p, w, rtol=2e-4, atol=2e-5)
        assert state <= full_state / size * 1.05 + 64


def test_tensor_parallel_mlp_matches_dense():
    """Column-parallel -> GELU -> row-parallel == the dense MLP: forward output, input gradient and the
    gradients of the weight shards (one allreduce in forward, one in backward)."""
    size = 4
    torch.manual_seed(11)
    dense1, dense2 = torch.nn.Linear(32, 64), torch.nn.Linear(64, 16)
    x = torch.randn(6, 32, requires_grad=True)
    target = torch.randn(6, 16)
    torch.nn.functional.mse_loss(dense2(torch.nn.functional.gelu(dense1(x))), target).backward()
    want_out = dense2(torch.nn.functional.gelu(dense1(x))).detach()

    def fn(ctx):
        tp = TensorParallel(ctx)
        col, row = tp.column_linear(32, 64), tp.row_linear(64, 16)
        col.load_full(dense1.weight.detach(), dense1.bias.detach())
        row.load_full(dense2.weight.detach(), dense2.bias.detach())
        xi = x.detach().clone().requires_grad_(True)
        out = row(torch.nn.functional.gelu(col(xi)))
        torch.nn.functional.mse_loss(out, target).backward()
        return (out.detach(), xi.grad, col.weight.grad, col.bias.grad, row.weight.grad, row.bias.grad, (col.lo, col.hi), (row.lo, row.hi))

    for out, gx, gcw, gcb, grw, grb, (clo, chi), (rlo, rhi) in gb.spawn_threads(size, fn):
        torch.testing.assert_close(out, want_out, rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(gx, x.grad, rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(gcw, dense1.weight.grad[clo:chi], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(gcb, dense1.bias.grad[clo:chi], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(grw, dense2.weight.grad[:, rlo:rhi], rtol=1e-4, atol=1e-6)
        torch.testing.assert_close(grb, dense2.bias.grad, rtol=1e-4, atol=1e-6)


def test_ring_attention_matches_full_attention():
    """RingExchange.attention: K/V blocks rotate around the ring, online softmax per block == attention over
    the whole sequence (plain and causal)."""
    size, S, Hh, D = 4, 64, 3, 16
    torch.manual_seed(13)
    q, k, v = (torch.randn(S, Hh, D) for _ in range(3))

    def full(causal):
        qt, kt, vt = (t.transpose(0, 1) for t in (q, k, v))                # [H, S, D]
        return torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=causal).transpose(0, 1)

    def fn(ctx):
        ring = RingExchange(ctx)
        lo, hi = ctx.rank * S // size, (ctx.rank + 1) * S // size
        return [ring.attention(q[lo:hi], k[lo:hi], v[lo:hi], causal=c) for c in (False, True)], (lo, hi)

    for (plain, causal), (lo, hi) in gb.spawn_threads(size, fn):
        torch.testing.assert_close(plain, full(False)[lo:hi], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(causal, full(True)[lo:hi], rtol=1e-4, atol=1e-5)


def test_ulysses_attention_matches_full_attention():
    size, S, Hh, D = 4, 32, 8, 16
    torch.manual_seed(17)
    q, k, v = (torch.randn(S, Hh, D) for _ in range(3))

    def full(causal):
        qt, kt, vt = (t.transpose(0, 1) for t in (q, k, v))
        return torch.nn.functional.scaled_dot_product_attention(qt, kt, vt, is_causal=causal).transpose(0, 1)

    def fn(ctx):
        u = UlyssesAttention(ctx)
        lo, hi = ctx.rank * S // size, (ctx.rank + 1) * S // size
        return [u.attention(q[lo:hi].contiguous(), k[lo:hi].contiguous(), v[lo:hi].contiguous(), causal=c) for c in (False, True)], (lo, hi)

    for (plain, causal), (lo, hi) in gb.spawn_threads(size, fn):
        torch.testing.assert_close(plain, full(False)[lo:hi], rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(causal, full(True)[lo:hi], rtol=1e-4, atol=1e-5)


def test_tensor_parallel_transformer_block_matches_dense():
    """models.TPBlock (heads and MLP columns split over the ranks, two allreduces per forward) == DenseBlock:
    output, input gradient, and every weight-shard gradient."""
    from gloo_b200.models import DenseBlock, TPBlock

    size = 2
    torch.manual_seed(19)
    dense = DenseBlock(d_model=32, n_heads=4, d_ff=64)
    x = torch.randn(2, 10, 32, requires_grad=True)
    out_d = dense(x)
    out_d.square().mean().backward()

    def fn(ctx):
        tp = TensorParallel(ctx)
        blk = TPBlock(tp, d_model=32, n_heads=4, d_ff=64)
        blk.load_dense(dense)
        xi = x.detach().clone().requires_grad_(True)
        out = blk(xi)
        out.square().mean().backward()
        return out.detach(), xi.grad, blk.up.weight.grad, (blk.up.lo, blk.up.hi), blk.proj.weight.grad, (blk.proj.lo, blk.proj.hi), blk.ln1.weight.grad

    for out, gx, gup, (ulo, uhi), gproj, (plo, phi), gln in gb.spawn_threads(size, fn):
        torch.testing.assert_close(out, out_d.detach(), rtol=1e-4, atol=1e-5)
        torch.testing.assert_close(gx, x.grad, rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(gup, dense.up.weight.grad[ulo:uhi], rtol=1e-3, atol=1e-6)
        torch.testing.assert_close(gproj, dense.proj.weight.grad[:, plo:phi], rtol=1e-3, atol=1e-6)
        # replicated parameters (layer norms) see the full gradient on every rank: their input gradient is
        # summed by f, their own gradient is computed from replicated activations
        torch.testing.assert_close(gln, dense.ln1.weight.grad, rtol=1e-3, atol=1e-6)


def test_pipeline_parallel_matches_sequential():
    """3 stages x 4 micro-batches (GPipe): parameter gradients == one process running the whole model on the
    whole batch; losses come back on the last stage."""
    from gloo_b200.parallel import PipelineParallel

    size, n_mb, mb = 3, 4, 5
    torch.manual_seed(23)
    stages = [torch.nn.Sequential(torch.nn.Linear(16, 32), torch.nn.Tanh()),
              torch.nn.Sequential(torch.nn.Linear(32, 32), torch.nn.Tanh()),
              torch.nn.Linear(32, 4)]
    x, y = torch.randn(n_mb * mb, 16), torch.randn(n_mb * mb, 4)
    seq = torch.nn.Sequential(*[torch.nn.Sequential(*s) if isinstance(s, torch.nn.Sequential) else s for s in stages])
    want_losses = []
    for i in range(n_mb):
        l = torch.nn.functional.mse_loss(seq(x[i * mb:(i + 1) * mb]), y[i * mb:(i + 1) * mb]) / n_mb
        l.backward()
        want_losses.append(float(l.detach()))
    want = [[p.grad.clone() for p in s.parameters()] for s in stages]
    for s in stages:
        s.zero_grad()

    def fn(ctx):
        import copy

        stage = copy.deepcopy(stages[ctx.rank])
        pp = PipelineParallel(ctx)
        mbs = [x[i * mb:(i + 1) * mb] for i in range(n_mb)]
        losses = pp.run(stage, mbs, in_shape=(mb, 16 if ctx.rank == 0 else 32), dtype=torch.float32,
                        loss_fn=lambda o, t: torch.nn.functional.mse_loss(o, t) / n_mb,
                        targets=[y[i * mb:(i + 1) * mb] for i in range(n_mb)])
        return [p.grad.clone() for p in stage.parameters()], losses

    res = gb.spawn_threads(size, fn)
    for r, (grads, losses) in enumerate(res):
        for g, w in zip(grads, want[r]):
            torch.testing.assert_close(g, w, rtol=1e-4, atol=1e-6)
        assert (losses is None) == (r != size - 1)
    assert res[-1][1] == pytest.approx(want_losses, rel=1e-5)
