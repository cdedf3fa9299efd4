"""One process per GPU: GradientBucketer with the bucket allreduce launched from the autograd hooks
(the overlapped path; threads-as-ranks on one GPU can only run the deferred one).
usage: bucketer_worker.py STORE_DIR RANK SIZE"""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

import gloo_b200 as gb  # noqa: E402
from gloo_b200.models import DDPMLP  # noqa: E402
from gloo_b200.ops import cuda as gcu  # noqa: E402
from gloo_b200.parallel import GradientBucketer  # noqa: E402


def main():
    store_dir, rank, size = sys.argv[1], int(sys.argv[2]), int(sys.argv[3])
    dev = rank % torch.cuda.device_count()
    torch.cuda.set_device(dev)
    ctx = gb.init_context(rank, size, path=store_dir, timeout_ms=60000)
    cc = gcu.CudaContext(ctx, dev, stage_bytes=32 << 20)
    torch.manual_seed(5)
    ref = DDPMLP(d_hidden=1024)
    x, y = torch.randn(size * 16, 64), torch.randn(size * 16, 8)
    single = DDPMLP(d_hidden=1024)
    single.load_state_dict(ref.state_dict())
    torch.nn.functional.mse_loss(single(x), y).backward()
    want = [p.grad.clone() for p in single.parameters()]
    m = DDPMLP(d_hidden=1024).cuda()
    m.load_state_dict(ref.state_dict())
    gbk = GradientBucketer(ctx, cc, m.parameters(), bucket_bytes=64 << 10)
    assert not gbk._defer and len(gbk.buckets) > 1
    xs, ys = x[rank * 16:(rank + 1) * 16].cuda(), y[rank * 16:(rank + 1) * 16].cuda()
    for step in range(3):
        torch.nn.functional.mse_loss(m(xs), ys).backward()
        launched_in_backward = sum(1 for b in gbk.buckets if b["launched"])
        gbk.finish()
        torch.cuda.synchronize()
        assert launched_in_backward == len(gbk.buckets), "every bucket should have been launched by its hooks"
        for p, w in zip(m.parameters(), want):
            torch.testing.assert_close(p.grad.cpu(), w, rtol=1e-4, atol=1e-5)
        gbk.zero_grad()
    cc.check_health()
    cc.pc.host_barrier()
    print(f"WORKER {rank} OK", flush=True)


if __name__ == "__main__":
    main()
