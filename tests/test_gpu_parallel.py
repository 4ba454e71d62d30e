"""Two ranks over NCCL (skipped on a single-GPU box): the ONE exchange step of the data-parallel path - the all-reduce
of the flat gradient - and the rank-independent GradBoost noise (reference: nn.DataParallel,
Classification/train.py:89-92)."""
import os
import socket

import pytest
import torch

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    import torch.distributed as dist
    import torch.nn.functional as Fn
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    torch.cuda.set_device(rank)
    dev = torch.device("cuda", rank)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    try:
        import frostnet_b200 as F
        torch.manual_seed(7)                                   # same initial weights on every rank
        model = F.FrostNet(nclass=16, mode="small", width_mult=0.35, quantized=True, drop_rate=0.0)
        model.train()
        model.fuse_model()
        F.prepare_qat(model)
        model.to(dev)
        F.parallel.broadcast_parameters(model)
        g = torch.Generator().manual_seed(100 + rank)          # different data per rank
        x = torch.randn(8, 3, 64, 64, generator=g).to(dev)
        y = torch.randint(0, 16, (8,), generator=g).to(dev)
        sd = {k: v.clone() for k, v in model.state_dict().items()}
        # (a) local gradient, no exchange
        Fn.cross_entropy(model(x), y).backward()
        local = torch.cat([p.grad.flatten() for p in model.parameters()]).clone()
        gathered = [torch.empty_like(local) for _ in range(world)]
        dist.all_gather(gathered, local)
        mean = torch.stack(gathered).double().mean(0)
        # (b) the same step with the all-reduce hooked into backward
        model.load_state_dict(sd)
        model.zero_grad()
        F.parallel.distribute(model)
        Fn.cross_entropy(model(x), y).backward()
        synced = torch.cat([p.grad.flatten() for p in model.parameters()])
        err = float((synced.double() - mean).abs().max() / mean.abs().max())
        # (c) three GradBoost steps with the noise on: weights stay bit-identical across ranks
        opt = F.QSGD(model.parameters(), lr=5e-3, momentum=0.9, nesterov=True, weight_decay=1e-5)
        opt.is_warmup = False
        for _ in range(3):
            opt.zero_grad()
            Fn.cross_entropy(model(x), y).backward()
            opt.step()
        w = torch.cat([p.detach().flatten() for p in model.parameters()])
        ws = [torch.empty_like(w) for _ in range(world)]
        dist.all_gather(ws, w)
        same = all(torch.equal(ws[0], t) for t in ws[1:])
        moved = float((w - torch.cat([sd[k].flatten() for k, _ in model.named_parameters()]).to(dev)).abs().max()) > 0
        out.put((rank, err, same, moved))
    finally:
        dist.destroy_process_group()


@pytest.mark.skipif(torch.cuda.device_count() < 2, reason="needs 2 GPUs")
def test_two_rank_nccl_gradient_mean_and_identical_weights():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=300) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    for rank, err, same, moved in res:
        assert err <= 1e-6, "rank %d: all-reduced gradient differs from the mean of the local gradients by %g" % (rank, err)
        assert same, "rank %d: weights diverged across ranks after 3 GradBoost steps" % rank
        assert moved
