"""CPU, world_size 2 over gloo: the host-side data-parallel logic (batch sharding, flat gradient
averaging, parameter broadcast)."""
import os
import socket

import torch
import torch.distributed as dist
import torch.multiprocessing as mp


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from frostnet_b200 import parallel
    flat = torch.full((1000,), float(rank + 1))
    parallel.average_gradients(flat)
    lo, hi = parallel.shard_batch(513)
    lin = torch.nn.Linear(4, 4)
    torch.manual_seed(rank)
    with torch.no_grad():
        lin.weight.normal_()
    parallel.broadcast_parameters(lin)
    gathered = [torch.zeros_like(lin.weight) for _ in range(world)]
    dist.all_gather(gathered, lin.weight.data)
    ok = bool(torch.equal(flat, torch.full((1000,), 1.5))) and torch.equal(gathered[0], gathered[1])
    out.put((rank, ok, lo, hi))
    dist.destroy_process_group()


def test_two_rank_gloo_gradient_average_and_sharding():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted(q.get(timeout=120) for _ in range(2))
    for p in procs:
        p.join(timeout=60)
    assert all(r[1] for r in res)
    assert (res[0][2], res[0][3]) == (0, 257) and (res[1][2], res[1][3]) == (257, 513)


def test_single_process_is_identity():
    from frostnet_b200 import parallel
    t = torch.ones(4)
    assert parallel.average_gradients(t) is t
    assert parallel.shard_batch(256, 0, 1) == (0, 256)
    assert parallel.shard_batch(2048, 7, 8) == (1792, 2048)
