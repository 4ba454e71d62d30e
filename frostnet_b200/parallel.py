"""Data parallelism: one process per GPU, ONE all-reduce of the flat fp32 gradient per step.

The reference uses single-process ``torch.nn.DataParallel`` (Classification/train.py:89-92): replicas
compute independent batch shards with per-replica BN statistics / observer state, gradients are
summed onto device 0.  Here every rank holds the model, the engine writes all parameter gradients
into one flat buffer (5.81 M fp32 = 23.2 MB for FrostNet-L) and ``average_gradients`` all-reduces it
over NCCL/NVLink before the (redundant, bit-identical on every rank) GradBoost step.  BN running
stats and observer min/max stay per replica, like DataParallel's replicas within a step;
``broadcast_buffers`` restores rank 0's copy (DataParallel's end state) when asked.
"""
import torch
import torch.distributed as dist


def average_gradients(flat, group=None):
    """In-place mean over ranks of a flat gradient buffer (NCCL: ReduceOp.AVG; gloo: SUM then scale)."""
    if not (dist.is_available() and dist.is_initialized()):
        return flat
    ws = dist.get_world_size(group)
    if ws == 1:
        return flat
    if dist.get_backend(group) == "nccl":
        dist.all_reduce(flat, op=dist.ReduceOp.AVG, group=group)
    else:
        dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
        flat.mul_(1.0 / ws)
    return flat


def shard_batch(global_batch, rank=None, world_size=None):
    """[start, stop) of this rank's slice of a global batch (batch sharding is the only partitioning)."""
    if rank is None:
        rank = dist.get_rank() if dist.is_initialized() else 0
    if world_size is None:
        world_size = dist.get_world_size() if dist.is_initialized() else 1
    per = global_batch // world_size
    rem = global_batch % world_size
    start = rank * per + min(rank, rem)
    return start, start + per + (1 if rank < rem else 0)


def broadcast_parameters(model, src=0, group=None):
    """Make every rank start from rank `src`'s weights and buffers."""
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in list(model.parameters()) + list(model.buffers()):
        dist.broadcast(t.data, src, group=group)


def broadcast_buffers(model, src=0, group=None):
    if not (dist.is_available() and dist.is_initialized()) or dist.get_world_size(group) == 1:
        return
    for t in model.buffers():
        dist.broadcast(t.data, src, group=group)


def distribute(model, group=None):
    """Hook the prepared model's engine so that backward() ends with the gradient all-reduce."""
    eng = model.__dict__.get("_frost_engine")
    if eng is None:
        raise RuntimeError("distribute() needs a model prepared with frostnet_b200.prepare_qat")
    eng.grad_sync = lambda flat: average_gradients(flat, group)
    return model
