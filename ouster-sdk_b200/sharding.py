"""Multi-GPU plumbing of the path: independent sensor streams shard across ranks (stream i ->
GPU i mod G) with no data-path collective; the only exchange is a one-time broadcast
of each distinct LUT from rank 0 (torch.distributed: NCCL over NVLink on GPUs, gloo in CPU tests)."""
import numpy as np


def streams_of_rank(n_streams, world_size, rank):
    """Stream ids owned by `rank`: round-robin, every stream owned by exactly one rank."""
    if world_size <= 0 or not (0 <= rank < world_size):
        raise ValueError("invalid rank/world_size")
    return list(range(rank, n_streams, world_size))


def rank_of_stream(stream_id, world_size):
    return stream_id % world_size


def broadcast_lut(direction, offset, dist, src=0, device=None):
    """Broadcast the (h*w, 3) LUT tables from `src`; returns torch tensors on `device`.
    Non-source ranks pass arrays/tensors of the right shape and dtype (contents ignored)."""
    import torch
    td = torch.as_tensor(np.ascontiguousarray(direction) if isinstance(direction, np.ndarray) else direction)
    to = torch.as_tensor(np.ascontiguousarray(offset) if isinstance(offset, np.ndarray) else offset)
    if device is not None:
        td, to = td.to(device), to.to(device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.broadcast(td, src)
        dist.broadcast(to, src)
    return td, to


def max_over_ranks(value, dist, device=None):
    """MAX all-reduce of a python float (device-timed milliseconds) across ranks."""
    import torch
    t = torch.tensor([float(value)], dtype=torch.float64, device=device)
    if dist is not None and dist.is_initialized() and dist.get_world_size() > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())
