"""Multi-GPU plumbing.  Agents are independent given the (replicated) environment and
cell parameters, so the step path needs NO collective: every rank steps a contiguous
shard of the global agent range with ``id_offset`` = its first global id (the Philox
streams are keyed on global ids, so results do not depend on the number of ranks).
The only communication is the optional gather of history slabs to one rank
(NCCL over NVLink on GPUs; gloo in the CPU tests)."""
import numpy as np


def shard_range(n_total, rank, world):
    """Contiguous [start, stop) of the agents owned by ``rank`` (sizes differ by at most 1)."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(int(n_total), int(world))
    start = rank * base + min(rank, rem)
    return start, start + base + (1 if rank < rem else 0)


def agent_params_for_rank(params, n_total, rank, world):
    """Agent params of this rank's shard: n_agents and id_offset filled in."""
    start, stop = shard_range(n_total, rank, world)
    p = dict(params)
    p["n_agents"], p["id_offset"] = stop - start, start
    return p


def gather_agent_axis(x, n_total, axis=0, dst=0, group=None):
    """Gather per-rank shards (split along ``axis`` by shard_range) on rank ``dst``.
    ``x``: torch tensor (CUDA with the nccl backend, CPU with gloo) or NumPy array.
    Returns the full tensor on ``dst`` and None elsewhere."""
    import torch
    import torch.distributed as dist
    was_np = isinstance(x, np.ndarray)
    t = torch.as_tensor(x) if was_np else x
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    t = t.movedim(axis, 0).contiguous()
    sizes = [shard_range(n_total, r, world)[1] - shard_range(n_total, r, world)[0] for r in range(world)]
    assert t.shape[0] == sizes[rank], "local shard size does not match shard_range"
    mx = max(sizes)
    pad = torch.zeros((mx,) + tuple(t.shape[1:]), dtype=t.dtype, device=t.device)
    pad[: t.shape[0]] = t
    bufs = [torch.empty_like(pad) for _ in range(world)]
    dist.all_gather(bufs, pad, group=group)          # one collective; history only, never on the step path
    if rank != dst:
        return None
    full = torch.cat([b[:s] for b, s in zip(bufs, sizes)], dim=0).movedim(0, axis)
    return full.cpu().numpy() if was_np else full
