"""Multi-GPU plumbing for the paths that shard: one process per GPU over
torch.distributed.  MSM shards by contiguous point range (each rank holds its own generator slice and emits
one 96-byte partial point; partials are all-gathered and summed on every rank); batched MinRoot verification
shards by chain index with no collective.  R1CS / fold kernels and the fold step itself are replicas only."""
from __future__ import annotations

from typing import Callable, List, Optional, Tuple


def shard_range(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Contiguous [first, first+count) of rank's share of n items; sizes differ by at most one."""
    if not (0 <= rank < world):
        raise ValueError("rank out of range")
    base, rem = divmod(n, world)
    first = rank * base + min(rank, rem)
    return first, base + (1 if rank < rem else 0)


def all_gather_bytes(payload: bytes, group=None) -> List[bytes]:
    """All-gather equal-length byte strings (the 96-byte partial points) over the default process group:
    NCCL on GPUs (tensor on the current CUDA device), gloo on CPU."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    backend = dist.get_backend(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if backend == "nccl" else torch.device("cpu")
    src = torch.frombuffer(bytearray(payload), dtype=torch.uint8).to(dev)
    out = torch.empty(world * len(payload), dtype=torch.uint8, device=dev)
    dist.all_gather_into_tensor(out, src, group=group)
    raw = out.cpu().numpy().tobytes()
    return [raw[k * len(payload):(k + 1) * len(payload)] for k in range(world)]


def sharded_commit(commit_shard: Callable[[], bytes], combine: Callable[[List[bytes]], bytes],
                   group=None) -> bytes:
    """commit_shard() -> this rank's 96-byte partial; combine(partials) -> their sum.  On GPUs:
    commit_shard = gens.commit_bytes(scalars) and combine = lambda ps: msm.point_sum(curve, b''.join(ps))."""
    import torch.distributed as dist
    part = commit_shard()
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        return part
    return combine(all_gather_bytes(part, group))


def shard_chains(n: int, rank: int, world: int) -> Tuple[int, int]:
    """Batched MinRoot verification: chains are independent, shard by index, no collective."""
    return shard_range(n, rank, world)


def sharded_verify(check_shard: Callable[[int, int], bytes], n: int, group=None) -> bytes:
    """Batched MinRoot verification of n independent (result, t, original) triples over the process group:
    check_shard(first, count) -> one verdict byte per chain of this rank's slice (on GPUs:
    PallasVDF.check_batch on results[first:first+count]); the verdict bytes are all-gathered so that every rank
    returns the n verdicts in chain order.  The checks themselves need no exchange."""
    import torch.distributed as dist
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out = check_shard(0, n)
        if len(out) != n:
            raise ValueError("check_shard returned the wrong number of verdicts")
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    first, count = shard_range(n, rank, world)
    mine = check_shard(first, count)
    if len(mine) != count:
        raise ValueError("check_shard returned the wrong number of verdicts")
    width = (n + world - 1) // world                     # slices differ by at most one chain: pad to equal length
    parts = all_gather_bytes(mine + bytes(width - count), group)
    return b"".join(parts[r][:shard_range(n, r, world)[1]] for r in range(world))
