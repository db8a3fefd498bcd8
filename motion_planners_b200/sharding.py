"""Host-side plumbing for more than one GPU (one process per GPU, torch.distributed for the rendezvous).

The data path has exactly two exchanges per iteration in rollout-sharded mode, both inside the C
library's weights_update_peer_kernel as tagged 8-byte words written into peer-mapped mailboxes over NVLink:
  A. the min / max of the rank's own cumulative costs per joint                [D][2] per rank
  B. the rank's unnormalised partial sums (update row, adaptation numerator, sum of weights)   [D][T+2] per rank
(with STOMP_B200_EXCHANGE=nccl, or where the mailboxes cannot be mapped: an NCCL all-gather of the per-rollout cost
scalars [K/G][1+3D] and an all-reduce of [D][T+2]),
and none in query-sharded mode.  This module holds what the *host* has to get right: the partition of
rollouts / queries over ranks, the broadcast of the NCCL unique id, and max-over-ranks timing.  The
functions take any initialised torch.distributed backend (nccl on GPUs, gloo in the CPU tests).
"""
from __future__ import annotations

import numpy as np


def rollout_shard(num_rollouts: int, world_size: int, rank: int):
    """(offset, count) of the generated rollouts rank owns.  Sharding needs K divisible by the world size
    (the library returns STOMP_B200_ERR_UNSUPPORTED otherwise)."""
    if num_rollouts % world_size != 0:
        raise ValueError(f"{num_rollouts} rollouts do not split evenly over {world_size} ranks")
    count = num_rollouts // world_size
    return rank * count, count


def query_shard(num_queries: int, world_size: int, rank: int):
    """(offset, count) of the queries rank owns: balanced contiguous blocks — floor(Q / G) each, the first Q mod G ranks
    one more (mirrors stomp_b200_create, csrc/engine.cu)."""
    per, rem = divmod(num_queries, world_size)
    return rank * per + min(rank, rem), per + (1 if rank < rem else 0)


def global_slot(local_slot: int, num_generated_local: int, rank: int, world_size: int, has_noiseless: bool):
    """Slot of a local rollout in the rollout-indexed tables every rank holds in full (sums, probabilities):
    generated rollouts are laid out rank after rank, the noise-less rollout comes last."""
    if has_noiseless and local_slot == num_generated_local:
        return num_generated_local * world_size
    if not 0 <= local_slot < num_generated_local:
        raise IndexError(local_slot)
    return rank * num_generated_local + local_slot


def broadcast_bytes(dist, payload: bytes, nbytes: int, device="cpu") -> bytes:
    """Rank 0's payload on every rank (the 128-byte NCCL unique id)."""
    import torch
    t = torch.zeros(nbytes, dtype=torch.uint8, device=device)
    if dist.get_rank() == 0:
        t.copy_(torch.frombuffer(bytearray(payload.ljust(nbytes, b"\0")), dtype=torch.uint8))
    dist.broadcast(t, 0)
    return bytes(t.cpu().numpy().tobytes())


def max_over_ranks(dist, value: float, device="cpu") -> float:
    import torch
    t = torch.tensor([value], dtype=torch.float64, device=device)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def gather_cost_scalars(dist, local_sums: np.ndarray, device="cpu") -> np.ndarray:
    """Exchange 1 as the host would do it (the library does the same with ncclAllGather): [K/G][W] -> [K][W]."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(local_sums)).to(device)
    out = [torch.empty_like(t) for _ in range(dist.get_world_size())]
    dist.all_gather(out, t)
    return torch.cat(out, 0).cpu().numpy()


def reduce_update(dist, local_update: np.ndarray, device="cpu") -> np.ndarray:
    """Exchange 2: sum of the per-rank partial updates [D][T+1]."""
    import torch
    t = torch.from_numpy(np.ascontiguousarray(local_update)).to(device)
    dist.all_reduce(t, op=dist.ReduceOp.SUM)
    return t.cpu().numpy()
