"""Host-side logic of the multi-GPU paths (no CUDA calls here, so it is testable on CPU with gloo).

* Sampling shards the sample batch across ranks: every sample is independent in eval mode
  (diffusion.py:256), weights are replicated, there is no data-path collective.
* Training is data parallel with rank-local BatchNorm statistics (what DistributedDataParallel
  around the reference module would do): parameter gradients live in one flat fp32 buffer, cut into
  contiguous buckets; each bucket is all-reduced (sum) as soon as the backward-plan entry that
  finalises its last gradient has been enqueued, and the fused Adam pass scales by 1/world_size.
"""
from __future__ import annotations

from typing import List, Tuple

import torch


def shard_range(n: int, world: int, rank: int) -> Tuple[int, int]:
    """Contiguous [lo, hi) slice of ``n`` samples owned by ``rank`` (sizes differ by at most one)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def grad_ready_index(param_names: List[str], bwd_op_names: List[str]) -> List[int]:
    """For every parameter (state_dict key) the index of the backward-plan entry after which its
    gradient is final (-1: never written, stays zero -- conv biases in front of a train-mode BN)."""
    pos = {n: i for i, n in enumerate(bwd_op_names)}
    last = pos.get("embed:bwd", len(bwd_op_names) - 1)
    out = []
    for k in param_names:
        mod, _, leaf = k.rpartition(".")
        if mod in ("final_conv", "initial_conv"):
            out.append(pos[f"{mod}:{'wgrad' if leaf == 'weight' else 'dbias'}"])
            continue
        parts = mod.split(".")
        if len(parts) == 2 and parts[1].isdigit():
            blk, idx = parts[0], int(parts[1])
            if f"{blk}.{idx}:wgrad" in pos:                       # a 3x3 conv of a block
                out.append(pos[f"{blk}.{idx}:wgrad"] if leaf == "weight" else -1)
                continue
            if f"bn:{blk}.{idx - 1}:bwd" in pos:                  # its BatchNorm
                out.append(pos[f"bn:{blk}.{idx - 1}:bwd"])
                continue
        out.append(last)                                          # conditioning head
    return out


def plan_buckets(offsets: List[int], sizes: List[int], ready: List[int], cap_elems: int) -> List[Tuple[int, int, int]]:
    """Greedy contiguous buckets over the flat gradient buffer: (lo, hi, ready_index).  A bucket may
    be all-reduced once the backward-plan entry ``ready_index`` has been enqueued."""
    buckets: List[Tuple[int, int, int]] = []
    lo, hi, rdy = None, None, -1
    for off, n, r in zip(offsets, sizes, ready):
        if lo is None:
            lo, hi, rdy = off, off, -1
        hi = off + (n + 3) // 4 * 4
        rdy = max(rdy, r)
        if hi - lo >= cap_elems:
            buckets.append((lo, hi, rdy))
            lo = None
    if lo is not None:
        buckets.append((lo, hi, rdy))
    return buckets


def allreduce_mean_(flat: torch.Tensor, buckets: List[Tuple[int, int, int]], world: int, group=None) -> None:
    """Reference semantics of the data-parallel exchange: every bucket summed over ranks, then / world."""
    for lo, hi, _ in buckets:
        torch.distributed.all_reduce(flat[lo:hi], group=group)
    flat.div_(world)


