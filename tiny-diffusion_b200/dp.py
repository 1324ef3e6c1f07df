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


def plan_buckets(offsets: List[int], sizes: List[int], ready: List[int], cap_elems: int,
                 tail_fraction: float = 0.85, tail_div: int = 4) -> List[Tuple[int, int, int]]:
    """Greedy contiguous buckets over the flat gradient buffer: (lo, hi, ready_index).  A bucket may
    be all-reduced once the backward-plan entry ``ready_index`` has been enqueued.

    Gradients that only become final in the last ``1 - tail_fraction`` of the backward plan (the first layers and the
    conditioning head) go into buckets of ``cap_elems / tail_div``: the all-reduce of the LAST bucket cannot overlap anything
    and the optimizer waits for it, so it should be small."""
    buckets: List[Tuple[int, int, int]] = []
    late = tail_fraction * max([r for r in ready if r >= 0] or [0])
    lo, hi, rdy = None, None, -1
    for off, n, r in zip(offsets, sizes, ready):
        if lo is not None and ((r >= late) != (rdy >= late)) and rdy >= 0 and r >= 0:
            buckets.append((lo, hi, rdy))             # never mix late and early gradients in one bucket
            lo = None
        if lo is None:
            lo, hi, rdy = off, off, -1
        hi = off + (n + 3) // 4 * 4
        rdy = max(rdy, r)
        cap = cap_elems // tail_div if rdy >= late else cap_elems
        if hi - lo >= cap:
            buckets.append((lo, hi, rdy))
            lo = None
    if lo is not None:
        buckets.append((lo, hi, rdy))
    # adjacent buckets that become final at the SAME plan entry are enqueued back to back anyway: one larger all-reduce costs
    # one launch latency instead of several (the dense denoisers: every gradient is final at the end -> ONE bucket)
    merged: List[Tuple[int, int, int]] = []
    for b in buckets:
        if merged and merged[-1][2] == b[2] and merged[-1][1] == b[0]:
            merged[-1] = (merged[-1][0], b[1], b[2])
        else:
            merged.append(b)
    return merged


class BucketReducer:
    """The data-parallel gradient exchange of ``train.TrainStep``: the flat gradient buffer cut into
    ``plan_buckets`` ranges; ``enqueue_ready(i)`` starts the asynchronous all-reduce (sum) of every
    bucket whose last gradient is final once backward-plan entry ``i`` has been enqueued, ``wait()``
    makes the current stream (CUDA) / the caller (CPU) wait for all of them.  The mean is taken by the
    consumer (``grad_scale = 1 / world`` in the fused Adam / clip kernels), not here.

    Backend-agnostic on purpose (NCCL on the GPUs, gloo in the CPU tests): no CUDA calls, the caller
    picks the stream the collectives are enqueued from."""

    def __init__(self, flat: torch.Tensor, buckets: List[Tuple[int, int, int]], group=None):
        self.flat, self.buckets, self.group = flat, buckets, group
        self.world = torch.distributed.get_world_size(group)
        self.ready_at = {}
        for lo, hi, r in buckets:
            self.ready_at.setdefault(max(r, 0), []).append((lo, hi))
        self.works = []

    def has_ready(self, i: int) -> bool:
        return i in self.ready_at

    def enqueue_ready(self, i: int) -> int:
        """Start the all-reduce of the buckets that become final at plan entry ``i``; returns how many."""
        spans = self.ready_at.get(i, ())
        for lo, hi in spans:
            self.works.append(torch.distributed.all_reduce(self.flat[lo:hi], group=self.group, async_op=True))
        return len(spans)

    def enqueue_all(self) -> None:
        for i in sorted(self.ready_at):
            self.enqueue_ready(i)

    def wait(self) -> None:
        for w in self.works:
            w.wait()
        self.works = []

    @property
    def grad_scale(self) -> float:
        return 1.0 / self.world
