"""Multi-GPU host logic on CPU: bucket planning against a synthetic backward plan, and the
data-parallel gradient exchange over a world_size-2 gloo group against the single-process mean."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from tinydiff.dp import BucketReducer, grad_ready_index, plan_buckets, shard_range


def test_shard_range_partitions():
    for n in (1, 7, 128, 1024, 1000):
        for world in (1, 2, 3, 8):
            spans = [shard_range(n, world, r) for r in range(world)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def _fake_plan():
    blocks = [("enc1", (0, 3)), ("enc2", (0, 3)), ("enc3", (0, 3)), ("bottleneck", (0,)), ("dec3", (0, 3)),
              ("dec2", (0, 3)), ("dec1", (0, 3))]
    names = ["time_embedding.0.weight", "time_embedding.0.bias", "time_embedding.2.weight", "time_embedding.2.bias",
             "class_embedding.weight", "initial_conv.weight", "initial_conv.bias"]
    for blk, idxs in blocks:
        for i in idxs:
            names += [f"{blk}.{i}.weight", f"{blk}.{i}.bias", f"{blk}.{i + 1}.weight", f"{blk}.{i + 1}.bias"]
    names += ["final_conv.weight", "final_conv.bias"] + [f"time_proj{i}.{l}" for i in (1, 2, 3) for l in ("weight", "bias")]
    ops = ["final_conv:dbias", "final_conv:wgrad", "final_conv:dgrad"]
    for blk, idxs in reversed(blocks):
        for i in reversed(idxs):
            ops += [f"bn:{blk}.{i}:bwd", f"{blk}.{i}:wgrad", f"{blk}.{i}:dgrad"]
    ops += ["initial_conv:dbias", "initial_conv:wgrad", "embed:bwd"]
    return names, ops


def test_ready_index_and_buckets():
    names, ops = _fake_plan()
    ready = grad_ready_index(names, ops)
    r = dict(zip(names, ready))
    assert r["final_conv.weight"] == ops.index("final_conv:wgrad")
    assert r["dec1.3.weight"] == ops.index("dec1.3:wgrad")
    assert r["dec1.4.weight"] == r["dec1.4.bias"] == ops.index("bn:dec1.3:bwd")
    assert r["enc2.0.bias"] == -1                                  # zero gradient, never written
    assert r["time_proj2.weight"] == r["time_embedding.0.weight"] == ops.index("embed:bwd")
    sizes = [1000 + 37 * i for i in range(len(names))]
    offsets, off = [], 0
    for n in sizes:
        offsets.append(off)
        off += (n + 3) // 4 * 4
    buckets = plan_buckets(offsets, sizes, ready, cap_elems=20000)
    assert buckets[0][0] == 0 and buckets[-1][1] == off
    assert all(a[1] == b[0] for a, b in zip(buckets, buckets[1:]))          # contiguous cover, no overlap
    for lo, hi, rdy in buckets:                                             # ready only after every member
        members = [ready[i] for i, o in enumerate(offsets) if lo <= o < hi]
        assert rdy == max(members)


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _worker(rank, world, port, q):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        names, ops = _fake_plan()
        sizes = [64 + i for i in range(len(names))]
        offsets, off = [], 0
        for n in sizes:
            offsets.append(off)
            off += (n + 3) // 4 * 4
        buckets = plan_buckets(offsets, sizes, grad_ready_index(names, ops), cap_elems=512)
        g = torch.Generator().manual_seed(100 + rank)
        flat = torch.randn(off, generator=g)
        # the exchange exactly as train.TrainStep drives it: buckets all-reduced as the backward plan reaches them
        red = BucketReducer(flat, buckets)
        started = 0
        for i in range(len(ops)):
            started += red.enqueue_ready(i)
        assert started == len(buckets)
        red.wait()
        flat.mul_(red.grad_scale)
        want = sum(torch.randn(off, generator=torch.Generator().manual_seed(100 + r)) for r in range(world)) / world
        q.put((rank, float((flat - want).abs().max())))
    finally:
        dist.destroy_process_group()


def test_gradient_exchange_world2_gloo():
    """N-rank result == single-process emulation (N replicas, averaged gradients)."""
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, q)) for r in range(world)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r for r, _ in res) == [0, 1]
    assert all(err < 1e-6 for _, err in res)
