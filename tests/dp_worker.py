"""Worker of tests/test_gpu_dp.py, launched under ``python -m torch.distributed.run --nproc-per-node N``: the N-rank
data-parallel train step on real NCCL against N oracle replicas with averaged gradients and one Adam step
(SURVEY.md 8e last row; diffusion.py:220-236 semantics under DistributedDataParallel: rank-local BatchNorm statistics,
gradient mean).  Rank 0 writes a JSON report; every rank exits non-zero on a mismatch."""
import argparse
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import torch                      # noqa: E402
import torch.distributed as dist  # noqa: E402


def rel(a, b):
    a, b = a.double().cpu(), b.double().cpu()
    return float((a - b).norm() / b.norm().clamp_min(1e-30))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default="")
    ap.add_argument("--batch", type=int, default=8, help="per-rank batch")
    ap.add_argument("--steps", type=int, default=2)
    ap.add_argument("--precision", default="fp32")
    ap.add_argument("--name", default="conditional_diffusion")
    args = ap.parse_args()
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    from oracle import ddpm_oracle as O                       # checker only
    from oracle.fixtures import init_state_dict, make_inputs
    from tinydiff import _lib as L
    from tinydiff.dp import shard_range
    from tinydiff.train import TrainStep
    import importlib
    mod = importlib.import_module(f"tinydiff.{args.name}")
    dev = L.require_device(f"cuda:{local}")
    dist.init_process_group("nccl", device_id=dev)
    name, Bn = args.name, args.batch
    spec = {"diffusion": O.UNET_MNIST, "conditional_diffusion": O.UNET_COND, "conditional_diffusion_laion": O.UNET_LAION}[name]
    init = init_state_dict(name)
    model = mod.NoiseModel()
    model.load_state_dict(init, strict=True)
    model.precision = args.precision
    model = model.to(dev).train()
    fp = mod.ForwardProcess()
    ts = TrainStep(model, fp, Bn, dev, lr=1e-3, use_graph=True, bucket_mb=4.0)
    assert ts.world == world and len(ts.buckets) > 1

    # ---- oracle emulation: `world` replicas, each on its shard, averaged gradients, one Adam step per step
    bn_blocks = ("enc1", "enc2", "enc3", "bottleneck", "dec3", "dec2", "dec1")
    sd = [{k: v.clone() for k, v in init.items()} for _ in range(world)]
    m = {k: torch.zeros_like(v) for k, v in init.items() if O.is_param(k)}
    v_ = {k: torch.zeros_like(v) for k, v in init.items() if O.is_param(k)}
    report = {"world": world, "per_rank_batch": Bn, "precision": args.precision, "steps": []}
    ok = True
    for step in range(1, args.steps + 1):
        inp = make_inputs(name, world * Bn, seed=700 + step)
        lo, hi = shard_range(world * Bn, world, rank)
        cond = inp.get("cond")
        loss = ts(inp["x0"][lo:hi], None if cond is None else cond[lo:hi], t=inp["t"][lo:hi], noise=inp["noise"][lo:hi])
        torch.cuda.synchronize()
        grads, losses, cal = [], [], []
        for r in range(world):
            a, b = shard_range(world * Bn, world, r)
            c = None if cond is None else cond[a:b]
            lr_, g_, st_, _ = O.unet_loss_and_grads(spec, sd[r], inp["x0"][a:b], inp["t"][a:b], inp["noise"][a:b],
                                                    fp.alphas_cumprod, c)
            if args.precision == "bf16":
                _, gc_, _, _ = O.unet_loss_and_grads(spec, sd[r], inp["x0"][a:b], inp["t"][a:b], inp["noise"][a:b],
                                                     fp.alphas_cumprod, c, autocast_bf16=True)
                cal.append(gc_)
            grads.append(g_)
            losses.append(float(lr_))
            sd[r].update(st_)                                      # rank-local BatchNorm running statistics
        mean = {k: sum(g[k] for g in grads) / world for k in m}
        mean_cal = {k: sum(g[k] for g in cal) / world for k in m} if cal else None
        for k in m:
            parts = k.split(".")
            if k.endswith(".bias") and parts[0] in bn_blocks and parts[1] in ("0", "3"):
                mean[k] = torch.zeros_like(mean[k])               # conv bias in front of a train-mode BN: exactly zero
        # (1) this rank's loss and BatchNorm buffers are those of ITS shard
        lerr = abs(float(loss) - losses[rank]) / losses[rank]
        bufs = dict(model.named_buffers())
        berr = max(rel(bufs[k], sd[rank][k]) for k in bufs if not k.endswith("num_batches_tracked"))
        # (2) the all-reduced gradient / world is the mean of the replicas' gradients
        gerr, gbad = {}, {}
        for k in ts.names:
            got = ts.eng.pgrad[k] / world
            if float(mean[k].norm()) < 1e-7:
                if float(got.abs().max()) > 1e-5:
                    gbad[k] = (float(got.abs().max()), 1e-5)
                continue
            e = rel(got, mean[k])
            gerr[k] = e
            tol = (1e-5 if k.startswith("final_conv") else 3e-2) if args.precision == "fp32" else \
                max(1.5 * rel(mean_cal[k], mean[k]), 3e-2)
            if e > tol:
                gbad[k] = (e, tol)
        # (3) every rank holds the same parameters after the step (bit for bit)
        flat = torch.cat([p.detach().flatten() for p in model.parameters()])
        ref = flat.clone()
        dist.broadcast(ref, src=0)
        same = bool(torch.equal(flat, ref))
        # (4) parameters after Adam against the oracle's update computed from OUR mean gradient (isolates the optimizer)
        perr = 0.0
        for k, p in model.named_parameters():
            new, m[k], v_[k] = O.adam_step(sd[0][k], (ts.eng.pgrad[k] / world).cpu(), m[k], v_[k], step)
            perr = max(perr, rel(p, new))
            for r in range(world):
                sd[r][k] = p.detach().cpu().clone()               # continue from the device's parameters
        med = sorted(gerr.values())[len(gerr) // 2]
        rec = {"step": step, "rank": rank, "loss_rel_err": lerr, "bn_buffer_max_rel_err": berr,
               "grad_rel_l2_median": med, "grad_rel_l2_max": max(gerr.values()), "grad_bad": {k: list(v) for k, v in gbad.items()},
               "params_identical_across_ranks": same, "param_after_adam_max_rel_err": perr,
               "buckets": len(ts.buckets), "launches_per_step": ts.launches_per_step}
        good = (lerr < (1e-5 if args.precision == "fp32" else 1e-2) and berr < (1e-5 if args.precision == "fp32" else 2e-3)
                and not gbad and same and perr < 1e-5)
        ok &= good
        allrec = [None] * world
        dist.all_gather_object(allrec, rec)
        if rank == 0:
            report["steps"].append(allrec)
    flag = torch.tensor([int(ok)], device=dev)
    dist.all_reduce(flag, op=dist.ReduceOp.MIN)
    report["ok"] = bool(int(flag))
    if rank == 0:
        text = json.dumps(report, indent=1)
        print(text)
        if args.out:
            os.makedirs(os.path.dirname(os.path.abspath(args.out)), exist_ok=True)
            with open(args.out, "w") as f:
                f.write(text + "\n")
    ts.close()
    torch.cuda.synchronize()
    dist.barrier()
    import threading
    th = threading.Thread(target=dist.destroy_process_group, daemon=True)
    th.start()
    th.join(20.0)
    sys.stdout.flush()
    os._exit(0 if int(flag) else 1)


if __name__ == "__main__":
    main()
