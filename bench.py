#!/usr/bin/env python
"""bench.py -- headline benchmark of the DDPM hot path (BASELINE.json metric).

    python bench.py --gpus N --steps K --warmup W            # this repo's sm_100a kernels
    python bench.py --impl reference --steps K --warmup W     # the reference algorithm on host cores

Workload (BASELINE.json configs[1]): conditional_diffusion.py class-conditional MNIST-shape UNet,
1000-step ancestral sampling, batch 128 per GPU (global batch 128*N; weak scaling, the sample batch
is sharded across ranks with no data-path collective).  One bench "step" = one full 1000-step
reverse loop over the rank's batch.  Random-init weights (seed 0), synthetic inputs.

Prints ONE JSON line on rank 0.  `value` = samples/s with x_T resident in HBM; `e2e` = the same
through the public `sample()` call with host buffers (pinned x_T/labels H2D and the D2H read of the
samples inside the timed region).
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

import torch  # noqa: E402

MODEL = "conditional_diffusion"
PER_GPU_BATCH = 128
T_STEPS = 1000
METRIC = "ddpm_1000step_samples_per_sec"
UNIT = "samples/s"


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.isfile(p):
        with open(p) as f:
            d = json.load(f)
        return {"hbm_gbs": d["hbm_gbs"], "bf16_burst": d["bf16_tflops"],
                "bf16_sustained": d.get("bf16_tflops_sustained", d["bf16_tflops"]), "source": "measured"}
    return {"hbm_gbs": 6650.0, "bf16_burst": 1590.0, "bf16_sustained": 1400.0, "source": "fallback"}


class ClockSampler:
    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index: int):
        self.tmp = tempfile.NamedTemporaryFile("w+", suffix=".csv", delete=False)
        self.proc = None
        try:
            self.proc = subprocess.Popen(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits",
                                          "-i", str(gpu_index), "-lms", "200"], stdout=self.tmp,
                                         stderr=subprocess.DEVNULL)
        except Exception:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": []}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except Exception:
            self.proc.kill()
        self.tmp.flush()
        self.tmp.seek(0)
        sm, mx, reasons = [], [], set()
        for line in self.tmp.read().splitlines():
            f = [x.strip() for x in line.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(("hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"), f[5:9]):
                if v.lower().startswith("active"):
                    reasons.add(name)
        try:
            os.unlink(self.tmp.name)
        except OSError:
            pass
        if sm:
            load = [s for s in sm if s >= 0.5 * max(sm)] or sm
            out = {"sm_mhz": statistics.median(load), "sm_max_mhz": max(mx), "reasons": sorted(reasons),
                   "samples": len(sm)}
        return out


# ------------------------------------------------------------------------------------------
# CPU reference arm (oracle port; the ONLY place outside tests/smoke that may execute oracle/)
# ------------------------------------------------------------------------------------------
def cpu_reference_samples_per_sec(reverse_steps: int, batch: int = PER_GPU_BATCH):
    """Time `reverse_steps` iterations of the reference sampler loop body (conditional_diffusion.py:
    369-384: eval-mode NoiseModel forward + the x_{t-1} update) at batch `batch` on all host cores
    and extrapolate to the 1000 steps one sample batch needs."""
    from oracle import ddpm_oracle as O
    from oracle.fixtures import init_state_dict, make_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = init_state_dict(MODEL)
    inp = make_inputs(MODEL, batch)
    betas, alphas, ac = O.make_schedule(T_STEPS)
    g = torch.Generator().manual_seed(5)
    x = torch.randn(batch, 1, 28, 28, generator=g)
    y = inp["cond"]
    with torch.no_grad():
        O.unet_forward(O.UNET_COND, sd, x, torch.full((batch,), T_STEPS - 1), y)      # warm-up
        t0 = time.perf_counter()
        for i in range(reverse_steps):
            t = T_STEPS - 1 - i
            eps = O.unet_forward(O.UNET_COND, sd, x, torch.full((batch,), t, dtype=torch.long), y)
            x = O.p_sample_step(x, eps, torch.randn(x.shape, generator=g), t, betas, alphas, ac)
        dt = time.perf_counter() - t0
    per_step = dt / reverse_steps
    return batch / (per_step * T_STEPS), cores, dt


def cpu_reference_train_imgs_per_sec(steps: int, batch: int = PER_GPU_BATCH):
    """Time `steps` iterations of the reference train-step body (conditional_diffusion.py:254-263:
    q_sample, train-mode forward, MSE, backward, Adam) at batch `batch` on all host cores."""
    from oracle import ddpm_oracle as O
    from oracle.fixtures import init_state_dict, make_inputs
    cores = os.cpu_count() or 1
    torch.set_num_threads(cores)
    sd = init_state_dict(MODEL)
    inp = make_inputs(MODEL, batch)
    _, _, ac = O.make_schedule(T_STEPS)
    m = {k: torch.zeros_like(v) for k, v in sd.items() if O.is_param(k)}
    v = {k: torch.zeros_like(x) for k, x in sd.items() if O.is_param(k)}
    t0 = time.perf_counter()
    for step in range(1, steps + 1):
        _, grads, stats, _ = O.unet_loss_and_grads(O.UNET_COND, sd, inp["x0"], inp["t"], inp["noise"], ac, inp["cond"])
        for k in m:
            sd[k], m[k], v[k] = O.adam_step(sd[k], grads[k], m[k], v[k], step)
        sd.update(stats)
    dt = time.perf_counter() - t0
    return batch * steps / dt, cores, dt


def run_reference(args):
    rank = int(os.environ.get("RANK", "0"))
    if rank != 0:
        return
    vals = []
    for _ in range(args.warmup):
        cpu_reference_samples_per_sec(1)
    t0 = time.perf_counter()
    for _ in range(args.steps):
        v, cores, _ = cpu_reference_samples_per_sec(args.ref_reverse_steps)
        vals.append(v)
    wall = time.perf_counter() - t0
    value = statistics.mean(vals)
    sample = (f"{args.ref_reverse_steps} of the {T_STEPS} reverse steps per bench step at batch {PER_GPU_BATCH}, "
              f"extrapolated x{T_STEPS // args.ref_reverse_steps}")
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(args.steps, 1),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
        "config": {"workload": f"{MODEL} UNet 1x28x28, T={T_STEPS} ancestral sampling, batch {PER_GPU_BATCH} "
                               f"(CPU oracle PORT of the reference loop body on the host cores; {args.ref_reverse_steps} of the "
                               f"{T_STEPS} reverse steps timed per bench step, extrapolated x{T_STEPS // args.ref_reverse_steps})",
                   "per_gpu_batch": PER_GPU_BATCH, "reverse_steps_timed": args.ref_reverse_steps,
                   "extrapolated": True, "kind": "port"},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit_line(line)


# ------------------------------------------------------------------------------------------
# GPU arm
# ------------------------------------------------------------------------------------------
def run_gpu(args):
    import torch.distributed as dist
    from tinydiff import _lib as L
    from tinydiff.conditional_diffusion import ForwardProcess, NoiseModel, sample

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus and world > 1:
        raise SystemExit(f"--gpus {args.gpus} but WORLD_SIZE={world}")
    torch.cuda.set_device(local_rank)
    dev = L.require_device(f"cuda:{local_rank}")
    if world > 1:
        # NCCL keeps its default channel count: capping it at 8 CTAs and sizing the persistent grids for the remaining 140 SMs
        # measured SLOWER on 8 GPUs (2.30 vs 2.24 ms per data-parallel step; profiles/r02_dp_settings.txt)
        dist.init_process_group("nccl", device_id=dev)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    B = PER_GPU_BATCH
    torch.manual_seed(0)
    model = NoiseModel().to(dev).eval()
    model.precision = args.precision
    fp = ForwardProcess(T_STEPS)
    gen = torch.Generator().manual_seed(1234 + rank)
    y_host = torch.randint(0, 10, (B,), generator=gen).pin_memory()
    xT_host = torch.randn(B, 1, 28, 28, generator=gen).pin_memory()
    out_host = torch.empty(B, 1, 28, 28).pin_memory()

    eng = model.engine(B, dev)
    eng.refresh_weights()
    launches_per_reverse_step = eng.num_launches() + 2          # + p_sample + step counter

    # ---- device-resident measurement: x_T, y already in HBM; the loop is replayed from captured graphs ----
    def device_step():
        eng.x_in.copy_(xT_dev)
        eng.y_in.copy_(y_dev)
        eng.use_t_dev = True
        eng.prepare_sampler_embed()
        loop.run(z=None, seed=7)

    from tinydiff.process import ReverseLoop
    xT_dev, y_dev = xT_host.to(dev), y_host.to(dev)
    eng.y_in.copy_(y_dev)
    eng.use_t_dev = True
    eng.prepare_sampler_embed()
    loop = ReverseLoop(fp, eng.x_in, eng.eps, eng.t_dev, eng.launch, use_graph=True)
    eng._reverse_loop = loop
    for _ in range(args.warmup):
        device_step()
    barrier()
    clocks = ClockSampler(local_rank) if rank == 0 else None
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(args.steps):
        device_step()
    e1.record()
    barrier()
    dev_ms = e0.elapsed_time(e1)
    assert torch.isfinite(eng.x_in).all(), "non-finite samples"

    # ---- end to end through the public API with host buffers -----------------------------------
    def e2e_step():
        x0 = sample(model, fp, dev, n_samples=B, y=y_host, x_T=xT_host, seed=7)
        out_host.copy_(x0, non_blocking=False)

    for _ in range(max(1, min(args.warmup, 2))):
        e2e_step()
    barrier()
    f0, f1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    f0.record()
    for _ in range(args.steps):
        e2e_step()
    f1.record()
    barrier()
    e2e_ms = f0.elapsed_time(f1)
    # ---- training step (second half of the BASELINE metric: train imgs/sec) -----------------------
    from tinydiff.train import TrainStep
    torch.manual_seed(0)
    tmodel = NoiseModel().to(dev).train()
    tmodel.precision = args.precision
    tfp = ForwardProcess(T_STEPS)
    ts = TrainStep(tmodel, tfp, B, dev, lr=1e-3, use_graph=(world == 1 or os.environ.get("TD_DP_GRAPH", "1") != "0"))
    x0_host = (torch.rand(B, 1, 28, 28, generator=gen) * 2 - 1).pin_memory()
    x0_dev = x0_host.to(dev)
    TS = args.train_steps

    def train_dev_step():                      # inputs resident in HBM; t and noise drawn on the device
        for _ in range(TS):
            ts(x0_dev, y_dev)

    def train_e2e_step():                      # host batch in, loss value out, every step
        for _ in range(TS):
            loss = ts(x0_host, y_host)
            loss_host.copy_(loss)
            torch.cuda.current_stream().synchronize()

    loss_host = torch.zeros(1).pin_memory()
    for _ in range(args.warmup):
        train_dev_step()
    barrier()
    g0, g1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    g0.record()
    for _ in range(args.steps):
        train_dev_step()
    g1.record()
    barrier()
    train_ms = g0.elapsed_time(g1)
    train_e2e_step()
    barrier()
    h0, h1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    h0.record()
    for _ in range(args.steps):
        train_e2e_step()
    h1.record()
    barrier()
    train_e2e_ms = h0.elapsed_time(h1)
    final_loss = float(loss_host)
    assert final_loss == final_loss and final_loss < 1e6, f"training diverged: loss {final_loss}"
    clk = clocks.stop() if clocks else None

    t = torch.tensor([dev_ms, e2e_ms, train_ms, train_e2e_ms], device=dev, dtype=torch.float64)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    dev_ms, e2e_ms, train_ms, train_e2e_ms = (float(v) for v in t)
    n_train = args.steps * TS
    fwd_l, bwd_l = ts.eng.num_launches()
    train = {
        "metric": "ddpm_train_imgs_per_sec", "unit": "img/s",
        "value": world * B * n_train / (train_ms / 1e3), "ms_per_step": train_ms / n_train, "train_steps": n_train,
        "config": {"workload": f"{MODEL} UNet train step (randint t, q_sample, train-mode forward, MSE, backward, Adam), "
                               f"batch {B}/GPU, data parallel x{world}"
                               + (", bucketed NCCL gradient all-reduce overlapped with backward" if world > 1 else
                                  ", whole step replayed as one CUDA graph")},
        "e2e": {"value": world * B * n_train / (train_e2e_ms / 1e3), "unit": "img/s",
                "h2d_bytes_per_step": int(x0_host.numel() * 4 + y_host.numel() * 8), "d2h_bytes_per_step": 4,
                "ms_per_step": train_e2e_ms / n_train},
        "gpu_launches_per_step": int(getattr(ts, "launches_per_step", 0)) or (fwd_l + bwd_l + 5), "final_loss": final_loss,
        "model_flops_per_step": ts.eng.conv_flops(),
        "achieved_model_tflops": ts.eng.conv_flops() * n_train / (train_ms / 1e3) / 1e12,
    }
    value = world * B * args.steps / (dev_ms / 1e3)
    e2e_value = world * B * args.steps / (e2e_ms / 1e3)
    dit_dp = None
    if world > 1 and not args.no_configs:
        # BASELINE config 3: the DiT train step, data parallel over the ranks (bucketed NCCL all-reduce inside the captured step)
        dit_dp = dit_data_parallel(dev, world, dist, ts)

    line = None
    if rank == 0:
        peaks = measured_peaks()
        roof = conv_roofline(eng, peaks, iters=max(3, args.steps), reverse_step_ms=dev_ms / args.steps / T_STEPS)
        ws_bytes = sum(b.numel() * b.element_size() for b in eng.bufs.values())
        line = {
            "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": dev_ms / args.steps, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "bf16" if args.precision == "bf16" else "f32",
            "data": "synthetic",
            "config": {"workload": f"{MODEL} UNet (11.18M params) 1x28x28, T={T_STEPS} ancestral sampling, "
                                   f"batch {B}/GPU (global {B * world}), random-init weights",
                       "per_gpu_batch": B, "global_batch": B * world, "parallelism": f"sample-sharded x{world}",
                       "l2": f"no flush: activation working set {ws_bytes / 2**20:.0f} MiB per forward > 126 MB L2",
                       "cuda_graph": f"{ReverseLoop.STEPS_PER_GRAPH} reverse steps captured per graph, "
                                     f"{T_STEPS // ReverseLoop.STEPS_PER_GRAPH} replays per bench step; programmatic dependent launch between kernels"},
            "e2e": {"value": e2e_value, "unit": UNIT, "h2d_bytes_per_step": int(xT_host.numel() * 4 + y_host.numel() * 8),
                    "d2h_bytes_per_step": int(out_host.numel() * 4), "ms_per_step": e2e_ms / args.steps},
            "gpu_launches": int(args.steps * T_STEPS * (loop.launches_per_step or launches_per_reverse_step + 5)),
            "launches_per_reverse_step": int(loop.launches_per_step or launches_per_reverse_step + 5),
            "clocks": clk,
            "roofline": roof,
            "model_flops_per_sample": eng.conv_flops() / B * T_STEPS,
            "achieved_model_tflops": value * eng.conv_flops() / B * T_STEPS / 1e12 / world,
            "train": train,
        }
        if dit_dp is not None:
            line["configs"] = dit_dp
        train["roofline_frac_of_sustained_bf16"] = train["achieved_model_tflops"] / peaks["bf16_sustained"]
        if world == 1:
            train["roofline"] = train_roofline(ts, peaks, train_ms / n_train)
            if not args.no_gpu_eager:
                line["gpu_eager_baseline"] = gpu_eager_baseline(dev, B)
            if not args.no_configs:
                ts.close()
                line["configs"] = other_configs(dev, peaks)
        if world == 1 and not args.no_cpu_baseline:
            v, cores, dt = cpu_reference_samples_per_sec(args.cpu_reverse_steps)
            line["cpu_baseline"] = {"value": v, "unit": UNIT, "cores": cores, "kind": "port",
                                    "sample": f"{args.cpu_reverse_steps} of the {T_STEPS} reverse steps at batch {B} "
                                              f"({dt:.1f} s of CPU work), extrapolated to 1000"}
            tv, cores, tdt = cpu_reference_train_imgs_per_sec(args.cpu_train_steps)
            train["cpu_baseline"] = {"value": tv, "unit": "img/s", "cores": cores, "kind": "port",
                                     "sample": f"{args.cpu_train_steps} train steps at batch {B} ({tdt:.1f} s of CPU work)"}
        emit_line(line)
    if world > 1:
        # the captured data-parallel step holds NCCL kernels: release the graphs before the communicator goes away, and
        # never let a slow communicator teardown keep the (already reported) run alive
        ts.close()
        torch.cuda.synchronize()
        dist.barrier()
        import threading
        th = threading.Thread(target=dist.destroy_process_group, daemon=True)
        th.start()
        th.join(20.0)
        sys.stdout.flush()
        sys.stderr.flush()
        os._exit(0)
    return line


def conv_roofline(eng, peaks, iters=3, reverse_step_ms=None):
    """Per-entry CUDA-event timing of one eval forward (same buffers as the timed loop) -> achieved TFLOP/s of the
    dominant kernels, the tcgen05 implicit-GEMM convolutions (conv3x3_halo_kernel, and conv3x3_tc_kernel on the
    4x4 / 8x8 maps) = algorithmic conv FLOPs / summed conv launch time.  Every plan entry is launched REPS times
    back to back between one event pair, so the host's launch latency is hidden behind the previous launch and
    the figure is device time (a single launch between events also counts the idle gap before it starts).
    These are launches timed alone (clocks near max), so `peak` is the BURST bf16 figure of MEASURED_PEAKS.json;
    `in_loop` is the same FLOP count over the driver-timed reverse step (graph replay, power-capped clocks, every
    non-convolution kernel included) against the SUSTAINED figure -- a lower bound on the in-loop conv fraction."""
    from tinydiff import _lib as L
    st = L.stream_ptr()
    names = [n for n, _ in eng.ops]
    acc = {n: 0.0 for n in names}
    eng.use_t_dev = False
    REPS = 4
    eng.launch()
    torch.cuda.synchronize()
    for it in range(iters):
        for n, fn in eng.ops:
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            for _ in range(REPS):
                fn(st)
            b.record()
            torch.cuda.synchronize()
            acc[n] += a.elapsed_time(b) / REPS / iters
    tc = [n for n in names if n in eng.plans and eng.engines.get(n) == L.CONV_TC]
    tc_ms = sum(acc[n] for n in tc)
    tc_flops = sum(eng.plans[n].flops for n in tc)
    total_ms = sum(acc.values())
    achieved = tc_flops / (tc_ms * 1e-3) / 1e12 if tc_ms > 0 else 0.0
    traffic, traffic_src = None, None
    summ = os.path.join(ROOT, "profiles", "ncu_summary.json")
    if os.path.isfile(summ):
        try:
            with open(summ) as f:
                d = json.load(f)
            traffic, traffic_src = d.get("conv_tc_dram_bytes_per_launch"), d.get("source")
        except Exception:
            traffic = None
    per_layer = []
    for n in tc:
        dsc = eng.plans[n].desc
        P = dsc.batch * dsc.height * dsc.width
        byts = 2 * P * dsc.cin + 2 * P * dsc.cout + 2 * 9 * dsc.cin * dsc.cout       # bf16 in + out + weights, each once
        us = acc[n] * 1e3
        tf = eng.plans[n].flops / (acc[n] * 1e-3) / 1e12 if acc[n] > 0 else 0.0
        per_layer.append({"layer": n, "shape": f"{dsc.cin}->{dsc.cout}@{dsc.height}x{dsc.width}", "gflop": round(eng.plans[n].flops / 1e9, 3),
                          "algorithmic_mb": round(byts / 1e6, 2), "us": round(us, 2), "tflops": round(tf, 1),
                          "frac_burst": round(tf / peaks["bf16_burst"], 3)})
    out = {"bound": "tensor", "kernel": "conv3x3_halo_kernel + conv3x3_tc_kernel (tcgen05 implicit GEMM, 13 launches per forward)",
           "achieved": achieved, "peak": peaks["bf16_burst"],
           "unit": "TFLOP/s", "frac": achieved / peaks["bf16_burst"],
           "peak_source": peaks["source"] + " burst bf16 (launches timed alone)",
           "frac_of_sustained": achieved / peaks["bf16_sustained"],
           "traffic": traffic, "traffic_source": traffic_src, "launches_per_forward": len(tc), "flops_per_forward": tc_flops,
           "conv_ms_per_forward": tc_ms, "all_kernels_ms_per_forward": total_ms,
           "conv_share_of_forward": tc_ms / total_ms if total_ms else None,
           "per_layer": per_layer,
           "per_kernel_us": {n: round(acc[n] * 1e3, 2) for n in names}}
    fused_pool = [n for n in tc if getattr(eng.plans[n].desc, "pool_y", None)
                  and int(L.load().td_conv3x3_pool_fused(eng.plans[n].handle))]
    if fused_pool:
        out["fused_into_conv_launches"] = {
            "layers": fused_pool,
            "note": "MaxPool2d(2) of these layers is written by their conv epilogue (no separate pooling launch): their launch time, "
                    "and so `achieved` / `frac`, includes that work (+2..4 us each, profiles/r02_pool_fuse_sweep.txt)"}
    if reverse_step_ms:
        tfl = tc_flops / (reverse_step_ms * 1e-3) / 1e12
        out["in_loop"] = {"achieved": tfl, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s", "frac": tfl / peaks["bf16_sustained"],
                          "note": "conv FLOPs / whole driver-timed reverse step (graph replay; glue kernels included in the time)"}
    return out


def train_roofline(ts, peaks, step_ms, iters=3):
    """Per-entry timing of the train plan (eager launches, each entry followed by a join of the weight-gradient side stream
    so the events on the main stream see it) -> which kernels the step spends its time in.  The headline fraction is model
    FLOPs (forward conv + data gradient + weight gradient) over the driver-timed graph step against the sustained peak."""
    from tinydiff import _lib as L
    eng = ts.eng
    st = L.stream_ptr()
    eng.launch_forward()
    eng.launch_backward()
    torch.cuda.synchronize()
    acc = {}
    for tag, ops in (("fwd", eng.fwd_ops), ("bwd", eng.bwd_ops)):
        for n, fn in ops:
            if n in ("embed", "embed:bwd", "wgrad:join"):
                continue
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            fn(st)
            eng.sync_wgrad_stream()
            torch.cuda.synchronize()
            a.record()
            for _ in range(iters):
                fn(st)
                eng.sync_wgrad_stream()
            b.record()
            torch.cuda.synchronize()
            acc[f"{tag}:{n}"] = a.elapsed_time(b) / iters * 1e3
    def group(k):
        if ":wgrad" in k:
            return "weight gradient (wgrad_tc + split-K reduce)"
        if ":dgrad" in k:
            return "data gradient conv (tcgen05)"
        if k.startswith("fwd:bn:"):
            return "BatchNorm forward (finalize + apply)"
        if k.startswith("bwd:bn:"):
            return "BatchNorm backward (reduce + finalize + apply)"
        if "upcat" in k or "resize" in k or "pool" in k:
            return "pool / upsample / concat glue"
        if k.startswith("fwd:") and k.split(":")[1] in ("initial_conv", "final_conv") or k.startswith("bwd:final_conv") \
                or k.startswith("bwd:initial_conv"):
            return "1-channel boundary convs (direct kernels)"
        if k.startswith("fwd:") and (k.split(":")[1] in eng.conv_plans):
            return "forward conv (tcgen05)"
        return "other"
    groups = {}
    for k, v in acc.items():
        groups[group(k)] = groups.get(group(k), 0.0) + v
    total = sum(acc.values())
    top = sorted(acc.items(), key=lambda kv: -kv[1])[:8]
    flops = eng.conv_flops()
    tfl = flops / (step_ms * 1e-3) / 1e12
    tensor_us = sum(v for g, v in groups.items() if "tcgen05" in g or "weight gradient" in g)
    limiting = max(((g, v) for g, v in groups.items() if "tcgen05" not in g and "weight gradient" not in g),
                   key=lambda gv: gv[1], default=("", 0.0))
    return {"bound": "tensor", "achieved": tfl, "peak": peaks["bf16_sustained"], "unit": "TFLOP/s",
            "frac": tfl / peaks["bf16_sustained"], "peak_source": peaks["source"] + " sustained bf16 (kernels timed inside a long step)",
            "model_flops_per_step": flops, "eager_entry_sum_us": round(total, 1),
            "tensor_kernel_share_of_entry_time": round(tensor_us / total, 3) if total else None,
            "limiting_non_tensor_group": {"name": limiting[0], "us_per_step": round(limiting[1], 1)},
            "group_us": {g: round(v, 1) for g, v in sorted(groups.items(), key=lambda gv: -gv[1])},
            "top_entries_us": {k: round(v, 1) for k, v in top}}


def gpu_eager_baseline(dev, B, reverse_steps=10, train_steps=10):
    """The on-box GPU bar (SURVEY.md 8d last sentence): the reference's own ops -- the functional
    restatement in oracle/, i.e. F.conv2d / F.batch_norm / F.interpolate ... through cuDNN / cuBLAS / ATen -- run EAGERLY on
    this B200, (i) fp32 with PyTorch's defaults (TF32 convolutions, as the reference runs on a GPU) and (ii) under
    torch.autocast(bfloat16) with channels_last tensors.  A baseline, not part of the product path."""
    from oracle import ddpm_oracle as O
    from oracle.fixtures import init_state_dict, make_inputs
    import torch.nn.functional as F
    out = {"what": "oracle/ddpm_oracle.py (functional restatement of the reference's NoiseModel / train step) with CUDA tensors: "
                   "library kernels (cuDNN / cuBLAS / ATen), eager, no CUDA graph",
           "batch": B, "reverse_steps_timed": reverse_steps, "train_steps_timed": train_steps, "extrapolated": True}
    sd0 = {k: v.to(dev) for k, v in init_state_dict(MODEL).items()}
    inp = {k: v.to(dev) for k, v in make_inputs(MODEL, B).items()}
    betas, alphas, ac = (v.to(dev) for v in O.make_schedule(T_STEPS))
    for mode in ("fp32_tf32_default", "bf16_autocast_channels_last"):
        bf = mode.startswith("bf16")
        sd = dict(sd0)
        if bf:
            sd = {k: (v.contiguous(memory_format=torch.channels_last) if v.dim() == 4 else v) for k, v in sd0.items()}
        x = inp["noise"].clone()
        if bf:
            x = x.contiguous(memory_format=torch.channels_last)
        ctx = lambda: torch.autocast("cuda", dtype=torch.bfloat16, enabled=bf)

        def rev_step(x, t):
            with torch.no_grad(), ctx():
                eps = O.unet_forward(O.UNET_COND, sd, x, torch.full((B,), t, device=dev, dtype=torch.long), inp["cond"]).float()
            noise = torch.randn_like(x) if t > 0 else torch.zeros_like(x)
            return (1 / torch.sqrt(alphas[t])) * (x - ((1 - alphas[t]) / torch.sqrt(1 - ac[t])) * eps) + torch.sqrt(betas[t]) * noise
        for i in range(3):
            x = rev_step(x, T_STEPS - 1 - i)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(reverse_steps):
            x = rev_step(x, T_STEPS - 4 - i)
        e1.record()
        torch.cuda.synchronize()
        rev_ms = e0.elapsed_time(e1) / reverse_steps
        # train step: diffusion.py:220-236 with torch autograd + torch.optim.Adam
        leaf = {k: (v.detach().clone().requires_grad_(True) if O.is_param(k) else v.clone()) for k, v in sd.items()}
        opt = torch.optim.Adam([v for k, v in leaf.items() if O.is_param(k)], lr=1e-3)

        def train_step():
            t = torch.randint(0, T_STEPS, (B,), device=dev)
            noise = torch.randn_like(inp["x0"])
            x_t = torch.sqrt(ac[t]).view(-1, 1, 1, 1) * inp["x0"] + torch.sqrt(1 - ac[t]).view(-1, 1, 1, 1) * noise
            if bf:
                x_t = x_t.contiguous(memory_format=torch.channels_last)
            stats = {}
            with ctx():
                pred = O.unet_forward(O.UNET_COND, leaf, x_t, t, inp["cond"], training=True, new_stats=stats)
            loss = F.mse_loss(pred.float(), noise)
            opt.zero_grad()
            loss.backward()
            opt.step()
            for k, v in stats.items():
                leaf[k] = v
            return loss
        for _ in range(3):
            train_step()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(train_steps):
            train_step()
        e1.record()
        torch.cuda.synchronize()
        tr_ms = e0.elapsed_time(e1) / train_steps
        out[mode] = {"reverse_step_ms": rev_ms, "samples_per_sec_1000step": B / (rev_ms * T_STEPS * 1e-3),
                     "train_step_ms": tr_ms, "train_imgs_per_sec": B / (tr_ms * 1e-3)}
        del leaf, opt
    torch.backends.cudnn.benchmark = False
    return out


def dit_data_parallel(dev, world, dist, prev_ts):
    """BASELINE config 3 under data parallelism: fused DiT train step (one graph per rank, NCCL all-reduce of the 13 MB of
    gradients inside it) at the reference batch and at 4096 per GPU; max over ranks, whole-job samples/s."""
    from tinydiff.diffusion_transformer import ForwardProcess, NoiseModel
    from tinydiff.train import TrainStep
    prev_ts.close()
    out = []
    for Bn in (128, 4096):
        torch.manual_seed(0)
        model = NoiseModel(dropout=0.0).to(dev).train()
        fp = ForwardProcess()
        g = torch.Generator().manual_seed(1)
        x0 = torch.randn(Bn, 20, generator=g).to(dev)
        y = torch.randint(0, 10, (Bn,), generator=g).to(dev)
        ts = TrainStep(model, fp, Bn, dev, lr=3e-4, use_graph=True)
        ts.load(x0, y)
        for _ in range(3):
            ts.run()
        dist.barrier()
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            ts.run()
        e1.record()
        dist.barrier()
        torch.cuda.synchronize()
        t = torch.tensor([e0.elapsed_time(e1) / 20], device=dev, dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        ms = float(t)
        out.append({"config": 3, "model": "diffusion_transformer", "batch_per_gpu": Bn, "n_gpus": world,
                    "train": {"ms_per_step": ms, "samples_per_sec": world * Bn / ms * 1e3, "buckets": len(ts.buckets)}})
        ts.close()
        del ts, model
    return out


def other_configs(dev, peaks):
    """BASELINE.json configs 3-5 (DiT, latent MLP, LAION latent UNet) at the reference batch and at the largest point of the
    SURVEY.md 8d sweep: fused train step (one CUDA graph, 10 timed replays) and graph-captured reverse steps through the public
    `sample()` (the full T = 1000 schedule at batch <= 256, T = 60 at 65536 samples and scaled; whole 20-step graphs either way).  Model FLOPs / time against the sustained bf16 peak (the dense denoisers compute in
    fp32 / tf32; their fraction is quoted against the same bf16 denominator and is latency-bound at the reference batch)."""
    import importlib
    from tinydiff.train import TrainStep
    res = []
    for cfg_id, name, batches in ((3, "diffusion_transformer", (128, 65536)), (4, "latent_diffusion", (128, 65536)),
                                  (5, "conditional_diffusion_laion", (8, 256))):
        for Bn in batches:
            rec = {"config": cfg_id, "model": name, "batch": Bn}
            T = 1000 if Bn <= 256 else 60
            try:
                mod = importlib.import_module(f"tinydiff.{name}")
                torch.manual_seed(0)
                kw = {"dropout": 0.0} if name == "diffusion_transformer" else {}
                model = mod.NoiseModel(**kw).to(dev)
                fp = mod.ForwardProcess(num_timesteps=T)
                g = torch.Generator().manual_seed(1)
                if name == "conditional_diffusion_laion":
                    x0 = (0.18215 * torch.randn(Bn, 4, 32, 32, generator=g)).to(dev)
                    cond = torch.randn(Bn, 768, generator=g).to(dev)
                else:
                    x0 = torch.randn(Bn, 20, generator=g).to(dev)
                    cond = torch.randint(0, 10, (Bn,), generator=g).to(dev)
                model.train()
                ts = TrainStep(model, fp, Bn, dev, use_graph=True,
                               max_grad_norm=10.0 if name == "conditional_diffusion_laion" else None)
                ts.load(x0, cond)
                for _ in range(3):
                    ts.run()
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for _ in range(10):
                    ts.run()
                e1.record()
                torch.cuda.synchronize()
                ms = e0.elapsed_time(e1) / 10
                fl = ts.eng.conv_flops()
                rec["train"] = {"ms_per_step": ms, "samples_per_sec": Bn / ms * 1e3, "model_tflops": fl / (ms * 1e-3) / 1e12,
                                "frac_of_sustained_bf16": fl / (ms * 1e-3) / 1e12 / peaks["bf16_sustained"],
                                "launches_per_step": ts.launches_per_step}
                ts.close()
                model.eval()
                if name == "conditional_diffusion_laion":
                    f = lambda: mod.sample(model, fp, dev, text_embeds=cond, seed=3)
                    ffl = model.engine(Bn, dev).conv_flops()
                else:
                    f = lambda: mod.sample(None, model, fp, dev, n_samples=Bn, y=cond, seed=3)
                    ffl = model.engine(Bn, dev, training=False).fwd_flops
                f()
                torch.cuda.synchronize()
                e0.record()
                for _ in range(2):
                    f()
                e1.record()
                torch.cuda.synchronize()
                us = e0.elapsed_time(e1) / 2 / T * 1e3
                rec["sampler"] = {"reverse_step_us": us, "samples_per_sec_1000step": Bn / (us * 1e-6 * 1000),
                                  "model_tflops": ffl / (us * 1e-6) / 1e12,
                                  "frac_of_sustained_bf16": ffl / (us * 1e-6) / 1e12 / peaks["bf16_sustained"]}
                del ts, model
                torch.cuda.empty_cache()
            except Exception as e:              # keep the sweep going; the failure is part of the record
                rec["error"] = f"{type(e).__name__}: {str(e)[:200]}"
            res.append(rec)
    return res


class _StdoutGuard:
    """Keep stdout clean for the ONE JSON line: libraries (NCCL's version banner, ...) write to fd 1,
    so fd 1 is pointed at stderr for the whole run and the line goes to the saved descriptor."""

    def __enter__(self):
        sys.stdout.flush()
        self.real = os.dup(1)
        os.dup2(2, 1)
        return self

    def emit(self, text: str):
        sys.stdout.flush()
        os.write(self.real, (text + "\n").encode())

    def __exit__(self, *exc):
        sys.stdout.flush()
        os.dup2(self.real, 1)
        os.close(self.real)


_GUARD = None


def emit_line(line: dict):
    text = json.dumps(line)
    if _GUARD is not None:
        _GUARD.emit(text)
    else:
        print(text, flush=True)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="tinydiff", choices=["tinydiff", "reference"])
    ap.add_argument("--precision", default="bf16", choices=["bf16", "fp32"])
    ap.add_argument("--cpu-reverse-steps", type=int, default=48)
    ap.add_argument("--ref-reverse-steps", type=int, default=10)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-gpu-eager", action="store_true", help="skip the eager cuDNN/ATen baseline on the GPU")
    ap.add_argument("--no-configs", action="store_true", help="skip the BASELINE configs 3-5 block")
    ap.add_argument("--train-steps", type=int, default=20, help="train steps per bench step")
    ap.add_argument("--cpu-train-steps", type=int, default=3)
    args = ap.parse_args()
    global _GUARD
    with _StdoutGuard() as g:
        _GUARD = g
        if args.impl == "reference":
            run_reference(args)
        else:
            run_gpu(args)
        _GUARD = None


if __name__ == "__main__":
    main()
