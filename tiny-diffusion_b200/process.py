"""ForwardProcess (diffusion.py:165-190) and the ancestral reverse loop (diffusion.py:254-276)
on top of libtinydiff.

``ForwardProcess`` keeps the reference's attribute contract -- ``num_timesteps`` and the fp32 CPU
tensors ``betas`` / ``alphas`` / ``alphas_cumprod`` built by the same op sequence -- and adds
private device-side tables so that q_sample is one kernel and a p_sample step is one kernel whose
timestep comes from a device counter (the loop is CUDA-graph capturable).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib as L


class ForwardProcess:
    def __init__(self, num_timesteps: int = 1000, beta_start: float = 1e-4, beta_end: float = 0.02):
        # diffusion.py:172-175 -- identical op sequence (fp32, CPU)
        self.num_timesteps = num_timesteps
        self.betas = torch.linspace(beta_start, beta_end, num_timesteps)
        self.alphas = 1.0 - self.betas
        self.alphas_cumprod = torch.cumprod(self.alphas, dim=0)
        self._dev: Dict[torch.device, Dict[str, torch.Tensor]] = {}

    # device-side tables -------------------------------------------------------------------
    def _tables(self, device: torch.device) -> Dict[str, torch.Tensor]:
        tab = self._dev.get(device)
        if tab is None:
            # coefficients of diffusion.py:272-274 in the reference's own fp32 op order:
            # (1 - alpha), not beta; sigma_t = sqrt(beta_t)
            c1 = 1 / torch.sqrt(self.alphas)
            c2 = (1 - self.alphas) / torch.sqrt(1 - self.alphas_cumprod)
            c3 = torch.sqrt(self.betas)
            coef = torch.stack([c1, c2, c3, torch.zeros_like(c1)], dim=1).contiguous()
            tab = {"abar": self.alphas_cumprod.to(device).contiguous(), "coef": coef.to(device)}
            self._dev[device] = tab
        return tab

    def q_sample(self, device, x_0: torch.Tensor, t: torch.Tensor, noise: Optional[torch.Tensor] = None):
        """diffusion.py:177-190.  Returns ``(x_t, noise)``.  ``noise`` may be injected; by default it is
        drawn with ``torch.randn_like`` exactly where the reference draws it (:178) so that the
        global RNG stream is consumed identically."""
        device = L.require_device(device)
        x_0 = x_0.to(device=device, dtype=torch.float32).contiguous()
        t = t.to(device=device, dtype=torch.int64).contiguous()
        if noise is None:
            noise = torch.randn_like(x_0)
        noise = noise.to(device=device, dtype=torch.float32).contiguous()
        if x_0.shape[0] != t.shape[0]:
            raise ValueError("t must have one entry per sample")
        x_t = torch.empty_like(x_0)
        B = x_0.shape[0]
        per = x_0.numel() // max(B, 1)
        tab = self._tables(device)
        L.check(L.load().td_qsample(x_0.data_ptr(), noise.data_ptr(), t.data_ptr(), tab["abar"].data_ptr(),
                                    x_t.data_ptr(), B, per, self.num_timesteps, None, L.stream_ptr()), "td_qsample")
        return x_t, noise


class ReverseLoop:
    """x_{t-1} = c1*(x_t - c2*eps) + c3*z for t = T-1 .. 0, one fused kernel per step, the step
    index held on the device.  ``eps_launch`` enqueues the denoiser for the current ``x``/``t``."""

    STEPS_PER_GRAPH = 20      # reverse steps unrolled into one captured graph (fewer, longer replays)

    def __init__(self, process: ForwardProcess, x: torch.Tensor, eps: torch.Tensor, t_dev: torch.Tensor,
                 eps_launch, use_graph: bool = True, guidance: Optional[float] = None, fused_step=None):
        """``guidance`` (extension, classifier-free guidance): x / eps hold a doubled batch, rows [0, n) conditional and
        rows [n, 2n) null-label; every step combines eps_u + w*(eps_c - eps_u) (td_psample_step_cfg)."""
        self.p, self.x, self.eps, self.t_dev, self.eps_launch = process, x, eps, t_dev, eps_launch
        self.use_graph = use_graph
        self.guidance = guidance
        # optional: ``fused_step(z_ptr, z_stride, seed_ptr, coef_ptr, T, ticket_ptr)`` runs denoiser + update + counter in ONE launch
        # (the dense engines' cluster kernel, dense.DenseEngine.fused_step)
        self.fused_step = fused_step if guidance is None else None
        self.n = x.numel() if guidance is None else x.numel() // 2      # elements the noise stream covers
        self.graphs = {}
        self.lib = L.load()
        self.tab = process._tables(x.device)
        self.seed = torch.zeros(2, device=x.device, dtype=torch.int64)
        self.ticket = torch.zeros(1, device=x.device, dtype=torch.int32)      # last-block ticket of the step kernel (t_dev -= 1)
        self._key = None
        self._z_ref = None
        self.launches_per_step = 0

    def _step(self, z_ptr, z_stride, seed_ptr):
        """One reverse step: the denoiser, then the update kernel, whose last block also decrements the step counter."""
        st = L.stream_ptr()
        if self.fused_step is not None:
            self.fused_step(z_ptr, z_stride, seed_ptr, self.tab["coef"].data_ptr(), self.p.num_timesteps, self.ticket.data_ptr())
            return
        self.eps_launch()
        if self.guidance is None:
            L.check(self.lib.td_psample_step_advance(self.x.data_ptr(), self.eps.data_ptr(), z_ptr, z_stride,
                                                     self.tab["coef"].data_ptr(), self.t_dev.data_ptr(), self.x.numel(),
                                                     self.p.num_timesteps, seed_ptr, self.ticket.data_ptr(), st),
                    "td_psample_step_advance")
        else:
            L.check(self.lib.td_psample_step_cfg_advance(self.x.data_ptr(), self.eps.data_ptr(), self.n, float(self.guidance),
                                                         z_ptr, z_stride, self.tab["coef"].data_ptr(), self.t_dev.data_ptr(),
                                                         self.p.num_timesteps, seed_ptr, self.ticket.data_ptr(), st),
                    "td_psample_step_cfg_advance")

    def run(self, z: Optional[torch.Tensor] = None, seed: int = 0, steps: Optional[int] = None) -> None:
        """Run ``steps`` (default all T) reverse steps in place on ``self.x``.
        z: optional injected noise table [T, *x.shape] (row t used at step t; row 0 unused)."""
        T = self.p.num_timesteps
        steps = T if steps is None else steps
        if not 0 <= steps <= T:
            raise ValueError(f"steps must be in [0, {T}] (the loop starts at t = T-1 and ends at t = 0)")
        n = self.n
        if z is not None:
            assert z.is_cuda and z.dtype == torch.float32 and z.is_contiguous() and z.shape[0] == T
            assert z.numel() == T * n
            z_ptr, z_stride, seed_ptr = z.data_ptr(), n, None
        else:
            self.seed[0] = seed
            self.seed[1] = 0
            z_ptr, z_stride, seed_ptr = None, 0, self.seed.data_ptr()
        if not self.use_graph:
            self.t_dev.fill_(T - 1)
            for _ in range(steps):
                self._step(z_ptr, z_stride, seed_ptr)
            return
        key = (z_ptr, z_stride, seed_ptr)
        if self._key != key:
            self.graphs, self._key, self._z_ref = {}, key, z
        self.t_dev.fill_(T - 1)
        unroll = max(1, int(self.STEPS_PER_GRAPH))
        done = 0
        while done < steps:
            k = unroll if steps - done >= unroll else 1
            self._graph(k, z_ptr, z_stride, seed_ptr).replay()
            done += k

    def _graph(self, k: int, z_ptr, z_stride, seed_ptr) -> "torch.cuda.CUDAGraph":
        """The captured graph of ``k`` consecutive reverse steps (the step index lives on the device, so the same graph
        serves every position of the loop)."""
        g = self.graphs.get(k)
        if g is not None:
            return g
        # one eager step first (lazy module loading and cudaFuncSetAttribute are not capturable), then the capture;
        # x and the step counter are restored afterwards
        x_saved, t_saved = self.x.clone(), self.t_dev.clone()
        if not self.graphs:
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._step(z_ptr, z_stride, seed_ptr)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = int(self.lib.td_launch_count())
        with torch.cuda.graph(g):
            for _ in range(k):
                self._step(z_ptr, z_stride, seed_ptr)
        self.launches_per_step = (int(self.lib.td_launch_count()) - n0) // k      # library kernels per reverse step
        self.x.copy_(x_saved)
        self.t_dev.copy_(t_saved)
        self.graphs[k] = g
        return g


class SamplerChains:
    """The reverse loop of ONE ``sample()`` call run as several independent sub-batches ("chains"), each on its own stream
    inside the same captured graph.

    Every sample is independent in eval mode (diffusion.py:256), so splitting the batch changes no result -- it is the
    sample sharding of the multi-GPU path applied inside one GPU.  Why it pays: a reverse step is a dependent chain of ~30
    kernels, 13 of them persistent one-CTA-per-SM convolutions; at batch 128 their (n-tile, subtile) units do not divide by
    148 SMs (896 / 512 / 256 units: the last round runs a handful of CTAs), every launch pays ~6-8 us of pipeline fill and
    drain, and the glue kernels between them leave most SMs idle.  With two chains the other chain's next kernel fills
    those holes: its CTAs start as soon as an SM frees.

    ``chains``: list of dicts with ``x`` / ``eps`` / ``t_dev`` tensors and a ``launch`` callable (one denoiser engine each).
    Injected noise ``z`` is the full table [T, n_total, ...]; chain c reads its slice of every row.  In-kernel Philox noise is
    keyed per chain (seed, c << 40)."""

    STEPS_PER_GRAPH = ReverseLoop.STEPS_PER_GRAPH

    def __init__(self, process: ForwardProcess, chains, use_graph: bool = True):
        self.p, self.chains, self.use_graph = process, chains, use_graph
        self.lib = L.load()
        dev = chains[0]["x"].device
        self.tab = process._tables(dev)
        self.seeds = [torch.zeros(2, device=dev, dtype=torch.int64) for _ in chains]
        self.tickets = [torch.zeros(1, device=dev, dtype=torch.int32) for _ in chains]
        self.sizes = [c["x"].numel() for c in chains]
        self.n_total = sum(self.sizes)
        self.side = [None] + [torch.cuda.Stream(device=dev) for _ in chains[1:]]
        self.graphs = {}
        self._key = None
        self._z_ref = None
        self.launches_per_step = 0

    def _step(self, c: int, z_ptr, z_stride, use_seed: bool):
        ch = self.chains[c]
        st = L.stream_ptr()
        ch["launch"]()
        zp = None if z_ptr is None else z_ptr + 4 * sum(self.sizes[:c])
        L.check(self.lib.td_psample_step_advance(ch["x"].data_ptr(), ch["eps"].data_ptr(), zp, z_stride, self.tab["coef"].data_ptr(),
                                                 ch["t_dev"].data_ptr(), self.sizes[c], self.p.num_timesteps,
                                                 self.seeds[c].data_ptr() if use_seed else None, self.tickets[c].data_ptr(), st),
                "td_psample_step_advance")

    def _steps_all(self, k: int, z_ptr, z_stride, use_seed: bool):
        """k reverse steps of every chain: chain 0 on the current stream, the others on their side streams (forked here and
        joined at the end; inside a capture these are parallel branches of the graph)."""
        main = torch.cuda.current_stream()
        fork = torch.cuda.Event()
        fork.record(main)
        for c in range(len(self.chains)):
            if c == 0:
                for _ in range(k):
                    self._step(0, z_ptr, z_stride, use_seed)
                continue
            self.side[c].wait_event(fork)
            with torch.cuda.stream(self.side[c]):
                for _ in range(k):
                    self._step(c, z_ptr, z_stride, use_seed)
        for c in range(1, len(self.chains)):
            ev = torch.cuda.Event()
            ev.record(self.side[c])
            main.wait_event(ev)

    def run(self, z: Optional[torch.Tensor] = None, seed: int = 0, steps: Optional[int] = None) -> None:
        T = self.p.num_timesteps
        steps = T if steps is None else steps
        if not 0 <= steps <= T:
            raise ValueError(f"steps must be in [0, {T}] (the loop starts at t = T-1 and ends at t = 0)")
        if z is not None:
            assert z.is_cuda and z.dtype == torch.float32 and z.is_contiguous() and z.shape[0] == T
            assert z.numel() == T * self.n_total
            z_ptr, z_stride, use_seed = z.data_ptr(), self.n_total, False
        else:
            for c, s in enumerate(self.seeds):
                s[0] = seed
                s[1] = c << 40
            z_ptr, z_stride, use_seed = None, 0, True
        for ch in self.chains:
            ch["t_dev"].fill_(T - 1)
        if not self.use_graph:
            self._steps_all(steps, z_ptr, z_stride, use_seed)
            return
        key = (z_ptr, z_stride, use_seed)
        if self._key != key:
            self.graphs, self._key, self._z_ref = {}, key, z
        unroll = max(1, int(self.STEPS_PER_GRAPH))
        done = 0
        while done < steps:
            k = unroll if steps - done >= unroll else 1
            self._graph(k, z_ptr, z_stride, use_seed).replay()
            done += k

    def _graph(self, k: int, z_ptr, z_stride, use_seed):
        g = self.graphs.get(k)
        if g is not None:
            return g
        saved = [(ch["x"].clone(), ch["t_dev"].clone()) for ch in self.chains]
        if not self.graphs:                       # one eager step first (lazy module loading, cudaFuncSetAttribute)
            side = torch.cuda.Stream()
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                self._steps_all(1, z_ptr, z_stride, use_seed)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize()
        g = torch.cuda.CUDAGraph()
        n0 = int(self.lib.td_launch_count())
        with torch.cuda.graph(g):
            self._steps_all(k, z_ptr, z_stride, use_seed)
        self.launches_per_step = (int(self.lib.td_launch_count()) - n0) // k
        for ch, (xs, ts) in zip(self.chains, saved):
            ch["x"].copy_(xs)
            ch["t_dev"].copy_(ts)
        self.graphs[k] = g
        return g
