"""The appearance-optimisation train step of ``Runner.train_clip`` (AvatarGen/AppearanceGen/main.py:345-566)
as a fused sequence of C-ABI calls -- no autograd graph, no eager PyTorch arithmetic on the hot path:

    H2D(view) -> avc_neus_render_fwd -> avc_loss_stage_fwd -> avc_clip_loss_fwd (texture + shading canvases, B=2)
             -> avc_clip_loss_bwd -> avc_loss_stage_bwd -> avc_neus_render_bwd
             -> [all-reduce of the flat gradient over view-sharded ranks] -> avc_adam_step

Multi-GPU: independent camera views shard one per rank; the only collective is one all-reduce(sum)
of the flat fp32 gradient per step, scaled by 1/world inside the fused Adam.
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Dict, Optional

import torch

from . import _lib, dist as avc_dist, losses
from .renderer import NeuSRenderer, render_backward_raw, render_forward_raw
from .workload import VIEW_FIELDS, HostView


class DeviceView:
    """Device-resident copy of a HostView.  All tensors are views of ONE device buffer with the HostView.pack()
    layout, so ``upload`` is a single (async, when the host side is pinned) H2D copy and the addresses never change
    -- which is what lets a captured CUDA graph of the step be replayed on every new view."""

    def __init__(self, hv: HostView, device):
        self.device = torch.device(device)
        self.H, self.W = hv.H, hv.W
        hv.pack()
        self._layout, total = hv.layout()
        self.flat = torch.empty(total, dtype=torch.uint8, device=self.device)
        for name in VIEW_FIELDS:
            setattr(self, name, None)
        for name, off, shape, dtype, nb in self._layout:
            setattr(self, name, self.flat[off:off + nb].view(dtype).reshape(shape))
        self.upload(hv)

    def upload(self, hv: HostView):
        hv.pack()
        if hv.flat.numel() != self.flat.numel() or hv.layout()[0] != self._layout:
            raise ValueError("DeviceView.upload: the view's tensor shapes differ from the ones this buffer was built for")
        self.flat.copy_(hv.flat, non_blocking=True)
        self.bg_choice, self.light_dir, self.ambience = hv.bg_choice, hv.light_dir, hv.ambience

    def copy_from(self, other: "DeviceView"):
        """Device-to-device refresh (one copy) from another resident view of the same layout."""
        self.flat.copy_(other.flat, non_blocking=True)
        self.bg_choice, self.light_dir, self.ambience = other.bg_choice, other.light_dir, other.ambience


class AppearanceTrainer:
    def __init__(self, renderer: NeuSRenderer, clip_tower, text_emb: torch.Tensor, lr: float = 5e-4,
                 igr_weight: float = 0.1, mask_weight: float = 0.5, clip_weight: float = 1.0,
                 betas=(0.9, 0.999), eps: float = 1e-8, process_group=None, device="cuda",
                 texture_cast_light: bool = True, add_no_texture: bool = True):
        self.renderer, self.clip = renderer, clip_tower
        # train.texture_cast_light / train.add_no_texture (main.py:509-534; off in confs/ablation/*_0..2.conf): without the
        # first, canvas 0 carries the un-shaded extra colour; without the second, the loss has no CLIP term on the shading
        # canvas.  The CLIP batch stays B = 2 in every configuration (one launch structure); the loss stage's backward
        # ignores the second canvas' cotangent when its term is absent.
        self.texture_cast_light, self.add_no_texture = bool(texture_cast_light), bool(add_no_texture)
        self.device = torch.device(device)
        self.fp = renderer.flat_params(self.device)
        self.exp_avg = torch.zeros_like(self.fp.flat)
        self.exp_avg_sq = torch.zeros_like(self.fp.flat)
        self.grad = torch.zeros_like(self.fp.flat)
        self.lr, self.betas, self.eps = lr, betas, eps
        self.igr_weight, self.mask_weight, self.clip_weight = igr_weight, mask_weight, clip_weight
        self.text = text_emb.detach().float().reshape(-1, clip_tower.cfg.out_dim).to(self.device)
        if self.text.shape[0] == 1:
            self.text = self.text.expand(2, -1)
        self.text = self.text.contiguous()          # [2, 512]: texture prompt, no-texture prompt (main.py:500-508)
        self.g_cos = torch.full((2,), -clip_weight, dtype=torch.float32, device=self.device)   # d(1-cos)*w / d cos
        self.pg = process_group
        self.world = 1 if process_group is None else torch.distributed.get_world_size(process_group)
        self.iter_step = 0
        self._out = None
        self._clip_ws = None
        self.cos = torch.zeros(2, dtype=torch.float32, device=self.device)
        self.emb = torch.zeros(2, clip_tower.cfg.out_dim, dtype=torch.float32, device=self.device)
        self.scalars = None
        self.loss = torch.zeros((), dtype=torch.float32, device=self.device)
        self.phase_events = None      # set to [] to record a CUDA event after every phase of the next step
        # CUDA-graph mode (capture() / replay()): Adam's step counter and lr live on the device
        self._graph = None
        self._graph_key = None
        self._graph_loss = None
        self._adam_state = torch.zeros(4, dtype=torch.float32, device=self.device)
        self._graph_has_adam = True
        self._lr_on_device = None
        self._dev_step = -1

    # ------------------------------------------------------------------------------------------
    def forward_backward(self, dv: DeviceView, cos_anneal: float = 1.0) -> torch.Tensor:
        """Everything up to (and including) the flat gradient; returns self.grad."""
        L = _lib.lib()
        r = self.renderer

        def mark(name):
            if self.phase_events is not None:
                e = torch.cuda.Event(enable_timing=True)
                e.record()
                self.phase_events.append((name, e))

        mark("start")
        bg = dv.ray_background if dv.bg_choice in (1, 2) else (torch.ones(3, device=self.device) if dv.bg_choice == 0 else None)
        bg_kind = 2 if dv.bg_choice in (1, 2) else (1 if dv.bg_choice == 0 else 0)
        jit = dv.jitter if r.perturb > 0 else None
        # the output buffers are reused from step to step only while the ray count stays the same (the real loop's
        # silhouettes give a different R every step)
        reuse = self._out if (self._out is not None and self._out["color_fine"].shape[0] == dv.rays_o.shape[0]) else None
        out, ws, chunk = render_forward_raw(r, dv.rays_o, dv.rays_d, dv.near, dv.far, jit, bg, bg_kind, cos_anneal,
                                            None, keep_ws=False, out=reuse)
        self._out = out
        mark("render_fwd")
        si = losses.StepInputs(dv.pix, dv.in_mask, dv.true_rgb, dv.mask, dv.H, dv.W, dv.light_dir, dv.ambience,
                               dv.bg_choice, dv.canvas_background, self.igr_weight, self.mask_weight, self.clip_weight,
                               view_scalars=dv.scalars, texture_cast_light=self.texture_cast_light,
                               add_no_texture=self.add_no_texture)
        canv, scal = losses.stage_forward(out, si)
        self.scalars = scal
        mark("loss_stage_fwd")
        # CLIP on both canvases at once (main.py:509-526): B = 2
        tower = self.clip
        if self._clip_ws is None:
            self._clip_ws = tower._workspace(2)
        cws = self._clip_ws
        _lib.check(L.avc_clip_loss_fwd(C.byref(tower.cfg), C.byref(tower.w), _lib.ptr(canv), dv.H, dv.W, 2, 0,
                                       _lib.ptr(self.text), _lib.ptr(self.emb), _lib.ptr(self.cos), _lib.ptr(cws),
                                       cws.numel(), _lib.stream_ptr()), "avc_clip_loss_fwd")
        mark("clip_fwd")
        d_canv = torch.empty_like(canv)
        _lib.check(L.avc_clip_loss_bwd(C.byref(tower.cfg), C.byref(tower.w), dv.H, dv.W, 2, 0, _lib.ptr(self.text),
                                       _lib.ptr(self.g_cos), None, _lib.ptr(d_canv), _lib.ptr(cws), cws.numel(),
                                       _lib.stream_ptr()), "avc_clip_loss_bwd")
        mark("clip_bwd")
        cot = losses.stage_backward(out, si, d_canv, scal)
        mark("loss_stage_bwd")
        render_backward_raw(r, dv.rays_o, dv.rays_d, bg, bg_kind, cos_anneal, out, ws, chunk, cot, grad=self.grad)
        mark("render_bwd")
        return self.grad

    def set_text(self, emb: torch.Tensor):
        """Select the cached text embedding(s) of the coming step (main.py:499-507: body / face / back prompt) by
        copying into the resident [2,512] buffer (addresses stay fixed: valid under CUDA-graph replay)."""
        e = emb.detach().float().reshape(-1, self.text.shape[1]).to(self.device)
        self.text.copy_(e.expand(2, -1) if e.shape[0] == 1 else e)

    def loss_value(self) -> torch.Tensor:
        """Total loss of main.py:528-534 as a device scalar (one tiny kernel; read it with .item() to sync)."""
        if not self.add_no_texture:                                   # main.py:528-531 without :533-534
            return self.scalars[losses.S_BASE] + (1.0 - self.cos[0]) * self.clip_weight
        return self.scalars[losses.S_BASE] + ((1.0 - self.cos) * self.clip_weight).sum()

    def optimizer_step(self, lr: Optional[float] = None):
        avc_dist.allreduce_sum_(self.grad, self.pg)      # the step's only collective (NCCL over NVLink)
        self.iter_step += 1
        b1, b2 = self.betas
        _lib.check(_lib.lib().avc_adam_step(_lib.ptr(self.fp.flat), _lib.ptr(self.grad), _lib.ptr(self.exp_avg),
                                            _lib.ptr(self.exp_avg_sq), self.fp.n, float(self.lr if lr is None else lr),
                                            b1, b2, self.eps, self.iter_step, 1.0 / self.world, _lib.stream_ptr()),
                   "avc_adam_step")
        if self.phase_events is not None:
            e = torch.cuda.Event(enable_timing=True)
            e.record()
            self.phase_events.append(("allreduce_adam", e))

    def step(self, dv: DeviceView, lr: Optional[float] = None, cos_anneal: float = 1.0) -> torch.Tensor:
        self.forward_backward(dv, cos_anneal)
        self.optimizer_step(lr)
        return self.loss_value()

    # ------------------------------------------------------------------------------------------
    # CUDA-graph mode: the ~370 launches of a step are captured once on a *static* DeviceView and replayed; the
    # caller refreshes that view in place (upload / copy_from) between replays.  Single GPU: the whole step including
    # Adam is in the graph.  View-sharded multi-GPU: the graph ends at the flat gradient, the all-reduce and Adam
    # are launched after it (NCCL stays outside the capture).
    def _graph_body(self, dv: DeviceView, cos_anneal: float, with_adam: bool):
        self.forward_backward(dv, cos_anneal)
        if with_adam:
            if self.world > 1:        # the step's only collective, captured into the graph (NCCL supports stream capture)
                avc_dist.allreduce_sum_(self.grad, self.pg)
            b1, b2 = self.betas
            _lib.check(_lib.lib().avc_adam_step_dev(_lib.ptr(self.fp.flat), _lib.ptr(self.grad),
                                                    _lib.ptr(self.exp_avg), _lib.ptr(self.exp_avg_sq), self.fp.n,
                                                    _lib.ptr(self._adam_state), b1, b2, self.eps, 1.0 / self.world,
                                                    _lib.stream_ptr()), "avc_adam_step_dev")
        return self.loss_value()

    def capture(self, dv: DeviceView, cos_anneal: float = 1.0):
        """Capture the step on ``dv`` (whose buffer addresses, bg_choice and cos_anneal are frozen into the graph)."""
        if self.phase_events is not None:
            raise RuntimeError("capture(): phase events cannot be recorded inside a graph")
        cur = torch.cuda.current_stream()
        side = torch.cuda.Stream(device=self.device)
        side.wait_stream(cur)
        with torch.cuda.stream(side):               # warm-up off the capture: workspaces, tensor maps, attributes
            self.forward_backward(dv, cos_anneal)
            self.loss_value()
        cur.wait_stream(side)
        torch.cuda.synchronize(self.device)
        self._set_device_adam(self.lr)
        # Single GPU: the whole step including Adam is one graph.  View-sharded ranks: the all-reduce and Adam are captured
        # too (AVC_GRAPH_ALLREDUCE=0, or a failing capture, falls back to launching them after the graph).
        import os
        self._graph_has_adam = True
        # (opt-in at N > 1: AVC_GRAPH_ALLREDUCE=1; the default keeps NCCL outside the capture, the structure measured at
        # 2 / 4 / 8 GPUs in round 1)
        want_inside = self.world == 1 or os.environ.get("AVC_GRAPH_ALLREDUCE", "0") == "1"
        if want_inside and self.world > 1:
            # NCCL sets its channels up lazily on the first collective: that must not happen inside a capture
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):
                avc_dist.allreduce_sum_(torch.zeros(1024, device=self.device), self.pg)
            torch.cuda.current_stream().wait_stream(side)
            torch.cuda.synchronize(self.device)
        g = None
        if want_inside:
            try:
                g = torch.cuda.CUDAGraph()
                with torch.cuda.graph(g):
                    self._graph_loss = self._graph_body(dv, cos_anneal, with_adam=True)
            except Exception:
                if self.world == 1:
                    raise
                g = None
                torch.cuda.synchronize(self.device)
        if g is None:
            self._graph_has_adam = False
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                self._graph_loss = self._graph_body(dv, cos_anneal, with_adam=False)
        self._graph, self._graph_key = g, (id(dv), dv.bg_choice, float(cos_anneal))
        return g

    def release_graph(self):
        """Drop the captured step graph.  MUST run before the process group is destroyed when the graph holds the NCCL
        all-reduce: a communicator cannot be torn down while a captured graph still references it
        (`destroy_process_group()` then never returns)."""
        if self._graph is not None:
            torch.cuda.synchronize(self.device)
            self._graph = None
            self._graph_key = None
            self._graph_loss = None
            torch.cuda.synchronize(self.device)

    def _set_device_adam(self, lr: float):
        self._adam_state[0].fill_(float(self.iter_step))
        self._adam_state[1].fill_(float(lr))
        self._lr_on_device = float(lr)
        self._dev_step = self.iter_step

    def replay(self, dv: DeviceView, lr: Optional[float] = None, cos_anneal: float = 1.0) -> torch.Tensor:
        """One step through the captured graph (captures first / again when the frozen arguments changed)."""
        if self._graph is None or self._graph_key != (id(dv), dv.bg_choice, float(cos_anneal)):
            self.capture(dv, cos_anneal)
        lr = self.lr if lr is None else lr
        if self._graph_has_adam:
            if self._dev_step != self.iter_step:      # eager steps ran in between: resynchronise the device counter
                self._set_device_adam(lr)
            elif lr != self._lr_on_device:
                self._adam_state[1].fill_(float(lr))
                self._lr_on_device = float(lr)
            self._graph.replay()
            self.iter_step += 1
            self._dev_step = self.iter_step
        else:
            self._graph.replay()
            self.optimizer_step(lr)
        return self._graph_loss
