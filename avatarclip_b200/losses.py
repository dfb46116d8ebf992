"""Host-side wrapper of the shading / canvas / loss-stage kernels (``avc_loss_stage_*``), i.e. of
AvatarGen/AppearanceGen/main.py:417-497,528-534 for use_silhouettes = True (every shipped train_clip conf), with the
add_no_texture / texture_cast_light switches of the ablation confs.  ``shade_and_losses`` is an autograd function so the
stage can sit between ``NeuSRenderer.render`` and the CLIP tower in an unmodified training loop; the fused
train step (``avatarclip_b200.trainer``) calls the same C entry points directly.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from ._lib import NeusCotangents

LOSS_SCALARS = 16
S_COLOR, S_EIK, S_BCE, S_PSNR, S_BASE = 0, 1, 2, 3, 4


class LossInputs(C.Structure):
    _fields_ = [(k, C.c_void_p) for k in ("color_fine", "extra_color_fine", "gradients", "weights", "weight_sum",
                                          "gradient_error", "pix", "in_mask", "true_rgb", "mask", "background")] + \
               [("bg_choice", C.c_int32), ("light_dir", C.c_float * 3), ("ambience", C.c_float),
                ("view_scalars", C.c_void_p), ("igr_weight", C.c_float), ("mask_weight", C.c_float), ("clip_weight", C.c_float),
                ("R", C.c_int32), ("S", C.c_int32), ("H", C.c_int32), ("W", C.c_int32),
                ("plain_texture", C.c_int32), ("no_shading_term", C.c_int32)]


def _bind(L):
    if getattr(L, "_loss_bound", False):
        return
    P = C.POINTER
    L.avc_loss_stage_fwd.argtypes = [P(LossInputs), C.c_void_p, C.c_void_p, C.c_void_p]
    L.avc_loss_stage_bwd.argtypes = [P(LossInputs), C.c_void_p, C.c_void_p, P(NeusCotangents), C.c_void_p]
    L.avc_loss_stage_fwd.restype = C.c_int
    L.avc_loss_stage_bwd.restype = C.c_int
    L._loss_bound = True


class StepInputs:
    """Per-step, non-differentiable inputs of the stage (all CUDA tensors)."""

    def __init__(self, pix, in_mask, true_rgb, mask, H, W, light_dir, ambience, bg_choice=3, background=None,
                 igr_weight=0.1, mask_weight=0.5, clip_weight=1.0, view_scalars=None,
                 texture_cast_light: bool = True, add_no_texture: bool = True):
        self.pix = pix.to(torch.int32).contiguous()
        self.in_mask = in_mask.to(torch.uint8).contiguous().reshape(-1)
        self.true_rgb = true_rgb.float().contiguous().reshape(-1, 3)
        self.mask = mask.float().contiguous().reshape(-1)
        self.background = None if background is None else background.float().contiguous().reshape(-1)
        self.H, self.W = int(H), int(W)
        self.light_dir = [float(v) for v in light_dir]
        self.ambience = float(ambience)
        self.bg_choice = int(bg_choice)
        # optional DEVICE [4] copy of (light_dir, ambience): read by the kernels instead of the host values, so a
        # captured CUDA graph of the step can be replayed on a new view
        self.view_scalars = view_scalars
        self.igr_weight, self.mask_weight, self.clip_weight = float(igr_weight), float(mask_weight), float(clip_weight)
        # the two switches the shipped ablation confs turn off (main.py:509-526): without texture_cast_light canvas 0 is
        # the un-shaded extra colour; without add_no_texture the shading canvas carries no CLIP term
        self.texture_cast_light, self.add_no_texture = bool(texture_cast_light), bool(add_no_texture)


def make_inputs(render_out: Dict[str, torch.Tensor], si: StepInputs) -> LossInputs:
    R, S = render_out["weights"].shape
    li = LossInputs()
    for k in ("color_fine", "extra_color_fine", "gradients", "weights", "weight_sum", "gradient_error"):
        setattr(li, k, render_out[k].data_ptr())
    li.pix, li.in_mask = si.pix.data_ptr(), si.in_mask.data_ptr()
    li.true_rgb, li.mask = si.true_rgb.data_ptr(), si.mask.data_ptr()
    li.background = None if si.background is None else si.background.data_ptr()
    li.bg_choice = si.bg_choice
    li.light_dir = (C.c_float * 3)(*si.light_dir)
    li.ambience, li.igr_weight, li.mask_weight, li.clip_weight = si.ambience, si.igr_weight, si.mask_weight, si.clip_weight
    li.view_scalars = None if si.view_scalars is None else si.view_scalars.data_ptr()
    li.R, li.S, li.H, li.W = R, S, si.H, si.W
    li.plain_texture, li.no_shading_term = int(not si.texture_cast_light), int(not si.add_no_texture)
    return li


def stage_forward(render_out, si: StepInputs):
    """-> canvases [2,H,W,3] (texture_shading, rand_shading_rgb), scalars [16]."""
    L = _lib.lib()
    _bind(L)
    dev = render_out["weights"].device
    canv = torch.empty(2, si.H, si.W, 3, dtype=torch.float32, device=dev)
    scal = torch.empty(LOSS_SCALARS, dtype=torch.float32, device=dev)
    li = make_inputs(render_out, si)
    _lib.check(L.avc_loss_stage_fwd(C.byref(li), _lib.ptr(canv), _lib.ptr(scal), _lib.stream_ptr()), "avc_loss_stage_fwd")
    return canv, scal


def stage_backward(render_out, si: StepInputs, d_canvases, scalars):
    """-> dict of cotangents for NeuSRenderer.render outputs."""
    L = _lib.lib()
    _bind(L)
    cot = {k: torch.empty_like(render_out[k]) for k in ("color_fine", "extra_color_fine", "gradients", "weights",
                                                         "weight_sum", "gradient_error")}
    c = NeusCotangents(**{k: v.data_ptr() for k, v in cot.items()})
    li = make_inputs(render_out, si)
    _lib.check(L.avc_loss_stage_bwd(C.byref(li), _lib.ptr(d_canvases), _lib.ptr(scalars), C.byref(c),
                                    _lib.stream_ptr()), "avc_loss_stage_bwd")
    return cot


class _StageFn(torch.autograd.Function):
    @staticmethod
    def forward(ctx, si, color, extra, gradients, weights, weight_sum, gerr):
        ro = {"color_fine": color.contiguous(), "extra_color_fine": extra.contiguous(),
              "gradients": gradients.contiguous(), "weights": weights.contiguous(),
              "weight_sum": weight_sum.contiguous(), "gradient_error": gerr.contiguous()}
        canv, scal = stage_forward(ro, si)
        ctx.si, ctx.ro, ctx.scal = si, ro, scal
        return canv, scal[S_BASE].clone(), scal.clone()

    @staticmethod
    def backward(ctx, g_canv, g_base, g_scal):
        # cotangents are linear in (d_canvases, g_base): run the kernel on d_canvases and scale the direct
        # loss terms by g_base (== 1 in the reference's loss)
        d_canv = g_canv.contiguous().float() if g_canv is not None else torch.zeros(2, ctx.si.H, ctx.si.W, 3, device=ctx.scal.device)
        cot = stage_backward(ctx.ro, ctx.si, d_canv, ctx.scal)
        gb = 1.0 if g_base is None else g_base
        if g_base is not None:
            zero = stage_backward(ctx.ro, ctx.si, torch.zeros_like(d_canv), ctx.scal)   # direct terms only
            for k in cot:
                cot[k] = (cot[k] - zero[k]) + zero[k] * gb
        return (None, cot["color_fine"], cot["extra_color_fine"], cot["gradients"], cot["weights"],
                cot["weight_sum"], cot["gradient_error"])


def shade_and_losses(render_out, si: StepInputs):
    """Differentiable stage: returns (canvases [2,H,W,3], base_loss scalar, scalars [16] (no grad))."""
    canv, base, scal = _StageFn.apply(si, render_out["color_fine"], render_out["extra_color_fine"],
                                      render_out["gradients"], render_out["weights"], render_out["weight_sum"],
                                      render_out["gradient_error"])
    return canv, base, scal.detach()
