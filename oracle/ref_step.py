"""Timed reference arm (measurement infrastructure, never on the product path): one appearance-optimisation
step of ``Runner.train_clip`` (AvatarGen/AppearanceGen/main.py:345-566) with the reference's OWN, UNMODIFIED
ray-march -- ``models/{embedder,fields,renderer}.py`` staged byte for byte into the git-ignored ``oracle/_ref``
by ``oracle/make_ref.py`` -- around which the parts of the step that cannot be imported here are the pinned
restatements of this package:

    reference NeuSRenderer.render (unmodified file)            renderer.py:302-397
    -> oracle.loss.shading_and_losses / total_loss             main.py:417-534 (pinned by oracle/pin_loss_stage.py)
    -> oracle.clip_vit (CLIP ViT-B/32 stand-in, unpinned)       main.py:509-526
    -> loss.backward() (autograd, incl. the double backward through SDFNetwork.gradient) -> torch.optim.Adam

``device='cpu'``: the reference on the host cores (fp32), what ``bench.py --impl reference`` and the native
line's ``cpu_baseline`` time.  ``device='cuda'``: the same files under
``torch.set_default_tensor_type('torch.cuda.FloatTensor')`` exactly as the reference runs on a GPU
(main.py:948; required because renderer.py:103,211 use the legacy ``torch.Tensor([...])`` constructor), with the
CLIP stand-in cast to fp16 like ``clip.load`` does on CUDA -- the "reference single-GPU PyTorch" denominator of
BASELINE.json's north_star (``ref_gpu`` in the native line).
"""
from __future__ import annotations

from typing import Dict

import torch

from . import clip_vit, loss as oloss
from .make_ref import load_reference_models


class default_cuda_tensors:
    """Context manager: main.py:948 ``torch.set_default_tensor_type('torch.cuda.FloatTensor')`` and back."""

    def __init__(self, on: bool):
        self.on = on

    def __enter__(self):
        if self.on:
            torch.set_default_tensor_type("torch.cuda.FloatTensor")

    def __exit__(self, *a):
        if self.on:
            torch.set_default_tensor_type("torch.FloatTensor")


class ReferenceTrainer:
    def __init__(self, sdf_kw: dict, col_kw: dict, ren_kw: dict, sdf_state: Dict[str, torch.Tensor],
                 col_state: Dict[str, torch.Tensor], variance: float, clip_state: Dict[str, torch.Tensor],
                 text_emb: torch.Tensor, lr: float = 5e-4, igr_weight: float = 0.1, mask_weight: float = 0.5,
                 clip_weight: float = 1.0, device: str = "cpu", clip_half: bool = False):
        ref_fields, ref_renderer = load_reference_models()
        self.device = torch.device(device)
        self.cuda = self.device.type == "cuda"
        with default_cuda_tensors(self.cuda):
            self.sdf = ref_fields.SDFNetwork(**sdf_kw)
            self.col = ref_fields.RenderingNetwork(**col_kw)
            self.var = ref_fields.SingleVarianceNetwork(float(variance))
            self.sdf.load_state_dict({k: v.to(self.device) for k, v in sdf_state.items()})
            self.col.load_state_dict({k: v.to(self.device) for k, v in col_state.items()}, strict=False)
            self.sdf, self.col, self.var = self.sdf.to(self.device), self.col.to(self.device), self.var.to(self.device)
            self.renderer = ref_renderer.NeuSRenderer(None, self.sdf, self.var, self.col, **ren_kw)
            params = list(self.sdf.parameters()) + list(self.var.parameters()) + list(self.col.parameters())   # main.py:141-143
            self.opt = torch.optim.Adam(params, lr=lr)
        self.clip_dtype = torch.float16 if clip_half else torch.float32
        self.clip_state = {k: v.to(self.device, self.clip_dtype) for k, v in clip_state.items()}
        self.text = text_emb.float().reshape(-1, text_emb.shape[-1]).to(self.device)
        if self.text.shape[0] == 1:
            self.text = self.text.expand(2, -1)
        self.w = (igr_weight, mask_weight, clip_weight)

    def named_params(self):
        return ([("sdf." + k, v) for k, v in self.sdf.named_parameters()]
                + [("col." + k, v) for k, v in self.col.named_parameters()]
                + [("var." + k, v) for k, v in self.var.named_parameters()])

    def _cos(self, canvas, text):
        emb = clip_vit.encode_image(self.clip_state, clip_vit.preprocess(canvas).to(self.clip_dtype)).float()
        return torch.cosine_similarity(emb.mean(0), text.reshape(-1, text.shape[-1]).mean(0), dim=0)

    def loss(self, view, cos_anneal: float = 1.0, jitter_seed=None):
        """``jitter_seed``: seed torch's RNG so that the draw inside the unmodified ``render`` (renderer.py:317-319,
        ``torch.rand([batch_size, 1]) - 0.5``) is reproducible by the caller; None = whatever the RNG gives."""
        dev = self.device
        igr, mw, cw = self.w
        with default_cuda_tensors(self.cuda):
            o, d = view.rays_o.to(dev), view.rays_d.to(dev)
            near, far = view.near.to(dev).reshape(-1, 1), view.far.to(dev).reshape(-1, 1)
            if view.bg_choice == 0:
                bg = torch.ones([1, 3])
            elif view.bg_choice in (1, 2):
                bg = view.ray_background.to(dev).reshape(-1, 1)
            else:
                bg = None
            if jitter_seed is not None:
                torch.manual_seed(jitter_seed)
            out = self.renderer.render(o, d, near, far, background_rgb=bg, cos_anneal_ratio=cos_anneal)
            H, W = view.H, view.W
            dm = view.in_mask.reshape(H, W).bool().to(dev)
            cbg = None if view.canvas_background is None else view.canvas_background.to(dev).reshape(-1, 1)
            stage = oloss.shading_and_losses(out, dm, H, W, view.true_rgb.to(dev), view.mask.to(dev).reshape(-1, 1),
                                             torch.as_tensor(view.light_dir, dtype=torch.float32).to(dev),
                                             view.ambience, background_choice=view.bg_choice, background_rgb=cbg,
                                             igr_weight=igr, mask_weight=mw)
            cos_t = self._cos(stage["texture_canvas"], self.text[0])
            cos_s = self._cos(stage["shading_canvas"], self.text[1])
            total = oloss.total_loss(stage, cos_t, cos_s, cw)
        return total, {"out": out, "stage": stage, "cos": torch.stack([cos_t, cos_s])}

    def step(self, view, cos_anneal: float = 1.0, jitter_seed=None):
        """zero_grad / backward / Adam.step (main.py:536-538).  Returns (loss, aux)."""
        total, aux = self.loss(view, cos_anneal, jitter_seed)
        with default_cuda_tensors(self.cuda):
            self.opt.zero_grad()
            total.backward()
            self.opt.step()
        return total.detach(), aux
