"""CPU oracle (test infrastructure, never on the product path): the shading / canvas / loss stage of
``Runner.train_clip`` (AvatarGen/AppearanceGen/main.py:417-534) restated with plain torch ops, plus the
camera helpers it needs (models/utils.py:9-70, models/dataset.py:277-293).

Every random draw of the reference (light direction jitter, ambience, background choice) is an explicit
argument so that both sides of a parity test see identical values.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import numpy as np
import torch
import torch.nn.functional as F


# ----------------------------------------------------------------------------- cameras / rays
def lookat(eye, at, up=(0.0, 1.0, 0.0)):
    """models/utils.py:9-27: camera-to-world matrix with columns (x, y, z axes, eye)."""
    eye, at, up = (np.asarray(v, dtype=np.float64) for v in (eye, at, up))
    z = eye - at
    z = z / np.linalg.norm(z)
    x = np.cross(up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
    return m


def sphere_coord(theta, phi, r=1.0):
    """models/utils.py:59-64."""
    return np.array([r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(theta)])


def pinhole_rays(pose, H, W, full_res=256, fov=math.pi / 3):
    """gen_rays_pose / gen_rays_silhouettes (models/dataset.py:252-293): pixel grid linspace(0, full-1, n),
    p = ((x-cx)/f, -(y-cy)/f, -1) normalised, rotated by pose[:3,:3]; origin pose[:3,3].  Returns
    rays_o, rays_d as [H, W, 3] (row = y)."""
    f = 0.5 * full_res / math.tan(0.5 * fov)
    cx = cy = 0.5 * full_res
    tx = torch.linspace(0, full_res - 1, W)
    ty = torch.linspace(0, full_res - 1, H)
    px, py = torch.meshgrid(tx, ty, indexing="ij")
    px, py = px.t(), py.t()
    p = torch.stack([(px - cx) / f, -(py - cy) / f, -torch.ones_like(px)], -1).float()
    v = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    pose = torch.as_tensor(pose, dtype=torch.float32)
    v = torch.sum(v[..., None, :] * pose[:3, :3], -1)
    o = pose[None, None, :3, 3].expand(v.shape)
    return o.contiguous(), v.contiguous()


# ----------------------------------------------------------------------------- loss stage
def shading_and_losses(render_out: Dict[str, torch.Tensor], dilated_mask: torch.Tensor, H: int, W: int,
                       true_rgb: torch.Tensor, mask: torch.Tensor, light_dir: torch.Tensor, ambience: float,
                       background_choice: int = 3, background_rgb: Optional[torch.Tensor] = None,
                       igr_weight: float = 0.1, mask_weight: float = 0.1, add_no_texture: bool = True,
                       texture_cast_light: bool = True):
    """main.py:417-497 with use_silhouettes=True.  ``add_no_texture`` / ``texture_cast_light`` are the switches of
    main.py:425,466,509-526 (both True in confs/examples*, switched off in confs/ablation/*_0..2.conf): "texture_canvas" is
    the image the first CLIP term sees (texture_shading, or the extra-colour canvas of :475-477 without cast light),
    "shading_canvas" the one of the second term (None without add_no_texture).

    render_out : dict from NeuSRenderer.render for the R rays of the dilated mask (row-major order of True pixels)
    dilated_mask : bool [H, W] with R True entries
    true_rgb : [H*W, 3] template render resized to the canvas; mask : [H*W, 1] (already thresholded, :407-410)
    light_dir : [3] (the sphere_coord(...) of :433); ambience : scalar (:440)
    background_choice : 0 white, 1/2 per-pixel grey given in background_rgb [H*W,1], 3 black (:387-415)
    Returns dict with the two CLIP input canvases [H,W,3], the non-CLIP loss terms and psnr.
    """
    color_fine = render_out["color_fine"]
    extra = render_out["extra_color_fine"]
    dt = color_fine.dtype
    shading = add_no_texture or texture_cast_light                       # :425
    # ---- cast light (:425-453)
    normals = (render_out["gradients"] * render_out["weights"][:, :, None]).sum(dim=1)
    normals = normals / (torch.norm(normals, dim=-1, keepdim=True) + 1e-7)
    ld = torch.zeros_like(normals) + light_dir.to(dt)
    ld = ld / (torch.norm(ld, dim=-1, keepdim=True) + 1e-7)
    diffuse_shading = (normals * ld).sum(-1, keepdim=True).clamp(min=0, max=1)
    diffuse_shading = torch.where(torch.isnan(diffuse_shading), torch.ones_like(diffuse_shading), diffuse_shading)
    rand_shading = ambience + (1 - ambience) * diffuse_shading
    weight_sum_r = render_out["weight_sum"].reshape(-1)
    low = weight_sum_r < 0.5
    rand_shading_rgb = rand_shading.reshape(-1, 1).repeat(1, 3)
    rand_shading_rgb = torch.where(low[:, None], extra, rand_shading_rgb)          # :448
    rand_shading = torch.where(low[:, None], torch.ones_like(rand_shading), rand_shading)  # :450-452 (l_ratio = 1)
    texture_shading = (extra * rand_shading).clamp(min=0, max=1)
    # ---- canvases (:461-487)
    background = torch.zeros(H, W, 3, dtype=dt)
    if background_choice == 0:
        background[:] = 1
    elif background_choice in (1, 2):
        bgf = background_rgb.reshape(H, W, 1).repeat(1, 1, 3)
        background = torch.where(dilated_mask[..., None], background, bgf)

    def scatter(base, vals):
        out = base.clone()
        out[dilated_mask] = vals
        return out

    full_texture = scatter(background, texture_shading if texture_cast_light else extra)     # :466-477, :509-520
    full_shading = scatter(background, rand_shading_rgb) if (shading and add_no_texture) else None
    full_color = scatter(torch.zeros(H, W, 3, dtype=dt), color_fine).reshape(-1, 3)
    full_wsum = scatter(torch.zeros(H, W, 1, dtype=dt), render_out["weight_sum"]).reshape(-1, 1)
    # ---- losses (:489-497)
    mask_sum = mask.sum() + 1e-5
    color_error = (full_color - true_rgb) * mask
    color_loss = F.l1_loss(color_error, torch.zeros_like(color_error), reduction="sum") / mask_sum
    psnr = 20.0 * torch.log10(1.0 / (((full_color - true_rgb) ** 2 * mask).sum() / (mask_sum * 3.0)).sqrt())
    eikonal = render_out["gradient_error"]
    mask_loss = F.binary_cross_entropy(full_wsum.clip(1e-3, 1.0 - 1e-3), mask)
    base = color_loss + eikonal * igr_weight + mask_loss * mask_weight
    return {"texture_canvas": full_texture, "shading_canvas": full_shading, "color_loss": color_loss,
            "eikonal_loss": eikonal, "mask_loss": mask_loss, "base_loss": base, "psnr": psnr}


def total_loss(stage: Dict[str, torch.Tensor], cosine_texture, cosine_shading, clip_weight: float = 1.0):
    """main.py:528-534; ``cosine_shading`` None = add_no_texture off (no second CLIP term)."""
    loss = stage["base_loss"] + (1.0 - cosine_texture) * clip_weight
    return loss if cosine_shading is None else loss + (1.0 - cosine_shading) * clip_weight


# ----------------------------------------------------------------------------- schedules
def learning_rate_factor(iter_step, warm_up_end, end_iter, alpha):
    """Runner.update_learning_rate, main.py:577-586."""
    if iter_step < warm_up_end:
        return iter_step / warm_up_end
    progress = (iter_step - warm_up_end) / (end_iter - warm_up_end)
    return (np.cos(np.pi * progress) + 1.0) * 0.5 * (1 - alpha) + alpha


def cos_anneal_ratio(iter_step, anneal_end):
    """Runner.get_cos_anneal_ratio, main.py:571-575."""
    return 1.0 if anneal_end == 0.0 else float(np.min([1.0, iter_step / anneal_end]))
