"""Synthetic appearance-optimisation views (host side, numpy / CPU torch only).

The reference draws a random camera per step, rasterises the SMPL template with neural_renderer and keeps the
rays inside the dilated silhouette (AvatarGen/AppearanceGen/main.py:348-385).  neural_renderer, SMPL and the
template raster are not available here, so measurements and parity tests use a *synthetic* view with the same
data layout: the R rays are the pixels of a disc-shaped "dilated mask" in an H x W canvas,
the "template render" is a smaller white disc, cameras sit on a circle around the origin.  Ray generation follows
``SMPL_Dataset.gen_rays_silhouettes`` / ``near_far_from_sphere`` (models/dataset.py:252-275,331-342) and
``lookat`` (models/utils.py:9-27).
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Optional

import numpy as np
import torch


VIEW_FIELDS = ("rays_o", "rays_d", "near", "far", "jitter", "pix", "in_mask", "true_rgb", "mask", "ray_background",
               "canvas_background", "scalars")


@dataclass
class HostView:
    """Everything one train_clip step consumes, as CPU tensors (pinned when requested)."""
    rays_o: torch.Tensor       # [R,3]
    rays_d: torch.Tensor       # [R,3]
    near: torch.Tensor         # [R]
    far: torch.Tensor          # [R]
    jitter: torch.Tensor       # [R]   (u - 0.5) of renderer.py:317-319
    pix: torch.Tensor          # [R] int32 canvas index
    in_mask: torch.Tensor      # [H*W] uint8
    true_rgb: torch.Tensor     # [H*W,3]
    mask: torch.Tensor         # [H*W]
    ray_background: Optional[torch.Tensor]   # [R] grey per ray (bg_choice 1/2) or None
    canvas_background: Optional[torch.Tensor]  # [H*W] grey per pixel or None
    bg_choice: int
    light_dir: np.ndarray      # [3]
    ambience: float
    H: int
    W: int
    scalars: Optional[torch.Tensor] = None   # [4] = light_dir, ambience (what the loss stage reads on the device)
    flat: Optional[torch.Tensor] = None      # uint8 buffer all the tensors above are views of, once pack() ran

    def __post_init__(self):
        if self.scalars is None:
            self.scalars = torch.tensor([float(self.light_dir[0]), float(self.light_dir[1]), float(self.light_dir[2]),
                                         float(self.ambience)], dtype=torch.float32)

    def layout(self):
        """[(field, byte offset, shape, dtype, bytes)], total bytes: one 256 B-aligned slot per present tensor."""
        items, off = [], 0
        for name in VIEW_FIELDS:
            t = getattr(self, name)
            if t is None:
                continue
            nb = t.numel() * t.element_size()
            items.append((name, off, tuple(t.shape), t.dtype, nb))
            off = (off + nb + 255) // 256 * 256
        return items, off

    def pack(self, pin: bool = False) -> "HostView":
        """Move every tensor into ONE contiguous (optionally pinned) buffer so a step's inputs are a single H2D copy."""
        if self.flat is not None:
            return self
        items, total = self.layout()
        flat = torch.zeros(total, dtype=torch.uint8)
        if pin and torch.cuda.is_available():
            flat = flat.pin_memory()
        for name, off, shape, dtype, nb in items:
            dst = flat[off:off + nb].view(dtype).reshape(shape)
            dst.copy_(getattr(self, name))
            setattr(self, name, dst)
        self.flat = flat
        return self

    def tensors(self):
        return [t for t in (getattr(self, n) for n in VIEW_FIELDS) if t is not None]

    def h2d_bytes(self) -> int:
        return sum(t.numel() * t.element_size() for t in self.tensors())


def lookat(eye, at, up=(0.0, 1.0, 0.0)):
    """models/utils.py:9-27 (camera-to-world)."""
    eye, at, up = (np.asarray(v, dtype=np.float64) for v in (eye, at, up))
    z = eye - at
    z /= np.linalg.norm(z)
    x = np.cross(up, z)
    x /= np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
    return m


def disc_mask(H: int, W: int, n_pixels: int) -> np.ndarray:
    """Boolean [H,W] with exactly n_pixels True: the pixels nearest to the canvas centre."""
    yy, xx = np.meshgrid(np.arange(H), np.arange(W), indexing="ij")
    d2 = (yy - (H - 1) / 2.0) ** 2 + (xx - (W - 1) / 2.0) ** 2
    order = np.argsort(d2.reshape(-1), kind="stable")
    m = np.zeros(H * W, dtype=bool)
    m[order[:n_pixels]] = True
    return m.reshape(H, W)


def make_view(index: int, n_rays: int = 512, H: int = 224, W: int = 224, seed: int = 0, bg_choice: int = 3,
              radius: float = 1.5, full_res: int = 256, fov: float = math.pi / 3, pin: bool = False) -> HostView:
    rng = np.random.RandomState(seed * 100003 + index)
    ang = 2.0 * math.pi * (index % 8) / 8.0 + 0.05 * rng.randn()
    eye = np.array([radius * math.sin(ang), 0.15 * rng.randn(), radius * math.cos(ang)])
    pose = lookat(eye, np.zeros(3))
    # pixel grid of gen_rays_silhouettes (dataset.py:259-268)
    f = 0.5 * full_res / math.tan(0.5 * fov)
    tx = torch.linspace(0, full_res - 1, W)
    ty = torch.linspace(0, full_res - 1, H)
    px, py = torch.meshgrid(tx, ty, indexing="ij")
    px, py = px.t(), py.t()
    p = torch.stack([(px - 0.5 * full_res) / f, -(py - 0.5 * full_res) / f, -torch.ones_like(px)], -1).float()
    v = p / torch.linalg.norm(p, ord=2, dim=-1, keepdim=True)
    pose_t = torch.from_numpy(pose).float()
    v = torch.sum(v[..., None, :] * pose_t[:3, :3], -1)                    # [H,W,3]
    o = pose_t[None, None, :3, 3].expand(v.shape)
    dm = disc_mask(H, W, n_rays)
    dmt = torch.from_numpy(dm)
    rays_d = v[dmt].contiguous()
    rays_o = o[dmt].contiguous()
    # near / far (dataset.py:331-342)
    a = (rays_d ** 2).sum(-1)
    b = 2.0 * (rays_o * rays_d).sum(-1)
    mid = 0.5 * (-b) / a
    near = (mid - 1.0).clamp(min=0.0)
    far = mid + 1.0
    # "template render": white inner disc; mask = render != 0 (main.py:361-364,377-380)
    inner = disc_mask(H, W, max(1, int(n_rays * 0.6)))
    true_rgb = torch.from_numpy(np.repeat(inner.reshape(-1, 1).astype(np.float32), 3, axis=1))
    mask = torch.from_numpy(inner.reshape(-1).astype(np.float32))
    pix = torch.from_numpy(np.nonzero(dm.reshape(-1))[0].astype(np.int32))
    jitter = torch.from_numpy(rng.rand(n_rays).astype(np.float32) - 0.5)
    canvas_bg = ray_bg = None
    if bg_choice in (1, 2):
        g = np.clip(0.5 + 0.2 * rng.randn(H * W), 0, 1).astype(np.float32)   # main.py:394-396
        canvas_bg = torch.from_numpy(g)
        ray_bg = canvas_bg[pix.long()].contiguous()
    theta, phi = 0.3 * rng.randn(), ang + rng.uniform(-math.pi / 4, math.pi / 4)
    light = np.array([math.sin(theta + math.pi / 2) * math.sin(phi), math.cos(theta + math.pi / 2), math.sin(theta + math.pi / 2) * math.cos(phi)])
    hv = HostView(rays_o, rays_d, near.contiguous(), far.contiguous(), jitter, pix,
                  torch.from_numpy(dm.reshape(-1).astype(np.uint8)), true_rgb, mask, ray_bg, canvas_bg, bg_choice,
                  light.astype(np.float32), float(rng.uniform(0, 0.2)), H, W)
    return hv.pack(pin=pin)
