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


# ------------------------------------------------------------------------------------------------------------------
# Synthetic weights (seeded) for measurements: neither the real ViT-B-32.pt nor a trained B2-sized NeuS exists on disk
# ------------------------------------------------------------------------------------------------------------------
B2_SDF_KW = dict(d_in=3, d_out=257, d_hidden=256, n_layers=8, skip_in=[4], multires=6, bias=0.5, scale=1.0,
                 geometric_init=True, weight_norm=True)
B2_COL_KW = dict(d_feature=256, mode="no_view_dir", d_in=6, d_out=3, d_hidden=256, n_layers=4, weight_norm=True,
                 multires_view=0, squeeze_out=True, extra_color=True)
B2_REN_KW = dict(n_samples=64, n_importance=64, n_outside=0, up_sample_steps=4, perturb=1.0, extra_color=True)
# the networks of every shipped conf (confs/examples/*.conf; pretrained_models/zero_beta_stand_pose.pth)
S_SDF_KW = dict(B2_SDF_KW, n_layers=4)
S_COL_KW = dict(B2_COL_KW, n_layers=2)
S_REN_KW = dict(B2_REN_KW, n_samples=32, n_importance=32)


def synth_states(sdf_kw: dict, col_kw: dict, seed: int = 0, tame: bool = True):
    """(sdf_state, col_state): the product modules' own constructors (geometric init of models/fields.py:45-63, bit
    identical to the reference under the same torch seed) run under ``torch.manual_seed(seed)``.  ``tame`` damps the raw
    sin/cos columns when the skip concat feeds the LAST linear so that the field stays SDF-like (DESIGN.md "parity
    definition") and perturbs every direction matrix a little (so no gradient is exactly zero)."""
    from .fields import RenderingNetwork, SDFNetwork
    with torch.random.fork_rng(devices=[]):
        torch.manual_seed(seed)
        sdf = SDFNetwork(**sdf_kw)
        col = RenderingNetwork(**col_kw)
        sp = {k: v.detach().clone() for k, v in sdf.state_dict().items()}
        cp = {k: v.detach().clone() for k, v in col.state_dict().items()}
        if tame:
            L = sdf_kw["n_layers"]
            d_enc = 3 * (1 + 2 * sdf_kw["multires"])
            if L in tuple(sdf_kw["skip_in"]):
                sp[f"lin{L}.weight_v"][:, -(d_enc - 3):] *= 0.02
                sp[f"lin{L}.weight_g"] = sp[f"lin{L}.weight_v"].norm(dim=1, keepdim=True)
            for k in sp:
                if k.endswith("weight_v"):
                    sp[k] = sp[k] + 0.01 * torch.randn(sp[k].shape)
    return sp, cp


def random_vit_state(seed: int = 0, width: int = 768, layers: int = 12, patch: int = 32, tokens: int = 50,
                     mlp: int = 3072, out_dim: int = 512):
    """Seeded random CLIP ViT-B/32 visual state dict (openai/CLIP key names, ``visual.`` prefix stripped) with the
    initialisation scales of openai/CLIP, values rounded to fp16 (``clip.load`` keeps fp16 weights on CUDA)."""
    g = torch.Generator().manual_seed(seed)
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    W = width
    sd = {"conv1.weight": rn(W, 3, patch, patch, std=(3 * patch * patch) ** -0.5),
          "class_embedding": rn(W, std=W ** -0.5), "positional_embedding": rn(tokens, W, std=W ** -0.5)}
    for name in ("ln_pre", "ln_post"):
        sd[f"{name}.weight"] = 1.0 + 0.05 * rn(W)
        sd[f"{name}.bias"] = 0.05 * rn(W)
    proj_std, attn_std, fc_std = (W ** -0.5) * ((2 * layers) ** -0.5), W ** -0.5, (2 * W) ** -0.5
    for i in range(layers):
        p = f"transformer.resblocks.{i}."
        for ln in ("ln_1", "ln_2"):
            sd[p + ln + ".weight"] = 1.0 + 0.05 * rn(W)
            sd[p + ln + ".bias"] = 0.05 * rn(W)
        sd[p + "attn.in_proj_weight"] = rn(3 * W, W, std=attn_std)
        sd[p + "attn.in_proj_bias"] = 0.02 * rn(3 * W)
        sd[p + "attn.out_proj.weight"] = rn(W, W, std=proj_std)
        sd[p + "attn.out_proj.bias"] = 0.02 * rn(W)
        sd[p + "mlp.c_fc.weight"] = rn(mlp, W, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = 0.02 * rn(mlp)
        sd[p + "mlp.c_proj.weight"] = rn(W, mlp, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = 0.02 * rn(W)
    sd["proj"] = rn(W, out_dim, std=W ** -0.5)
    return {k: v.half().float() for k, v in sd.items()}


def build_networks(sdf_kw, col_kw, ren_kw, sdf_state, col_state, variance, device, engine: int = 1, chunk: int = 4096):
    """Product modules + renderer from state dicts -> (sdf, col, var, renderer)."""
    from .fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork
    from .renderer import NeuSRenderer
    sdf, col, var = SDFNetwork(**sdf_kw), RenderingNetwork(**col_kw), SingleVarianceNetwork(float(variance))
    sdf.load_state_dict(sdf_state)
    col.load_state_dict(col_state, strict=False)
    sdf, col, var = sdf.to(device), col.to(device), var.to(device)
    ren = NeuSRenderer(None, sdf, var, col, engine=engine, max_rays_per_chunk=chunk, **ren_kw)
    return sdf, col, var, ren


def synthetic_body_mesh(n_lat: int = 24, n_lon: int = 32):
    """A procedural stand-in for the SMPL template (the real one is licence-gated and lives under /root/reference only):
    a union of ellipsoids -- torso, head, two arms, two legs -- inside the unit sphere.  Built y-up / facing +z (the
    NeuS world frame) and returned in the frame of ``Runner.v`` (the posed SMPL output whose root orientation makes it
    z-up: render_one_batch maps it back with (x, y, z) -> (x, z, -y), models/utils.py:115-119), i.e. as (x, -z, y).
    Returns (verts [V,3] float32, faces [F,3] int32); 24 x 32 -> 9 216 triangles."""
    parts = [((0.0, 0.10, 0.0), (0.17, 0.30, 0.11)), ((0.0, 0.55, 0.01), (0.09, 0.11, 0.10)),
             ((-0.30, 0.20, 0.0), (0.05, 0.26, 0.05)), ((0.30, 0.20, 0.0), (0.05, 0.26, 0.05)),
             ((-0.09, -0.50, 0.0), (0.07, 0.36, 0.07)), ((0.09, -0.50, 0.0), (0.07, 0.36, 0.07))]
    vs, fs, base = [], [], 0
    th = np.linspace(0.0, np.pi, n_lat + 1)
    ph = np.linspace(0.0, 2 * np.pi, n_lon, endpoint=False)
    for c, r in parts:
        T, P = np.meshgrid(th, ph, indexing="ij")
        v = np.stack([c[0] + r[0] * np.sin(T) * np.cos(P), c[1] + r[1] * np.cos(T), c[2] + r[2] * np.sin(T) * np.sin(P)], -1)
        vs.append(v.reshape(-1, 3))
        for i in range(n_lat):
            for j in range(n_lon):
                a, b = base + i * n_lon + j, base + i * n_lon + (j + 1) % n_lon
                c2, d = a + n_lon, b + n_lon
                fs.append([a, c2, b])
                fs.append([b, c2, d])
        base += (n_lat + 1) * n_lon
    v = np.concatenate(vs)
    v = np.stack([v[:, 0], -v[:, 2], v[:, 1]], 1)
    return v.astype(np.float32), np.asarray(fs, dtype=np.int32)


SMPL_PARENTS = (-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21)


def synthetic_smpl(V: int = 6890, seed: int = 0):
    """SMPL-shaped random tensors for measurements (SMPL_NEUTRAL.pkl is licence-gated and absent): body-like
    vertices, sparse skin weights (<= 4 joints per vertex, rows sum to 1), J_regressor rows summing to 1, small pose blend
    shapes [207, 3V], the standard 24-joint parent table, a moderate pose [1, 72] (axis-angle)."""
    g = torch.Generator().manual_seed(seed)
    nj = len(SMPL_PARENTS)
    v = torch.randn(V, 3, generator=g) * torch.tensor([0.25, 0.55, 0.12])
    w = torch.zeros(V, nj)
    w.scatter_add_(1, torch.randint(0, nj, (V, 4), generator=g), torch.rand(V, 4, generator=g) + 0.05)
    w = w / w.sum(1, keepdim=True)
    jr = torch.rand(nj, V, generator=g) ** 8
    jr = jr / jr.sum(1, keepdim=True)
    return dict(v_shaped=v[None], pose=torch.randn(1, nj * 3, generator=g) * 0.25,
                posedirs=torch.randn((nj - 1) * 9, V * 3, generator=g) * 1e-3, J_regressor=jr,
                parents=torch.tensor(SMPL_PARENTS, dtype=torch.long), lbs_weights=w)
