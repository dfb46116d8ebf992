"""Per-step view preparation of ``Runner.train_clip`` (AvatarGen/AppearanceGen/main.py:348-415) on the device.

For every step the reference (1) draws a camera, (2) rasterises the SMPL template with neural_renderer
(``render_one_batch``, models/utils.py:108-125), (3) dilates the silhouette on the HOST with scipy, sizes a square
canvas for ~``max_ray_num`` rays and masks the ray grid (``gen_rays_silhouettes``, models/dataset.py:252-275),
(4) resizes the template render to the canvas (main.py:375-380), (5) draws a background (main.py:387-405).
Here all of it is kernels of libavc_b200.so (csrc/avc_view.cu, csrc/avc_rays.cu) on a SIDE stream:

    submit(draw)   raster -> 21 x 21 box dilation + pixel count                          (async; 4-byte D2H into pinned memory)
    finish(handle) canvas size from the count -> nearest mask resize + ordered compaction (4-byte D2H: ray count R)
                   -> rays / near / far of the R pixels, resized target + mask, background field, per-ray jitter

``Runner.train_clip`` submits the view of step i+1 before it launches step i and finishes it while step i runs, so the
two tiny read-backs never leave the GPU idle (the reference blocks on scipy + several boolean-mask syncs per step).
"""
from __future__ import annotations

import ctypes as C
import math
from typing import Optional

import numpy as np
import torch

from . import _lib
from .sampling import StepDraw

_SIGS_DONE = False


def _sigs():
    global _SIGS_DONE
    L = _lib.lib()
    if _SIGS_DONE:
        return L
    vp, i32, u32, f32, sz = C.c_void_p, C.c_int32, C.c_uint32, C.c_float, C.c_size_t
    fp = C.POINTER(C.c_float)
    L.avc_raster_workspace_bytes.argtypes = [i32, i32, i32, C.POINTER(sz)]
    L.avc_raster_template.argtypes = [vp, vp, i32, i32, fp, fp, i32, i32, vp, vp, vp, sz, vp]
    L.avc_dilate_count.argtypes = [vp, i32, i32, vp, vp, vp]
    L.avc_mask_compact.argtypes = [vp, i32, i32, i32, vp, vp, vp, vp]
    L.avc_view_targets.argtypes = [vp, i32, i32, i32, vp, vp, vp]
    L.avc_background_field.argtypes = [i32, i32, i32, u32, i32, f32, vp, vp, i32, vp, vp]
    L.avc_uniform_fill.argtypes = [u32, i32, f32, f32, vp, vp]
    L.avc_gen_rays.argtypes = [fp, f32, f32, f32, f32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp]
    for n in ("avc_raster_workspace_bytes", "avc_raster_template", "avc_dilate_count", "avc_mask_compact",
              "avc_view_targets", "avc_background_field", "avc_uniform_fill", "avc_gen_rays"):
        getattr(L, n).restype = C.c_int
    _SIGS_DONE = True
    return L


class PreparedView:
    """What one train_clip step consumes, device resident (same attribute names as ``trainer.DeviceView``)."""
    __slots__ = ("rays_o", "rays_d", "near", "far", "jitter", "pix", "in_mask", "true_rgb", "mask", "ray_background",
                 "canvas_background", "scalars", "bg_choice", "light_dir", "ambience", "H", "W", "R", "draw",
                 "template_rgb")


class _Pending:
    __slots__ = ("draw", "rgb", "mask256", "dil", "count_host", "event", "eye", "at")


class ViewBuilder:
    def __init__(self, verts: torch.Tensor, faces, max_ray_num: int, mask_weight: float, device, image_size: int = 256,
                 camera_angle_x: float = math.pi / 3, supersample: int = 2, dilate_iterations: int = 10):
        """``verts`` [V,3] / [1,V,3] template vertices (``Runner.v``, main.py:322-328), ``faces`` [F,3] ints."""
        self.L = _sigs()
        self.device = torch.device(device)
        self.verts = torch.as_tensor(verts, dtype=torch.float32).reshape(-1, 3).contiguous().to(self.device)
        self.faces = torch.as_tensor(np.asarray(faces).astype(np.int32)).reshape(-1, 3).contiguous().to(self.device)
        self.n, self.ss, self.it = int(image_size), int(supersample), int(dilate_iterations)
        self.max_ray_num, self.mask_weight = int(max_ray_num), float(mask_weight)
        self.focal = 0.5 * self.n / math.tan(0.5 * camera_angle_x)            # dataset.py:235-236
        size = C.c_size_t()
        _lib.check(self.L.avc_raster_workspace_bytes(self.verts.shape[0], self.n, self.ss, C.byref(size)), "raster ws")
        self.ws = torch.empty(size.value, dtype=torch.uint8, device=self.device)
        self.stream = torch.cuda.Stream(device=self.device)
        self._counts = torch.zeros(2, dtype=torch.int32, device=self.device)

    # ------------------------------------------------------------------------------------------ stage 1
    def submit(self, draw: StepDraw) -> _Pending:
        L, n = self.L, self.n
        p = _Pending()
        p.draw = draw
        p.eye = np.ascontiguousarray(draw.eye, dtype=np.float32)
        p.at = np.ascontiguousarray(draw.at, dtype=np.float32)
        fp = C.POINTER(C.c_float)
        with torch.cuda.stream(self.stream):
            sp = C.c_void_p(self.stream.cuda_stream)
            p.rgb = torch.empty(n, n, 3, dtype=torch.float32, device=self.device)
            p.mask256 = torch.empty(n, n, dtype=torch.uint8, device=self.device)
            p.dil = torch.empty(n, n, dtype=torch.uint8, device=self.device)
            cnt = torch.zeros(1, dtype=torch.int32, device=self.device)
            _lib.check(L.avc_raster_template(_lib.ptr(self.verts), _lib.ptr(self.faces), self.verts.shape[0],
                                             self.faces.shape[0], p.eye.ctypes.data_as(fp), p.at.ctypes.data_as(fp), n,
                                             self.ss, _lib.ptr(p.rgb), _lib.ptr(p.mask256), _lib.ptr(self.ws),
                                             self.ws.numel(), sp), "avc_raster_template")
            _lib.check(L.avc_dilate_count(_lib.ptr(p.mask256), n, self.it, _lib.ptr(p.dil), _lib.ptr(cnt), sp),
                       "avc_dilate_count")
            p.count_host = torch.empty(1, dtype=torch.int32).pin_memory()
            p.count_host.copy_(cnt, non_blocking=True)
            p.event = torch.cuda.Event()
            p.event.record(self.stream)
        return p

    # ------------------------------------------------------------------------------------------ stage 2
    def finish(self, p: _Pending) -> PreparedView:
        L, n, dev = self.L, self.n, self.device
        d = p.draw
        p.event.synchronize()
        count = int(p.count_host[0])
        v = PreparedView()
        v.draw, v.template_rgb = d, p.rgb
        fp = C.POINTER(C.c_float)
        pose = np.ascontiguousarray(np.asarray(d.pose, dtype=np.float32).reshape(4, 4))
        with torch.cuda.stream(self.stream):
            sp = C.c_void_p(self.stream.cuda_stream)
            f32 = dict(dtype=torch.float32, device=dev)
            if count == 0:
                # dataset.py:253-254 returns gen_rays_pose(resolution_level=4) -- a 2-tuple the caller cannot unpack
                # (main.py:366 crashes).  Here: the full 64 x 64 grid, every pixel a ray.
                W = n // 4
                in_mask = torch.ones(W * W, dtype=torch.uint8, device=dev)
                pix = torch.arange(W * W, dtype=torch.int32, device=dev)
                R = W * W
            else:
                ratio = count / float(n * n)
                W = min(n, int(np.sqrt(self.max_ray_num / ratio)))                      # dataset.py:257-258
                in_mask = torch.empty(W * W, dtype=torch.uint8, device=dev)
                pix = torch.empty(W * W, dtype=torch.int32, device=dev)
                cnt = torch.zeros(1, dtype=torch.int32, device=dev)
                _lib.check(L.avc_mask_compact(_lib.ptr(p.dil), n, W, W * W, _lib.ptr(in_mask), _lib.ptr(pix),
                                              _lib.ptr(cnt), sp), "avc_mask_compact")
                rh = torch.empty(1, dtype=torch.int32).pin_memory()
                rh.copy_(cnt, non_blocking=True)
                self.stream.synchronize()
                R = int(rh[0])
                pix = pix[:R]
            v.H = v.W = W
            v.R = R
            v.in_mask, v.pix = in_mask, pix
            v.rays_o, v.rays_d = torch.empty(R, 3, **f32), torch.empty(R, 3, **f32)
            v.near, v.far = torch.empty(R, **f32), torch.empty(R, **f32)
            _lib.check(L.avc_gen_rays(pose.ctypes.data_as(fp), self.focal, self.focal, 0.5 * n, 0.5 * n, n, n, W, W,
                                      _lib.ptr(pix), R, _lib.ptr(v.rays_o), _lib.ptr(v.rays_d), _lib.ptr(v.near),
                                      _lib.ptr(v.far), sp), "avc_gen_rays")
            v.true_rgb, v.mask = torch.empty(W * W, 3, **f32), torch.empty(W * W, **f32)
            _lib.check(L.avc_view_targets(_lib.ptr(p.rgb), n, W, 1 if self.mask_weight > 0.0 else 0,
                                          _lib.ptr(v.true_rgb), _lib.ptr(v.mask), sp), "avc_view_targets")
            v.bg_choice = int(d.bg_choice)
            v.canvas_background = v.ray_background = None
            if v.bg_choice in (1, 2):
                v.canvas_background, v.ray_background = torch.empty(W * W, **f32), torch.empty(R, **f32)
                chess_len = max(1, W // max(1, d.chess_div))                              # main.py:399
                _lib.check(L.avc_background_field(v.bg_choice, W, W, d.field_seed & 0xFFFFFFFF, chess_len,
                                                  float(d.blur_sigma), _lib.ptr(v.canvas_background), _lib.ptr(pix), R,
                                                  _lib.ptr(v.ray_background), sp), "avc_background_field")
            v.jitter = torch.empty(R, **f32)
            _lib.check(L.avc_uniform_fill((d.field_seed * 2654435761 + 12345) & 0xFFFFFFFF, R, -0.5, 0.5,
                                          _lib.ptr(v.jitter), sp), "avc_uniform_fill")
            v.light_dir, v.ambience = np.asarray(d.light_dir, dtype=np.float32), float(d.ambience)
            sc = torch.tensor([float(v.light_dir[0]), float(v.light_dir[1]), float(v.light_dir[2]), v.ambience],
                              dtype=torch.float32).pin_memory()
            v.scalars = sc.to(dev, non_blocking=True)
            ready = torch.cuda.Event()
            ready.record(self.stream)
        cur = torch.cuda.current_stream(dev)
        cur.wait_event(ready)                                   # the step (on the caller's stream) starts after the view
        # The view's tensors were allocated on the side stream and are consumed on the caller's: tell the caching allocator,
        # or a block freed when the caller drops the view could be handed to the NEXT view's preparation while the step
        # that reads it is still running (the loop does not synchronise every step when no scalar writer is attached).
        for name in PreparedView.__slots__:
            t = getattr(v, name, None)
            if torch.is_tensor(t) and t.is_cuda:
                t.record_stream(cur)
        return v

    def build(self, draw: StepDraw) -> PreparedView:
        return self.finish(self.submit(draw))


def read_obj(path: str):
    """Vertices [V,3] float32 and triangle faces [F,3] int32 of a Wavefront OBJ (``readOBJ`` in the reference's
    models/utils.py is used the same way at main.py:316: positions and the vertex index of every face corner)."""
    vs, fs = [], []
    with open(path) as f:
        for line in f:
            if line.startswith("v "):
                vs.append([float(t) for t in line.split()[1:4]])
            elif line.startswith("f "):
                idx = [int(t.split("/")[0]) - 1 for t in line.split()[1:]]
                for k in range(1, len(idx) - 1):
                    fs.append([idx[0], idx[k], idx[k + 1]])
    return np.asarray(vs, dtype=np.float32), np.asarray(fs, dtype=np.int32)
