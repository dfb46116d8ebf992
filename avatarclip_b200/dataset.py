"""Host-side mirror of the ray generators of ``SMPL_Dataset`` (AvatarGen/AppearanceGen/models/dataset.py:252-293,
331-342) computed by ``avc_gen_rays``.  The file loaders of the reference class (108 PNG renders + transforms_train.json)
are not part of the hot path; only the camera model is kept: H = W = 256, focal = .5 W / tan(.5 camera_angle_x)."""
from __future__ import annotations

import ctypes as C
import math

import numpy as np
import torch

from . import _lib


def silhouette_canvas(mask: torch.Tensor, max_ray_num: int, full: int = 256):
    """Host part of ``gen_rays_silhouettes`` (dataset.py:253-258, 269-271): dilate the template silhouette 10 times with
    the full 3 x 3 structure (``ndimage.generate_binary_structure(2, 2)``), pick the square canvas whose masked pixel
    count is about ``max_ray_num`` (never above ``full``), and resize the dilated mask to it (nearest, as
    ``F.interpolate``'s default).  ``mask``: [full, full], non-zero = inside.  Returns (canvas, bool [canvas, canvas]).
    Ten 3 x 3 max-pools are the ten binary dilations (outside the image counts as background in both)."""
    m = (torch.as_tensor(mask) != 0).float().reshape(1, 1, full, full)
    for _ in range(10):
        m = torch.nn.functional.max_pool2d(m, kernel_size=3, stride=1, padding=1)
    current_ratio = float(m.sum()) / float(full * full)
    canvas = min(full, int(math.sqrt(max_ray_num / current_ratio)))
    resized = torch.nn.functional.interpolate(m, size=(canvas, canvas)).reshape(canvas, canvas)
    return canvas, resized > 0


class RayGenerator:
    def __init__(self, H: int = 256, W: int = 256, camera_angle_x: float = math.pi / 3, device="cuda"):
        self.H, self.W = int(H), int(W)
        self.focal = 0.5 * self.W / math.tan(0.5 * camera_angle_x)          # dataset.py:235-236
        self.device = torch.device(device)
        L = _lib.lib()
        f32, i32, vp = C.c_float, C.c_int32, C.c_void_p
        L.avc_gen_rays.argtypes = [C.POINTER(C.c_float), f32, f32, f32, f32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp]
        L.avc_gen_rays.restype = C.c_int

    def _call(self, pose, Wc, Hc, pix):
        L = _lib.lib()
        pose = np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(4, 4))
        R = int(pix.numel()) if pix is not None else Wc * Hc
        f = dict(dtype=torch.float32, device=self.device)
        ro, rd = torch.empty(R, 3, **f), torch.empty(R, 3, **f)
        near, far = torch.empty(R, 1, **f), torch.empty(R, 1, **f)
        _lib.check(L.avc_gen_rays(pose.ctypes.data_as(C.POINTER(C.c_float)), self.focal, self.focal, 0.5 * self.W,
                                  0.5 * self.H, self.W, self.H, Wc, Hc, _lib.ptr(pix), R, _lib.ptr(ro), _lib.ptr(rd),
                                  _lib.ptr(near), _lib.ptr(far), _lib.stream_ptr()), "avc_gen_rays")
        return ro, rd, near, far

    def gen_rays_pose(self, pose, resolution_level=1):
        """dataset.py:277-293 -> rays_o, rays_d [H', W', 3] (+ near, far [H'*W', 1] as extra returns)."""
        Wc, Hc = int(self.W // resolution_level), int(self.H // resolution_level)
        ro, rd, near, far = self._call(pose, Wc, Hc, None)
        return ro.reshape(Hc, Wc, 3), rd.reshape(Hc, Wc, 3), near, far

    def gen_rays_pixels(self, pose, canvas: int, dilated_mask: torch.Tensor):
        """The ray part of gen_rays_silhouettes (dataset.py:259-275) for a canvas x canvas grid and a boolean
        [canvas, canvas] mask already resized to it: rays of the True pixels in row-major order."""
        pix = torch.nonzero(dilated_mask.reshape(-1).to(self.device), as_tuple=False).reshape(-1).to(torch.int32)
        ro, rd, near, far = self._call(pose, canvas, canvas, pix)
        return ro, rd, near, far, pix

    def gen_rays_silhouettes(self, pose, max_ray_num: int, mask: torch.Tensor):
        """dataset.py:252-275: rays of the dilated-silhouette pixels on a canvas sized for ~max_ray_num rays.
        Returns (rays_o [R,3], rays_d [R,3], W, dilated_mask bool [W,W]) like the reference, plus near / far / pix."""
        if float((torch.as_tensor(mask) != 0).sum()) == 0:            # :253-254
            ro, rd, near, far = self.gen_rays_pose(pose, resolution_level=4)
            return ro, rd, ro.shape[1], None, near, far, None
        canvas, dm = silhouette_canvas(mask, max_ray_num, self.H)
        ro, rd, near, far, pix = self.gen_rays_pixels(pose, canvas, dm)
        return ro, rd, canvas, dm, near, far, pix
