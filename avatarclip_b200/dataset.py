"""Host-side mirror of ``SMPL_Dataset`` (AvatarGen/AppearanceGen/models/dataset.py:199-347): same constructor (a conf
subtree with ``data_dir``), attributes (``H, W, focal, K, poses, images, masks, n_images, object_bbox_min/max``) and method
names / return arities, with the rays computed by libavc_b200.so (``avc_gen_rays``, csrc/avc_rays.cu).

``data_dir`` holds ``transforms_train.json`` + ``img/*.png`` (the 108-view template render written by
ShapeGen/render.py:118-139, or by ``avatarclip_b200.handoff.render_coarse_shape``).  ``train_clip`` only needs the
intrinsics; when the directory is absent the dataset is built "headless" (256 x 256, 60 degree field of view -- the values
every shipped transforms_train.json holds) and says so.
"""
from __future__ import annotations

import ctypes as C
import json
import math
import os

import numpy as np
import torch

from . import _lib


def silhouette_canvas(mask: torch.Tensor, max_ray_num: int, full: int = 256):
    """Host restatement of the mask handling in ``gen_rays_silhouettes`` (dataset.py:253-258, 269-271): dilate the template
    silhouette 10 times with the full 3 x 3 structure (``ndimage.generate_binary_structure(2, 2)``), pick the square
    canvas whose masked pixel count is about ``max_ray_num`` (never above ``full``), and resize the dilated mask to it
    (nearest, as ``F.interpolate``'s default).  ``mask``: [full, full], non-zero = inside.  Returns (canvas, bool
    [canvas, canvas]).  Ten 3 x 3 max-pools are the ten binary dilations (outside the image counts as background in
    both).  The training loop does the same on the device (``avatarclip_b200.views``); this is the API-level path."""
    m = (torch.as_tensor(mask) != 0).float().reshape(1, 1, full, full)
    for _ in range(10):
        m = torch.nn.functional.max_pool2d(m, kernel_size=3, stride=1, padding=1)
    current_ratio = float(m.sum()) / float(full * full)
    canvas = min(full, int(math.sqrt(max_ray_num / current_ratio)))
    resized = torch.nn.functional.interpolate(m, size=(canvas, canvas)).reshape(canvas, canvas)
    return canvas, resized > 0


class RayGenerator:
    """Pinhole rays of a camera-to-world pose on a W x H grid spanning the full-resolution image (dataset.py:259-268,
    277-293) + ``near_far_from_sphere`` (:331-342), one kernel."""

    def __init__(self, H: int = 256, W: int = 256, camera_angle_x: float = math.pi / 3, device="cuda"):
        self.H, self.W = int(H), int(W)
        self.focal = 0.5 * self.W / math.tan(0.5 * camera_angle_x)          # dataset.py:235-236
        self.device = torch.device(device)
        L = _lib.lib()
        f32, i32, vp = C.c_float, C.c_int32, C.c_void_p
        L.avc_gen_rays.argtypes = [C.POINTER(C.c_float), f32, f32, f32, f32, i32, i32, i32, i32, vp, i32, vp, vp, vp, vp, vp]
        L.avc_gen_rays.restype = C.c_int

    def rays(self, pose, Wc, Hc, pix=None):
        """-> rays_o [R,3], rays_d [R,3], near [R,1], far [R,1]; R = len(pix) or Wc * Hc (row-major)."""
        L = _lib.lib()
        if torch.is_tensor(pose):
            pose = pose.detach().cpu().numpy()
        pose = np.ascontiguousarray(np.asarray(pose, dtype=np.float32).reshape(4, 4))
        R = int(pix.numel()) if pix is not None else Wc * Hc
        f = dict(dtype=torch.float32, device=self.device)
        ro, rd = torch.empty(R, 3, **f), torch.empty(R, 3, **f)
        near, far = torch.empty(R, 1, **f), torch.empty(R, 1, **f)
        _lib.check(L.avc_gen_rays(pose.ctypes.data_as(C.POINTER(C.c_float)), self.focal, self.focal, 0.5 * self.W,
                                  0.5 * self.H, self.W, self.H, Wc, Hc, _lib.ptr(pix), R, _lib.ptr(ro), _lib.ptr(rd),
                                  _lib.ptr(near), _lib.ptr(far), _lib.stream_ptr()), "avc_gen_rays")
        return ro, rd, near, far

    # kept for callers of round 1 (extended returns)
    def gen_rays_pose_ex(self, pose, resolution_level=1):
        Wc, Hc = int(self.W // resolution_level), int(self.H // resolution_level)
        ro, rd, near, far = self.rays(pose, Wc, Hc)
        return ro.reshape(Hc, Wc, 3), rd.reshape(Hc, Wc, 3), near, far


class SMPL_Dataset(RayGenerator):
    def __init__(self, conf, device="cuda"):
        self.conf = conf
        self.data_dir = conf.get_string("data_dir") if hasattr(conf, "get_string") else conf["data_dir"]
        meta_path = os.path.join(self.data_dir, "transforms_train.json")
        self.images_lis, images, poses = [], [], []
        camera_angle_x = math.pi / 3
        self.headless = not os.path.exists(meta_path)
        if not self.headless:
            import cv2 as cv
            with open(meta_path) as fp:
                meta = json.load(fp)
            camera_angle_x = float(meta["camera_angle_x"])
            for frame in meta["frames"]:
                fname = os.path.join(self.data_dir, frame["file_path"] + ".png")
                img = cv.imread(fname, cv.IMREAD_UNCHANGED)
                if img is None:
                    raise FileNotFoundError(fname)
                if img.ndim == 2:
                    img = np.repeat(img[:, :, None], 3, axis=2)
                img = img[:, :, [2, 1, 0] + ([3] if img.shape[2] == 4 else [])]      # BGR(A) -> RGB(A), as imageio reads
                images.append(img)
                self.images_lis.append(fname)
                poses.append(np.array(frame["transform_matrix"]))
        self.n_images = len(images)
        if self.n_images:
            im = (np.array(images) / 255.0).astype(np.float32)
            im = im[:, :, ::-1]                                                      # dataset.py:222 (flips the W axis)
            self.images = torch.from_numpy(im.copy()).cpu()
            self.masks = torch.zeros_like(self.images)
            self.masks[self.images != 0] = 1.0
            H, W = self.images[0].shape[:2]
            self.poses = torch.from_numpy(np.array(poses).astype(np.float32)).to(device)
        else:
            print(f"[avatarclip_b200] dataset: {meta_path} not found -- headless intrinsics (256 x 256, 60 deg); "
                  "validate_image / --mode train need the rendered views")
            self.images = torch.zeros(0, 256, 256, 3)
            self.masks = torch.zeros(0, 256, 256, 3)
            self.poses = torch.zeros(0, 4, 4, device=device)
            H = W = 256
        super().__init__(H, W, camera_angle_x, device)
        self.image_pixels = self.H * self.W
        self.object_bbox_min = np.array([-1.01, -1.01, -1.01])
        self.object_bbox_max = np.array([1.01, 1.01, 1.01])
        self.K = torch.from_numpy(np.array([[self.focal, 0, 0.5 * self.W], [0, self.focal, 0.5 * self.H], [0, 0, 1]])).cpu()

    # ------------------------------------------------------------------ reference API (same return arities)
    def gen_rays_pose(self, pose, resolution_level=1):
        """dataset.py:277-293 -> rays_o, rays_d [H', W', 3]."""
        Wc, Hc = int(self.W // resolution_level), int(self.H // resolution_level)
        ro, rd, _, _ = self.rays(pose, Wc, Hc)
        return ro.reshape(Hc, Wc, 3), rd.reshape(Hc, Wc, 3)

    def gen_rays_at(self, img_idx, resolution_level=1):
        """dataset.py:295-311."""
        return self.gen_rays_pose(self.poses[int(img_idx)], resolution_level)

    def gen_rays_silhouettes(self, pose, max_ray_num, mask):
        """dataset.py:252-275 -> (rays_o [R,3], rays_d [R,3], W, dilated_mask bool [W,W])."""
        if float((torch.as_tensor(mask) != 0).sum()) == 0:            # :253-254 (the reference returns a 2-tuple here)
            return self.gen_rays_pose(pose, resolution_level=4)
        canvas, dm = silhouette_canvas(mask, max_ray_num, self.H)
        pix = torch.nonzero(dm.reshape(-1).to(self.device), as_tuple=False).reshape(-1).to(torch.int32)
        ro, rd, _, _ = self.rays(pose, canvas, canvas, pix)
        return ro, rd, canvas, dm.to(self.device)

    def gen_random_rays_at(self, img_idx, batch_size):
        """dataset.py:313-329 -> [batch, 10] = rays_o, rays_d, colour, mask on the device."""
        img_idx = int(img_idx)
        px = torch.randint(low=0, high=self.W, size=[batch_size])
        py = torch.randint(low=0, high=self.H, size=[batch_size])
        color = self.images[img_idx][(py, px)][:, :3]
        mask = self.masks[img_idx][(py, px)][:, :1]
        pix = (py * self.W + px).to(torch.int32).to(self.device)
        ro, rd, _, _ = self.rays(self.poses[img_idx], self.W, self.H, pix)
        return torch.cat([ro, rd, color.to(self.device), mask.to(self.device)], dim=-1)

    def near_far_from_sphere(self, rays_o, rays_d, is_sphere=False):
        """dataset.py:331-342 (``is_sphere`` selects a commented-out branch in the reference: both values do the same)."""
        a = torch.sum(rays_d ** 2, dim=-1, keepdim=True)
        b = 2.0 * torch.sum(rays_o * rays_d, dim=-1, keepdim=True)
        mid = 0.5 * (-b) / a
        near = (mid - 1).clamp_min(0)
        far = mid + 1
        return near, far

    def image_at(self, idx, resolution_level):
        """dataset.py:344-347."""
        import cv2 as cv
        img = cv.imread(self.images_lis[idx])
        img = img[:, ::-1, :]
        return (cv.resize(img, (self.W // resolution_level, self.H // resolution_level))).clip(0, 255)
