"""On-disk formats either side of the hot path.

* ``write_ply`` / ``read_ply``: the vertex-coloured triangle mesh ``Runner.validate_mesh`` exports (main.py:913-916,
  ``trimesh.exchange.export.export_mesh(..., file_type='ply')``: binary little-endian PLY, vertex x y z (float) + red
  green blue alpha (uchar), faces as uchar-counted int32 index lists) -- what Avatar2FBX consumes.
* ``render_coarse_shape``: the ShapeGen -> AppearanceGen hand-off (AvatarGen/ShapeGen/render.py:32-58,107-139): the posed
  coarse shape rendered from 18 azimuths x 6 elevations at distance 2.2 with the template rasteriser
  (``avc_raster_template``; the reference uses neural_renderer's ``look_at`` mode: eye on a sphere, looking at the
  origin) into ``img/0000.png ...`` + ``transforms_train.json`` (``camera_angle_x`` = 60 degrees, camera-to-world
  matrices from ``lookat``) -- the directory ``SMPL_Dataset`` / ``--mode train`` read.
"""
from __future__ import annotations

import json
import os
import struct

import numpy as np


def write_ply(path: str, vertices, triangles, vertex_colors=None):
    v = np.asarray(vertices, dtype=np.float32).reshape(-1, 3)
    f = np.asarray(triangles, dtype=np.int32).reshape(-1, 3)
    header = ["ply", "format binary_little_endian 1.0", f"element vertex {v.shape[0]}",
              "property float x", "property float y", "property float z"]
    if vertex_colors is not None:
        c = np.asarray(vertex_colors, dtype=np.uint8).reshape(-1, 3)
        c = np.concatenate([c, np.full((c.shape[0], 1), 255, dtype=np.uint8)], 1)
        header += ["property uchar red", "property uchar green", "property uchar blue", "property uchar alpha"]
    header += [f"element face {f.shape[0]}", "property list uchar int vertex_indices", "end_header"]
    with open(path, "wb") as fp:
        fp.write(("\n".join(header) + "\n").encode("ascii"))
        if vertex_colors is not None:
            rec = np.zeros(v.shape[0], dtype=[("p", "<f4", 3), ("c", "u1", 4)])
            rec["p"], rec["c"] = v, c
            fp.write(rec.tobytes())
        else:
            fp.write(v.astype("<f4").tobytes())
        frec = np.zeros(f.shape[0], dtype=[("n", "u1"), ("i", "<i4", 3)])
        frec["n"], frec["i"] = 3, f
        fp.write(frec.tobytes())
    return path


def read_ply(path: str):
    """Reads what ``write_ply`` (or trimesh's binary PLY export of a vertex-coloured triangle mesh) wrote."""
    with open(path, "rb") as fp:
        nv = nf = 0
        has_color = False
        while True:
            line = fp.readline().decode("ascii").strip()
            if line.startswith("element vertex"):
                nv = int(line.split()[-1])
            elif line.startswith("element face"):
                nf = int(line.split()[-1])
            elif line.startswith("property uchar red"):
                has_color = True
            elif line == "end_header":
                break
        if has_color:
            rec = np.frombuffer(fp.read(nv * 16), dtype=[("p", "<f4", 3), ("c", "u1", 4)])
            v, c = rec["p"].copy(), rec["c"][:, :3].copy()
        else:
            v, c = np.frombuffer(fp.read(nv * 12), dtype="<f4").reshape(-1, 3).copy(), None
        frec = np.frombuffer(fp.read(nf * 13), dtype=[("n", "u1"), ("i", "<i4", 3)])
        return v, frec["i"].copy(), c


def get_points_from_angles(distance, elevation, azimuth):
    """neural_renderer.get_points_from_angles (degrees): eye on a sphere around the origin."""
    e, a = np.radians(elevation), np.radians(azimuth)
    return np.array([distance * np.cos(e) * np.sin(a), distance * np.sin(e), -distance * np.cos(e) * np.cos(a)])


def lookat_inverse_view(eye, target, up):
    """ShapeGen/render.py:19-30: inverse of the world-to-camera view matrix (= camera-to-world)."""
    z = (eye - target) / np.linalg.norm(eye - target)
    x = np.cross(up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    view = np.array([[x[0], x[1], x[2], -np.dot(x, eye)], [y[0], y[1], y[2], -np.dot(y, eye)],
                     [z[0], z[1], z[2], -np.dot(z, eye)], [0, 0, 0, 1]])
    return np.linalg.inv(view)


def render_coarse_shape(vertices, faces, output_dir: str, device="cuda", image_size: int = 256,
                        camera_distance: float = 2.2):
    """ShapeGen/render.py:32-58,118-139 -> ``output_dir/img/NNNN.png`` + ``transforms_train.json`` (108 views).
    ``vertices``: the posed coarse shape [V,3] (``my_lbs`` output, the frame of ``Runner.v``)."""
    import cv2 as cv
    import torch
    from .sampling import StepDraw
    from .views import ViewBuilder
    vb = ViewBuilder(vertices, faces, 10 ** 9, 0.0, device, image_size=image_size)
    os.makedirs(os.path.join(output_dir, "img"), exist_ok=True)
    frames, i = [], 0
    for angle in range(0, 360, 20):
        for elevation in range(-60, 60, 20):
            eye = get_points_from_angles(camera_distance, elevation, angle)
            d = StepDraw(i, False, eye.astype(np.float32), np.zeros(3, dtype=np.float32), 0.0, 0.0, 1, np.eye(4), 3, 10, 1.0,
                         np.zeros(3, dtype=np.float32), 0.0, 0)
            p = vb.submit(d)
            torch.cuda.synchronize()
            img = p.rgb.flip(1).cpu().numpy()                     # the hand-off images are NOT flipped (render.py:52-57)
            cv.imwrite(os.path.join(output_dir, "img", "{}.png".format(str(i).zfill(4))),
                       cv.cvtColor((img * 255).astype(np.uint8), cv.COLOR_RGB2BGR))
            t = lookat_inverse_view(eye, np.array([0, 0, 0]), np.array([0, 1, 0]))
            frames.append({"file_path": "img/{}".format(str(i).zfill(4)), "transform_matrix": t.tolist()})
            i += 1
    with open(os.path.join(output_dir, "transforms_train.json"), "w") as f:
        json.dump({"camera_angle_x": 60 / 180 * np.pi, "frames": frames}, f)
    return output_dir
