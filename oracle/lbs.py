"""CPU oracle (test infrastructure, never on the product path): SMPL linear-blend skinning as AvatarCLIP uses it.

``my_lbs`` / ``batch_rodrigues`` follow AvatarGen/AppearanceGen/models/utils.py:72-106,176-224; the smplx helpers
they call (``vertices2joints``, ``batch_rigid_transform``, ``transform_mat``) are restated from their verbatim
in-tree copies at AvatarGen/AppearanceGen/drive.py:51-160 (``smplx`` itself is an un-vendored, unpinned dependency,
requirements.txt:13, and is not installed).

PARITY UNPINNED beyond that: the reference has no test or golden vector for LBS and SMPL_NEUTRAL.pkl is licence
gated / absent, so tests use synthetic SMPL-shaped tensors (``synthetic_smpl``) with the real template vertex count.
"""
from __future__ import annotations

import numpy as np
import torch
import torch.nn.functional as F

SMPL_PARENTS = [-1, 0, 0, 0, 1, 2, 3, 4, 5, 6, 7, 8, 9, 9, 9, 12, 13, 14, 16, 17, 18, 19, 20, 21]   # SMPL kinematic tree


def batch_rodrigues(rot_vecs: torch.Tensor, epsilon: float = 1e-8) -> torch.Tensor:
    """models/utils.py:72-106."""
    n = rot_vecs.shape[0]
    angle = torch.norm(rot_vecs + epsilon, dim=1, keepdim=True, p=2)
    rot_dir = rot_vecs / angle
    cos = torch.cos(angle)[:, None]
    sin = torch.sin(angle)[:, None]
    rx, ry, rz = torch.split(rot_dir, 1, dim=1)
    zeros = torch.zeros((n, 1), dtype=rot_vecs.dtype)
    K = torch.cat([zeros, -rz, ry, rz, zeros, -rx, -ry, rx, zeros], dim=1).view(n, 3, 3)
    ident = torch.eye(3, dtype=rot_vecs.dtype)[None]
    return ident + sin * K + (1 - cos) * torch.bmm(K, K)


def batch_rigid_transform(rot_mats, joints, parents):
    """drive.py:96-145 (copy of smplx.lbs.batch_rigid_transform)."""
    joints = joints.unsqueeze(-1)
    rel = joints.clone()
    rel[:, 1:] -= joints[:, parents[1:]]
    tm = torch.cat([F.pad(rot_mats.reshape(-1, 3, 3), [0, 0, 0, 1]),
                    F.pad(rel.reshape(-1, 3, 1), [0, 0, 0, 1], value=1)], dim=2).reshape(-1, joints.shape[1], 4, 4)
    chain = [tm[:, 0]]
    for i in range(1, parents.shape[0]):
        chain.append(torch.matmul(chain[parents[i]], tm[:, i]))
    T = torch.stack(chain, dim=1)
    posed = T[:, :, :3, 3]
    jh = F.pad(joints, [0, 0, 0, 1])
    A = T - F.pad(torch.matmul(T, jh), [3, 0, 0, 0, 0, 0, 0, 0])
    return posed, A


def my_lbs(v_shaped, pose, posedirs, J_regressor, parents, lbs_weights, pose2rot: bool = True):
    """models/utils.py:176-224 (v_template / shapedirs are unused there: the shape blend is applied upstream)."""
    B = pose.shape[0]
    dt = pose.dtype
    J = torch.einsum("bik,ji->bjk", v_shaped, J_regressor)                    # vertices2joints (drive.py:51-70)
    ident = torch.eye(3, dtype=dt)
    if pose2rot:
        rot = batch_rodrigues(pose.view(-1, 3)).view(B, -1, 3, 3)
        feat = (rot[:, 1:] - ident).view(B, -1)
    else:
        rot = pose.view(B, -1, 3, 3)
        feat = (rot[:, 1:] - ident).view(B, -1)
    v_posed = torch.matmul(feat, posedirs).view(B, -1, 3) + v_shaped
    Jt, A = batch_rigid_transform(rot, J, parents)
    nj = J_regressor.shape[0]
    T = torch.matmul(lbs_weights[None].expand(B, -1, -1), A.view(B, nj, 16)).view(B, -1, 4, 4)
    vh = torch.cat([v_posed, torch.ones(B, v_posed.shape[1], 1, dtype=dt)], dim=2)
    verts = torch.matmul(T, vh.unsqueeze(-1))[:, :, :3, 0]
    return verts, Jt


def synthetic_smpl(V: int = 6890, seed: int = 0, dtype=torch.float32):
    """SMPL-shaped random tensors: body-like vertices, sparse skin weights (<= 4 joints per vertex, rows sum to 1),
    J_regressor rows summing to 1, small pose blend shapes, the standard 24-joint parent table, a stand-ish pose."""
    g = torch.Generator().manual_seed(seed)
    nj = 24
    v = torch.randn(V, 3, generator=g, dtype=dtype) * torch.tensor([0.25, 0.55, 0.12], dtype=dtype)
    w = torch.zeros(V, nj, dtype=dtype)
    idx = torch.randint(0, nj, (V, 4), generator=g)
    val = torch.rand(V, 4, generator=g, dtype=dtype) + 0.05
    w.scatter_add_(1, idx, val)
    w = w / w.sum(1, keepdim=True)
    jr = torch.rand(nj, V, generator=g, dtype=dtype) ** 8
    jr = jr / jr.sum(1, keepdim=True)
    posedirs = torch.randn(207, V * 3, generator=g, dtype=dtype) * 1e-3
    pose = torch.randn(1, 72, generator=g, dtype=dtype) * 0.25
    return dict(v_shaped=v[None], pose=pose, posedirs=posedirs, J_regressor=jr,
                parents=torch.tensor(SMPL_PARENTS, dtype=torch.long), lbs_weights=w)
