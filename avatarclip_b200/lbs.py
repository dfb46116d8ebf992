"""Host-side mirror of ``my_lbs`` (AvatarGen/AppearanceGen/models/utils.py:176-224): same name, argument order and
return value, computed by libavc_b200.so (``avc_lbs_fwd``).  Batch size 1, as in the reference's only call site
(main.py:322-328)."""
from __future__ import annotations

import ctypes as C

import torch

from . import _lib


def my_lbs(v_shaped, pose, v_template, shapedirs, posedirs, J_regressor, parents, lbs_weights, pose2rot: bool = True):
    if not v_shaped.is_cuda:
        raise _lib.AvcError("avatarclip_b200 has no CPU path")
    if v_shaped.shape[0] != 1:
        raise NotImplementedError("batch size 1 (the reference's only use, main.py:322-328)")
    L = _lib.lib()
    vp, i32, sz = C.c_void_p, C.c_int32, C.c_size_t
    L.avc_lbs_workspace_bytes.argtypes = [i32, C.POINTER(sz)]
    L.avc_lbs_fwd.argtypes = [vp, vp, i32, vp, vp, vp, vp, i32, i32, vp, vp, vp, sz, vp]
    L.avc_lbs_fwd.restype = C.c_int
    dev = v_shaped.device
    f = lambda t: t.detach().to(dev, torch.float32).contiguous()
    V, nj = v_shaped.shape[1], J_regressor.shape[0]
    vs, ps, jr, pd, lw = f(v_shaped), f(pose), f(J_regressor), f(posedirs), f(lbs_weights)
    par = parents.detach().to(dev, torch.int32).contiguous()
    verts = torch.empty(1, V, 3, dtype=torch.float32, device=dev)
    joints = torch.empty(1, nj, 3, dtype=torch.float32, device=dev)
    need = sz()
    _lib.check(L.avc_lbs_workspace_bytes(nj, C.byref(need)), "avc_lbs_workspace_bytes")
    ws = torch.empty(need.value, dtype=torch.uint8, device=dev)
    _lib.check(L.avc_lbs_fwd(_lib.ptr(vs), _lib.ptr(ps), 1 if pose2rot else 0, _lib.ptr(jr), _lib.ptr(par), _lib.ptr(pd),
                             _lib.ptr(lw), V, nj, _lib.ptr(verts), _lib.ptr(joints), _lib.ptr(ws), ws.numel(),
                             _lib.stream_ptr()), "avc_lbs_fwd")
    return verts, joints
