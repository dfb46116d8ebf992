"""The hand-derived backward (oracle.neus_manual) against autograd of the restated
reference (oracle.neus), in fp64 so that only derivation errors can show."""
import math

import pytest
import torch

from oracle import neus, neus_manual


def _case(skip_last: bool, seed: int, bg_kind: str):
    g = torch.Generator().manual_seed(seed)
    dt = torch.float64
    if skip_last:
        sconf = neus.SDFConf(d_out=33, d_hidden=32 + 16, n_layers=3, skip_in=(3,), multires=6)
    else:
        sconf = neus.SDFConf(d_out=49, d_hidden=48, n_layers=5, skip_in=(2,), multires=6)
    F = sconf.d_out - 1
    cconf = neus.ColorConf(d_feature=F, d_hidden=40, n_layers=2 if skip_last else 3)
    rconf = neus.RenderConf(n_samples=8, n_importance=8, up_sample_steps=4)
    sp = {k: v.requires_grad_(True) for k, v in neus.init_sdf_params(sconf, g, dt).items()}
    # perturb the geometric init so no weight block is exactly zero
    with torch.no_grad():
        for k, v in sp.items():
            if k.endswith("weight_v"):
                v.add_(0.05 * torch.randn(v.shape, generator=g, dtype=dt))
    cp = {k: v.requires_grad_(True) for k, v in neus.init_color_params(cconf, g, dt).items()}
    var = torch.tensor(0.25, dtype=dt, requires_grad=True)
    R = 12
    d = torch.randn(R, 3, generator=g, dtype=dt) * 0.2 + torch.tensor([0.0, 0.0, -1.0], dtype=dt)
    d = d / d.norm(dim=-1, keepdim=True)
    o = torch.tensor([0.0, 0.0, 1.6], dtype=dt) + 0.1 * torch.randn(R, 3, generator=g, dtype=dt)
    near, far = neus.near_far_from_sphere(o, d)
    jit = torch.rand(R, 1, generator=g, dtype=dt) - 0.5
    bg = {"none": None, "white": torch.ones(1, 3, dtype=dt), "ray": torch.rand(R, 1, generator=g, dtype=dt)}[bg_kind]
    return sconf, cconf, rconf, sp, cp, var, o, d, near, far, jit, bg, g


@pytest.mark.parametrize("skip_last,bg_kind,anneal", [(False, "none", 0.3), (True, "ray", 1.0), (False, "white", 0.0)])
def test_manual_backward_matches_autograd(skip_last, bg_kind, anneal):
    sconf, cconf, rconf, sp, cp, var, o, d, near, far, jit, bg, g = _case(skip_last, 5, bg_kind)
    out = neus.render(sp, cp, var, sconf, cconf, rconf, o, d, near, far, jitter=jit, background_rgb=bg,
                      cos_anneal_ratio=anneal)
    z = out["z_vals"].detach()
    keys = ["color_fine", "extra_color_fine", "s_val", "cdf_fine", "weight_sum", "weight_max", "gradients",
            "weights", "gradient_error"]
    cot = {k: torch.randn(out[k].shape, generator=g, dtype=torch.float64) for k in keys}
    loss = sum((out[k] * cot[k]).sum() for k in keys)
    names = [("sdf." + k, v) for k, v in sp.items()] + [("col." + k, v) for k, v in cp.items()] + [("var.variance", var)]
    auto = torch.autograd.grad(loss, [v for _, v in names])

    with torch.no_grad():
        mout, stash = neus_manual.fine_forward(sp, cp, var, sconf, cconf, rconf, o, d, z, background_rgb=bg,
                                               cos_anneal_ratio=anneal)
        for k in keys + ["mid_z_vals", "inside_sphere"]:
            assert torch.allclose(mout[k], out[k].detach(), rtol=1e-9, atol=1e-11), k
        grads = neus_manual.fine_backward(sp, cp, var, sconf, cconf, rconf, stash, cot)
    for (k, _), ga in zip(names, auto):
        gm = grads[k]
        scale = ga.abs().max().item() + 1e-30
        err = (gm.reshape(ga.shape) - ga).abs().max().item() / scale
        assert err < 1e-8, (k, err)
