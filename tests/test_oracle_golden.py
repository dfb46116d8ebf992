"""oracle.neus (the restated reference) against the golden vectors that
oracle/pin_against_reference.py produced by running the UNMODIFIED reference."""
import os

import pytest
import torch

from oracle import neus
from conftest import GOLDEN


def _load(name):
    return torch.load(os.path.join(GOLDEN, f"neus_{name}.pt"), map_location="cpu", weights_only=False)


@pytest.mark.parametrize("name", ["cfg1", "small"])
def test_restatement_matches_reference_outputs_and_grads(name):
    b = _load(name)
    sconf = neus.SDFConf(**{k: (tuple(v) if k == "skip_in" else v) for k, v in b["sdf_kw"].items()})
    cconf = neus.ColorConf(**b["col_kw"])
    rconf = neus.RenderConf(**b["ren_kw"])
    sp = {k: v.clone().requires_grad_(True) for k, v in b["sdf_state"].items()}
    cp = {k: v.clone().requires_grad_(True) for k, v in b["col_state"].items()}
    var = b["variance"].clone().requires_grad_(True)
    out = neus.render(sp, cp, var, sconf, cconf, rconf, b["rays_o"], b["rays_d"], b["near"], b["far"],
                      jitter=b["jitter"], background_rgb=b["background"], cos_anneal_ratio=b["cos_anneal"],
                      z_vals=b["z_vals"])
    for k, ref in b["out"].items():
        err = (out[k].detach() - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
        assert err < 1e-4, (k, err)
    loss = sum((out[k] * w).sum() for k, w in b["loss_weights"].items())
    names = [("sdf." + k, v) for k, v in sp.items()] + [("col." + k, v) for k, v in cp.items()] + [("var.variance", var)]
    grads = torch.autograd.grad(loss, [v for _, v in names], allow_unused=True)
    for (k, v), g in zip(names, grads):
        ref = b["grads"][k]
        g = torch.zeros_like(v) if g is None else g
        err = (g - ref).abs().max().item() / (ref.abs().max().item() + 1e-12)
        assert err < 2e-3, (k, err)


@pytest.mark.parametrize("name", ["cfg1", "small"])
def test_sample_placement_matches_reference(name):
    """Placement is discontinuous (bin search, radius<1 masks): require the bulk of the rays
    to agree to 1e-4 and report the rest ('parity definition')."""
    b = _load(name)
    sconf = neus.SDFConf(**{k: (tuple(v) if k == "skip_in" else v) for k, v in b["sdf_kw"].items()})
    rconf = neus.RenderConf(**b["ren_kw"])
    with torch.no_grad():
        z = neus.hierarchical_z(lambda x: neus.sdf_value(b["sdf_state"], sconf, x), rconf, b["rays_o"], b["rays_d"],
                                b["near"], b["far"], b["jitter"])
    dz = (z - b["z_vals"]).abs().max(dim=1)[0]
    assert (dz < 1e-4).float().mean().item() >= 0.95
    assert torch.all(z[:, 1:] >= z[:, :-1])


def test_lbs_restatement_matches_reference_golden():
    """oracle.lbs.my_lbs vs the vector recorded from the reference's own my_lbs (oracle/pin_lbs.py)."""
    from oracle import lbs
    g = torch.load(os.path.join(GOLDEN, "lbs_small.pt"), map_location="cpu", weights_only=False)
    v, j = lbs.my_lbs(**g["inputs"])
    assert (v - g["verts"]).abs().max().item() < 1e-6
    assert (j - g["joints"]).abs().max().item() < 1e-6
