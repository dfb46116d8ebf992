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


# ------------------------------------------------------------------------------------------------------------------
# Loss stage / ray generators: golden vectors written by oracle/pin_loss_stage.py, which EXECUTES the reference's own
# source lines (main.py:417-534, models/dataset.py:252-293,331-342) in the build container.
def _loss_golden():
    return torch.load(os.path.join(GOLDEN, "loss_stage.pt"), map_location="cpu", weights_only=False)


def _rel(a, b):
    a, b = a.double(), b.double()
    return float((a - b).abs().max() / max(float(b.abs().max()), 1e-30))


@pytest.mark.parametrize("case", [0, 1, 2, 3, 4, 5, 6])
def test_loss_stage_restatement_matches_reference_lines(case):
    """Cases 0-2: add_no_texture = texture_cast_light = True (confs/examples*); 3-6: the switch combinations of
    confs/ablation/*_0..2.conf (main.py:425,466,509-534)."""
    from oracle import clip_vit as cv
    from oracle import loss as ol
    c = _loss_golden()["cases"][case]
    H, choice = c["H"], c["choice_i"]
    igr_w, mask_w, clip_w = c["weights"]
    no_tex, cast = c["flags"]
    leaves = {k: v.clone().requires_grad_(v.is_floating_point()) for k, v in c["render_out"].items()}
    mask = (c["mask"] > 0.5).float() if mask_w > 0.0 else torch.ones_like(c["mask"])          # main.py:407-410
    st = ol.shading_and_losses(leaves, c["dilated_mask"], H, H, c["true_rgb"], mask, c["light_dir"].float(), c["ambience"],
                               choice, c["background_rgb"] if choice in (1, 2) else None, igr_w, mask_w,
                               add_no_texture=no_tex, texture_cast_light=cast)
    clip_state = cv.random_vit_state(seed=c["clip_seed"])
    c_tex = cv.clip_cosine(clip_state, st["texture_canvas"], c["text"])
    c_sh = cv.clip_cosine(clip_state, st["shading_canvas"], c["text"]) if no_tex else None
    loss = ol.total_loss(st, c_tex, c_sh, clip_w)
    ref = c["ref"]
    got = {"loss": loss, "color_fine_loss": st["color_loss"], "mask_loss": st["mask_loss"], "eikonal_loss": st["eikonal_loss"],
           "psnr": st["psnr"], "cosine": c_tex, "texture_shading": st["texture_canvas"].reshape(-1, 3)}
    if no_tex:
        got.update({"cosine_shading": c_sh, "rand_shading_rgb": st["shading_canvas"].reshape(-1, 3)})
    else:
        assert ref["cosine_shading"] is None and st["shading_canvas"] is None
    for k, v in got.items():
        assert _rel(v.detach(), ref[k]) < 5e-6, k
    names = ["color_fine", "extra_color_fine", "gradients", "weights", "weight_sum", "gradient_error"]
    grads = torch.autograd.grad(loss, [leaves[k] for k in names], allow_unused=True)
    for k, g in zip(names, grads):
        g = torch.zeros_like(leaves[k]) if g is None else g
        assert _rel(g, c["ref_grads"][k]) < 5e-5, k


def test_ray_generators_match_reference_lines():
    import numpy as np
    from oracle import loss as ol
    from oracle import neus as on
    r = _loss_golden()["rays"]
    Wc = r["canvas"]
    o_all, v_all = ol.pinhole_rays(r["pose"], Wc, Wc)
    dm = r["dilated_mask"]
    assert int(dm.sum()) == r["rays_o"].shape[0] <= r["max_ray_num"] * 1.05
    assert _rel(v_all[dm], r["rays_d"]) < 1e-6 and _rel(o_all[dm], r["rays_o"]) == 0.0
    near, far = on.near_far_from_sphere(r["rays_o"], r["rays_d"])
    assert _rel(near, r["near"]) < 1e-6 and _rel(far, r["far"]) < 1e-6
    # the synthetic workload generator follows the same formulas (avatarclip_b200/workload.py cites dataset.py:259-268)
    from avatarclip_b200.workload import lookat
    assert np.allclose(lookat((0.4, 0.1, 1.5), (0.0, 0.0, 0.0)), ol.lookat((0.4, 0.1, 1.5), (0.0, 0.0, 0.0)))


def test_silhouette_canvas_host_logic_matches_reference_lines():
    """avatarclip_b200.dataset.silhouette_canvas (dilation x10, canvas sizing, nearest resize) against what the
    reference's gen_rays_silhouettes produced for the same silhouette (golden written by oracle/pin_loss_stage.py)."""
    import numpy as np
    from avatarclip_b200.dataset import silhouette_canvas
    r = _loss_golden()["rays"]
    cy, cx = r["sil_center"]
    yy, xx = np.meshgrid(np.arange(256), np.arange(256), indexing="ij")
    sil = (((yy - cy) ** 2 + (xx - cx) ** 2) < r["sil_radius"] ** 2).astype(np.float32)
    canvas, dm = silhouette_canvas(torch.from_numpy(sil), r["max_ray_num"])
    assert canvas == r["canvas"]
    assert torch.equal(dm, r["dilated_mask"])
