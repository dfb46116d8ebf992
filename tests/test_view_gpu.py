"""Per-step view preparation on the device (csrc/avc_view.cu through avatarclip_b200.views) against the numpy oracle
(oracle/raster.py; rasteriser parity is UNPINNED -- neural_renderer is third-party and absent -- the mask handling is
pinned against dataset.py's own lines by oracle/pin_loss_stage.py)."""
import ctypes as C

import numpy as np
import pytest
import torch

from oracle import loss as ol, raster

pytestmark = pytest.mark.gpu


def _builder(max_ray_num=3000, n=128, mask_weight=0.5, lat=12, lon=16):
    from avatarclip_b200.views import ViewBuilder
    from avatarclip_b200.workload import synthetic_body_mesh
    v, f = synthetic_body_mesh(lat, lon)
    return ViewBuilder(v, f, max_ray_num, mask_weight, "cuda", image_size=n), v, f


def _draw(seed=3, face=False, bg=None):
    from avatarclip_b200.sampling import StepSampler
    s = StepSampler(seed=seed, use_face_prompt=face, head_height=0.55)
    d = s.draw(0)
    while bg is not None and d.bg_choice != bg:
        d = s.draw(1)
    return d


@pytest.mark.parametrize("seed,face", [(3, False), (5, False), (9, True)])
def test_raster_dilate_compact_match_oracle(seed, face):
    vb, verts, faces = _builder()
    d = _draw(seed, face)
    p = vb.submit(d)
    torch.cuda.synchronize()
    rgb_o, mask_o = raster.raster_template(verts, faces, d.eye, d.at, n=vb.n, ss=2)
    rgb = p.rgb.cpu().numpy()
    mask = p.mask256.cpu().numpy().astype(bool)
    # fp32 (device) vs fp64 (oracle) coverage tests differ on a few silhouette-edge supersamples only
    assert (mask != mask_o).mean() < 2e-3, (mask != mask_o).sum()
    assert (np.abs(rgb - rgb_o).max(axis=2) > 1e-4).mean() < 5e-3
    assert mask.sum() > 50
    # dilation + count: exact on the device's own silhouette
    dil_o, W_o = raster.dilate_and_canvas(mask, vb.max_ray_num)
    assert np.array_equal(p.dil.cpu().numpy().astype(bool), dil_o)
    assert int(p.count_host[0]) == int(dil_o.sum())
    v = vb.finish(p)
    assert v.W == W_o == v.H
    rm = raster.nearest_resize(dil_o, v.W)
    assert np.array_equal(v.in_mask.cpu().numpy().astype(bool).reshape(v.W, v.W), rm)
    assert np.array_equal(v.pix.cpu().numpy(), np.nonzero(rm.reshape(-1))[0].astype(np.int32))
    assert v.R == int(rm.sum()) and abs(v.R - vb.max_ray_num) < 0.1 * vb.max_ray_num
    # rays of those pixels: the pinned pinhole-ray restatement
    ro, rd = ol.pinhole_rays(d.pose, v.W, v.W, full_res=vb.n)
    sel = torch.from_numpy(rm)
    assert torch.allclose(v.rays_d.cpu(), rd[sel], atol=2e-6) and torch.allclose(v.rays_o.cpu(), ro[sel], atol=1e-6)
    # targets: nearest resize of the template render, mask = channel 0 != 0 (main.py:375-380)
    t_o = raster.nearest_resize(rgb, v.W).reshape(-1, 3)
    assert np.array_equal(v.true_rgb.cpu().numpy(), t_o)
    assert np.array_equal(v.mask.cpu().numpy(), (t_o[:, 0] != 0).astype(np.float32))
    j = v.jitter.cpu()
    assert j.min() >= -0.5 and j.max() < 0.5 and abs(float(j.mean())) < 0.05 and j.unique().numel() > 0.99 * v.R


def test_background_fields():
    vb, _, _ = _builder()
    d1 = _draw(3, bg=1)
    v1 = vb.build(d1)
    g = v1.canvas_background.cpu()
    assert g.min() >= 0 and g.max() <= 1 and abs(float(g.mean()) - 0.5) < 0.02 and abs(float(g.std()) - 0.2) < 0.02
    assert torch.equal(v1.ray_background.cpu(), g[v1.pix.cpu().long()])                 # main.py:412-413
    d2 = _draw(4, bg=2)
    v2 = vb.build(d2)
    want = raster.chessboard_blurred(v2.H, v2.W, max(1, v2.W // d2.chess_div), d2.blur_sigma)
    got = v2.canvas_background.cpu().numpy().reshape(v2.H, v2.W)
    assert np.abs(got - want).max() < 2e-6
    # torchvision's own GaussianBlur on the same board (the object main.py:400 builds), fixed sigma
    import torchvision.transforms as T
    yy, xx = np.meshgrid(np.arange(v2.H), np.arange(v2.W), indexing="ij")
    L = max(1, v2.W // d2.chess_div)
    board = torch.from_numpy(np.where(((yy // L) + (xx // L)) % 2 == 0, 0.8, 0.2).astype(np.float32))
    tv = T.GaussianBlur(kernel_size=(5, 9), sigma=(d2.blur_sigma, d2.blur_sigma))(board[None, None])[0, 0].numpy()
    assert np.abs(got - tv).max() < 2e-6


def test_prepared_view_drives_a_fused_step():
    """ViewBuilder output is what AppearanceTrainer consumes: one finite step with the tcgen05 engine."""
    from avatarclip_b200 import workload as WL
    from avatarclip_b200.clip_vit import ClipImageTower
    from avatarclip_b200.trainer import AppearanceTrainer
    vb, _, _ = _builder(max_ray_num=1500)
    sp, cp = WL.synth_states(WL.S_SDF_KW, WL.S_COL_KW, seed=0)
    _, _, _, ren = WL.build_networks(WL.S_SDF_KW, WL.S_COL_KW, WL.S_REN_KW, sp, cp, 0.3, "cuda", engine=1)
    tower = ClipImageTower(WL.random_vit_state(seed=0), device="cuda")
    tr = AppearanceTrainer(ren, tower, torch.randn(1, 512, generator=torch.Generator().manual_seed(1)))
    losses = []
    for i, bg in enumerate((3, 1, 2, 0)):
        v = vb.build(_draw(20 + i, bg=bg))
        losses.append(float(tr.step(v)))
    assert all(np.isfinite(losses)), losses
