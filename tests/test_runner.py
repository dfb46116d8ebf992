"""Runner mirror: conf-driven construction, LR schedule, checkpoint layout (CPU) and a short train_clip run (GPU)."""
import os

import numpy as np
import pytest
import torch

from oracle import loss as ol

HERE = os.path.dirname(os.path.abspath(__file__))
CONF = os.path.join(HERE, "runner_conf_sample.conf")


def _runner(tmp_path, device):
    from avatarclip_b200.runner import Runner
    r = Runner(CONF, mode="train_clip", case="smpl", device=device)
    r.base_exp_dir = str(tmp_path / "exp")
    return r


def test_runner_construction_schedule_and_checkpoint_layout(tmp_path):
    r = _runner(tmp_path, "cpu")
    assert r.conf["general.base_exp_dir"] == "./exp/smpl/demo"          # CASE_NAME substitution (main.py:41)
    for it in (0, 100, 499, 500, 5000, 99999):
        r.iter_step = it
        want = 5e-4 * ol.learning_rate_factor(it, 500.0, 100000, 0.05)
        assert abs(r.current_lr() - want) < 1e-12
    assert r.get_cos_anneal_ratio() == 1.0
    r.iter_step = 7
    path = r.save_checkpoint()
    ck = torch.load(path, weights_only=False)
    assert set(ck.keys()) == {"sdf_network_fine", "variance_network_fine", "color_network_fine", "optimizer", "iter_step"}
    assert list(ck["sdf_network_fine"].keys())[:3] == ["lin0.bias", "lin0.weight_g", "lin0.weight_v"]
    assert ck["iter_step"] == 7 and os.path.basename(path) == "ckpt_000007.pth"
    # the reference's optimizer (main.py:141-145) accepts the stored optimizer state
    params = list(r.sdf_network.parameters()) + list(r.deviation_network.parameters()) + list(r.color_network.parameters())
    torch.optim.Adam(params, lr=5e-4).load_state_dict(ck["optimizer"])
    r2 = _runner(tmp_path, "cpu")
    r2.load_checkpoint(path)
    assert r2.iter_step == 7
    for a, b in zip(r.sdf_network.parameters(), r2.sdf_network.parameters()):
        assert torch.equal(a, b)


@pytest.mark.gpu
def test_runner_train_clip_runs_and_resumes(tmp_path):
    from oracle import clip_vit as cv
    r = _runner(tmp_path, "cuda")
    text = torch.randn(1, 512, generator=torch.Generator().manual_seed(0))
    r.init_clip(cv.random_vit_state(seed=0), text)
    r.init_smpl()
    logs = []
    before = r.sdf_network.lin1.weight_v.detach().clone()
    r.report_freq = 2
    assert r.train_clip(max_steps=4, log=logs.append) == 4
    assert not torch.equal(before, r.sdf_network.lin1.weight_v)         # parameters moved
    assert any("loss" in str(l) for l in logs)
    path = r.save_checkpoint()
    ck = torch.load(path, weights_only=False)
    params = list(r.sdf_network.parameters()) + list(r.deviation_network.parameters()) + list(r.color_network.parameters())
    opt = torch.optim.Adam(params, lr=5e-4)
    opt.load_state_dict(ck["optimizer"])                                # reference-format Adam state
    assert len(ck["optimizer"]["state"]) == len(params)
    r2 = _runner(tmp_path, "cuda")
    r2.init_clip(cv.random_vit_state(seed=0), text)
    r2.load_checkpoint(path)
    assert r2.iter_step == 4 and torch.equal(r2.trainer.exp_avg, r.trainer.exp_avg)
    img = r2.render_image(ol.lookat([0.0, 0.0, 1.6], [0.0, 0.0, 0.0]), resolution_level=8)
    assert img.shape == (32, 32, 3) and torch.isfinite(img).all()
