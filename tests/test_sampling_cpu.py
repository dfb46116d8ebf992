"""Per-step draws of Runner.train_clip (cameras, background mode, light, ambience): avatarclip_b200.sampling against the
values the reference's OWN lines produced under the same numpy seed (tests/golden/sampling.json, written by
oracle/pin_sampling.py from main.py:348-359,388-399,433,440 and models/utils.py:6-70).  Bit-exact."""
import json
import os

import numpy as np

from avatarclip_b200.sampling import StepSampler, lookat

GOLDEN = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "sampling.json")


def test_step_draws_match_reference_lines_bit_exactly():
    g = json.load(open(GOLDEN))
    s = StepSampler(seed=g["seed"], use_face_prompt=True, head_height=g["head_height"], use_bg_aug=True)
    modes = set()
    for i, r in enumerate(g["steps"]):
        d = s.draw(i)
        assert d.face_step == (i % 4 == 0)
        assert np.array_equal(d.eye, np.asarray(r["eye"], dtype=np.float32))
        assert np.array_equal(d.at, np.asarray(r["at"], dtype=np.float32))
        assert (d.theta, d.phi, d.is_front) == (r["theta"], r["phi"], r["is_front"])
        assert np.array_equal(d.pose, np.asarray(r["pose"]))
        assert d.bg_choice == r["choice_i"]
        if r["choice_i"] == 2:
            assert 224 // d.chess_div == r["chess_length"]
        assert np.array_equal(d.light_dir, np.asarray(r["light_dir"]).astype(np.float32))
        assert d.ambience == r["ambience"]
        modes.add(d.bg_choice)
        assert d.prompt == ("face" if i % 4 == 0 else ("back" if r["is_front"] == 0 else "body"))
    assert modes == {0, 1, 2, 3}


def test_ablation_conf_streams_match_reference_lines_bit_exactly():
    """confs/ablation/*_0.conf (no background augmentation) and *_1.conf: add_no_texture = texture_cast_light = False, so
    the step has no shading block and the numpy stream no light / ambience draws (main.py:425); no face prompt."""
    g = json.load(open(GOLDEN))
    for name, rec in g["ablation"].items():
        s = StepSampler(seed=g["seed"], use_face_prompt=rec["face"], head_height=g["head_height"], use_bg_aug=rec["bg_aug"],
                        cast_light=rec["shading"])
        for i, r in enumerate(rec["steps"]):
            d = s.draw(i)
            assert not d.face_step
            assert np.array_equal(d.eye, np.asarray(r["eye"], dtype=np.float32)), (name, i)
            assert np.array_equal(d.at, np.asarray(r["at"], dtype=np.float32)), (name, i)
            assert np.array_equal(d.pose, np.asarray(r["pose"])), (name, i)
            assert d.bg_choice == r["choice_i"], (name, i)
            if r["choice_i"] == 2:
                assert 224 // d.chess_div == r["chess_length"], (name, i)
        if not rec["bg_aug"]:
            assert all(r["choice_i"] == 3 for r in rec["steps"])


def test_rank_strided_stream_equals_single_rank_stream():
    """N ranks x 1 view == 1 rank x N views: rank r of N gets draw number step * N + r of the same seeded stream."""
    one = StepSampler(seed=7)
    flat = [one.draw(i) for i in range(8)]
    for rank in range(4):
        s = StepSampler(seed=7)
        for step in range(2):
            d = s.draw_for_rank(step, rank, 4)
            ref = flat[step * 4 + rank]
            assert d.index == ref.index and np.array_equal(d.pose, ref.pose) and d.ambience == ref.ambience


def test_lookat_is_a_rigid_camera_frame():
    m = lookat(np.array([0.3, 0.2, 1.5], dtype=np.float32), np.array([0.0, 0.1, 0.0], dtype=np.float32))
    R = m[:3, :3]
    assert np.allclose(R.T @ R, np.eye(3), atol=1e-6) and abs(np.linalg.det(R) - 1) < 1e-6
    assert np.allclose(m[:3, 3], [0.3, 0.2, 1.5], atol=1e-7)
