"""Pin avatarclip_b200/sampling.py (cameras and per-step numpy draws of Runner.train_clip) against the UNMODIFIED
reference source, executed in place (test infrastructure; build container only: python -m oracle.pin_sampling).

Like oracle/pin_loss_stage.py this cuts the cited LINE RANGES out of the reference files and ``exec``s them -- the
arithmetic and the order of the numpy draws are the reference's own text:

    models/utils.py:6-70         norm_np_arr, lookat, random_eye_normal, random_eye, sphere_coord, random_at
    main.py:348-359              camera block of the step (face step / normal step, eye += at, lookat)
    main.py:388-391,399          background mode draw, chessboard divisor draw
    main.py:433,440              light direction, ambience

and replays ``np.random.seed(s)`` on both sides for 24 consecutive steps with use_face_prompt on (every 4th step is a
face step) and background augmentation on.  Bit-exact equality is asserted for every numpy-derived quantity; the
reference's values are written to tests/golden/sampling.json so the CPU test-suite replays them without /root/reference.
The streams of the ablation confs (no face prompt; with / without background augmentation; no shading block, i.e. no
light / ambience draws, main.py:425) are replayed the same way for 12 steps each.
"""
from __future__ import annotations

import json
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from oracle.pin_loss_stage import cut          # noqa: E402
from avatarclip_b200.sampling import StepSampler  # noqa: E402

GOLDEN = os.path.join(ROOT, "tests", "golden", "sampling.json")


def reference_draws(seed: int, n_steps: int, head_height: float, face: bool = True, bg_aug: bool = True,
                    shading: bool = True):
    ns = {"np": np}
    exec(cut("models/utils.py", 6, 70, "def norm_np_arr", ").clip(-0.3, 0.3)"), ns)
    cam = cut("main.py", 348, 359, "if self.use_face_prompt and iter_i % 4 == 0", "pose = lookat(eye, at")
    bgc = cut("main.py", 388, 391, "if self.use_bg_aug", "choice_i = 3")
    chess = cut("main.py", 399, 399, "chess_length = H // np.random.choice", "np.arange(10,20))")
    light = cut("main.py", 433, 433, "light_dir = sphere_coord(theta + np.random.uniform", "")
    amb = cut("main.py", 440, 440, "ambience = np.random.uniform(0, 0.2)", "")

    class Self:
        use_face_prompt, use_bg_aug = face, bg_aug
    Self.head_height = head_height
    np.random.seed(seed)
    out = []
    for iter_i in range(n_steps):
        loc = dict(ns, self=Self, iter_i=iter_i, H=224)
        exec(cam, loc)
        exec(bgc, loc)
        if loc["choice_i"] == 2:
            exec(chess, loc)
        if shading:               # main.py:425: `if self.add_no_texture or self.texture_cast_light:` guards both draws
            exec(light, loc)
            exec(amb, loc)
        out.append({"eye": loc["eye"].tolist(), "at": np.asarray(loc["at"]).tolist(), "theta": float(loc["theta"]),
                    "phi": float(loc["phi"]), "is_front": int(loc["is_front"]), "pose": loc["pose"].tolist(),
                    "choice_i": int(loc["choice_i"]),
                    "chess_length": int(loc["chess_length"]) if loc["choice_i"] == 2 else None,
                    "light_dir": loc["light_dir"].tolist() if shading else None,
                    "ambience": float(loc["ambience"]) if shading else None})
    return out


def compare(ref, seed, head_height, face=True, bg_aug=True, shading=True):
    s = StepSampler(seed=seed, use_face_prompt=face, head_height=head_height, use_bg_aug=bg_aug, cast_light=shading)
    for i, r in enumerate(ref):
        d = s.draw(i)
        assert d.face_step == (face and i % 4 == 0)
        assert np.array_equal(d.eye, np.asarray(r["eye"], dtype=np.float32)), (i, "eye")
        assert np.array_equal(d.at, np.asarray(r["at"], dtype=np.float32)), (i, "at")
        assert d.theta == r["theta"] and d.phi == r["phi"] and d.is_front == r["is_front"], (i, "angles")
        assert np.array_equal(d.pose, np.asarray(r["pose"])), (i, "pose")
        assert d.bg_choice == r["choice_i"], (i, "bg")
        if r["choice_i"] == 2:
            assert 224 // d.chess_div == r["chess_length"], (i, "chess")
        if shading:
            assert np.array_equal(d.light_dir, np.asarray(r["light_dir"]).astype(np.float32)), (i, "light")
            assert d.ambience == r["ambience"], (i, "ambience")


def main():
    seed, n, hh = 2022, 24, 0.65
    ref = reference_draws(seed, n, hh)
    compare(ref, seed, hh)
    # the streams of the ablation confs (confs/ablation/*_0.conf: no background augmentation, no shading block;
    # *_1.conf: background augmentation, no shading block; neither uses the face / back prompts)
    ablation = {}
    for name, kw in (("ablation_0", dict(face=False, bg_aug=False, shading=False)),
                     ("ablation_1", dict(face=False, bg_aug=True, shading=False))):
        r = reference_draws(seed, 12, hh, **kw)
        compare(r, seed, hh, **kw)
        ablation[name] = dict(kw, steps=r)
    json.dump({"seed": seed, "head_height": hh, "steps": ref, "ablation": ablation}, open(GOLDEN, "w"))
    modes = sorted(set(r["choice_i"] for r in ref))
    print(f"[pin] sampling: {n} steps bit-exact against main.py:348-359,388-399,433,440 + utils.py:6-70 "
          f"(background modes seen: {modes}); wrote {GOLDEN}")


if __name__ == "__main__":
    main()
