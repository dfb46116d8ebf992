"""Per-step random draws of ``Runner.train_clip`` (AvatarGen/AppearanceGen/main.py:348-358,387-405,433,440,498-508) and
the camera helpers they use (models/utils.py:9-70), host side (numpy, fp64 like the reference).

The reference interleaves these draws with the rest of the step on numpy's GLOBAL generator (seeded by ``train.seed``,
main.py:104-114).  ``StepSampler`` reproduces the numpy stream draw for draw, in the reference's order:

    camera       face step:  uniform(0, 2pi) [phi], normal(0, pi/12) [theta]                (utils.py:45-57, main.py:350-353)
                 otherwise:  uniform(1, 2) [distance], uniform(0, 2pi), normal(0, pi/3)     (utils.py:29-43)
                             normal(0, 0.1) x 3 [look-at point]                             (utils.py:66-70)
    background   choice(4) when use_bg_aug; mode 2 additionally choice(arange(10, 20))      (main.py:388-400)
    light        uniform(-pi/4, pi/4) x 2, then ambience uniform(0, 0.2)                    (main.py:433,440)
                 -- only when add_no_texture or texture_cast_light (the block of main.py:425-453; the ablation confs
                 *_0 / *_1 switch both off and the stream then has no light draws)

so that with the same seed the same cameras, backgrounds modes, light directions and ambiences come out
(``oracle/pin_sampling.py`` executes the reference's own lines against this file).  Draws the reference takes from
torch's generator ON THE GPU (Gaussian background field, blur sigma, the per-ray jitter of renderer.py:317-319) cannot
be matched bit for bit across devices anyway; here they come from the sampler's own seeded torch CPU generator (sigma)
or a counter-based generator on the device (fields), keyed by (seed, draw index).

View-sharded multi-GPU: draw number ``step * world + rank`` of ONE seeded stream belongs to rank
``rank`` -- every rank runs the same sampler and keeps its own draws.
"""
from __future__ import annotations

from dataclasses import dataclass
from typing import Optional

import numpy as np

UP = np.array([0, 1, 0])


def lookat(eye, at, up=UP):
    """Camera-to-world matrix (models/utils.py:9-27): columns = camera x, y, z axes and the eye point.  The inputs keep
    their dtypes (the step passes float32 eye / at: the z axis is normalised in float32, the rest promotes to float64,
    exactly as numpy does for the reference), the result is float64."""
    eye, at, up = np.asarray(eye), np.asarray(at), np.asarray(up)
    z = eye - at
    z = z / np.linalg.norm(z)
    x = np.cross(up, z)
    x = x / np.linalg.norm(x)
    y = np.cross(z, x)
    m = np.eye(4)
    m[:3, 0], m[:3, 1], m[:3, 2], m[:3, 3] = x, y, z, eye
    return m


def sphere_coord(theta, phi, r=1.0):
    """models/utils.py:59-64 (polar angle theta measured from +z, azimuth phi in the x-y plane)."""
    return np.array([r * np.sin(theta) * np.cos(phi), r * np.sin(theta) * np.sin(phi), r * np.cos(theta)])


def random_eye_normal(rng=np.random):
    """models/utils.py:29-43 -> (eye, theta, phi, is_front)."""
    distance = rng.uniform(1, 2)
    phi = rng.uniform(0, 2 * np.pi)
    theta = rng.normal(0, np.pi / 3)
    is_front = 0 if (theta > np.pi / 2 or theta < -np.pi / 2) else 1
    return sphere_coord(theta, phi, distance), theta, phi, is_front


def random_eye(is_front=None, distance=None, theta_std=None, rng=np.random):
    """models/utils.py:45-57 -> (eye, theta, phi, is_front)."""
    distance = rng.uniform(1, 2) if distance is None else distance
    phi = rng.uniform(0, 2 * np.pi)
    theta = np.clip(rng.normal(0, np.pi / 6 if theta_std is None else theta_std), -np.pi / 2, np.pi / 2)
    is_front = rng.choice(2) if is_front is None else is_front
    if is_front == 0:
        theta += np.pi
    return sphere_coord(theta, phi, distance), theta, phi, is_front


def random_at(rng=np.random):
    """models/utils.py:66-70."""
    return rng.normal(np.array([0, 0, 0]), np.array([0.1, 0.1, 0.1])).clip(-0.3, 0.3)


@dataclass
class StepDraw:
    index: int                 # position in the seeded stream (= step * world + rank)
    face_step: bool            # main.py:349: use_face_prompt and iter_i % 4 == 0
    eye: np.ndarray            # [3] float32 (already offset by `at`, main.py:353,358)
    at: np.ndarray             # [3] float32
    theta: float
    phi: float
    is_front: int
    pose: np.ndarray           # [4,4] float64 camera-to-world
    bg_choice: int             # 0 white, 1 Gaussian field, 2 blurred chessboard, 3 black (main.py:387-405)
    chess_div: int             # main.py:398: H // chess_div squares (only meaningful for bg_choice 2)
    blur_sigma: float          # torchvision GaussianBlur.get_params: uniform(0.1, 2.0) (only for bg_choice 2)
    light_dir: np.ndarray      # [3] float32 (main.py:433)
    ambience: float            # main.py:440
    field_seed: int            # key of the device-side counter-based generator (Gaussian field, jitter)

    @property
    def prompt(self) -> str:
        """Which cached text embedding the step uses (main.py:499-507): 'face' | 'back' (caller checks use_back_prompt)
        | 'body'."""
        return "face" if self.face_step else ("back" if self.is_front == 0 else "body")


class StepSampler:
    def __init__(self, seed: Optional[int] = None, use_face_prompt: bool = False, head_height: float = 0.65,
                 use_bg_aug: bool = True, rng=None, cast_light: bool = True):
        """``seed`` None: numpy's global generator, as the reference when ``train.seed`` is absent; an int: a private
        ``RandomState(seed)`` -- the same stream ``np.random.seed(seed)`` gives the reference (main.py:104-110)."""
        self.rng = rng if rng is not None else (np.random if seed is None else np.random.RandomState(seed))
        self.seed = 0 if seed is None else int(seed)
        self.use_face_prompt, self.head_height, self.use_bg_aug = bool(use_face_prompt), float(head_height), bool(use_bg_aug)
        self.cast_light = bool(cast_light)      # train.add_no_texture or train.texture_cast_light (main.py:425)
        import torch
        self._tgen = torch.Generator().manual_seed(self.seed)
        self.count = 0

    def draw(self, iter_i: int) -> StepDraw:
        """The draws of loop iteration ``iter_i`` (its parity with 4 selects face steps, main.py:349)."""
        import torch
        rng = self.rng
        face = self.use_face_prompt and iter_i % 4 == 0
        if face:
            eye, theta, phi, is_front = random_eye(is_front=1, distance=0.4, theta_std=np.pi / 12, rng=rng)
            at = np.array([0, self.head_height, 0.3]).astype(np.float32)
        else:
            eye, theta, phi, is_front = random_eye_normal(rng=rng)
            at = random_at(rng=rng).astype(np.float32)
        eye = eye.astype(np.float32)
        eye += at
        pose = lookat(eye, at, UP)
        bg_choice = int(rng.choice(4)) if self.use_bg_aug else 3
        chess_div, sigma = 10, 1.0
        if bg_choice == 2:
            chess_div = int(rng.choice(np.arange(10, 20)))
            sigma = float(torch.empty(1).uniform_(0.1, 2.0, generator=self._tgen))
        if self.cast_light:
            light = sphere_coord(theta + rng.uniform(-np.pi / 4, np.pi / 4), phi + rng.uniform(-np.pi / 4, np.pi / 4))
            ambience = float(rng.uniform(0, 0.2))
        else:                                   # no shading block in the step: no draws, the values are never used
            light, ambience = np.array([0.0, 0.0, 1.0]), 0.0
        d = StepDraw(self.count, face, eye, at, float(theta), float(phi), int(is_front), pose, bg_choice, chess_div, sigma,
                     light.astype(np.float32), ambience, (self.seed * 1000003 + self.count) & 0x7FFFFFFF)
        self.count += 1
        return d

    def draw_for_rank(self, step: int, rank: int = 0, world: int = 1) -> StepDraw:
        """Advance the shared stream by ``world`` draws and return this rank's (index step * world + rank)."""
        mine = None
        for r in range(world):
            d = self.draw(step * world + r)
            if r == rank:
                mine = d
        return mine
