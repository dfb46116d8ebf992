"""Host-side mirror of ``Runner`` (AvatarGen/AppearanceGen/main.py:30-945) for the hot path.

Same constructor arguments, conf keys, checkpoint file layout (``sdf_network_fine`` / ``variance_network_fine`` /
``color_network_fine`` / ``optimizer`` / ``iter_step``) and CLI flags (main.py:953-961) as the reference; the train
step itself is ``avatarclip_b200.trainer.AppearanceTrainer`` (fused C-ABI calls, fused Adam over the flat vector).

What is NOT here and why (SURVEY §8f "next rows"):
* the per-step template silhouette (``render_one_batch`` -> neural_renderer, main.py:360): the rasteriser is a
  third-party CUDA extension that is neither vendored nor installable.  ``train_clip`` therefore takes a
  ``view_source`` callable ``(step) -> avatarclip_b200.workload.HostView``; without one it uses the synthetic
  disc-silhouette views of ``workload.make_view`` and says so.
* ``clip.load`` / ``smplx``: ``init_clip`` takes the visual state dict and the already encoded prompts (or uses the
  ``clip`` package when importable), ``init_smpl`` takes template vertices.
* ``validate_mesh`` / ``render_geometry_cast_light`` (marching cubes, trimesh): next rows.
"""
from __future__ import annotations

import argparse
import logging
import os
from typing import Callable, Optional

import numpy as np
import torch

from . import conf as hocon
from .fields import RenderingNetwork, SDFNetwork, SingleVarianceNetwork
from .renderer import NeuSRenderer


class Runner:
    def __init__(self, conf_path, mode="train", case="CASE_NAME", is_continue=False, is_colab=False, conf=None,
                 device="cuda", engine: int = 1):
        self.device = torch.device(device)
        self.conf_path = conf_path
        if is_colab:
            self.conf = conf
        else:
            with open(self.conf_path) as f:
                conf_text = f.read().replace("CASE_NAME", case)          # main.py:39-42
            self.conf = hocon.parse_string(conf_text)
        c = self.conf
        self.base_exp_dir = c["general.base_exp_dir"]
        self.iter_step = 0
        # training parameters (main.py:49-64)
        self.end_iter = c.get_int("train.end_iter")
        self.save_freq = c.get_int("train.save_freq")
        self.report_freq = c.get_int("train.report_freq")
        self.val_freq = c.get_int("train.val_freq")
        self.batch_size = c.get_int("train.batch_size")
        self.learning_rate = c.get_float("train.learning_rate")
        self.learning_rate_alpha = c.get_float("train.learning_rate_alpha")
        self.warm_up_end = c.get_float("train.warm_up_end", default=0.0)
        self.anneal_end = c.get_float("train.anneal_end", default=0.0)
        self.max_ray_num = c.get_int("train.max_ray_num", default=112 * 112)
        self.igr_weight = c.get_float("train.igr_weight")
        self.mask_weight = c.get_float("train.mask_weight")
        self.clip_weight = c.get_float("train.clip_weight", default=None)
        self.add_no_texture = c.get_bool("train.add_no_texture", default=False)
        self.texture_cast_light = c.get_bool("train.texture_cast_light", default=False)
        self.use_silhouettes = c.get_bool("train.use_silhouettes", default=False)
        self.use_bg_aug = c.get_bool("train.use_bg_aug", default=True)
        self.is_continue, self.mode = is_continue, mode
        # networks (main.py:134-151): the conf subtrees are the constructor kwargs
        self.nerf_outside = None
        self.sdf_network = SDFNetwork(**c["model.sdf_network"]).to(self.device)
        self.deviation_network = SingleVarianceNetwork(**c["model.variance_network"]).to(self.device)
        self.color_network = RenderingNetwork(**c["model.rendering_network"]).to(self.device)
        self.renderer = NeuSRenderer(self.nerf_outside, self.sdf_network, self.deviation_network, self.color_network,
                                     engine=engine, **c["model.neus_renderer"])
        self.trainer = None
        self.clip_tower = None
        self.encoded_text = None
        self._pending_optimizer_state = None
        pretrain = c.get_string("train.pretrain", default=None)
        if pretrain is not None and os.path.exists(pretrain):
            logging.info("Load pretrain: %s", pretrain)
            self.load_pretrain(pretrain)
        if is_continue:
            ckdir = os.path.join(self.base_exp_dir, "checkpoints")
            names = sorted(n for n in os.listdir(ckdir) if n.endswith("pth") and int(n[5:-4]) <= self.end_iter)
            if names:
                self.load_checkpoint(names[-1])

    # ------------------------------------------------------------------ schedules (main.py:571-586)
    def get_cos_anneal_ratio(self):
        return 1.0 if self.anneal_end == 0.0 else float(np.min([1.0, self.iter_step / self.anneal_end]))

    def current_lr(self) -> float:
        if self.iter_step < self.warm_up_end:
            f = self.iter_step / self.warm_up_end
        else:
            a = self.learning_rate_alpha
            progress = (self.iter_step - self.warm_up_end) / (self.end_iter - self.warm_up_end)
            f = (np.cos(np.pi * progress) + 1.0) * 0.5 * (1 - a) + a
        return float(self.learning_rate * f)

    # ------------------------------------------------------------------ CLIP / SMPL seams
    def init_clip(self, visual_state_dict=None, encoded_text=None):
        """main.py:258-288.  With the ``clip`` package importable this does what the reference does; otherwise pass the
        ViT-B/32 visual state dict and the encoded prompt(s) [n,512] (texture prompt first)."""
        from .clip_vit import ClipImageTower
        if visual_state_dict is None:
            import clip                                                   # noqa: F401 (optional dependency)
            model, _ = clip.load("ViT-B/32", jit=False)
            model = model.eval().requires_grad_(False).to(self.device)
            prompt = self.conf.get_string("clip.prompt")
            encoded_text = model.encode_text(clip.tokenize([prompt]).to(self.device)).detach().float()
            visual_state_dict = model.visual.state_dict()
        self.clip_tower = ClipImageTower(visual_state_dict, device=self.device)
        self.encoded_text = encoded_text.detach().float().to(self.device)

    def init_smpl(self, v=None, f=None):
        """main.py:290-335 keeps the posed template (self.v [1,6890,3], self.f) for the silhouette rasteriser."""
        self.v, self.f = v, f

    def _ensure_trainer(self):
        from .trainer import AppearanceTrainer
        if self.trainer is None:
            if self.clip_tower is None:
                raise RuntimeError("call init_clip() first (main.py:970-972)")
            self.trainer = AppearanceTrainer(self.renderer, self.clip_tower, self.encoded_text, lr=self.learning_rate,
                                             igr_weight=self.igr_weight, mask_weight=self.mask_weight,
                                             clip_weight=1.0 if self.clip_weight is None else self.clip_weight,
                                             device=self.device)
            self.trainer.iter_step = self.iter_step
            if self._pending_optimizer_state is not None:      # checkpoint loaded before init_clip() (the CLI order)
                self._load_optimizer_state_dict(self._pending_optimizer_state)
                self._pending_optimizer_state = None
        return self.trainer

    # ------------------------------------------------------------------ train_clip (main.py:337-566)
    def train_clip(self, max_steps: Optional[int] = None, view_source: Optional[Callable] = None, log=print):
        from .trainer import DeviceView
        from .workload import make_view
        tr = self._ensure_trainer()
        if view_source is None:
            log("[avatarclip_b200] no silhouette source given: using synthetic disc silhouettes "
                "(the reference rasterises the SMPL template with neural_renderer, main.py:360)")
            n = min(self.max_ray_num, 112 * 112)
            view_source = lambda step: make_view(step, n_rays=n, H=224, W=224, seed=0,
                                                 bg_choice=int(np.random.choice(4)) if self.use_bg_aug else 3)
        res_step = self.end_iter - self.iter_step
        dv = None
        for it in range(res_step):
            if it == 30010 or (max_steps is not None and it >= max_steps):      # main.py:346-347
                break
            hv = view_source(self.iter_step)
            if dv is None or dv.rays_o.shape != hv.rays_o.shape or dv.H != hv.H or \
                    (dv.ray_background is None) != (hv.ray_background is None):
                dv = DeviceView(hv, self.device)
            else:
                dv.upload(hv)
            loss = tr.step(dv, lr=self.current_lr(), cos_anneal=self.get_cos_anneal_ratio())
            self.iter_step += 1
            if self.iter_step % self.report_freq == 0:
                log("iter:{:8>d} loss = {} lr={}".format(self.iter_step, float(loss), self.current_lr()))
            if self.iter_step % self.save_freq == 0:
                self.save_checkpoint()
        return self.iter_step

    # ------------------------------------------------------------------ checkpoints (main.py:601-632)
    def _optimizer_state_dict(self):
        """torch.optim.Adam-format state dict (parameter order of main.py:141-143: sdf, variance, colour) built from
        the fused Adam's flat moment vectors, so the reference can resume from our checkpoints."""
        params = list(self.sdf_network.parameters()) + list(self.deviation_network.parameters()) + \
            list(self.color_network.parameters())
        state = {}
        tr = self.trainer
        if tr is not None and tr.iter_step > 0:
            slot = {id(p): (o, m) for p, o, m in tr.fp.slots}
            for i, p in enumerate(params):
                o, m = slot[id(p)]
                state[i] = {"step": torch.tensor(float(tr.iter_step)),
                            "exp_avg": tr.exp_avg[o:o + m].view(p.shape).clone(),
                            "exp_avg_sq": tr.exp_avg_sq[o:o + m].view(p.shape).clone()}
        group = {"lr": self.current_lr(), "betas": (0.9, 0.999), "eps": 1e-8, "weight_decay": 0, "amsgrad": False,
                 "maximize": False, "foreach": None, "capturable": False, "differentiable": False, "fused": None,
                 "params": list(range(len(params)))}
        return {"state": state, "param_groups": [group]}

    def _load_optimizer_state_dict(self, sd):
        """optimizer.load_state_dict of main.py:606 for the fused Adam: per-parameter moments -> the flat moment vectors.
        When the trainer does not exist yet (``Runner(..., is_continue=True)`` runs before ``init_clip``, main.py:963-972)
        the state is kept and applied as soon as the trainer is built."""
        if not sd or not sd.get("state"):
            return
        if self.trainer is None:
            self._pending_optimizer_state = sd
            if self.clip_tower is not None:
                self._ensure_trainer()                 # builds the trainer and applies the pending state
            return
        tr = self.trainer
        params = list(self.sdf_network.parameters()) + list(self.deviation_network.parameters()) + \
            list(self.color_network.parameters())
        slot = {id(p): (o, m) for p, o, m in tr.fp.slots}
        for i, p in enumerate(params):
            st = sd["state"].get(i)
            if st is None:
                continue
            o, m = slot[id(p)]
            tr.exp_avg[o:o + m].copy_(st["exp_avg"].reshape(-1).to(self.device))
            tr.exp_avg_sq[o:o + m].copy_(st["exp_avg_sq"].reshape(-1).to(self.device))

    def save_checkpoint(self):
        checkpoint = {
            "sdf_network_fine": self.sdf_network.state_dict(),
            "variance_network_fine": self.deviation_network.state_dict(),
            "color_network_fine": self.color_network.state_dict(),
            "optimizer": self._optimizer_state_dict(),
            "iter_step": self.iter_step,
        }
        os.makedirs(os.path.join(self.base_exp_dir, "checkpoints"), exist_ok=True)
        path = os.path.join(self.base_exp_dir, "checkpoints", "ckpt_{:0>6d}.pth".format(self.iter_step))
        torch.save(checkpoint, path)
        return path

    def load_checkpoint(self, checkpoint_name):
        path = checkpoint_name if os.path.isabs(checkpoint_name) else \
            os.path.join(self.base_exp_dir, "checkpoints", checkpoint_name)
        ck = torch.load(path, map_location=self.device, weights_only=False)
        self.sdf_network.load_state_dict(ck["sdf_network_fine"])
        self.deviation_network.load_state_dict(ck["variance_network_fine"])
        self.color_network.load_state_dict(ck["color_network_fine"])
        self.iter_step = ck["iter_step"]
        if self.trainer is not None:
            self.trainer.iter_step = self.iter_step
        self._load_optimizer_state_dict(ck.get("optimizer", {}))

    def load_pretrain(self, checkpoint_name):
        ck = torch.load(checkpoint_name, map_location=self.device, weights_only=False)
        self.sdf_network.load_state_dict(ck["sdf_network_fine"])
        self.deviation_network.load_state_dict(ck["variance_network_fine"])
        self.color_network.load_state_dict(ck["color_network_fine"], strict=False)      # no extra_lin in the file

    # ------------------------------------------------------------------ validation render (main.py:741-820, render only)
    def render_image(self, pose, resolution_level=1):
        """Chunked full-image render of ``extra_color_fine`` for a camera-to-world pose -> [H, W, 3] tensor."""
        from .dataset import RayGenerator
        rg = RayGenerator(device=self.device)
        ro, rd, near, far = rg.gen_rays_pose(pose, resolution_level)
        H, W = ro.shape[:2]
        with torch.no_grad():
            out = self.renderer.render(ro.reshape(-1, 3), rd.reshape(-1, 3), near, far, perturb_overwrite=0,
                                       background_rgb=None, cos_anneal_ratio=self.get_cos_anneal_ratio())
        return out["extra_color_fine"].reshape(H, W, 3)


def main(argv=None):
    logging.basicConfig(level=logging.INFO, format="[%(filename)s:%(lineno)s - %(funcName)20s() ] %(message)s")
    p = argparse.ArgumentParser()                                   # main.py:953-961
    p.add_argument("--conf", type=str, default="./confs/base.conf")
    p.add_argument("--mode", type=str, default="train")
    p.add_argument("--mcube_threshold", type=float, default=0.0)
    p.add_argument("--is_continue", default=False, action="store_true")
    p.add_argument("--gpu", type=int, default=0)
    p.add_argument("--case", type=str, default="smpl")
    args = p.parse_args(argv)
    torch.cuda.set_device(args.gpu)
    runner = Runner(args.conf, args.mode, args.case, args.is_continue, device=f"cuda:{args.gpu}")
    if args.mode == "train_clip":
        runner.init_clip()
        runner.init_smpl()
        runner.train_clip()
    else:
        raise SystemExit(f"mode {args.mode!r}: only train_clip (the hot path) is implemented; see DESIGN.md")


if __name__ == "__main__":
    main()
