"""CPU oracle (test infrastructure, never on the product path): CLIP ViT-B/32 image tower.

PARITY UNPINNED.  The arithmetic lives in a third-party dependency that is NOT under
/root/reference: ``git+https://github.com/openai/CLIP.git`` (unpinned, AvatarCLIP
``requirements.txt:12``); the reference only calls it (AvatarGen/AppearanceGen/main.py:259-261
``clip.load('ViT-B/32')`` -> eval, frozen; :512,518,524 ``perceptor.encode_image``;
:513-526 cosine against the cached text embedding).  No reference test or golden vector exists
for this boundary and the weights are not on disk, so this file restates the published
architecture (openai/CLIP ``clip/model.py``: ``VisionTransformer``, ``ResidualAttentionBlock``,
``QuickGELU``, fp32 ``LayerNorm``) and ``oracle/pin_clip.py`` cross-checks the restatement against
the independent HuggingFace ``transformers`` implementation of the same architecture on seeded
random weights.

State-dict keys follow openai/CLIP (``visual.*`` prefix stripped): conv1.weight [768,3,32,32],
class_embedding [768], positional_embedding [50,768], ln_pre.{weight,bias},
transformer.resblocks.{i}.{ln_1,ln_2}.{weight,bias}, .attn.in_proj_{weight,bias},
.attn.out_proj.{weight,bias}, .mlp.c_fc.{weight,bias}, .mlp.c_proj.{weight,bias},
ln_post.{weight,bias}, proj [768,512].
"""
from __future__ import annotations

import math
from dataclasses import dataclass
from typing import Dict, Optional

import torch
import torch.nn.functional as F

CLIP_MEAN = (0.48145466, 0.4578275, 0.40821073)     # main.py:261
CLIP_STD = (0.26862954, 0.26130258, 0.27577711)


@dataclass
class ViTConf:
    image_size: int = 224
    patch: int = 32
    width: int = 768
    layers: int = 12
    heads: int = 12
    mlp: int = 3072
    out_dim: int = 512

    @property
    def grid(self):
        return self.image_size // self.patch

    @property
    def tokens(self):
        return self.grid * self.grid + 1


def random_vit_state(conf: ViTConf = ViTConf(), seed: int = 0, dtype=torch.float32) -> Dict[str, torch.Tensor]:
    """Seeded random weights with openai/CLIP's initialisation scales (clip/model.py
    ``VisionTransformer.__init__`` and ``CLIP.initialize_parameters``), rounded to fp16 values
    (``clip.load`` keeps fp16 weights on CUDA) and returned as ``dtype``."""
    g = torch.Generator().manual_seed(seed)
    W = conf.width
    scale = W ** -0.5
    rn = lambda *s, std=1.0: torch.randn(*s, generator=g) * std
    sd: Dict[str, torch.Tensor] = {}
    sd["conv1.weight"] = rn(W, 3, conf.patch, conf.patch, std=(3 * conf.patch * conf.patch) ** -0.5)
    sd["class_embedding"] = rn(W, std=scale)
    sd["positional_embedding"] = rn(conf.tokens, W, std=scale)
    for name in ("ln_pre", "ln_post"):
        sd[f"{name}.weight"] = 1.0 + 0.05 * rn(W)
        sd[f"{name}.bias"] = 0.05 * rn(W)
    proj_std = (W ** -0.5) * ((2 * conf.layers) ** -0.5)
    attn_std = W ** -0.5
    fc_std = (2 * W) ** -0.5
    for i in range(conf.layers):
        p = f"transformer.resblocks.{i}."
        for ln in ("ln_1", "ln_2"):
            sd[p + ln + ".weight"] = 1.0 + 0.05 * rn(W)
            sd[p + ln + ".bias"] = 0.05 * rn(W)
        sd[p + "attn.in_proj_weight"] = rn(3 * W, W, std=attn_std)
        sd[p + "attn.in_proj_bias"] = 0.02 * rn(3 * W)
        sd[p + "attn.out_proj.weight"] = rn(W, W, std=proj_std)
        sd[p + "attn.out_proj.bias"] = 0.02 * rn(W)
        sd[p + "mlp.c_fc.weight"] = rn(conf.mlp, W, std=fc_std)
        sd[p + "mlp.c_fc.bias"] = 0.02 * rn(conf.mlp)
        sd[p + "mlp.c_proj.weight"] = rn(W, conf.mlp, std=proj_std)
        sd[p + "mlp.c_proj.bias"] = 0.02 * rn(W)
    sd["proj"] = rn(W, conf.out_dim, std=scale)
    return {k: v.half().to(dtype) for k, v in sd.items()}


def quick_gelu(x):
    return x * torch.sigmoid(1.702 * x)


def encode_image(sd: Dict[str, torch.Tensor], image: torch.Tensor, conf: ViTConf = ViTConf()) -> torch.Tensor:
    """``VisionTransformer.forward`` (openai/CLIP clip/model.py): image [B,3,224,224] (already
    normalised) -> embedding [B, out_dim]."""
    B = image.shape[0]
    W, Hh = conf.width, conf.heads
    x = F.conv2d(image, sd["conv1.weight"], stride=conf.patch)              # [B, W, g, g]
    x = x.reshape(B, W, -1).permute(0, 2, 1)                                # [B, g*g, W]
    cls = sd["class_embedding"].to(x.dtype) + torch.zeros(B, 1, W, dtype=x.dtype)
    x = torch.cat([cls, x], dim=1) + sd["positional_embedding"]
    x = F.layer_norm(x, (W,), sd["ln_pre.weight"], sd["ln_pre.bias"], 1e-5)
    hd = W // Hh
    for i in range(conf.layers):
        p = f"transformer.resblocks.{i}."
        h = F.layer_norm(x, (W,), sd[p + "ln_1.weight"], sd[p + "ln_1.bias"], 1e-5)
        qkv = F.linear(h, sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"])
        q, k, v = qkv.chunk(3, dim=-1)
        sh = lambda t: t.reshape(B, -1, Hh, hd).permute(0, 2, 1, 3)       # [B, heads, T, hd]
        q, k, v = sh(q), sh(k), sh(v)
        att = torch.softmax((q @ k.transpose(-1, -2)) / math.sqrt(hd), dim=-1)
        o = (att @ v).permute(0, 2, 1, 3).reshape(B, -1, W)
        x = x + F.linear(o, sd[p + "attn.out_proj.weight"], sd[p + "attn.out_proj.bias"])
        h = F.layer_norm(x, (W,), sd[p + "ln_2.weight"], sd[p + "ln_2.bias"], 1e-5)
        h = quick_gelu(F.linear(h, sd[p + "mlp.c_fc.weight"], sd[p + "mlp.c_fc.bias"]))
        x = x + F.linear(h, sd[p + "mlp.c_proj.weight"], sd[p + "mlp.c_proj.bias"])
    x = F.layer_norm(x[:, 0, :], (W,), sd["ln_post.weight"], sd["ln_post.bias"], 1e-5)
    return x @ sd["proj"]


def preprocess(canvas_hw3: torch.Tensor, size: int = 224) -> torch.Tensor:
    """main.py:509-511: ``RandomResizedCrop(224, scale=(1,1))`` on a square [H,W,3] canvas is a
    whole-image bilinear resize (ratio range collapses to the full crop; SURVEY.md 3.1 probe) with the
    torchvision>=0.17 tensor default ``antialias=True``... the reference era (torchvision 0.8,
    README.md:122) resized tensors WITHOUT antialiasing; that is what is restated here
    (``align_corners=False``).  Then ``Normalize(CLIP_MEAN, CLIP_STD)``.  Returns [1,3,size,size]."""
    img = canvas_hw3.permute(2, 0, 1).unsqueeze(0)
    img = F.interpolate(img, size=(size, size), mode="bilinear", align_corners=False, antialias=False)
    mean = torch.tensor(CLIP_MEAN, dtype=img.dtype).view(1, 3, 1, 1)
    std = torch.tensor(CLIP_STD, dtype=img.dtype).view(1, 3, 1, 1)
    return (img - mean) / std


def clip_cosine(sd, canvas_hw3: torch.Tensor, text_emb: torch.Tensor, conf: ViTConf = ViTConf()) -> torch.Tensor:
    """main.py:510-514: cosine between mean_b(image embedding) and mean_b(text embedding)."""
    emb = encode_image(sd, preprocess(canvas_hw3, conf.image_size), conf)
    return torch.cosine_similarity(emb.mean(0), text_emb.reshape(-1, text_emb.shape[-1]).mean(0), dim=0)
