"""Host-side mirror of the reference's network classes (AvatarGen/AppearanceGen/models/fields.py).

Same class names, constructor keywords and state-dict keys (``linK.weight_g`` [out,1],
``linK.weight_v`` [out,in], ``linK.bias``, ``extra_lin.*``, ``variance``) as the reference, so the
shipped ``.pth`` checkpoints load unchanged and checkpoints written here load in the reference.
The modules hold parameters only; every forward / backward runs in libavc_b200.so through
``avatarclip_b200.renderer``.  There is no PyTorch compute path in this file.
"""
from __future__ import annotations

import math
from typing import Sequence

import numpy as np
import torch
import torch.nn as nn


class WNLinear(nn.Module):
    """Parameter holder with the layout of ``nn.utils.weight_norm(nn.Linear(i, o))``:
    parameters registered in the order bias, weight_g, weight_v (what the reference's modules
    expose to ``named_parameters`` / the optimizer)."""

    def __init__(self, in_features: int, out_features: int, weight: torch.Tensor, bias: torch.Tensor):
        super().__init__()
        self.in_features, self.out_features = in_features, out_features
        self.bias = nn.Parameter(bias.detach().clone())
        self.weight_g = nn.Parameter(weight.detach().norm(dim=1, keepdim=True).clone())
        self.weight_v = nn.Parameter(weight.detach().clone())

    def extra_repr(self):
        return f"in_features={self.in_features}, out_features={self.out_features}, weight_norm=True"


def _default_linear_init(in_f: int, out_f: int):
    """nn.Linear's default initialisation (kaiming_uniform(a=sqrt(5)) + uniform bias)."""
    lin = nn.Linear(in_f, out_f)
    return lin.weight.detach(), lin.bias.detach()


class SDFNetwork(nn.Module):
    """models/fields.py:9-107.  ``forward`` / ``sdf`` / ``sdf_hidden_appearance`` / ``gradient``
    evaluate through the CUDA library (no autograd through these convenience methods -- training
    goes through ``NeuSRenderer.render``)."""

    def __init__(self, d_in, d_out, d_hidden, n_layers, skip_in: Sequence[int] = (4,), multires=0, bias=0.5,
                 scale=1, geometric_init=True, weight_norm=True, inside_outside=False):
        super().__init__()
        if not weight_norm:
            raise NotImplementedError("avatarclip_b200 implements the weight_norm=True configuration "
                                      "(every shipped conf); see DESIGN.md")
        dims = [d_in] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.multires = int(multires)
        if multires > 0:
            dims[0] = d_in * (1 + 2 * multires)
        self.d_in, self.d_out, self.d_hidden, self.n_layers = d_in, d_out, d_hidden, n_layers
        self.num_layers = len(dims)
        self.skip_in = tuple(int(s) for s in skip_in)
        self.scale = float(scale)
        for l in range(0, self.num_layers - 1):
            out_dim = dims[l + 1] - dims[0] if (l + 1) in self.skip_in else dims[l + 1]
            w, b = _default_linear_init(dims[l], out_dim)
            if geometric_init:   # models/fields.py:45-63
                if l == self.num_layers - 2:
                    if not inside_outside:
                        torch.nn.init.normal_(w, mean=np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                        torch.nn.init.constant_(b, -bias)
                    else:
                        torch.nn.init.normal_(w, mean=-np.sqrt(np.pi) / np.sqrt(dims[l]), std=0.0001)
                        torch.nn.init.constant_(b, bias)
                elif multires > 0 and l == 0:
                    torch.nn.init.constant_(b, 0.0)
                    torch.nn.init.constant_(w[:, 3:], 0.0)
                    torch.nn.init.normal_(w[:, :3], 0.0, np.sqrt(2) / np.sqrt(out_dim))
                elif multires > 0 and l in self.skip_in:
                    torch.nn.init.constant_(b, 0.0)
                    torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
                    torch.nn.init.constant_(w[:, -(dims[0] - 3):], 0.0)
                else:
                    torch.nn.init.constant_(b, 0.0)
                    torch.nn.init.normal_(w, 0.0, np.sqrt(2) / np.sqrt(out_dim))
            setattr(self, "lin" + str(l), WNLinear(dims[l], out_dim, w, b))

    # -- convenience evaluators (inference only) -------------------------------------------
    def _binding(self):
        b = getattr(self, "_avc_binding", None)
        if b is None:
            raise RuntimeError("SDFNetwork is not attached to a NeuSRenderer yet; construct "
                               "avatarclip_b200.NeuSRenderer(nerf, sdf_network, ...) first")
        return b

    def sdf(self, x):
        return self._binding().sdf_query(x)

    def forward(self, inputs):
        """models/fields.py:72-88: [P,3] -> [P, d_out] = (sdf, feature vector)."""
        return self._binding().sdf_eval(inputs, want_features=True, want_gradient=False)[0]

    def sdf_hidden_appearance(self, x):
        """models/fields.py:93-94."""
        return self.forward(x)

    def gradient(self, x):
        """models/fields.py:96-107: d sdf / d x, [P,3] -> [P,1,3] (evaluated analytically by the reverse sweep of
        SURVEY Appendix B; inference only -- training differentiates through NeuSRenderer.render)."""
        return self._binding().sdf_eval(x, want_features=False, want_gradient=True)[1].unsqueeze(1)


class RenderingNetwork(nn.Module):
    """models/fields.py:111-185 (parameters only)."""

    def __init__(self, d_feature, mode, d_in, d_out, d_hidden, n_layers, weight_norm=True, multires_view=0,
                 squeeze_out=True, extra_color=False):
        super().__init__()
        if not weight_norm or mode != "no_view_dir" or multires_view != 0 or not squeeze_out or d_in != 6 or d_out != 3:
            raise NotImplementedError(
                "avatarclip_b200 implements mode='no_view_dir', multires_view=0, squeeze_out, weight_norm, d_in=6, "
                "d_out=3 -- the configuration of every shipped conf; see DESIGN.md")
        self.mode, self.squeeze_out, self.extra_color = mode, squeeze_out, bool(extra_color)
        self.d_feature, self.d_hidden, self.n_layers = d_feature, d_hidden, n_layers
        dims = [d_in + d_feature] + [d_hidden for _ in range(n_layers)] + [d_out]
        self.num_layers = len(dims)
        for l in range(0, self.num_layers - 1):
            w, b = _default_linear_init(dims[l], dims[l + 1])
            setattr(self, "lin" + str(l), WNLinear(dims[l], dims[l + 1], w, b))
        if self.extra_color:                                  # models/fields.py:147-150
            w, b = _default_linear_init(dims[self.num_layers - 2], d_out)
            self.extra_lin = WNLinear(dims[self.num_layers - 2], d_out, w, b)
        # extra_color=False (confs/base_models/astrongman.conf, the --mode train pre-fit): no second head, no extra
        # parameters, no extra draws from torch's generator -- exactly the reference's module.  The kernels' slot for the
        # head is then a constant zero map inside the flat parameter vector (renderer.FlatParams), not a Parameter.


class SingleVarianceNetwork(nn.Module):
    """models/fields.py:270-276."""

    def __init__(self, init_val):
        super().__init__()
        self.register_parameter("variance", nn.Parameter(torch.tensor(float(init_val))))

    def forward(self, x):
        return torch.ones([len(x), 1], device=self.variance.device) * torch.exp(self.variance * 10.0)
