"""LBS kernel against (a) the golden vector recorded from the reference's own my_lbs (oracle/pin_lbs.py) and
(b) the restated oracle on an SMPL-sized synthetic body (6890 vertices)."""
import os

import pytest
import torch

import util_neus as U
from oracle import lbs as olbs

pytestmark = pytest.mark.gpu


def _run(d):
    from avatarclip_b200.lbs import my_lbs
    c = lambda t: t.cuda()
    v, j = my_lbs(c(d["v_shaped"]), c(d["pose"]), None, None, c(d["posedirs"]), c(d["J_regressor"]), c(d["parents"]),
                  c(d["lbs_weights"]), pose2rot=True)
    return v.cpu(), j.cpu()


def test_lbs_matches_reference_golden():
    g = torch.load(os.path.join(U.GOLDEN, "lbs_small.pt"), map_location="cpu", weights_only=False)
    v, j = _run(g["inputs"])
    assert (v - g["verts"]).abs().max().item() < 1e-5      # <= 1e-5 abs
    assert (j - g["joints"]).abs().max().item() < 1e-5


def test_lbs_smpl_sized():
    d = olbs.synthetic_smpl(V=6890, seed=1)
    v, j = _run(d)
    vo, jo = olbs.my_lbs(**d)
    assert (v - vo).abs().max().item() < 1e-5
    assert (j - jo).abs().max().item() < 1e-5
