"""Pin oracle/lbs.py against the reference's own ``my_lbs`` (models/utils.py:176-224) run in THIS container.

``models/utils.py`` star-imports ``smplx.lbs`` (not installed) and ``neural_renderer`` (not installable); the
smplx helpers it needs exist verbatim in-tree in ``drive.py:51-160``.  This script executes exactly those in-tree
function definitions into a stand-in ``smplx.lbs`` module, stubs ``neural_renderer`` (unused by ``my_lbs``), imports
the UNMODIFIED ``models/utils.py`` and compares ``my_lbs`` on synthetic SMPL-shaped tensors with the restatement.
Nothing is copied into this repository; a golden vector is written to tests/golden/lbs_small.pt.

    python -m oracle.pin_lbs
"""
import os
import sys
import types

import torch

REF_AG = "/root/reference/AvatarGen/AppearanceGen"
GOLDEN = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests", "golden")


def main():
    src = open(os.path.join(REF_AG, "drive.py")).read().split("\n")
    # the block of smplx.lbs copies: from `def vertices2joints` up to (not including) the first non-lbs helper
    start = next(i for i, l in enumerate(src) if l.startswith("def vertices2joints"))
    end = next(i for i, l in enumerate(src) if i > start and l.startswith("def ") and
               l.split("(")[0][4:] not in ("vertices2joints", "blend_shapes", "batch_rigid_transform", "transform_mat",
                                           "batch_rodrigues"))
    ns = {}
    exec("import torch\nimport torch.nn.functional as F\nfrom torch import Tensor\nfrom typing import Tuple, List, Optional, Dict, Union\nimport numpy as np\n"
         + "\n".join(src[start:end]), ns)
    fake = types.ModuleType("smplx.lbs")
    for k, v in ns.items():
        if callable(v) and not k.startswith("_"):
            setattr(fake, k, v)
    fake.__all__ = [k for k in dir(fake) if not k.startswith("_")]
    pkg = types.ModuleType("smplx")
    pkg.lbs = fake
    sys.modules["smplx"], sys.modules["smplx.lbs"] = pkg, fake
    sys.modules["neural_renderer"] = types.ModuleType("neural_renderer")
    sys.path.insert(0, REF_AG)
    from models import utils as ref_utils          # the unmodified reference file

    from oracle import lbs
    d = lbs.synthetic_smpl(V=240, seed=3)
    v_ref, j_ref = ref_utils.my_lbs(d["v_shaped"], d["pose"], None, None, d["posedirs"], d["J_regressor"], d["parents"],
                                    d["lbs_weights"], pose2rot=True)
    v_or, j_or = lbs.my_lbs(**d)
    ev = (v_ref - v_or).abs().max().item()
    ej = (j_ref - j_or).abs().max().item()
    print(f"[pin_lbs] restated my_lbs vs reference my_lbs (in-tree smplx.lbs copies): max abs err verts {ev:.2e}, joints {ej:.2e}")
    assert ev < 1e-5 and ej < 1e-5
    os.makedirs(GOLDEN, exist_ok=True)
    torch.save({"inputs": d, "verts": v_ref, "joints": j_ref}, os.path.join(GOLDEN, "lbs_small.pt"))


if __name__ == "__main__":
    main()
