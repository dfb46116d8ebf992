"""avc_gen_rays against the restated gen_rays_pose / near_far_from_sphere (oracle.loss.pinhole_rays, oracle.neus)."""
import numpy as np
import pytest
import torch

from oracle import loss as ol, neus

pytestmark = pytest.mark.gpu


@pytest.mark.parametrize("canvas", [256, 112, 97])
def test_gen_rays_matches_oracle(canvas):
    from avatarclip_b200.dataset import RayGenerator
    pose = ol.lookat([1.1, 0.3, -1.2], [0.05, -0.1, 0.0])
    o, d = ol.pinhole_rays(pose, canvas, canvas, full_res=256)
    near, far = neus.near_far_from_sphere(o.reshape(-1, 3), d.reshape(-1, 3))
    g = torch.Generator().manual_seed(canvas)
    mask = torch.rand(canvas, canvas, generator=g) < 0.3
    rg = RayGenerator()
    pix = torch.nonzero(mask.reshape(-1).cuda(), as_tuple=False).reshape(-1).to(torch.int32)
    ro, rd, nr, fr = rg.rays(pose, canvas, canvas, pix)
    sel = mask.reshape(-1)
    assert torch.equal(pix.cpu().long(), torch.nonzero(sel).reshape(-1))
    assert (rd.cpu() - d.reshape(-1, 3)[sel]).abs().max().item() < 2e-6
    assert (ro.cpu() - o.reshape(-1, 3)[sel]).abs().max().item() < 1e-6
    assert (nr.cpu() - near[sel]).abs().max().item() < 5e-6
    assert (fr.cpu() - far[sel]).abs().max().item() < 5e-6
