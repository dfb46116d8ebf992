"""world_size-2 gloo test of the view-sharded data-parallel host logic (no GPU): two ranks each take the view
avatarclip_b200.dist.view_index assigns them, all-reduce their flat gradients, and must end up with exactly the
1-process 2-view gradient-accumulation result."""
import os
import socket
import sys

import torch
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))


def _tiny_world():
    import util_neus as U
    from oracle import clip_vit as cv
    from oracle.train_step import OracleTrainer
    sdf_kw, col_kw, ren_kw, _ = U.CASES["tiny"]
    sconf, cconf, rconf = U.confs_from_kw(sdf_kw, col_kw, ren_kw)
    sp, cp = U.synth_state(sdf_kw, col_kw, 0)
    vconf = cv.ViTConf(image_size=64, patch=32, width=64, layers=2, heads=1, mlp=128, out_dim=32)
    clip_sd = cv.random_vit_state(vconf, seed=0)
    text = torch.randn(2, 32, generator=torch.Generator().manual_seed(1))
    return OracleTrainer(sconf, cconf, rconf, sp, cp, 0.3, clip_sd, text, clip_conf=vconf)     # a tiny CLIP tower


def _flat_grad(orc, view):
    total, _ = orc.loss(view)
    gs = torch.autograd.grad(total, [p for _, p in orc.named_params()], allow_unused=True)
    return torch.cat([(torch.zeros_like(p) if g is None else g).reshape(-1) for (_, p), g in zip(orc.named_params(), gs)])


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    torch.set_num_threads(2)
    from avatarclip_b200 import dist as ad
    from avatarclip_b200.workload import make_view
    pg = ad.init_from_env("gloo")
    orc = _tiny_world()
    v = make_view(ad.view_index(0, rank, world), n_rays=40, H=48, W=48, seed=0)
    g = _flat_grad(orc, v)
    ad.allreduce_sum_(g, pg)
    g = g * ad.grad_scale(pg)
    if rank == 0:
        torch.save(g, out)
    torch.distributed.destroy_process_group()


def test_two_rank_view_sharding_equals_gradient_accumulation(tmp_path):
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    out = str(tmp_path / "g.pt")
    mp.spawn(_worker, args=(2, port, out), nprocs=2, join=True)
    got = torch.load(out)
    from avatarclip_b200 import dist as ad
    from avatarclip_b200.workload import make_view
    orc = _tiny_world()
    want = sum(_flat_grad(orc, make_view(ad.view_index(0, r, 2), n_rays=40, H=48, W=48, seed=0)) for r in range(2)) / 2
    assert torch.allclose(got, want, rtol=1e-5, atol=1e-7)
    assert ad.view_index(3, 1, 4) == 13


def test_host_view_packs_into_one_buffer():
    """HostView.pack(): every tensor becomes a view of one contiguous buffer (a step's inputs are a single H2D copy) and
    the layout is a pure function of the shapes, so views of equal shape can refresh the same device buffer."""
    from avatarclip_b200.workload import VIEW_FIELDS, make_view
    a = make_view(0, n_rays=40, H=16, W=16, seed=3, bg_choice=1)
    b = make_view(5, n_rays=40, H=16, W=16, seed=3, bg_choice=2)
    assert a.flat is not None and a.flat.dtype == torch.uint8
    la, ta = a.layout()
    lb, tb = b.layout()
    assert la == lb and ta == tb == a.flat.numel()
    base = a.flat.data_ptr()
    for name, off, shape, dtype, nb in la:
        t = getattr(a, name)
        assert t.data_ptr() == base + off and off % 256 == 0 and tuple(t.shape) == shape
    assert a.h2d_bytes() <= ta
    assert torch.allclose(a.scalars[:3], torch.as_tensor(a.light_dir, dtype=torch.float32))
    assert abs(float(a.scalars[3]) - a.ambience) < 1e-7
    c = make_view(0, n_rays=40, H=16, W=16, seed=3, bg_choice=3)          # no background tensors: another layout
    assert c.layout()[0] != la
    assert set(n for n, *_ in c.layout()[0]) <= set(VIEW_FIELDS)
