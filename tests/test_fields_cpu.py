"""Network modules against the UNMODIFIED reference modules (oracle/_ref = models/{embedder,fields,renderer}.py staged byte
for byte by oracle/make_ref.py; skipped where the staged copy is absent): same parameters under the same torch seed (the
geometric initialisation of models/fields.py:40-63 and nn.Linear's default draws, in the same order), same state-dict keys
in the same order, checkpoints interchangeable in both directions -- for the train_clip networks (extra_color = True) and
for the network of confs/base_models/astrongman.conf (--mode train, no extra head)."""
import pytest
import torch

from oracle import make_ref

pytestmark = pytest.mark.skipif(not make_ref.available(), reason="oracle/_ref not staged (python -m oracle.make_ref)")

SDF_S = dict(d_in=3, d_out=257, d_hidden=256, n_layers=4, skip_in=[4], multires=6, bias=0.5, scale=1.0,
             geometric_init=True, weight_norm=True)
SDF_B2 = dict(SDF_S, n_layers=8)
SDF_SMALL = dict(SDF_S, d_out=129, d_hidden=128, n_layers=3, skip_in=[3])
COL = dict(d_feature=256, mode="no_view_dir", d_in=6, d_out=3, d_hidden=256, n_layers=2, weight_norm=True, multires_view=0,
           squeeze_out=True)


def _pair(ref_cls, our_cls, kw, seed=0):
    torch.manual_seed(seed)
    ref = ref_cls(**kw)
    after_ref = torch.rand(1)
    torch.manual_seed(seed)
    ours = our_cls(**kw)
    after_ours = torch.rand(1)
    assert torch.equal(after_ref, after_ours), "the constructor must consume torch's generator exactly like the reference's"
    return ref, ours


def _same_state(ref, ours):
    a, b = ref.state_dict(), ours.state_dict()
    assert list(a.keys()) == list(b.keys())
    for k in a:
        assert a[k].shape == b[k].shape and torch.equal(a[k], b[k]), k
    assert [n for n, _ in ref.named_parameters()] == [n for n, _ in ours.named_parameters()]      # torch.optim order
    ours.load_state_dict(a)           # strict, both directions
    ref.load_state_dict(b)


@pytest.mark.parametrize("kw", [SDF_S, SDF_B2, SDF_SMALL], ids=["shipped", "b2", "small"])
def test_sdf_network_init_is_bit_identical(kw):
    import avatarclip_b200 as ab
    fields, _ = make_ref.load_reference_models()
    _same_state(*_pair(fields.SDFNetwork, ab.SDFNetwork, kw))


@pytest.mark.parametrize("extra_color", [True, False])
def test_rendering_network_init_is_bit_identical(extra_color):
    import avatarclip_b200 as ab
    fields, _ = make_ref.load_reference_models()
    ref, ours = _pair(fields.RenderingNetwork, ab.RenderingNetwork, dict(COL, extra_color=extra_color), seed=3)
    _same_state(ref, ours)
    assert hasattr(ours, "extra_lin") == extra_color == hasattr(ref, "extra_lin")


def test_variance_network_and_renderer_constructor():
    import avatarclip_b200 as ab
    fields, _ = make_ref.load_reference_models()
    _same_state(fields.SingleVarianceNetwork(0.3), ab.SingleVarianceNetwork(0.3))
    sdf, col, var = ab.SDFNetwork(**SDF_SMALL), ab.RenderingNetwork(**dict(COL, d_feature=128, d_hidden=128)), \
        ab.SingleVarianceNetwork(0.3)
    ren = ab.NeuSRenderer(None, sdf, var, col, n_samples=32, n_importance=32, n_outside=0, up_sample_steps=4, perturb=1.0)
    assert ren.extra_color is False          # the astrongman.conf renderer block: no extra_color key (renderer.py:83)
    with pytest.raises(ValueError):          # 6-channel colour net under a 3-channel renderer (renderer.py:227-232)
        ab.NeuSRenderer(None, sdf, var, ab.RenderingNetwork(**dict(COL, d_feature=128, d_hidden=128, extra_color=True)),
                        32, 32, 0, 4, 1.0)
    with pytest.raises(NotImplementedError):
        ab.NeuSRenderer(None, sdf, var, col, 32, 32, 8, 4, 1.0)       # n_outside > 0: NeRF background, dead in every conf
