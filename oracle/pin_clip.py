"""Cross-check the restated CLIP ViT-B/32 image tower (oracle/clip_vit.py) against the independent
HuggingFace ``transformers`` implementation of the same published architecture, on seeded random
weights (the real ViT-B-32.pt is not on disk and there is no network).  This validates the
restatement of the ARCHITECTURE; it does not pin parity with openai/CLIP's own code, which is
absent -- the CLIP row stays "parity unpinned" (see oracle/__init__.py).

    python -m oracle.pin_clip
"""
import torch


def main():
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    from oracle import clip_vit as cv

    conf = cv.ViTConf()
    sd = cv.random_vit_state(conf, seed=0)
    hf = CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=768, intermediate_size=3072, num_hidden_layers=12, num_attention_heads=12, image_size=224,
        patch_size=32, hidden_act="quick_gelu", projection_dim=512, layer_norm_eps=1e-5)).eval()
    m = hf.state_dict()
    W = conf.width

    def put(k, v):
        assert m[k].shape == v.shape, (k, m[k].shape, v.shape)
        m[k] = v.clone()

    put("vision_model.embeddings.class_embedding", sd["class_embedding"])
    put("vision_model.embeddings.patch_embedding.weight", sd["conv1.weight"])
    put("vision_model.embeddings.position_embedding.weight", sd["positional_embedding"])
    for a, b in (("pre_layrnorm", "ln_pre"), ("post_layernorm", "ln_post")):
        put(f"vision_model.{a}.weight", sd[f"{b}.weight"]); put(f"vision_model.{a}.bias", sd[f"{b}.bias"])
    for i in range(conf.layers):
        p, q = f"transformer.resblocks.{i}.", f"vision_model.encoder.layers.{i}."
        for a, b in (("layer_norm1", "ln_1"), ("layer_norm2", "ln_2")):
            put(q + a + ".weight", sd[p + b + ".weight"]); put(q + a + ".bias", sd[p + b + ".bias"])
        wi, bi = sd[p + "attn.in_proj_weight"], sd[p + "attn.in_proj_bias"]
        for j, n in enumerate(("q_proj", "k_proj", "v_proj")):
            put(q + f"self_attn.{n}.weight", wi[j * W:(j + 1) * W]); put(q + f"self_attn.{n}.bias", bi[j * W:(j + 1) * W])
        put(q + "self_attn.out_proj.weight", sd[p + "attn.out_proj.weight"])
        put(q + "self_attn.out_proj.bias", sd[p + "attn.out_proj.bias"])
        put(q + "mlp.fc1.weight", sd[p + "mlp.c_fc.weight"]); put(q + "mlp.fc1.bias", sd[p + "mlp.c_fc.bias"])
        put(q + "mlp.fc2.weight", sd[p + "mlp.c_proj.weight"]); put(q + "mlp.fc2.bias", sd[p + "mlp.c_proj.bias"])
    put("visual_projection.weight", sd["proj"].t().contiguous())
    hf.load_state_dict(m)

    g = torch.Generator().manual_seed(1)
    img = torch.randn(2, 3, 224, 224, generator=g, requires_grad=True)
    a = cv.encode_image(sd, img, conf)
    b = hf(pixel_values=img).image_embeds
    err = (a - b).abs().max().item() / b.abs().max().item()
    ga, = torch.autograd.grad(a.square().sum(), img)
    gb, = torch.autograd.grad(b.square().sum(), img)
    gerr = (ga - gb).abs().max().item() / gb.abs().max().item()
    print(f"[pin_clip] restated tower vs transformers CLIPVisionModelWithProjection: rel-to-max err "
          f"embedding {err:.2e}, input-gradient {gerr:.2e}")
    assert err < 1e-4 and gerr < 1e-3


if __name__ == "__main__":
    main()
