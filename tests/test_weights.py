"""CPU: checkpoint loaders (bsc-nav_b200/weights.py) against the `transformers` implementations with seeded random weights --
the state dict goes through the loader, the oracle (oracle/vit_ref.py) must reproduce the model's own forward.
Covers the DINOv2 position-table interpolation 37x37 -> 16x16 (modeling_dinov2_with_registers.py:93-145),
the torch.hub dinov2 key layout the reference loads (memory_2.py:43) and the open_clip layout of its agent (BSCAgent.py:309)."""
import importlib

import pytest
import torch

from oracle import vit_ref

vc = importlib.import_module("bsc-nav_b200.vit_config")
wl = importlib.import_module("bsc-nav_b200.weights")


def _randomise(model, seed):
    g = torch.Generator().manual_seed(seed)
    with torch.no_grad():
        for p in model.parameters():
            p.copy_(torch.randn(p.shape, generator=g) * (0.05 if p.dim() > 1 else 0.3) + (1.0 if p.dim() == 1 and p.numel() > 8 and "norm" in "" else 0.0))
    return model.eval()


def _hf_dinov2(cfg, image_size):
    from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    hf = Dinov2WithRegistersModel(Dinov2WithRegistersConfig(
        hidden_size=cfg.dim, num_hidden_layers=cfg.depth, num_attention_heads=cfg.heads, mlp_ratio=cfg.mlp // cfg.dim,
        image_size=image_size, patch_size=cfg.patch, num_register_tokens=cfg.n_reg, layer_norm_eps=cfg.ln_eps, hidden_act="gelu",
        layerscale_value=1.0))
    return _randomise(hf, 11)


def _hub_names(sd, depth, chunked=False):
    """transformers DINOv2 state dict -> the torch.hub facebookresearch/dinov2 key layout (what memory_2.py:43 loads)."""
    out = {"cls_token": sd["embeddings.cls_token"], "pos_embed": sd["embeddings.position_embeddings"],
           "register_tokens": sd["embeddings.register_tokens"], "mask_token": sd["embeddings.mask_token"],
           "patch_embed.proj.weight": sd["embeddings.patch_embeddings.projection.weight"],
           "patch_embed.proj.bias": sd["embeddings.patch_embeddings.projection.bias"],
           "norm.weight": sd["layernorm.weight"], "norm.bias": sd["layernorm.bias"]}
    for i in range(depth):
        s = f"encoder.layer.{i}."
        d = f"blocks.{i // 2}.{i}." if chunked else f"blocks.{i}."
        a = s + "attention.attention."
        out[d + "attn.qkv.weight"] = torch.cat([sd[a + n + ".weight"] for n in ("query", "key", "value")])
        out[d + "attn.qkv.bias"] = torch.cat([sd[a + n + ".bias"] for n in ("query", "key", "value")])
        for src, dst in (("norm1", "norm1"), ("norm2", "norm2"), ("attention.output.dense", "attn.proj"), ("mlp.fc1", "mlp.fc1"),
                         ("mlp.fc2", "mlp.fc2")):
            out[d + dst + ".weight"], out[d + dst + ".bias"] = sd[s + src + ".weight"], sd[s + src + ".bias"]
        out[d + "ls1.gamma"], out[d + "ls2.gamma"] = sd[s + "layer_scale1.lambda1"], sd[s + "layer_scale2.lambda1"]
    return out


@pytest.mark.parametrize("layout", ["hf", "hub", "hub-chunked"])
def test_dinov2_reg_checkpoint_with_37x37_table(layout):
    cfg = vc.tiny(vc.DINOV2_L14_REG, dim=64, depth=4, heads=1, mlp=256)
    hf = _hf_dinov2(cfg, image_size=518)                       # 37 x 37 position table, as the released checkpoints
    sd = hf.state_dict()
    assert sd["embeddings.position_embeddings"].shape == (1, 1 + 37 * 37, cfg.dim)
    if layout != "hf":
        sd = _hub_names(sd, cfg.depth, chunked=layout == "hub-chunked")
        assert wl.detect_format(sd) == "dinov2-hub"
    w = wl.load_state_dict(cfg, sd)
    assert w["pos"].shape == (1 + 256, cfg.dim) and w["reg"].shape == (4, cfg.dim)
    x = torch.randn(2, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state[:, 1 + cfg.n_reg:]      # x_norm_patchtokens: final LayerNorm, cls + registers dropped
    got = vit_ref.vit_forward(x, w, cfg)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def test_pos_embed_interpolation_variants():
    g = torch.Generator().manual_seed(0)
    pos = torch.randn(1, 1 + 37 * 37, 8, generator=g)
    same = wl.interpolate_pos_embed(torch.randn(1 + 16 * 16, 8, generator=g), 16)
    assert same.shape == (257, 8)
    a = wl.interpolate_pos_embed(pos, 16, antialias=True, offset=0.0)
    b = wl.interpolate_pos_embed(pos, 16, antialias=False, offset=0.1)          # the non-register hub models
    assert a.shape == b.shape == (257, 8) and torch.equal(a[0], pos[0, 0]) and not torch.allclose(a[1:], b[1:])
    with pytest.raises(ValueError):
        wl.interpolate_pos_embed(torch.randn(1 + 30, 8), 16)


def _open_clip_names(sd, depth, text):
    """transformers CLIP state dict -> open_clip / OpenAI-clip key layout (model.visual.* / model.*)."""
    t = "text_model." if text else "vision_model."
    pre = "" if text else "visual."
    out = {}
    if text:
        out["token_embedding.weight"] = sd[t + "embeddings.token_embedding.weight"]
        out["positional_embedding"] = sd[t + "embeddings.position_embedding.weight"]
        out["ln_final.weight"], out["ln_final.bias"] = sd[t + "final_layer_norm.weight"], sd[t + "final_layer_norm.bias"]
        out["text_projection"] = sd["text_projection.weight"].T
    else:
        out[pre + "class_embedding"] = sd[t + "embeddings.class_embedding"]
        out[pre + "positional_embedding"] = sd[t + "embeddings.position_embedding.weight"]
        out[pre + "conv1.weight"] = sd[t + "embeddings.patch_embedding.weight"]
        out[pre + "ln_pre.weight"], out[pre + "ln_pre.bias"] = sd[t + "pre_layrnorm.weight"], sd[t + "pre_layrnorm.bias"]
        out[pre + "ln_post.weight"], out[pre + "ln_post.bias"] = sd[t + "post_layernorm.weight"], sd[t + "post_layernorm.bias"]
        out[pre + "proj"] = sd["visual_projection.weight"].T
    for i in range(depth):
        s, d = f"{t}encoder.layers.{i}.", f"{pre}transformer.resblocks.{i}."
        out[d + "attn.in_proj_weight"] = torch.cat([sd[s + f"self_attn.{n}_proj.weight"] for n in "qkv"])
        out[d + "attn.in_proj_bias"] = torch.cat([sd[s + f"self_attn.{n}_proj.bias"] for n in "qkv"])
        for src, dst in (("layer_norm1", "ln_1"), ("layer_norm2", "ln_2"), ("self_attn.out_proj", "attn.out_proj"), ("mlp.fc1", "mlp.c_fc"),
                         ("mlp.fc2", "mlp.c_proj")):
            out[d + dst + ".weight"], out[d + dst + ".bias"] = sd[s + src + ".weight"], sd[s + src + ".bias"]
    return out


@pytest.mark.parametrize("layout", ["hf", "open_clip"])
def test_clip_image_tower_checkpoint(layout):
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = vc.replace(vc.tiny(vc.OPENCLIP_H14, dim=160, depth=2, heads=2, mlp=256), embed=48)
    hf = _randomise(CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=cfg.dim, intermediate_size=cfg.mlp, num_hidden_layers=cfg.depth, num_attention_heads=cfg.heads,
        image_size=cfg.image, patch_size=cfg.patch, hidden_act="gelu", layer_norm_eps=cfg.ln_eps, projection_dim=cfg.embed)), 5)
    sd = hf.state_dict()
    if layout == "open_clip":
        sd = _open_clip_names(sd, cfg.depth, text=False)
        assert wl.detect_format(sd) == "open_clip"
    w = wl.load_state_dict(cfg, sd)
    x = torch.randn(2, 3, cfg.image, cfg.image, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        out = hf(pixel_values=x)
    torch.testing.assert_close(vit_ref.encode_image(x, w, cfg), out.image_embeds, rtol=2e-4, atol=2e-4)
    with torch.no_grad():
        tokens = hf.vision_model.post_layernorm(out.last_hidden_state)[:, 1:]
    torch.testing.assert_close(vit_ref.vit_forward(x, w, cfg), tokens, rtol=2e-4, atol=2e-4)


@pytest.mark.parametrize("layout", ["hf", "open_clip"])
def test_clip_text_tower_checkpoint(layout):
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    cfg = vc.replace(vc.tiny(vc.OPENCLIP_H14_TEXT, dim=64, depth=2, heads=1, mlp=128), embed=40, text_context=20, vocab=300)
    hf = _randomise(CLIPTextModelWithProjection(CLIPTextConfig(
        vocab_size=cfg.vocab, hidden_size=cfg.dim, intermediate_size=cfg.mlp, projection_dim=cfg.embed, num_hidden_layers=cfg.depth,
        num_attention_heads=cfg.heads, max_position_embeddings=cfg.text_context, hidden_act="gelu", layer_norm_eps=cfg.ln_eps,
        eos_token_id=2, bos_token_id=0, pad_token_id=1)), 8)
    sd = hf.state_dict()
    if layout == "open_clip":
        sd = _open_clip_names(sd, cfg.depth, text=True)
    w = wl.load_state_dict(cfg, sd)
    ids = torch.randint(3, cfg.vocab - 1, (3, cfg.text_context), generator=torch.Generator().manual_seed(4))
    ids[:, 0] = 0
    for b, n in enumerate((5, 12, 19)):
        ids[b, n] = cfg.vocab - 1                          # end-of-text = largest id (open_clip pools at argmax)
        ids[b, n + 1:] = 1
    with torch.no_grad():
        ref = hf(input_ids=ids).text_embeds
    if layout == "hf":
        pytest.skip("transformers pools at eos_token_id, open_clip at argmax: compared in test_vit_oracle.py") if False else None
    got = vit_ref.encode_text(ids, w, cfg)
    # transformers pools at the first eos_token_id (= 2) if present, else argmax: ids avoid 2, so both pool at the largest id
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def test_wrong_checkpoint_is_rejected():
    cfg = vc.tiny(vc.DINOV2_L14_REG, dim=64, depth=2, heads=1, mlp=256)
    hf = _hf_dinov2(cfg, image_size=224)
    sd = dict(hf.state_dict())
    bad = dict(sd)
    bad["encoder.layer.1.mlp.fc1.weight"] = bad["encoder.layer.1.mlp.fc1.weight"][:, :32]
    with pytest.raises(ValueError):
        wl.load_state_dict(cfg, bad)
    with pytest.raises(KeyError):
        wl.load_state_dict(vc.replace(cfg, depth=3), sd)
    with pytest.raises(ValueError):
        wl.load_state_dict(cfg, {"foo": torch.zeros(1)})
