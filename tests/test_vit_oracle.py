"""Cross-check of the ViT oracle (oracle/vit_ref.py, the torch-fp32 restatement the GPU engine is tested against) with the
`transformers` implementations of the same architectures, seeded random weights at test size (CPU):
CLIPVisionModel (north-star CLIP ViT-L/14 switches: pre-LN, QuickGELU, no patch bias) and Dinov2WithRegistersModel
(reference-faithful DINOv2-reg switches: 4 register tokens, LayerScale, GELU(erf), patch bias).
The ViT arithmetic is third-party code absent from /root/reference (torch.hub dinov2, unpinned: memory_2.py:43), so this
cross-check is what stands behind "parity unpinned" in DESIGN.md."""
import importlib

import pytest
import torch

from oracle import vit_ref

vc = importlib.import_module("bsc-nav_b200.vit_config")


def _copy_layers(cfg, w, get):
    """get(i) -> dict of torch Parameters for layer i keyed like ours."""
    D = cfg.dim
    with torch.no_grad():
        for i in range(cfg.depth):
            p, tgt = f"l{i}.", get(i)
            tgt["ln1_w"].copy_(w[p + "ln1_w"]); tgt["ln1_b"].copy_(w[p + "ln1_b"])
            tgt["ln2_w"].copy_(w[p + "ln2_w"]); tgt["ln2_b"].copy_(w[p + "ln2_b"])
            for j, n in enumerate(("q", "k", "v")):
                tgt[n + "_w"].copy_(w[p + "qkv_w"][j * D:(j + 1) * D]); tgt[n + "_b"].copy_(w[p + "qkv_b"][j * D:(j + 1) * D])
            tgt["proj_w"].copy_(w[p + "proj_w"]); tgt["proj_b"].copy_(w[p + "proj_b"])
            tgt["fc1_w"].copy_(w[p + "fc1_w"]); tgt["fc1_b"].copy_(w[p + "fc1_b"])
            tgt["fc2_w"].copy_(w[p + "fc2_w"]); tgt["fc2_b"].copy_(w[p + "fc2_b"])
            if cfg.layerscale:
                tgt["ls1"].copy_(w[p + "ls1"]); tgt["ls2"].copy_(w[p + "ls2"])


def test_clip_oracle_matches_transformers():
    from transformers import CLIPVisionConfig, CLIPVisionModel
    cfg = vc.tiny(vc.CLIP_L14, dim=64, depth=2, heads=1, mlp=128)
    w = vc.random_weights(cfg, seed=4, perturb=True)
    hf = CLIPVisionModel(CLIPVisionConfig(hidden_size=cfg.dim, intermediate_size=cfg.mlp, num_hidden_layers=cfg.depth,
                                          num_attention_heads=cfg.heads, image_size=cfg.image, patch_size=cfg.patch,
                                          hidden_act="quick_gelu", layer_norm_eps=cfg.ln_eps)).eval()
    vm = hf.vision_model
    with torch.no_grad():
        vm.embeddings.patch_embedding.weight.copy_(w["patch_w"].view(cfg.dim, 3, cfg.patch, cfg.patch))
        vm.embeddings.class_embedding.copy_(w["cls"])
        vm.embeddings.position_embedding.weight.copy_(w["pos"])
        vm.pre_layrnorm.weight.copy_(w["ln_pre_w"]); vm.pre_layrnorm.bias.copy_(w["ln_pre_b"])
        vm.post_layernorm.weight.copy_(w["ln_post_w"]); vm.post_layernorm.bias.copy_(w["ln_post_b"])

    def get(i):
        L = vm.encoder.layers[i]
        a = L.self_attn
        return {"ln1_w": L.layer_norm1.weight, "ln1_b": L.layer_norm1.bias, "ln2_w": L.layer_norm2.weight, "ln2_b": L.layer_norm2.bias,
                "q_w": a.q_proj.weight, "q_b": a.q_proj.bias, "k_w": a.k_proj.weight, "k_b": a.k_proj.bias, "v_w": a.v_proj.weight,
                "v_b": a.v_proj.bias, "proj_w": a.out_proj.weight, "proj_b": a.out_proj.bias, "fc1_w": L.mlp.fc1.weight,
                "fc1_b": L.mlp.fc1.bias, "fc2_w": L.mlp.fc2.weight, "fc2_b": L.mlp.fc2.bias}
    _copy_layers(cfg, w, get)
    x = torch.randn(2, 3, cfg.image, cfg.image, generator=torch.Generator().manual_seed(1))
    with torch.no_grad():
        # the memory consumes the final-LayerNormed patch tokens (x_norm_patchtokens): post_layernorm over all tokens, drop cls
        ref = vm.post_layernorm(hf(pixel_values=x).last_hidden_state)[:, 1:]
    got = vit_ref.vit_forward(x, w, cfg)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def test_clip_image_embedding_head_dim_80_matches_transformers():
    """The agent's view-check encoder (BSCAgent.py:309, open_clip ViT-H/14: heads of dimension 80, GELU(erf), bias-free
    projection of the post-LayerNorm'd cls token) at test size against transformers.CLIPVisionModelWithProjection."""
    from transformers import CLIPVisionConfig, CLIPVisionModelWithProjection
    cfg = vc.tiny(vc.OPENCLIP_H14, dim=160, depth=2, heads=2, mlp=256)
    cfg = vc.replace(cfg, embed=48)
    assert cfg.head_dim == 80
    w = vc.random_weights(cfg, seed=6, perturb=True)
    hf = CLIPVisionModelWithProjection(CLIPVisionConfig(
        hidden_size=cfg.dim, intermediate_size=cfg.mlp, num_hidden_layers=cfg.depth, num_attention_heads=cfg.heads,
        image_size=cfg.image, patch_size=cfg.patch, hidden_act="gelu", layer_norm_eps=cfg.ln_eps, projection_dim=cfg.embed)).eval()
    vm = hf.vision_model
    with torch.no_grad():
        vm.embeddings.patch_embedding.weight.copy_(w["patch_w"].view(cfg.dim, 3, cfg.patch, cfg.patch))
        vm.embeddings.class_embedding.copy_(w["cls"])
        vm.embeddings.position_embedding.weight.copy_(w["pos"])
        vm.pre_layrnorm.weight.copy_(w["ln_pre_w"]); vm.pre_layrnorm.bias.copy_(w["ln_pre_b"])
        vm.post_layernorm.weight.copy_(w["ln_post_w"]); vm.post_layernorm.bias.copy_(w["ln_post_b"])
        hf.visual_projection.weight.copy_(w["proj"].T)

    def get(i):
        L = vm.encoder.layers[i]
        a = L.self_attn
        return {"ln1_w": L.layer_norm1.weight, "ln1_b": L.layer_norm1.bias, "ln2_w": L.layer_norm2.weight, "ln2_b": L.layer_norm2.bias,
                "q_w": a.q_proj.weight, "q_b": a.q_proj.bias, "k_w": a.k_proj.weight, "k_b": a.k_proj.bias, "v_w": a.v_proj.weight,
                "v_b": a.v_proj.bias, "proj_w": a.out_proj.weight, "proj_b": a.out_proj.bias, "fc1_w": L.mlp.fc1.weight,
                "fc1_b": L.mlp.fc1.bias, "fc2_w": L.mlp.fc2.weight, "fc2_b": L.mlp.fc2.bias}
    _copy_layers(cfg, w, get)
    x = torch.randn(3, 3, cfg.image, cfg.image, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = hf(pixel_values=x).image_embeds
    got = vit_ref.encode_image(x, w, cfg)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)
    # check_around's scoring on those embeddings (BSCAgent.py:713-718): literal torch restatement
    prob, best = vit_ref.score_views(got, got[1])
    assert int(best) == 1 and abs(float(prob.sum()) - 1.0) < 1e-6


def test_clip_text_tower_matches_transformers():
    """encode_text of the agent's CLIP (BSCAgent.py:545-546, 711): causal transformer, end-of-text pooling (argmax of the ids),
    bias-free projection -- against transformers.CLIPTextModelWithProjection at test size."""
    from transformers import CLIPTextConfig, CLIPTextModelWithProjection
    cfg = vc.replace(vc.tiny(vc.OPENCLIP_H14_TEXT, dim=64, depth=2, heads=1, mlp=128), embed=40, text_context=20, vocab=300)
    w = vc.random_weights(cfg, seed=7, perturb=True)
    # eos_token_id=2 selects transformers' argmax pooling (what open_clip does: text.argmax(dim=-1))
    hf = CLIPTextModelWithProjection(CLIPTextConfig(
        vocab_size=cfg.vocab, hidden_size=cfg.dim, intermediate_size=cfg.mlp, projection_dim=cfg.embed, num_hidden_layers=cfg.depth,
        num_attention_heads=cfg.heads, max_position_embeddings=cfg.text_context, hidden_act="gelu", layer_norm_eps=cfg.ln_eps,
        eos_token_id=2, bos_token_id=0, pad_token_id=1)).eval()
    tm = hf.text_model
    with torch.no_grad():
        tm.embeddings.token_embedding.weight.copy_(w["tok_emb"])
        tm.embeddings.position_embedding.weight.copy_(w["pos"])
        tm.final_layer_norm.weight.copy_(w["ln_post_w"]); tm.final_layer_norm.bias.copy_(w["ln_post_b"])
        hf.text_projection.weight.copy_(w["proj"].T)

    def get(i):
        L = tm.encoder.layers[i]
        a = L.self_attn
        return {"ln1_w": L.layer_norm1.weight, "ln1_b": L.layer_norm1.bias, "ln2_w": L.layer_norm2.weight, "ln2_b": L.layer_norm2.bias,
                "q_w": a.q_proj.weight, "q_b": a.q_proj.bias, "k_w": a.k_proj.weight, "k_b": a.k_proj.bias, "v_w": a.v_proj.weight,
                "v_b": a.v_proj.bias, "proj_w": a.out_proj.weight, "proj_b": a.out_proj.bias, "fc1_w": L.mlp.fc1.weight,
                "fc1_b": L.mlp.fc1.bias, "fc2_w": L.mlp.fc2.weight, "fc2_b": L.mlp.fc2.bias}
    _copy_layers(cfg, w, get)
    ids = _token_rows(cfg, [5, 11, 19, 1], seed=3)
    with torch.no_grad():
        ref = hf(input_ids=ids).text_embeds
    got = vit_ref.encode_text(ids, w, cfg)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)


def _token_rows(cfg, lengths, seed):
    """Token rows the way open_clip.tokenize lays them out: <start> words... <end> then zero padding; <end> = vocab - 1 is the
    largest id, so argmax finds it."""
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(len(lengths), cfg.text_context, dtype=torch.long)
    for r, n in enumerate(lengths):
        n = min(n, cfg.text_context - 2)
        ids[r, 0] = cfg.vocab - 2
        ids[r, 1:1 + n] = torch.randint(1, cfg.vocab - 2, (n,), generator=g)
        ids[r, 1 + n] = cfg.vocab - 1
    return ids


def test_dinov2_reg_oracle_matches_transformers():
    try:
        from transformers import Dinov2WithRegistersConfig, Dinov2WithRegistersModel
    except ImportError:
        pytest.skip("transformers without Dinov2WithRegisters")
    cfg = vc.tiny(vc.DINOV2_L14_REG, dim=64, depth=2, heads=1, mlp=128)
    w = vc.random_weights(cfg, seed=5, perturb=True)
    hf = Dinov2WithRegistersModel(Dinov2WithRegistersConfig(
        hidden_size=cfg.dim, num_hidden_layers=cfg.depth, num_attention_heads=cfg.heads, mlp_ratio=cfg.mlp // cfg.dim,
        image_size=cfg.image, patch_size=cfg.patch, num_register_tokens=cfg.n_reg, layer_norm_eps=cfg.ln_eps, hidden_act="gelu",
        layerscale_value=1.0, use_swiglu_ffn=False)).eval()
    emb = hf.embeddings
    with torch.no_grad():
        emb.patch_embeddings.projection.weight.copy_(w["patch_w"].view(cfg.dim, 3, cfg.patch, cfg.patch))
        emb.patch_embeddings.projection.bias.copy_(w["patch_b"])
        emb.cls_token.copy_(w["cls"].view(1, 1, -1))
        emb.register_tokens.copy_(w["reg"].view(1, cfg.n_reg, -1))
        # position table stored at the working resolution (image / patch)^2: no interpolation takes place
        assert emb.position_embeddings.shape == (1, 1 + cfg.n_patch, cfg.dim)
        emb.position_embeddings.copy_(w["pos"].view(1, 1 + cfg.n_patch, -1))
        hf.layernorm.weight.copy_(w["ln_post_w"]); hf.layernorm.bias.copy_(w["ln_post_b"])

    def get(i):
        L = hf.encoder.layer[i]
        a = L.attention
        return {"ln1_w": L.norm1.weight, "ln1_b": L.norm1.bias, "ln2_w": L.norm2.weight, "ln2_b": L.norm2.bias,
                "q_w": a.attention.query.weight, "q_b": a.attention.query.bias, "k_w": a.attention.key.weight,
                "k_b": a.attention.key.bias, "v_w": a.attention.value.weight, "v_b": a.attention.value.bias,
                "proj_w": a.output.dense.weight, "proj_b": a.output.dense.bias, "fc1_w": L.mlp.fc1.weight, "fc1_b": L.mlp.fc1.bias,
                "fc2_w": L.mlp.fc2.weight, "fc2_b": L.mlp.fc2.bias, "ls1": L.layer_scale1.lambda1, "ls2": L.layer_scale2.lambda1}
    _copy_layers(cfg, w, get)
    x = torch.randn(2, 3, cfg.image, cfg.image, generator=torch.Generator().manual_seed(2))
    with torch.no_grad():
        ref = hf(pixel_values=x).last_hidden_state[:, 1 + cfg.n_reg:]
    got = vit_ref.vit_forward(x, w, cfg)
    torch.testing.assert_close(got, ref, rtol=2e-4, atol=2e-4)
