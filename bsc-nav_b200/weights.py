"""Checkpoint -> engine weights: maps the state dicts of the feature extractors the reference loads onto the weight dict
`vit.pack_weights` / `ViTEngine` consume (layout of vit_config.random_weights).

The reference obtains its models from
  * torch.hub 'facebookresearch/dinov2', entry `dinov2_vitl14_reg`   (memory_2.py:43, create_memory_for_dataset.py:58)
  * open_clip 'ViT-H-14' pretrained 'metaclip_fullcc'                 (BSCAgent.py:309: image + text tower)
and BASELINE.json names CLIP ViT-L/14; the same architectures are also distributed in `transformers` naming.  Supported
key layouts (auto-detected by `load_state_dict`):

  dinov2-hub   cls_token, pos_embed, register_tokens, patch_embed.proj.*, blocks.N.{norm1,attn.qkv,attn.proj,ls1.gamma,norm2,
               mlp.fc1,mlp.fc2,ls2.gamma}, norm.*            (also the chunked form blocks.C.N.*)
  hf-dinov2    embeddings.{cls_token,position_embeddings,register_tokens,patch_embeddings.projection.*},
               encoder.layer.N.{norm1,attention.attention.{query,key,value},attention.output.dense,layer_scale1.lambda1,norm2,
               mlp.fc1,mlp.fc2,layer_scale2.lambda1}, layernorm.*
  open_clip    [visual.]{class_embedding,positional_embedding,proj,conv1.weight,ln_pre,transformer.resblocks.N.{ln_1,
               attn.in_proj_weight,attn.in_proj_bias,attn.out_proj,ln_2,mlp.c_fc,mlp.c_proj},ln_post}   (OpenAI clip: same names)
               text: token_embedding.weight, positional_embedding, transformer.resblocks.N.*, ln_final, text_projection
  hf-clip      vision_model.{embeddings.{class_embedding,patch_embedding.weight,position_embedding.weight},pre_layrnorm,
               encoder.layers.N.{layer_norm1,self_attn.{q,k,v,out}_proj,layer_norm2,mlp.fc1,mlp.fc2},post_layernorm},
               visual_projection.weight;  text_model.{embeddings.{token_embedding,position_embedding}.weight,encoder.layers.N.*,
               final_layer_norm}, text_projection.weight

DINOv2 stores its position table at the pre-training grid (37 x 37 for /14 at 518 px); at 224 px the model interpolates it to
16 x 16 once per forward (bicubic, antialias; transformers/models/dinov2_with_registers/modeling_dinov2_with_registers.py:93-145,
same arithmetic in the hub code for the register models: interpolate_antialias=True, interpolate_offset=0.0).  The engine keeps a
fixed grid, so the table is interpolated here, once, with the same torch call.
"""
from __future__ import annotations

import re

import torch
import torch.nn.functional as F

from .vit_config import ViTConfig


def interpolate_pos_embed(pos: torch.Tensor, grid: int, antialias: bool = True, offset: float = 0.0) -> torch.Tensor:
    """(1 + M*M, D) position table (cls first) -> (1 + grid*grid, D).

    antialias / offset follow the checkpoint's family: register models (the reference's dinov2_vitl14_reg) and transformers use
    antialias=True with an explicit output size; the original non-register hub models use antialias=False and a scale factor of
    (grid + 0.1) / M (`interpolate_offset=0.1`)."""
    pos = pos.reshape(-1, pos.shape[-1]).float()
    n = pos.shape[0] - 1
    m = int(round(n ** 0.5))
    if m * m != n:
        raise ValueError(f"position table of {n} patch positions is not square")
    if m == grid:
        return pos.clone()
    patch = pos[1:].reshape(1, m, m, -1).permute(0, 3, 1, 2)
    if offset:
        s = float(grid + offset) / m
        patch = F.interpolate(patch, scale_factor=(s, s), mode="bicubic", align_corners=False, antialias=antialias)
    else:
        patch = F.interpolate(patch, size=(grid, grid), mode="bicubic", align_corners=False, antialias=antialias)
    if patch.shape[-2:] != (grid, grid):
        raise ValueError("interpolated position table has the wrong size")
    return torch.cat([pos[:1], patch.permute(0, 2, 3, 1).reshape(grid * grid, -1)], dim=0)


def _f(t) -> torch.Tensor:
    return t.detach().to("cpu", torch.float32).contiguous()


def _strip(sd: dict, prefixes=("module.", "model.", "backbone.")) -> dict:
    out = {}
    for k, v in sd.items():
        for p in prefixes:
            if k.startswith(p):
                k = k[len(p):]
        out[k] = v
    return out


def detect_format(sd: dict) -> str:
    keys = set(sd)
    if any(k.startswith("vision_model.") or k.startswith("text_model.") for k in keys):
        return "hf-clip"
    if "embeddings.position_embeddings" in keys:
        return "hf-dinov2"
    if "pos_embed" in keys and "patch_embed.proj.weight" in keys:
        return "dinov2-hub"
    if any(k.endswith("transformer.resblocks.0.attn.in_proj_weight") for k in keys):
        return "open_clip"
    raise ValueError("unrecognised checkpoint layout (expected dinov2-hub, hf-dinov2, open_clip / OpenAI clip or hf-clip keys)")


def _check(cfg: ViTConfig, w: dict) -> dict:
    """Shape validation against the engine configuration: a wrong checkpoint fails here, not inside a kernel."""
    D, M = cfg.dim, cfg.mlp
    want = {"ln_post_w": (D,), "ln_post_b": (D,)}
    if cfg.text_context:
        want.update({"tok_emb": (cfg.vocab, D), "pos": (cfg.text_context, D)})
    else:
        want.update({"patch_w": (D, cfg.patch_k), "patch_b": (D,), "cls": (D,), "pos": (1 + cfg.n_patch, D), "reg": (cfg.n_reg, D)})
    if cfg.ln_pre:
        want.update({"ln_pre_w": (D,), "ln_pre_b": (D,)})
    if cfg.embed:
        want["proj"] = (D, cfg.embed)
    for i in range(cfg.depth):
        p = f"l{i}."
        want.update({p + "ln1_w": (D,), p + "ln1_b": (D,), p + "qkv_w": (3 * D, D), p + "qkv_b": (3 * D,), p + "proj_w": (D, D),
                     p + "proj_b": (D,), p + "ln2_w": (D,), p + "ln2_b": (D,), p + "fc1_w": (M, D), p + "fc1_b": (M,),
                     p + "fc2_w": (D, M), p + "fc2_b": (D,)})
        if cfg.layerscale:
            want.update({p + "ls1": (D,), p + "ls2": (D,)})
    for k, shape in want.items():
        if k not in w:
            raise KeyError(f"checkpoint is missing the tensor for '{k}' ({cfg.name})")
        if tuple(w[k].shape) != tuple(shape):
            raise ValueError(f"'{k}': checkpoint shape {tuple(w[k].shape)} != {tuple(shape)} expected by {cfg.name}")
        if not torch.isfinite(w[k]).all():
            raise ValueError(f"'{k}' holds non-finite values")
    return {k: w[k] for k in want}


def from_dinov2_hub(sd: dict, cfg: ViTConfig, antialias: bool | None = None, offset: float | None = None) -> dict:
    """torch.hub facebookresearch/dinov2 (memory_2.py:43): dinov2_vit{s,b,l,g}14[_reg]."""
    sd = _strip(sd)
    reg = cfg.n_reg > 0
    antialias = reg if antialias is None else antialias          # hub: *_reg models interpolate with antialias and offset 0
    offset = (0.0 if reg else 0.1) if offset is None else offset
    blocks = {}
    for k in sd:
        m = re.match(r"blocks\.(?:\d+\.)?(\d+)\.(.+)$", k)       # blocks.N.* or the chunked blocks.C.N.*
        if m:
            blocks.setdefault(int(m.group(1)), {})[m.group(2)] = sd[k]
    w = {"patch_w": _f(sd["patch_embed.proj.weight"]).reshape(cfg.dim, -1), "patch_b": _f(sd["patch_embed.proj.bias"]),
         "cls": _f(sd["cls_token"]).reshape(-1), "pos": interpolate_pos_embed(_f(sd["pos_embed"]), cfg.grid, antialias, offset),
         "reg": _f(sd["register_tokens"]).reshape(-1, cfg.dim) if reg else torch.zeros(0, cfg.dim),
         "ln_post_w": _f(sd["norm.weight"]), "ln_post_b": _f(sd["norm.bias"])}
    for i in range(cfg.depth):
        b, p = blocks[i], f"l{i}."
        w.update({p + "ln1_w": _f(b["norm1.weight"]), p + "ln1_b": _f(b["norm1.bias"]), p + "qkv_w": _f(b["attn.qkv.weight"]),
                  p + "qkv_b": _f(b["attn.qkv.bias"]), p + "proj_w": _f(b["attn.proj.weight"]), p + "proj_b": _f(b["attn.proj.bias"]),
                  p + "ln2_w": _f(b["norm2.weight"]), p + "ln2_b": _f(b["norm2.bias"]), p + "fc1_w": _f(b["mlp.fc1.weight"]),
                  p + "fc1_b": _f(b["mlp.fc1.bias"]), p + "fc2_w": _f(b["mlp.fc2.weight"]), p + "fc2_b": _f(b["mlp.fc2.bias"])})
        if cfg.layerscale:
            w[p + "ls1"], w[p + "ls2"] = _f(b["ls1.gamma"]), _f(b["ls2.gamma"])
    return _check(cfg, w)


def from_hf_dinov2(sd: dict, cfg: ViTConfig) -> dict:
    """transformers Dinov2WithRegistersModel / Dinov2Model."""
    sd = _strip(sd, prefixes=("module.", "dinov2_with_registers.", "dinov2."))
    e = "embeddings."
    w = {"patch_w": _f(sd[e + "patch_embeddings.projection.weight"]).reshape(cfg.dim, -1),
         "patch_b": _f(sd[e + "patch_embeddings.projection.bias"]), "cls": _f(sd[e + "cls_token"]).reshape(-1),
         "pos": interpolate_pos_embed(_f(sd[e + "position_embeddings"]), cfg.grid, antialias=True, offset=0.0),
         "reg": _f(sd[e + "register_tokens"]).reshape(-1, cfg.dim) if cfg.n_reg else torch.zeros(0, cfg.dim),
         "ln_post_w": _f(sd["layernorm.weight"]), "ln_post_b": _f(sd["layernorm.bias"])}
    for i in range(cfg.depth):
        s, p = f"encoder.layer.{i}.", f"l{i}."
        a = s + "attention.attention."
        w.update({p + "ln1_w": _f(sd[s + "norm1.weight"]), p + "ln1_b": _f(sd[s + "norm1.bias"]),
                  p + "qkv_w": torch.cat([_f(sd[a + n + ".weight"]) for n in ("query", "key", "value")]),
                  p + "qkv_b": torch.cat([_f(sd[a + n + ".bias"]) for n in ("query", "key", "value")]),
                  p + "proj_w": _f(sd[s + "attention.output.dense.weight"]), p + "proj_b": _f(sd[s + "attention.output.dense.bias"]),
                  p + "ln2_w": _f(sd[s + "norm2.weight"]), p + "ln2_b": _f(sd[s + "norm2.bias"]),
                  p + "fc1_w": _f(sd[s + "mlp.fc1.weight"]), p + "fc1_b": _f(sd[s + "mlp.fc1.bias"]),
                  p + "fc2_w": _f(sd[s + "mlp.fc2.weight"]), p + "fc2_b": _f(sd[s + "mlp.fc2.bias"])})
        if cfg.layerscale:
            w[p + "ls1"], w[p + "ls2"] = _f(sd[s + "layer_scale1.lambda1"]), _f(sd[s + "layer_scale2.lambda1"])
    return _check(cfg, w)


def _resblocks(sd: dict, prefix: str, cfg: ViTConfig, w: dict):
    for i in range(cfg.depth):
        s, p = f"{prefix}transformer.resblocks.{i}.", f"l{i}."
        w.update({p + "ln1_w": _f(sd[s + "ln_1.weight"]), p + "ln1_b": _f(sd[s + "ln_1.bias"]),
                  p + "qkv_w": _f(sd[s + "attn.in_proj_weight"]), p + "qkv_b": _f(sd[s + "attn.in_proj_bias"]),
                  p + "proj_w": _f(sd[s + "attn.out_proj.weight"]), p + "proj_b": _f(sd[s + "attn.out_proj.bias"]),
                  p + "ln2_w": _f(sd[s + "ln_2.weight"]), p + "ln2_b": _f(sd[s + "ln_2.bias"]),
                  p + "fc1_w": _f(sd[s + "mlp.c_fc.weight"]), p + "fc1_b": _f(sd[s + "mlp.c_fc.bias"]),
                  p + "fc2_w": _f(sd[s + "mlp.c_proj.weight"]), p + "fc2_b": _f(sd[s + "mlp.c_proj.bias"])})


def from_open_clip(sd: dict, cfg: ViTConfig) -> dict:
    """open_clip / OpenAI clip: the image tower (keys under `visual.` or bare) or, for a text configuration, the text tower."""
    sd = _strip(sd, prefixes=("module.",))
    w = {}
    if cfg.text_context:
        pre = "text." if "text.token_embedding.weight" in sd else ""
        w.update({"tok_emb": _f(sd[pre + "token_embedding.weight"]), "pos": _f(sd[pre + "positional_embedding"]),
                  "ln_post_w": _f(sd[pre + "ln_final.weight"]), "ln_post_b": _f(sd[pre + "ln_final.bias"])})
        if cfg.embed:
            w["proj"] = _f(sd[pre + "text_projection"])
        _resblocks(sd, pre, cfg, w)
        return _check(cfg, w)
    pre = "visual." if "visual.conv1.weight" in sd else ""
    pos = _f(sd[pre + "positional_embedding"])
    if pos.shape[0] != 1 + cfg.n_patch:                 # open_clip resizes with bicubic, no antialias, align_corners=False
        pos = interpolate_pos_embed(pos, cfg.grid, antialias=False, offset=0.0)
    w.update({"patch_w": _f(sd[pre + "conv1.weight"]).reshape(cfg.dim, -1),
              "patch_b": _f(sd[pre + "conv1.bias"]) if pre + "conv1.bias" in sd else torch.zeros(cfg.dim),
              "cls": _f(sd[pre + "class_embedding"]).reshape(-1), "pos": pos, "reg": torch.zeros(0, cfg.dim),
              "ln_pre_w": _f(sd[pre + "ln_pre.weight"]), "ln_pre_b": _f(sd[pre + "ln_pre.bias"]),
              "ln_post_w": _f(sd[pre + "ln_post.weight"]), "ln_post_b": _f(sd[pre + "ln_post.bias"])})
    if cfg.embed:
        w["proj"] = _f(sd[pre + "proj"])
    _resblocks(sd, pre, cfg, w)
    return _check(cfg, w)


def from_hf_clip(sd: dict, cfg: ViTConfig) -> dict:
    """transformers CLIPModel / CLIPVisionModel[WithProjection] / CLIPTextModel[WithProjection]."""
    sd = _strip(sd, prefixes=("module.",))
    text = bool(cfg.text_context)
    t = "text_model." if text else "vision_model."
    w = {}
    if text:
        w.update({"tok_emb": _f(sd[t + "embeddings.token_embedding.weight"]), "pos": _f(sd[t + "embeddings.position_embedding.weight"]),
                  "ln_post_w": _f(sd[t + "final_layer_norm.weight"]), "ln_post_b": _f(sd[t + "final_layer_norm.bias"])})
        if cfg.embed:
            w["proj"] = _f(sd["text_projection.weight"]).T.contiguous()
    else:
        w.update({"patch_w": _f(sd[t + "embeddings.patch_embedding.weight"]).reshape(cfg.dim, -1), "patch_b": torch.zeros(cfg.dim),
                  "cls": _f(sd[t + "embeddings.class_embedding"]).reshape(-1), "pos": _f(sd[t + "embeddings.position_embedding.weight"]),
                  "reg": torch.zeros(0, cfg.dim),
                  "ln_pre_w": _f(sd[t + "pre_layrnorm.weight"]), "ln_pre_b": _f(sd[t + "pre_layrnorm.bias"]),
                  "ln_post_w": _f(sd[t + "post_layernorm.weight"]), "ln_post_b": _f(sd[t + "post_layernorm.bias"])})
        if cfg.embed:
            w["proj"] = _f(sd["visual_projection.weight"]).T.contiguous()
    for i in range(cfg.depth):
        s, p = f"{t}encoder.layers.{i}.", f"l{i}."
        a = s + "self_attn."
        w.update({p + "ln1_w": _f(sd[s + "layer_norm1.weight"]), p + "ln1_b": _f(sd[s + "layer_norm1.bias"]),
                  p + "qkv_w": torch.cat([_f(sd[a + n + "_proj.weight"]) for n in "qkv"]),
                  p + "qkv_b": torch.cat([_f(sd[a + n + "_proj.bias"]) for n in "qkv"]),
                  p + "proj_w": _f(sd[a + "out_proj.weight"]), p + "proj_b": _f(sd[a + "out_proj.bias"]),
                  p + "ln2_w": _f(sd[s + "layer_norm2.weight"]), p + "ln2_b": _f(sd[s + "layer_norm2.bias"]),
                  p + "fc1_w": _f(sd[s + "mlp.fc1.weight"]), p + "fc1_b": _f(sd[s + "mlp.fc1.bias"]),
                  p + "fc2_w": _f(sd[s + "mlp.fc2.weight"]), p + "fc2_b": _f(sd[s + "mlp.fc2.bias"])})
    return _check(cfg, w)


_LOADERS = {"dinov2-hub": from_dinov2_hub, "hf-dinov2": from_hf_dinov2, "open_clip": from_open_clip, "hf-clip": from_hf_clip}


def load_state_dict(cfg: ViTConfig, sd, fmt: str | None = None) -> dict:
    """State dict (or a module with .state_dict()) of a supported checkpoint -> weight dict for ViTEngine(cfg, weights)."""
    if hasattr(sd, "state_dict"):
        sd = sd.state_dict()
    if "state_dict" in sd and isinstance(sd["state_dict"], dict):
        sd = sd["state_dict"]
    fmt = fmt or detect_format(_strip(sd))
    return _LOADERS[fmt](sd, cfg)


def load_checkpoint(cfg: ViTConfig, path: str, fmt: str | None = None) -> dict:
    """A checkpoint file (.pt / .pth / .bin via torch.load, .safetensors when the package is importable)."""
    if path.endswith(".safetensors"):
        from safetensors.torch import load_file
        sd = load_file(path)
    else:
        sd = torch.load(path, map_location="cpu", weights_only=True)
    return load_state_dict(cfg, sd, fmt)
