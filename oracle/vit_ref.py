"""TEST INFRASTRUCTURE ONLY -- plain-torch fp32 restatement of the ViT the memory uses as feature extractor.

PARITY UNPINNED for this file: the ViT arithmetic is third-party code that is NOT under
/root/reference -- `torch.hub.load('facebookresearch/dinov2', 'dinov2_vitl14_reg')`, unpinned
(reference call sites memory_2.py:43, :591-592, :738-739: `forward_features(x)['x_norm_patchtokens']`);
no reference test holds golden ViT outputs and no checkpoint is reachable offline.  The published
architecture is restated below and cross-checked (tests/test_vit_oracle.py) against the
implementations installed in this image, `transformers.Dinov2WithRegistersModel` and
`transformers.CLIPVisionModel`, on seeded random weights.

Architecture switches:
  dinov2-reg : patch-embed conv WITH bias, [cls]+patches +pos, 4 register tokens inserted after cls,
               blocks x += ls1*Attn(LN1 x); x += ls2*MLP(LN2 x), GELU(erf), LN eps 1e-6, final LN, drop cls+reg.
  clip       : patch-embed conv without bias, [cls]+patches +pos, pre-LN, blocks without LayerScale,
               QuickGELU, LN eps 1e-5, post-LN on all tokens, drop cls.
Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this.
"""
from __future__ import annotations

import math

import torch
import torch.nn.functional as F


def _ln(x, w, b, eps):
    return F.layer_norm(x, (x.shape[-1],), w, b, eps)


def _lin(x, w, b, quant):
    if quant is not None:
        x, w = quant(x), quant(w)
    return F.linear(x, w, b)


def patchify(x: torch.Tensor, patch: int) -> torch.Tensor:
    """(B,3,H,W) -> (B, nh*nw, 3*patch*patch), feature order (c, ky, kx) == conv weight flattening."""
    B, C, H, W = x.shape
    nh, nw = H // patch, W // patch
    x = x.reshape(B, C, nh, patch, nw, patch).permute(0, 2, 4, 1, 3, 5)
    return x.reshape(B, nh * nw, C * patch * patch)


def _blocks(h, w, cfg, quant=None, causal=False):
    """The transformer blocks (pre-norm residual attention + MLP); causal: keys <= query position (CLIP text tower)."""
    D, Hh = cfg.dim, cfg.heads
    dh = D // Hh
    B, T = h.shape[0], h.shape[1]
    mask = torch.full((T, T), float("-inf")).triu(1) if causal else None
    for i in range(cfg.depth):
        p = f"l{i}."
        y = _ln(h, w[p + "ln1_w"], w[p + "ln1_b"], cfg.ln_eps)
        qkv = _lin(y, w[p + "qkv_w"], w[p + "qkv_b"], quant).reshape(B, T, 3, Hh, dh).permute(2, 0, 3, 1, 4)
        q, k, v = qkv[0], qkv[1], qkv[2]
        if quant is not None:
            q, k, v = quant(q), quant(k), quant(v)
        s = torch.matmul(q, k.transpose(-1, -2)) / math.sqrt(dh)
        if mask is not None:
            s = s + mask
        a = torch.softmax(s, dim=-1)
        if quant is not None:
            a = quant(a)
        o = torch.matmul(a, v).permute(0, 2, 1, 3).reshape(B, T, D)
        o = _lin(o, w[p + "proj_w"], w[p + "proj_b"], quant)
        if cfg.layerscale:
            o = o * w[p + "ls1"]
        h = h + o
        y = _ln(h, w[p + "ln2_w"], w[p + "ln2_b"], cfg.ln_eps)
        y = _lin(y, w[p + "fc1_w"], w[p + "fc1_b"], quant)
        y = F.gelu(y) if cfg.act == "gelu" else y * torch.sigmoid(1.702 * y)
        y = _lin(y, w[p + "fc2_w"], w[p + "fc2_b"], quant)
        if cfg.layerscale:
            y = y * w[p + "ls2"]
        h = h + y
    return h


@torch.no_grad()
def vit_forward(x: torch.Tensor, w: dict, cfg, quant=None, return_all: bool = False) -> torch.Tensor:
    """fp32 forward.  x (B,3,S,S) normalised image -> patch tokens (B, Np, D).

    `quant`, if given, is applied to both operands of every matmul (operand-rounding emulation used
    to budget fp16/bf16/tf32 tensor-core error; accumulation stays fp32).
    """
    D, Hh = cfg.dim, cfg.heads
    dh = D // Hh
    B = x.shape[0]
    t = _lin(patchify(x, cfg.patch), w["patch_w"], w["patch_b"] if cfg.patch_bias else None, quant)
    t = t + w["pos"][1:]
    cls = (w["cls"] + w["pos"][0]).expand(B, 1, D)
    parts = [cls]
    if cfg.n_reg:
        parts.append(w["reg"].expand(B, cfg.n_reg, D))
    parts.append(t)
    h = torch.cat(parts, dim=1)
    if cfg.ln_pre:
        h = _ln(h, w["ln_pre_w"], w["ln_pre_b"], cfg.ln_eps)
    h = _blocks(h, w, cfg, quant)
    h = _ln(h, w["ln_post_w"], w["ln_post_b"], cfg.ln_eps)
    if return_all:
        return h
    return h[:, 1 + cfg.n_reg:]


@torch.no_grad()
def encode_image(x: torch.Tensor, w: dict, cfg, quant=None) -> torch.Tensor:
    """open_clip / CLIP `encode_image` (the agent's view-check encoder, BSCAgent.py:309, 710, 721-724; third-party, restated
    and cross-checked against transformers.CLIPVisionModelWithProjection in tests/test_vit_oracle.py): the post-LayerNorm'd
    cls token times the bias-free projection `proj` (D, E)."""
    return vit_forward(x, w, cfg, quant, return_all=True)[:, 0] @ w["proj"]


@torch.no_grad()
def encode_text(token_ids: torch.Tensor, w: dict, cfg, quant=None) -> torch.Tensor:
    """open_clip `encode_text` (BSCAgent.py:545-546, 711; third-party, restated and cross-checked against
    transformers.CLIPTextModelWithProjection): token + positional embedding, causal transformer, final LayerNorm, the row at
    argmax(token_ids) (end-of-text has the largest id) times the bias-free text projection (D, E)."""
    ids = token_ids.long()
    h = w["tok_emb"][ids] + w["pos"]
    h = _ln(_blocks(h, w, cfg, quant, causal=True), w["ln_post_w"], w["ln_post_b"], cfg.ln_eps)
    return h[torch.arange(h.shape[0]), ids.argmax(dim=-1)] @ w["proj"]


def score_views(image_features: torch.Tensor, goal_features: torch.Tensor):
    """check_around, BSCAgent.py:713-718 / 726-731: normalise, similarities, softmax over the views, argmax."""
    a = F.normalize(image_features, p=2, dim=-1)
    g = F.normalize(goal_features.reshape(1, -1), p=2, dim=-1)
    sims = torch.softmax(torch.matmul(a, g.T).squeeze(1), dim=0)
    return sims, torch.max(sims, dim=0)[1]


class TorchViT(torch.nn.Module):
    """Adapter with the reference's plug-point signature (`preload_dino`, memory_2.py:43-46)."""

    def __init__(self, weights: dict, cfg, quant=None):
        super().__init__()
        self.w, self.cfg, self.quant = weights, cfg, quant

    def forward_features(self, x):
        return {"x_norm_patchtokens": vit_forward(x.float(), self.w, self.cfg, self.quant)}


def preprocess_frame(rgb_u8: torch.Tensor, size: int = 224) -> torch.Tensor:
    """Frame path of memory_2.py:71-74,733-735: u8 HWC -> /255 -> Resize (bilinear, antialias, on a tensor)
    -> ImageNet normalise.  rgb_u8 (B,H,W,3) uint8 -> (B,3,size,size) f32."""
    x = rgb_u8.permute(0, 3, 1, 2).float() / 255
    x = F.interpolate(x, size=(size, size), mode="bilinear", align_corners=False, antialias=True)
    mean = torch.tensor([0.485, 0.456, 0.406]).view(1, 3, 1, 1)
    std = torch.tensor([0.229, 0.224, 0.225]).view(1, 3, 1, 1)
    return (x - mean) / std


def gaussian_pool(tokens: torch.Tensor) -> torch.Tensor:
    """Query embedding of memory_2.py:592-608: (B, n*n, D) -> (1, D)."""
    Bq, n_tok, _ = tokens.shape
    g = int(math.sqrt(n_tok))
    xs = torch.arange(g).repeat(g).view(1, n_tok)
    ys = torch.arange(g).repeat_interleave(g).view(1, n_tok)
    c = (g - 1) / 2
    d = (xs - c) ** 2 + (ys - c) ** 2
    wts = torch.exp(-d / (2 * (g / 2) ** 2))
    wts = (wts / wts.sum(dim=1, keepdim=True)).unsqueeze(-1)
    return (tokens * wts).sum(dim=1).mean(dim=0).unsqueeze(0)
