"""ViT feature-extractor configurations and weight containers (host side, no device code).

The memory's feature extractor plug-point is `preload_dino` (reference memory_2.py:43-46): any object with
`forward_features(x) -> {'x_norm_patchtokens': (B, n_patch, D)}`.  The reference loads DINOv2 ViT-L/14 with
4 register tokens (args.py:50); BASELINE.json's north_star names CLIP ViT-L/14.  One engine covers both via
the switches below.
"""
from __future__ import annotations

from dataclasses import dataclass, replace

import torch


@dataclass(frozen=True)
class ViTConfig:
    name: str
    image: int = 224
    patch: int = 14
    dim: int = 1024
    depth: int = 24
    heads: int = 16
    mlp: int = 4096
    n_reg: int = 0            # register tokens inserted after cls (DINOv2-reg: 4)
    layerscale: bool = False  # DINOv2
    act: str = "gelu"         # "gelu" (erf) | "quick_gelu" (x*sigmoid(1.702x), CLIP)
    ln_pre: bool = False      # CLIP pre-LayerNorm
    ln_eps: float = 1e-6
    patch_bias: bool = True
    embed: int = 0            # image-embedding width of the cls head (CLIP `proj`, D x embed, no bias); 0 = no head
    text_context: int = 0     # > 0: CLIP text tower over this many positions (causal attention, pooled at the end-of-text token)
    vocab: int = 0            # text tower: rows of the token-embedding table

    @property
    def head_dim(self) -> int:
        return self.dim // self.heads

    @property
    def grid(self) -> int:
        return self.image // self.patch

    @property
    def n_patch(self) -> int:
        return self.grid * self.grid

    @property
    def tokens(self) -> int:
        return self.text_context if self.text_context else 1 + self.n_reg + self.n_patch

    @property
    def patch_k(self) -> int:
        return 3 * self.patch * self.patch

    def flops_per_frame(self) -> float:
        """2*[Np*Kp*D + L*(T*D*3D + T*D*D + 2*T*D*M + 2*T^2*D)]."""
        T, D, L, M = self.tokens, self.dim, self.depth, self.mlp
        return 2.0 * (self.n_patch * self.patch_k * D + L * (T * D * 3 * D + T * D * D + 2 * T * D * M + 2 * T * T * D))


CLIP_L14 = ViTConfig("clip-vit-l14", act="quick_gelu", ln_pre=True, ln_eps=1e-5, patch_bias=False)
DINOV2_L14_REG = ViTConfig("dinov2-vitl14-reg", n_reg=4, layerscale=True, act="gelu", ln_eps=1e-6, patch_bias=True)
CLIP_B32 = ViTConfig("clip-vit-b32", patch=32, dim=768, depth=12, heads=12, mlp=3072, act="quick_gelu",
                     ln_pre=True, ln_eps=1e-5, patch_bias=False)

# the agent's view-check encoder (BSCAgent.py:309: open_clip 'ViT-H/14' -> config ViT-H-14: 32 layers, width 1280, 16 heads of
# dimension 80, MLP 5120, GELU(erf), 1024-d embedding); image tower only
OPENCLIP_H14 = ViTConfig("openclip-vit-h14", dim=1280, depth=32, heads=16, mlp=5120, act="gelu", ln_pre=True, ln_eps=1e-5,
                         patch_bias=False, embed=1024)
# its text tower (encode_text: goal text and long-term labels, BSCAgent.py:541-546, 708-711): width 1024, 24 layers, 16 heads,
# context 77, vocabulary 49 408, causal attention, pooled at the end-of-text token; the BPE tokenizer stays with the caller
OPENCLIP_H14_TEXT = ViTConfig("openclip-vit-h14-text", dim=1024, depth=24, heads=16, mlp=4096, act="gelu", ln_eps=1e-5,
                              patch_bias=False, embed=1024, text_context=77, vocab=49408)

CONFIGS = {c.name: c for c in (CLIP_L14, DINOV2_L14_REG, CLIP_B32, OPENCLIP_H14, OPENCLIP_H14_TEXT)}
# torch.hub entry names the reference passes as --dino_size (args.py:50)
CONFIGS["dinov2_vitl14_reg"] = DINOV2_L14_REG


def tiny(cfg: ViTConfig, dim=128, depth=2, heads=2, mlp=256) -> ViTConfig:
    """Same architecture switches at test size."""
    return replace(cfg, name=cfg.name + "-tiny", dim=dim, depth=depth, heads=heads, mlp=mlp)


def random_weights(cfg: ViTConfig, seed: int = 0, perturb: bool = False, std: float = 0.02) -> dict:
    """Seeded random-init weights: linear/embedding ~ N(0, std), LN = 1/0, LayerScale = 1.

    perturb=True additionally randomises LN affine / LayerScale / biases so every epilogue term is exercised
    by parity tests.  Layout (fp32, CPU): patch_w (D, 3*P*P) in (c,ky,kx) order, qkv_w (3D, D) = [q;k;v].
    """
    g = torch.Generator().manual_seed(seed)
    D, M = cfg.dim, cfg.mlp

    def n(*shape, s=std):
        return torch.randn(*shape, generator=g) * s

    def ln_pair():
        if perturb:
            return 1.0 + n(D, s=0.1), n(D, s=0.1)
        return torch.ones(D), torch.zeros(D)

    if cfg.text_context:
        w = {"tok_emb": n(cfg.vocab, D), "pos": n(cfg.text_context, D, s=0.01)}
    else:
        w = {
            "patch_w": n(D, cfg.patch_k),
            "patch_b": n(D) if cfg.patch_bias else torch.zeros(D),
            "cls": n(D),
            "reg": n(max(cfg.n_reg, 1), D)[: cfg.n_reg] if cfg.n_reg else torch.zeros(0, D),
            "pos": n(1 + cfg.n_patch, D),
        }
    if cfg.ln_pre:
        w["ln_pre_w"], w["ln_pre_b"] = ln_pair()
    for i in range(cfg.depth):
        p = f"l{i}."
        w[p + "ln1_w"], w[p + "ln1_b"] = ln_pair()
        w[p + "qkv_w"], w[p + "qkv_b"] = n(3 * D, D), n(3 * D)
        w[p + "proj_w"], w[p + "proj_b"] = n(D, D), n(D)
        w[p + "ln2_w"], w[p + "ln2_b"] = ln_pair()
        w[p + "fc1_w"], w[p + "fc1_b"] = n(M, D), n(M)
        w[p + "fc2_w"], w[p + "fc2_b"] = n(D, M), n(D)
        if cfg.layerscale:
            w[p + "ls1"] = 1.0 + n(D, s=0.1) if perturb else torch.ones(D)
            w[p + "ls2"] = 1.0 + n(D, s=0.1) if perturb else torch.ones(D)
    w["ln_post_w"], w["ln_post_b"] = ln_pair()
    if cfg.embed:
        w["proj"] = n(D, cfg.embed, s=D ** -0.5)
    return w
