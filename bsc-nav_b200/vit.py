"""ViT feature extractor engine (host side): weight packing and the `forward_features` plug-point.

The object returned by `ViTEngine(...)` can be passed as `preload_dino` to VoxelTokenMemory exactly like the
torch.hub DINOv2 module the reference passes (create_memory_for_dataset.py:58-70, memory_2.py:43-46): it exposes
`forward_features(x) -> {'x_norm_patchtokens': (B, n_patch, D)}`.  All arithmetic runs in libbscnav_b200.so
(tcgen05 GEMMs, fp16 operands / fp32 accumulation, fp32 residual stream, LayerNorm and softmax).
"""
from __future__ import annotations

import ctypes as C
import math

import torch

from . import _lib
from .vit_config import ViTConfig


PRECISIONS = {"fp16": 0, "bf16x3": 1}


def _cfg_struct(cfg: ViTConfig, precision: str = "fp16") -> _lib.VitCfg:
    k_pad = (cfg.patch_k + 63) // 64 * 64
    return _lib.VitCfg(image=cfg.image, patch=cfg.patch, dim=cfg.dim, depth=cfg.depth, heads=cfg.heads, mlp=cfg.mlp,
                       n_reg=cfg.n_reg, act=0 if cfg.act == "gelu" else 1, ln_pre=1 if cfg.ln_pre else 0,
                       patch_k_pad=k_pad, ln_eps=cfg.ln_eps, head_dim=cfg.dim // cfg.heads, text_context=cfg.text_context,
                       precision=PRECISIONS[precision])


def split_bf16(m: torch.Tensor) -> torch.Tensor:
    """fp32 [N, K] -> bf16 [N, 2K] rows (hi | lo): hi = bf16(m), lo = bf16(m - hi); m = hi + lo up to ~2^-17 relative."""
    m = m.float()
    hi = m.to(torch.bfloat16)
    lo = (m - hi.float()).to(torch.bfloat16)
    return torch.cat([hi, lo], dim=1)


def pack_weights(cfg: ViTConfig, w: dict, precision: str = "fp16"):
    """fp32 weight dict (vit_config.random_weights layout) -> (half blob, float blob), CPU tensors.
    precision="bf16x3" (the wide-range path): the 16-bit blob holds every matrix as (hi | lo) bf16 row pairs (split_bf16).

    Folded on the host (exact in real arithmetic): 1/sqrt(head_dim) into the q rows of qkv (a power of two for
    head_dim 64, so also exact in fp16; for head_dim 80 the scaled fp32 weight is rounded to fp16 once), LayerScale gamma
    into proj / fc2 weights and biases, cls + pos[0].
    Layout must match csrc/vit_engine.cu:vit_layout.
    """
    D, M = cfg.dim, cfg.mlp
    dh = D // cfg.heads
    if dh not in (64, 80):
        raise ValueError("engine supports head_dim 64 (tcgen05 attention) and 80 (ViT-H/14, mma.sync attention)")
    k_pad = (cfg.patch_k + 63) // 64 * 64
    half, flt = [], []
    if cfg.text_context:                      # text tower: no patch embedding / cls; the token table stays a separate tensor
        flt.append(w["pos"].reshape(-1))
    else:
        pw = torch.zeros(D, k_pad)
        pw[:, : cfg.patch_k] = w["patch_w"]
        half.append(pw)
        flt += [w["patch_b"] if cfg.patch_bias else torch.zeros(D), w["cls"] + w["pos"][0]]
        if cfg.n_reg:
            flt.append(w["reg"].reshape(-1))
        flt.append(w["pos"][1:].reshape(-1))
    if cfg.ln_pre:
        flt += [w["ln_pre_w"], w["ln_pre_b"]]
    qs = 1.0 / math.sqrt(dh)
    for i in range(cfg.depth):
        p = f"l{i}."
        qkv_w, qkv_b = w[p + "qkv_w"].clone(), w[p + "qkv_b"].clone()
        qkv_w[:D] *= qs
        qkv_b[:D] *= qs
        proj_w, proj_b, fc2_w, fc2_b = w[p + "proj_w"], w[p + "proj_b"], w[p + "fc2_w"], w[p + "fc2_b"]
        if cfg.layerscale:
            proj_w, proj_b = proj_w * w[p + "ls1"][:, None], proj_b * w[p + "ls1"]
            fc2_w, fc2_b = fc2_w * w[p + "ls2"][:, None], fc2_b * w[p + "ls2"]
        half += [qkv_w, proj_w, w[p + "fc1_w"], fc2_w]
        flt += [w[p + "ln1_w"], w[p + "ln1_b"], qkv_b, proj_b, w[p + "ln2_w"], w[p + "ln2_b"], w[p + "fc1_b"], fc2_b]
    flt += [w["ln_post_w"], w["ln_post_b"]]
    if precision == "bf16x3":
        if dh != 64:
            raise ValueError("the wide-range path supports head_dim 64")
        hb = torch.cat([split_bf16(t).reshape(-1) for t in half])
    else:
        hb = torch.cat([t.reshape(-1).float() for t in half]).to(torch.float16)
    fb = torch.cat([t.reshape(-1).float() for t in flt]).contiguous()
    return hb, fb


class ViTEngine:
    """Device-resident ViT: packed weights + workspace + TMA descriptors, one instance per GPU."""

    def __init__(self, cfg: ViTConfig, weights: dict, max_batch: int = 64, device="cuda", precision: str = "fp16"):
        """precision: "fp16" -- fp16 operands / fp32 accumulation, the throughput path (an activation beyond the fp16 range raises
        through check_range()); "bf16x3" -- wide-range path: split bf16 operands, three tensor-core passes per product, fp32
        attention: fp32 range and ~1e-5 relative error against the fp32 reference (memory_2.py:737-739) at ~1/3 of the rate."""
        self.lib = _lib.lib()
        self.cfg = cfg
        self.device = torch.device(device)
        self.max_batch = int(max_batch)
        if precision not in PRECISIONS:
            raise ValueError(f"precision must be one of {sorted(PRECISIONS)}")
        self.precision = precision
        self._c = _cfg_struct(cfg, precision)
        hb, fb = pack_weights(cfg, weights, precision)
        if hb.numel() != self.lib.bscnav_vit_half_elems(C.byref(self._c)) or fb.numel() != self.lib.bscnav_vit_float_elems(C.byref(self._c)):
            raise _lib.BscnavError("weight blob layout mismatch between vit.py and vit_engine.cu")
        self.w_half = hb.to(self.device)
        self.w_float = fb.to(self.device)
        self.proj = weights["proj"].to(self.device, torch.float32).contiguous() if "proj" in weights else None
        self.tok_emb = weights["tok_emb"].to(self.device, torch.float32).contiguous() if cfg.text_context else None
        nbytes = _lib.check(self.lib.bscnav_vit_workspace_bytes(C.byref(self._c), self.max_batch), "vit_workspace_bytes")
        self.workspace = torch.empty(nbytes, dtype=torch.uint8, device=self.device)
        h = C.c_void_p()
        with torch.cuda.device(self.device):
            _lib.check(self.lib.bscnav_vit_create(C.byref(self._c), _lib.ptr(self.w_half), _lib.ptr(self.w_float),
                                                  _lib.ptr(self.workspace), nbytes, self.max_batch, C.byref(h)), "vit_create")
        self._h = h

    @classmethod
    def auto(cls, cfg: ViTConfig, weights: dict, max_batch: int = 64, device="cuda", probe: torch.Tensor | None = None,
             tol: float = 5e-4):
        """Pick the precision from the weights: run `probe` (uint8 frames (B, H, W, 3|4) on the device; default: 2 seeded
        random 224 x 224 frames) through both paths and keep the fp16 throughput path only if its tokens agree with the
        wide-range path to `tol` per token (half the 1e-3 parity bar) and no activation left the fp16 range.  Trained checkpoints
        with outlier channels fail this and get the bf16x3 engine.  The measured error is kept in `.probe_error`."""
        dev = torch.device(device)
        if probe is None:
            probe = torch.randint(0, 256, (2, cfg.image, cfg.image, 3), dtype=torch.uint8, generator=torch.Generator().manual_seed(0)).to(dev)
        nb = min(int(probe.shape[0]), int(max_batch))
        wide = cls(cfg, weights, max_batch=max_batch, device=device, precision="bf16x3")
        ref = wide.forward_frames(probe[:nb].contiguous())
        fast = cls(cfg, weights, max_batch=max_batch, device=device, precision="fp16")
        got = fast.forward_frames(probe[:nb].contiguous())
        overflow = fast.range_overflow()
        err = float(((got - ref).norm(dim=-1) / ref.norm(dim=-1)).max()) if not overflow else float("inf")
        keep = fast if (not overflow and err < tol) else wide
        keep.probe_error = err
        return keep

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self.lib.bscnav_vit_destroy(h)

    # -- frames path: uint8 HWC (3 or 4 channels) -> patch tokens (B, n_patch, D) fp32, optionally the pre-processed image
    def forward_frames(self, rgb_u8: torch.Tensor, out: torch.Tensor | None = None, return_pre: bool = False):
        assert rgb_u8.dtype == torch.uint8 and rgb_u8.is_cuda and rgb_u8.dim() == 4 and rgb_u8.is_contiguous()
        B, H, W, ch = rgb_u8.shape
        if out is None:
            out = torch.empty(B, self.cfg.n_patch, self.cfg.dim, dtype=torch.float32, device=self.device)
        pre = torch.empty(B, 3, self.cfg.image, self.cfg.image, dtype=torch.float32, device=self.device) if return_pre else None
        for b0 in range(0, B, self.max_batch):
            nb = min(self.max_batch, B - b0)
            _lib.check(self.lib.bscnav_vit_forward_frames(self._h, _lib.ptr(rgb_u8[b0:]), nb, H, W, ch, _lib.ptr(out[b0:]),
                                                          _lib.ptr(pre[b0:]) if return_pre else None, _lib.stream_ptr()),
                       "vit_forward_frames")
        return (out, pre) if return_pre else out

    # -- the reference's plug-point: normalised float images (B,3,S,S) on the device
    @torch.no_grad()
    def forward_features(self, x: torch.Tensor) -> dict:
        x = x.to(self.device, torch.float32).contiguous()
        B = x.shape[0]
        assert x.shape[1:] == (3, self.cfg.image, self.cfg.image), x.shape
        out = torch.empty(B, self.cfg.n_patch, self.cfg.dim, dtype=torch.float32, device=self.device)
        for b0 in range(0, B, self.max_batch):
            nb = min(self.max_batch, B - b0)
            _lib.check(self.lib.bscnav_vit_forward_images(self._h, _lib.ptr(x[b0:]), nb, _lib.ptr(out[b0:]), _lib.stream_ptr()),
                       "vit_forward_images")
        return {"x_norm_patchtokens": out}

    # -- CUDA graphs for the launch-bound small calls (one goal image, one prompt: ~170-230 kernels of a few microseconds):
    #    the forward for a given (entry, batch) is captured once on static buffers and replayed
    def _graphed(self, key, run, x: torch.Tensor, out_shape) -> torch.Tensor:
        if not hasattr(self, "_graphs"):
            self._graphs = {}
        if key not in self._graphs:
            s_in = torch.empty_like(x)
            s_out = torch.empty(out_shape, dtype=torch.float32, device=self.device)
            s_in.copy_(x)
            side = torch.cuda.Stream(device=self.device)
            side.wait_stream(torch.cuda.current_stream())
            with torch.cuda.stream(side):                 # warm-up off the capture: one-time attribute calls, a0 clearing
                for _ in range(2):
                    run(s_in, s_out)
            torch.cuda.current_stream().wait_stream(side)
            g = torch.cuda.CUDAGraph()
            with torch.cuda.graph(g):
                run(s_in, s_out)
            self._graphs[key] = (g, s_in, s_out)
        g, s_in, s_out = self._graphs[key]
        s_in.copy_(x)
        g.replay()
        return s_out.clone()

    # -- open_clip `encode_image` (BSCAgent.py:710, 721-724): normalised float images -> (B, embed) image embeddings
    @torch.no_grad()
    def encode_image(self, x: torch.Tensor, graph: bool = False) -> torch.Tensor:
        if self.proj is None:
            raise _lib.BscnavError("this ViT has no image-embedding head (weights['proj'])")
        x = x.to(self.device, torch.float32).contiguous()
        B, E = x.shape[0], self.proj.shape[1]
        assert x.shape[1:] == (3, self.cfg.image, self.cfg.image) and self.proj.shape[0] == self.cfg.dim, x.shape

        def run(src, dst, nb=None):
            nb = src.shape[0] if nb is None else nb
            _lib.check(self.lib.bscnav_vit_encode_images(self._h, _lib.ptr(src), nb, _lib.ptr(self.proj), E, _lib.ptr(dst),
                                                         _lib.stream_ptr()), "vit_encode_images")
        if graph and B <= self.max_batch:
            return self._graphed(("image", B), run, x, (B, E))
        out = torch.empty(B, E, dtype=torch.float32, device=self.device)
        for b0 in range(0, B, self.max_batch):
            run(x[b0:], out[b0:], min(self.max_batch, B - b0))
        return out

    # -- open_clip `encode_text` (BSCAgent.py:545-546, 711): token ids (B, context) from the caller's tokenizer -> (B, embed)
    @torch.no_grad()
    def encode_text(self, token_ids: torch.Tensor, graph: bool = False) -> torch.Tensor:
        if not self.cfg.text_context or self.proj is None:
            raise _lib.BscnavError("not a text tower (cfg.text_context, weights['tok_emb'], weights['proj'])")
        ids = token_ids.to(self.device, torch.int32).contiguous()
        assert ids.dim() == 2 and ids.shape[1] == self.cfg.text_context, ids.shape
        B, E = ids.shape[0], self.proj.shape[1]

        def run(src, dst, nb=None):
            nb = src.shape[0] if nb is None else nb
            _lib.check(self.lib.bscnav_text_encode(self._h, _lib.ptr(src), nb, _lib.ptr(self.tok_emb), self.tok_emb.shape[0],
                                                   _lib.ptr(self.proj), E, _lib.ptr(dst), _lib.stream_ptr()), "text_encode")
        if graph and B <= self.max_batch:
            return self._graphed(("text", B), run, ids, (B, E))
        out = torch.empty(B, E, dtype=torch.float32, device=self.device)
        for b0 in range(0, B, self.max_batch):
            run(ids[b0:], out[b0:], min(self.max_batch, B - b0))
        return out

    def range_overflow(self, reset: bool = True) -> bool:
        """True when an fp16 activation (qkv or the MLP's hidden layer) overflowed since the last reset; synchronises."""
        v = C.c_int32(0)
        _lib.check(self.lib.bscnav_vit_range_status(self._h, 1 if reset else 0, C.byref(v), _lib.stream_ptr()), "vit_range_status")
        return bool(v.value)

    def check_range(self):
        """Raise if the fp16 activation range was exceeded: the tokens produced since the last check are not trustworthy."""
        if self.range_overflow():
            raise _lib.BscnavError(f"{self.cfg.name}: an activation left the fp16 range (|x| > 65504) inside the ViT; "
                                   "these weights need the wide-range path (ViTEngine(..., precision='bf16x3'))")

    def profile(self, enable: bool):
        _lib.check(self.lib.bscnav_vit_profile(self._h, 1 if enable else 0), "vit_profile")

    def profile_read(self) -> dict:
        """{category: (ms, work, launches)} for 'gemm' (FLOPs), 'attention' (FLOPs), 'layernorm' (bytes), 'preprocess' (bytes)."""
        ms, work, n = (C.c_double * 4)(), (C.c_double * 4)(), (C.c_int32 * 4)()
        _lib.check(self.lib.bscnav_vit_profile_read(self._h, ms, work, n), "vit_profile_read")
        return {k: (ms[i], work[i], n[i]) for i, k in enumerate(("gemm", "attention", "layernorm", "preprocess"))}

    def to(self, *_a, **_k):     # callers do `.to('cuda')` on the hub module (create_memory_for_dataset.py:58)
        return self

    def eval(self):
        return self


def gaussian_pool(tokens: torch.Tensor) -> torch.Tensor:
    """Query embedding (memory_2.py:592-608): (B, g*g, D) device tokens -> (1, D)."""
    B, n, D = tokens.shape
    g = int(math.isqrt(n))
    out = torch.empty(1, D, dtype=torch.float32, device=tokens.device)
    _lib.check(_lib.lib().bscnav_gaussian_pool(_lib.ptr(tokens.contiguous()), B, g, D, _lib.ptr(out), _lib.stream_ptr()),
               "gaussian_pool")
    return out


def gemm_f16(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, out_mode: int = 0, act: int = 0,
             out: torch.Tensor | None = None) -> torch.Tensor:
    """out = a @ w.T (+bias): fp16 operands, fp32 accumulation on tcgen05 (tests / micro-benchmarks)."""
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float16 if out_mode in (0, 1) else torch.float32, device=a.device)
        if out_mode == 2:
            out.zero_()
    _lib.check(_lib.lib().bscnav_gemm_f16(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), M, N, K, out_mode, act,
                                          _lib.stream_ptr()), "gemm_f16")
    return out


def gemm_f16_sliced(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None, out_mode: int = 0, act: int = 0,
                    out: torch.Tensor | None = None, scratch: torch.Tensor | None = None):
    """gemm_f16 on the engine's small-M path (128-wide tiles, K sliced over idle SMs when the launcher decides to); returns
    (out, slices).  `scratch`: zero-initialised uint8 device tensor, reusable across calls (tests)."""
    M, K = a.shape
    N = w.shape[0]
    if out is None:
        out = torch.empty(M, N, dtype=torch.float16 if out_mode in (0, 1) else torch.float32, device=a.device)
        if out_mode == 2:
            out.zero_()
    if scratch is None:
        scratch = torch.zeros(1024 + 152 * 128 * 128 * 4, dtype=torch.uint8, device=a.device)
    n = C.c_int32(0)
    _lib.check(_lib.lib().bscnav_gemm_f16_sliced(_lib.ptr(a), _lib.ptr(w), _lib.ptr(bias), _lib.ptr(out), M, N, K, out_mode, act,
                                                 _lib.ptr(scratch), scratch.numel(), C.byref(n), _lib.stream_ptr()), "gemm_f16_sliced")
    return out, int(n.value)


def gemm_bf16x3(a: torch.Tensor, w: torch.Tensor, bias: torch.Tensor | None = None) -> torch.Tensor:
    """out = a @ w.T (+bias) from fp32 inputs on the wide-range path: operands split into (hi | lo) bf16 pairs, three tcgen05
    passes, fp32 out (tests)."""
    M, K = a.shape
    N = w.shape[0]
    a2, w2 = split_bf16(a).contiguous(), split_bf16(w).contiguous()
    out = torch.empty(M, N, dtype=torch.float32, device=a.device)
    _lib.check(_lib.lib().bscnav_gemm_bf16x3(_lib.ptr(a2), _lib.ptr(w2), _lib.ptr(bias), _lib.ptr(out), M, N, K, _lib.stream_ptr()),
               "gemm_bf16x3")
    return out


def score_views(embeds: torch.Tensor, goal: torch.Tensor):
    """check_around's scoring (BSCAgent.py:713-718, 726-731): cosine of every view embedding (n, E) with the goal embedding
    (E,), softmax over the views -> (probabilities (n,) float32, index of the best view as a 0-d int64 tensor), on the device."""
    embeds = embeds.to(torch.float32).contiguous()
    goal = goal.to(embeds.device, torch.float32).reshape(-1).contiguous()
    n, E = embeds.shape
    assert goal.numel() == E
    prob = torch.empty(n, dtype=torch.float32, device=embeds.device)
    best = torch.empty(2, dtype=torch.int32, device=embeds.device)
    _lib.check(_lib.lib().bscnav_score_views(_lib.ptr(embeds), _lib.ptr(goal), n, E, _lib.ptr(prob), _lib.ptr(best), _lib.stream_ptr()),
               "score_views")
    return prob, best[0].long()


def attention_f16(qkv: torch.Tensor, batch: int, T: int, heads: int, impl: int = 1, head_dim: int = 64,
                  causal: bool = False) -> torch.Tensor:
    """Self-attention over a packed fp16 qkv tensor [batch*T, 3*heads*head_dim] (q pre-scaled); impl 0 = mma.sync (head_dim 64
    or 80, optional causal mask at head_dim 64), 1 / 2 = tcgen05 kernels (head_dim 64, no mask)."""
    out = torch.empty(batch * T, heads * head_dim, dtype=torch.float16, device=qkv.device)
    if head_dim != 64 or causal:
        if impl not in (0, 3) or (impl == 3 and (head_dim != 64 or not causal)):
            raise ValueError("head_dim != 64 runs on the mma.sync kernel only (impl=0); causal masks on impl 0 or 3")
        _lib.check(_lib.lib().bscnav_attention_f16_hd(_lib.ptr(qkv), _lib.ptr(out), batch, T, heads, head_dim,
                                                      (2 if impl == 3 else 1) if causal else 0, _lib.stream_ptr()), "attention_f16_hd")
        return out
    _lib.check(_lib.lib().bscnav_attention_f16(_lib.ptr(qkv), _lib.ptr(out), batch, T, heads, impl, _lib.stream_ptr()), "attention_f16")
    return out
