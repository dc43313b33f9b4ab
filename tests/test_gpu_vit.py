"""GPU parity of the ViT engine pieces (tcgen05 GEMM, LayerNorm, attention, pre-processing) and of the whole
forward against the torch-fp32 oracle (oracle/vit_ref.py).  Tolerance: 1e-3 relative, per token (L2), as
BASELINE.json's north_star states for feature vectors."""
import importlib
import os

import numpy as np
import pytest
import torch

from oracle import vit_ref

pytestmark = pytest.mark.gpu


def _rel(a, b):
    return ((a - b).norm(dim=-1) / b.norm(dim=-1)).max().item()


@pytest.mark.parametrize("M,N,K", [(128, 128, 64), (300, 256, 128), (1000, 1024, 1024), (257 * 3, 3072, 1024), (515, 4096, 1024),
                                   (200, 1024, 4096), (4112, 2304, 768), (514, 3840, 1280), (300, 1280, 5120), (77, 3072, 1024), (50, 1024, 4096)])
def test_gemm_modes(pkg, M, N, K):
    g = torch.Generator("cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda", generator=g)).half()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    out = pkg.gemm_f16(a, w, bias, out_mode=0)
    torch.testing.assert_close(out.float(), ref, rtol=2e-3, atol=2e-3)
    out = pkg.gemm_f16(a, w, bias, out_mode=1, act=0)
    torch.testing.assert_close(out.float(), torch.nn.functional.gelu(ref), rtol=2e-3, atol=2e-3)
    out = pkg.gemm_f16(a, w, bias, out_mode=1, act=1)
    torch.testing.assert_close(out.float(), ref * torch.sigmoid(1.702 * ref), rtol=2e-3, atol=2e-3)
    res = torch.randn(M, N, device="cuda", generator=g)
    out = pkg.gemm_f16(a, w, bias, out_mode=2, out=res.clone())
    torch.testing.assert_close(out, res + ref, rtol=1e-4, atol=1e-4)
    out = pkg.gemm_f16(a, w, None, out_mode=4, out=torch.empty(M, N, device="cuda"))
    torch.testing.assert_close(out, ref - bias, rtol=1e-4, atol=1e-4)


@pytest.mark.parametrize("M,N,K,want_slices", [(257, 1024, 4096, 6), (257, 1024, 2048, 4), (257, 3072, 1024, 1), (77, 1024, 4096, 8),
                                               (1028, 1024, 4096, 2), (130, 1280, 5120, 7), (257, 1024, 320, 1), (2056, 1024, 4096, 1)])
def test_gemm_k_slices(pkg, M, N, K, want_slices):
    """Small-M launches with a long K and fp32 output slice K over idle SMs (gemm_tcgen05.cu: partial tiles in scratch, the
    last-arriving slice sums them in slice order): the epilogues against torch, the slice count the launcher chose, bit-repeatable
    results whatever the arrival order, counters left re-armed; fp16-output modes run unsliced through the same entry."""
    g = torch.Generator("cuda").manual_seed(M + N + K)
    a = (torch.randn(M, K, device="cuda", generator=g)).half()
    w = (torch.randn(N, K, device="cuda", generator=g) * 0.05).half()
    bias = torch.randn(N, device="cuda", generator=g)
    ref = a.float() @ w.float().t() + bias
    scratch = torch.zeros(1024 + 152 * 128 * 128 * 4, dtype=torch.uint8, device="cuda")
    res = torch.randn(M, N, device="cuda", generator=g)
    first = None
    for _ in range(4):                                    # the arrival order of the slices varies, the sum must not
        out, s = pkg.gemm_f16_sliced(a, w, bias, out_mode=2, out=res.clone(), scratch=scratch)
        if torch.cuda.get_device_properties(0).multi_processor_count == 148 and os.environ.get("BSCNAV_KSLICE", "1") != "0":
            assert s == want_slices
        torch.testing.assert_close(out, res + ref, rtol=1e-4, atol=1e-4)
        first = out if first is None else first
        assert torch.equal(out, first)
    out, _ = pkg.gemm_f16_sliced(a, w, None, out_mode=4, out=torch.empty(M, N, device="cuda"), scratch=scratch)
    torch.testing.assert_close(out, ref - bias, rtol=1e-4, atol=1e-4)
    out, s = pkg.gemm_f16_sliced(a, w, bias, out_mode=0, scratch=scratch)
    assert s == 1
    torch.testing.assert_close(out.float(), ref, rtol=2e-3, atol=2e-3)
    out, _ = pkg.gemm_f16_sliced(a, w, bias, out_mode=1, act=0, scratch=scratch)
    torch.testing.assert_close(out.float(), torch.nn.functional.gelu(ref), rtol=2e-3, atol=2e-3)
    assert int(scratch[:1024].view(torch.int32).abs().sum()) == 0


def test_gemm_repeatable(pkg):
    g = torch.Generator("cuda").manual_seed(0)
    a = torch.randn(2056, 1024, device="cuda", generator=g).half()
    w = (torch.randn(4096, 1024, device="cuda", generator=g) * 0.05).half()
    o1 = pkg.gemm_f16(a, w)
    for _ in range(3):
        assert torch.equal(o1, pkg.gemm_f16(a, w))


@pytest.mark.parametrize("hw", [(480, 640), (680, 680), (128, 128), (96, 128)])
@pytest.mark.parametrize("ch", [3, 4])
def test_preprocess_matches_torch(pkg, hw, ch):
    cfg = pkg.tiny(pkg.CLIP_L14)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, 0), max_batch=2)
    rgb = torch.randint(0, 256, (2, hw[0], hw[1], ch), dtype=torch.uint8, generator=torch.Generator().manual_seed(1))
    _, pre = eng.forward_frames(rgb.cuda(), return_pre=True)
    ref = vit_ref.preprocess_frame(rgb[..., :3])
    torch.testing.assert_close(pre.cpu(), ref, rtol=0, atol=2e-6)


@pytest.mark.parametrize("base", ["clip-vit-l14", "dinov2-vitl14-reg"])
def test_tiny_vit_matches_oracle(pkg, base):
    cfg = pkg.tiny(pkg.CONFIGS[base], dim=128, depth=2, heads=2, mlp=256)
    w = pkg.random_weights(cfg, seed=3, perturb=True)
    eng = pkg.ViTEngine(cfg, w, max_batch=3)
    x = torch.randn(5, 3, 224, 224, generator=torch.Generator().manual_seed(2))
    got = eng.forward_features(x.cuda())["x_norm_patchtokens"].cpu()
    ref = vit_ref.vit_forward(x, w, cfg)
    assert got.shape == ref.shape == (5, 256, 128)
    assert _rel(got, ref) < 1e-3


@pytest.mark.parametrize("attn", ["0", "2"])
@pytest.mark.parametrize("base", ["clip-vit-l14", "dinov2-vitl14-reg", "clip-vit-b32"])
def test_full_vit_matches_oracle(pkg, base, attn, monkeypatch):
    monkeypatch.setenv("BSCNAV_ATTENTION", attn)      # read when the engine is created: 0 = mma.sync, 2 = tcgen05 warp-specialised
    cfg = pkg.CONFIGS[base]
    w = pkg.random_weights(cfg, seed=0)
    eng = pkg.ViTEngine(cfg, w, max_batch=4)
    rgb = torch.randint(0, 256, (3, 480, 640, 4), dtype=torch.uint8, generator=torch.Generator().manual_seed(5))
    got = eng.forward_frames(rgb.cuda()).cpu()
    torch.set_num_threads(max(1, torch.get_num_threads()))
    ref = vit_ref.vit_forward(vit_ref.preprocess_frame(rgb[..., :3]), w, cfg)
    rel = _rel(got, ref)
    print(base, "max per-token relative L2 error", rel)
    assert rel < 1e-3


@pytest.mark.parametrize("base", ["clip-vit-l14", "dinov2-vitl14-reg"])
def test_one_frame_per_call_matches_oracle(pkg, base):
    """The reference's call pattern (memory_2.py:732-742: one frame per forward) takes its own route through the engine -- blocks
    replayed from a CUDA graph, fc2 in K slices, the row-group-split attention kernel: full-size tokens of three frames passed one
    at a time (eager, captured, replayed) against the fp32 oracle at the bar, and next to the same frames as one batch."""
    cfg = pkg.CONFIGS[base]
    w = pkg.random_weights(cfg, seed=0)
    eng = pkg.ViTEngine(cfg, w, max_batch=4)
    rgb = torch.randint(0, 256, (3, 480, 640, 4), dtype=torch.uint8, generator=torch.Generator().manual_seed(6))
    one = torch.cat([eng.forward_frames(rgb[i:i + 1].cuda()) for i in range(3)]).cpu()
    batch = eng.forward_frames(rgb.cuda()).cpu()
    ref = vit_ref.vit_forward(vit_ref.preprocess_frame(rgb[..., :3]), w, cfg)
    rel1, relb = _rel(one, ref), _rel(batch, ref)
    print(base, "per-token relative L2 error, one frame per call", rel1, "batched", relb, "between them", _rel(one, batch))
    assert rel1 < 1e-3 and relb < 1e-3
    assert not eng.range_overflow()


def test_gaussian_pool(pkg):
    t = torch.randn(3, 256, 1024, generator=torch.Generator().manual_seed(0))
    got = pkg.gaussian_pool(t.cuda()).cpu()
    torch.testing.assert_close(got, vit_ref.gaussian_pool(t), rtol=1e-5, atol=1e-6)


@pytest.mark.parametrize("T,heads,batch", [(257, 16, 3), (261, 16, 2), (50, 12, 4), (128, 2, 2), (272, 4, 2), (200, 3, 1), (17, 1, 1),
                                           (129, 2, 2), (144, 3, 2), (145, 2, 1), (256, 2, 1)])
@pytest.mark.parametrize("impl", [0, 1, 2, 3])
def test_attention_kernels(pkg, T, heads, batch, impl):
    g = torch.Generator("cuda").manual_seed(T + heads)
    D = heads * 64
    qkv = (torch.randn(batch * T, 3 * D, device="cuda", generator=g) * 0.8).half()
    out = pkg.attention_f16(qkv, batch, T, heads, impl=impl).float().view(batch, T, heads, 64)
    q, k, v = qkv.float().view(batch, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    ref = torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v          # q arrives pre-scaled
    ref = ref.permute(0, 2, 1, 3)
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)


def test_attention_ws_many_items_matches_mma(pkg):
    """The persistent tcgen05 kernel over several waves of (frame, head) items (barrier phases wrap many times)."""
    B, T, heads = 23, 257, 16
    g = torch.Generator("cuda").manual_seed(9)
    qkv = (torch.randn(B * T, 3 * heads * 64, device="cuda", generator=g) * 0.8).half()
    a = pkg.attention_f16(qkv, B, T, heads, impl=0).float()
    for impl in (2, 3, 3, 3):
        b = pkg.attention_f16(qkv, B, T, heads, impl=impl).float()
        torch.testing.assert_close(b, a, rtol=2e-3, atol=2e-3)
    # DINOv2-reg length (5 tail rows shared by the two tail warps), several waves
    T = 261
    qkv = (torch.randn(B * T, 3 * heads * 64, device="cuda", generator=g) * 0.8).half()
    a = pkg.attention_f16(qkv, B, T, heads, impl=0).float()
    for _ in range(2):
        torch.testing.assert_close(pkg.attention_f16(qkv, B, T, heads, impl=3).float(), a, rtol=2e-3, atol=2e-3)


# ---- the agent's view-check encoder: open_clip ViT-H/14 image tower, heads of dimension 80 (BSCAgent.py:309, 700-731)
@pytest.mark.parametrize("T,heads,batch", [(257, 16, 2), (50, 3, 3), (17, 1, 1), (261, 2, 2)])
def test_attention_head_dim_80(pkg, T, heads, batch):
    g = torch.Generator("cuda").manual_seed(T + heads)
    D = heads * 80
    qkv = (torch.randn(batch * T, 3 * D, device="cuda", generator=g) * 0.8).half()
    out = pkg.attention_f16(qkv, batch, T, heads, impl=0, head_dim=80).float().view(batch, T, heads, 80)
    q, k, v = qkv.float().view(batch, T, 3, heads, 80).permute(2, 0, 3, 1, 4)
    ref = (torch.softmax(q @ k.transpose(-1, -2), dim=-1) @ v).permute(0, 2, 1, 3)
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)


def test_tiny_h14_image_embedding_matches_oracle(pkg):
    import dataclasses
    cfg = dataclasses.replace(pkg.tiny(pkg.OPENCLIP_H14, dim=640, depth=2, heads=8, mlp=1280), embed=96)
    w = pkg.random_weights(cfg, seed=8, perturb=True)
    eng = pkg.ViTEngine(cfg, w, max_batch=3)
    x = torch.randn(5, 3, 224, 224, generator=torch.Generator().manual_seed(3))
    got = eng.encode_image(x.cuda()).cpu()
    ref = vit_ref.encode_image(x, w, cfg)
    assert got.shape == ref.shape == (5, 96)
    assert _rel(got, ref) < 1e-3
    # the patch-token output of the same engine (head dim 80 through forward_features)
    tok = eng.forward_features(x.cuda())["x_norm_patchtokens"].cpu()
    assert _rel(tok, vit_ref.vit_forward(x, w, cfg)) < 1e-3


def test_full_h14_image_embedding_and_view_scores(pkg):
    cfg = pkg.OPENCLIP_H14
    w = pkg.random_weights(cfg, seed=0)
    eng = pkg.ViTEngine(cfg, w, max_batch=4)
    x = torch.randn(3, 3, 224, 224, generator=torch.Generator().manual_seed(4))
    got = eng.encode_image(x.cuda())
    ref = vit_ref.encode_image(x, w, cfg)
    rel = _rel(got.cpu(), ref)
    print("ViT-H/14 image embedding, max relative L2 error", rel)
    assert got.shape == (3, 1024) and rel < 1e-3
    # check_around: the views against an image goal (BSCAgent.py:721-731)
    checker = pkg.ViewChecker(eng, turn_left=30)
    best, prob = checker.best_view(x.cuda(), x[2])
    ref_prob, ref_best = vit_ref.score_views(ref, ref[2])
    assert best == int(ref_best) == 2
    torch.testing.assert_close(prob.cpu(), ref_prob, rtol=1e-3, atol=1e-5)
    assert checker.turn_actions(2) == ["turn_left"] * 2 and checker.turn_actions(9) == ["turn_right"] * 3
    assert checker.turn_actions(14) == ["turn_left"] * 2          # second lap of views (BSCAgent.py:736-737)


@pytest.mark.parametrize("n,E", [(12, 1024), (24, 1024), (1, 16), (37, 100)])
def test_score_views_matches_torch(pkg, n, E):
    g = torch.Generator().manual_seed(n + E)
    emb, goal = torch.randn(n, E, generator=g), torch.randn(E, generator=g)
    emb[n // 2] = goal * 3 + 0.2 * torch.randn(E, generator=g)
    prob, best = pkg.score_views(emb.cuda(), goal.cuda())
    ref_prob, ref_best = vit_ref.score_views(emb, goal)
    assert int(best) == int(ref_best) == n // 2
    torch.testing.assert_close(prob.cpu(), ref_prob, rtol=1e-5, atol=1e-7)


# ---- the same CLIP's text tower (encode_text: goal text, long-term labels; BSCAgent.py:541-557, 708-711)
@pytest.mark.parametrize("impl", [0, 3])
@pytest.mark.parametrize("T,heads,batch", [(77, 16, 3), (20, 2, 2), (130, 3, 2), (64, 1, 1), (257, 2, 1), (261, 2, 2), (200, 1, 1)])
def test_attention_causal(pkg, T, heads, batch, impl):
    """Causal mask (text tower): the mma.sync kernel and the default tcgen05 kernel (tile rows, halves without a visible key, tail rows)."""
    g = torch.Generator("cuda").manual_seed(T * 3 + heads)
    D = heads * 64
    qkv = (torch.randn(batch * T, 3 * D, device="cuda", generator=g) * 0.8).half()
    out = pkg.attention_f16(qkv, batch, T, heads, impl=impl, causal=True).float().view(batch, T, heads, 64)
    q, k, v = qkv.float().view(batch, T, 3, heads, 64).permute(2, 0, 3, 1, 4)
    mask = torch.full((T, T), float("-inf"), device="cuda").triu(1)
    ref = (torch.softmax(q @ k.transpose(-1, -2) + mask, dim=-1) @ v).permute(0, 2, 1, 3)
    torch.testing.assert_close(out, ref, rtol=2e-3, atol=2e-3)


def _token_rows(cfg, lengths, seed):
    g = torch.Generator().manual_seed(seed)
    ids = torch.zeros(len(lengths), cfg.text_context, dtype=torch.long)
    for r, n in enumerate(lengths):
        n = min(n, cfg.text_context - 2)
        ids[r, 0] = cfg.vocab - 2
        ids[r, 1:1 + n] = torch.randint(1, cfg.vocab - 2, (n,), generator=g)
        ids[r, 1 + n] = cfg.vocab - 1
    return ids


def test_tiny_text_tower_matches_oracle(pkg):
    import dataclasses
    cfg = dataclasses.replace(pkg.tiny(pkg.OPENCLIP_H14_TEXT, dim=128, depth=2, heads=2, mlp=256), embed=64, vocab=1000)
    w = pkg.random_weights(cfg, seed=9, perturb=True)
    eng = pkg.ViTEngine(cfg, w, max_batch=2)
    ids = _token_rows(cfg, [3, 75, 1, 40, 12], seed=1)          # more rows than max_batch: chunked
    got = eng.encode_text(ids.cuda()).cpu()
    ref = vit_ref.encode_text(ids, w, cfg)
    assert got.shape == ref.shape == (5, 64)
    assert _rel(got, ref) < 1e-3
    with pytest.raises(Exception):
        eng.forward_features(torch.zeros(1, 3, 224, 224))      # a text engine has no image path


def test_full_text_tower_and_label_matching(pkg):
    cfg = pkg.OPENCLIP_H14_TEXT
    w = pkg.random_weights(cfg, seed=0)
    eng = pkg.ViTEngine(cfg, w, max_batch=8)
    ids = _token_rows(cfg, [4, 9, 2, 30, 6, 4], seed=2)
    ids[5] = ids[0]                                             # the prompt repeats label 0
    got = eng.encode_text(ids.cuda())
    ref = vit_ref.encode_text(ids, w, cfg)
    rel = _rel(got.cpu(), ref)
    print("CLIP text tower (24 x 1024, context 77), max relative L2 error", rel)
    assert got.shape == (6, 1024) and rel < 1e-3
    # BSCAgent.py:548-557: the prompt against the stored labels
    assert pkg.match_long_memory_label(got[5], got[:5]) == 0
    a, b = torch.nn.functional.normalize(ref[5:6], dim=-1), torch.nn.functional.normalize(ref[:5], dim=-1)
    assert int(torch.argmax((a @ b.T).squeeze(0))) == 0


def test_graph_replay_equals_eager(pkg):
    """The captured forward (CUDA graph on static buffers) of the small CLIP calls returns exactly what the eager launches do,
    for new inputs on every replay."""
    import dataclasses
    tcfg = dataclasses.replace(pkg.tiny(pkg.OPENCLIP_H14_TEXT, dim=128, depth=2, heads=2, mlp=256), embed=64, vocab=1000)
    txt = pkg.ViTEngine(tcfg, pkg.random_weights(tcfg, seed=9, perturb=True), max_batch=4)
    icfg = dataclasses.replace(pkg.tiny(pkg.OPENCLIP_H14, dim=640, depth=2, heads=8, mlp=1280), embed=96)
    img = pkg.ViTEngine(icfg, pkg.random_weights(icfg, seed=8, perturb=True), max_batch=2)
    for rep in range(3):
        ids = _token_rows(tcfg, [3 + rep, 40, 7], seed=rep).cuda()
        assert torch.equal(txt.encode_text(ids, graph=True), txt.encode_text(ids))
        x = torch.randn(1, 3, 224, 224, generator=torch.Generator().manual_seed(rep)).cuda()
        assert torch.equal(img.encode_image(x, graph=True), img.encode_image(x))
    assert set(txt._graphs) == {("text", 3)} and set(img._graphs) == {("image", 1)}


@pytest.mark.parametrize("B", [1, 3])
def test_block_graph_equals_eager(pkg, B):
    """Small batches replay the transformer blocks from a CUDA graph held by the engine (vit_engine.cu run_blocks): first call eager,
    second captured, later ones replayed.  Every call must return exactly what the eager launches return (the profiling hook
    forces those), for new frames on every call, also after a different batch size has been through the engine."""
    cfg = pkg.tiny(pkg.CLIP_L14, dim=256, depth=3, heads=4, mlp=512)
    eng = pkg.ViTEngine(cfg, pkg.random_weights(cfg, seed=4, perturb=True), max_batch=4)
    g = torch.Generator().manual_seed(B)
    for rep in range(5):
        nb = B if rep != 3 else B + 1                     # a second batch size in between: its own graph, same workspace
        rgb = torch.randint(0, 256, (nb, 96, 128, 3), dtype=torch.uint8, generator=g).cuda()
        got = eng.forward_frames(rgb).clone()
        eng.profile(True)
        want = eng.forward_frames(rgb).clone()
        eng.profile(False)
        assert torch.equal(got, want), (B, rep)
        assert not eng.range_overflow()
