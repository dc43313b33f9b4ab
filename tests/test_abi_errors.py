"""CPU: argument validation of the C ABI.  Every entry point checks its arguments before touching the device, so these calls
return BSCNAV_EINVAL / BSCNAV_EUNSUPPORTED without a GPU (error behaviour is part of the boundary: SURVEY 8b)."""
import ctypes as C
import importlib

import pytest

EINVAL, EUNSUPPORTED = -1, -4


@pytest.fixture(scope="module")
def env():
    import __graft_entry__ as ge
    ge.build()
    pkg = importlib.import_module("bsc-nav_b200")
    L = importlib.import_module("bsc-nav_b200._lib")
    store = importlib.import_module("bsc-nav_b200.store")
    return pkg, L, L.lib(), store


def _geom(store, **over):
    kw = dict(height=96, width=128, grid_size=64, token_dim=128, cache_size=2)
    kw.update(over)
    return store.StoreConfig(**kw).geom()


def test_geometry_validation(env):
    _pkg, L, lib, store = env
    ok = _geom(store)
    assert lib.bscnav_query_scratch_bytes(C.byref(ok), 4, 100) > 0
    assert lib.bscnav_shadow_bytes(C.byref(ok)) == 64 * 64 * 2 * 128 * 2
    for bad in (_geom(store, token_dim=100), _geom(store, cache_size=0), _geom(store, cache_size=65), _geom(store, grid_size=0)):
        assert lib.bscnav_query_scratch_bytes(C.byref(bad), 4, 100) == EINVAL
        assert lib.bscnav_shadow_bytes(C.byref(bad)) == EINVAL
    assert lib.bscnav_query_scratch_bytes(C.byref(ok), 0, 100) == EINVAL        # no queries
    assert lib.bscnav_query_scratch_bytes(C.byref(ok), 4, 0) == EINVAL          # K = 0
    assert lib.bscnav_query_scratch_bytes(C.byref(ok), 4, 1025) == EINVAL       # K above the supported maximum
    assert lib.bscnav_fuse_scratch_bytes(0, 100) == EINVAL and lib.bscnav_fuse_scratch_bytes(2, 0) == EINVAL
    assert lib.bscnav_query_batch_scratch_bytes(C.byref(ok), 0, 10, 0) == EINVAL
    assert lib.bscnav_colour_scratch_bytes(C.byref(ok), 0, 10) == EINVAL


def test_null_pointers_are_rejected(env):
    _pkg, L, lib, store = env
    g = _geom(store)
    st = L.Store()                                                               # all pointers null
    assert lib.bscnav_store_reset(C.byref(g), C.byref(st), None) == EINVAL
    assert lib.bscnav_fuse_frames(C.byref(g), C.byref(st), None, None, None, None, 1, 96 * 128, None, None, 0, None, None) == EINVAL
    assert lib.bscnav_query_topk(C.byref(g), C.byref(st), None, 1, 10, None, None, 0, None, None, None, None) == EINVAL
    assert lib.bscnav_query_topk_batch(C.byref(g), C.byref(st), None, None, 1, 10, None, 0, None, 0, None, None, None, None) == EINVAL
    assert lib.bscnav_merge_delta(C.byref(g), C.byref(st), None, None, None, None, 5, None, 0, None) == EINVAL
    assert lib.bscnav_store_build_shadow(C.byref(g), C.byref(st), None, None) == EINVAL
    assert lib.bscnav_cluster_centers(None, None, None, 1, 100, 10.0, 5, 8, None, None, None, None, None, None) == EINVAL
    assert lib.bscnav_gemm_f16(None, None, None, None, 128, 128, 64, 0, 0, None) == EINVAL
    assert lib.bscnav_attention_f16(None, None, 1, 257, 16, 2, None) == EINVAL
    assert lib.bscnav_gemm_f16_sliced(None, None, None, None, 257, 1024, 4096, 2, 0, None, 0, None, None) == EINVAL
    assert lib.bscnav_store_publish_counts(C.byref(st), None, None) == EINVAL      # null counters, null destination
    assert lib.bscnav_store_publish_counts(None, None, None) == EINVAL
    col = L.Colour()
    assert lib.bscnav_colour_reset(C.byref(g), C.byref(col), None) == EINVAL


def test_shape_limits(env):
    _pkg, L, lib, store = env
    buf = (C.c_char * 64)()
    p = C.cast(buf, C.c_void_p)
    # cluster centres: K above 1024, non-positive min_samples, negative eps
    for k, eps, ms in ((1025, 10.0, 5), (100, 10.0, 0), (100, -1.0, 5)):
        assert lib.bscnav_cluster_centers(p, p, p, 1, k, eps, ms, 8, p, p, p, p, p, None) == EINVAL
    # GEMM: N must be a multiple of 128, K of 64; unknown output mode
    assert lib.bscnav_gemm_f16(p, p, None, p, 128, 100, 64, 0, 0, None) == EINVAL
    assert lib.bscnav_gemm_f16(p, p, None, p, 128, 128, 60, 0, 0, None) == EINVAL
    assert lib.bscnav_gemm_f16(p, p, None, p, 128, 128, 64, 9, 0, None) == EINVAL
    # the sliced entry: scratch must hold the counters and at least two partial tiles; fp16-mode numbers only
    assert lib.bscnav_gemm_f16_sliced(p, p, None, p, 257, 1024, 4096, 2, 0, p, 1024, None, None) == EINVAL
    assert lib.bscnav_gemm_f16_sliced(p, p, None, p, 257, 1000, 4096, 2, 0, p, 1 << 24, None, None) == EINVAL
    assert lib.bscnav_gemm_f16_sliced(p, p, None, p, 257, 1024, 4096, 6, 0, p, 1 << 24, None, None) == EINVAL
    # dense fusion needs n_points == H * W when no sample index list is given
    g = _geom(store)
    st = L.Store(*[p] * 9)
    assert lib.bscnav_fuse_frames(C.byref(g), C.byref(st), p, p, p, None, 1, 17, p, p, 1 << 30, None, None) == EINVAL
    # colour channels are defined per voxel: a BEV-keyed store is refused
    gb = _geom(store, bev=True)
    col = L.Colour(*[p] * 4)
    assert lib.bscnav_colour_frames(C.byref(gb), C.byref(st), C.byref(col), p, p, 3, p, p, None, 1, 96 * 128, p, 1 << 30, None) == EUNSUPPORTED
    assert lib.bscnav_colour_frames(C.byref(g), C.byref(st), C.byref(col), p, p, 5, p, p, None, 1, 96 * 128, p, 1 << 30, None) == EINVAL


def test_product_path_has_no_cpu_fallback(env, monkeypatch):
    """The facade refuses to run without the CUDA library (no oracle / PyTorch fallback on the product path)."""
    _pkg, L, _lib, _store = env
    monkeypatch.setattr(L, "_lib", None)
    monkeypatch.setattr(L, "LIB_PATH", "/nonexistent/libbscnav_b200.so")
    with pytest.raises(L.BscnavError):
        L.lib()


def test_clip_tower_entry_points_validate(env):
    """The agent-side CLIP entries (view check / label matching): nulls, sizes, and the ViT config rules (head dim 64 | 80,
    text tower: head dim 64, no registers / pre-LN)."""
    pkg, L, lib, _store = env
    vit = importlib.import_module("bsc-nav_b200.vit")
    buf = (C.c_char * 64)()
    p = C.cast(buf, C.c_void_p)
    assert lib.bscnav_vit_encode_images(None, p, 1, p, 1024, p, None) == EINVAL
    assert lib.bscnav_text_encode(None, p, 1, p, 49408, p, 1024, p, None) == EINVAL
    assert lib.bscnav_score_views(None, p, 12, 1024, p, p, None) == EINVAL
    assert lib.bscnav_score_views(p, p, 0, 1024, p, p, None) == EINVAL
    assert lib.bscnav_score_views(p, p, 1025, 1024, p, p, None) == EINVAL           # more views than one block scores
    assert lib.bscnav_attention_f16_hd(None, None, 1, 77, 16, 64, 1, None) == EINVAL
    # layouts: image tower with heads of 80, text tower; sizes follow the documented blob layout
    for cfg in (pkg.OPENCLIP_H14, pkg.OPENCLIP_H14_TEXT, pkg.CLIP_L14):
        c = vit._cfg_struct(cfg)
        D, M, L_ = cfg.dim, cfg.mlp, cfg.depth
        per_layer_half = 3 * D * D + D * D + 2 * M * D
        per_layer_float = 2 * D + 3 * D + D + 2 * D + M + D
        if cfg.text_context:
            half, flt = L_ * per_layer_half, cfg.text_context * D + L_ * per_layer_float + 2 * D
        else:
            half = D * c.patch_k_pad + L_ * per_layer_half
            flt = D + D + cfg.n_patch * D + (2 * D if cfg.ln_pre else 0) + L_ * per_layer_float + 2 * D
        assert lib.bscnav_vit_half_elems(C.byref(c)) == half
        assert lib.bscnav_vit_float_elems(C.byref(c)) == flt
        assert lib.bscnav_vit_workspace_bytes(C.byref(c), 4) > 0
    bad = vit._cfg_struct(pkg.OPENCLIP_H14)
    bad.head_dim = 96                                                               # dim != heads * head_dim, and unsupported
    assert lib.bscnav_vit_half_elems(C.byref(bad)) == EINVAL
    bad = vit._cfg_struct(pkg.OPENCLIP_H14_TEXT)
    bad.ln_pre = 1                                                                  # the text tower has no pre-LN
    assert lib.bscnav_vit_float_elems(C.byref(bad)) == EINVAL
    bad = vit._cfg_struct(pkg.OPENCLIP_H14_TEXT)
    bad.text_context = -3
    assert lib.bscnav_vit_workspace_bytes(C.byref(bad), 4) == EINVAL
