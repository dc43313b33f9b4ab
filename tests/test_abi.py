"""CPU: the C-ABI library loads and exports every symbol include/bscnav_b200.h declares (no compute calls)."""
import ctypes
import importlib
import os
import re

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    import __graft_entry__ as ge
    ge.build()
    _lib = importlib.import_module("bsc-nav_b200._lib")
    header = open(os.path.join(ROOT, "include", "bscnav_b200.h")).read()
    declared = set(re.findall(r"\b(bscnav_[a-z0-9_]+)\s*\(", header))
    assert declared, "no declarations parsed"
    lib = ctypes.CDLL(_lib.LIB_PATH)
    for name in sorted(declared):
        assert hasattr(lib, name), f"{name} declared in the header but not exported"
    assert declared == set(_lib.EXPORTS), declared ^ set(_lib.EXPORTS)
    assert lib.bscnav_abi_version() == 3


def test_struct_sizes_match_header():
    _lib = importlib.import_module("bsc-nav_b200._lib")
    assert ctypes.sizeof(_lib.Geom) == 12 * 4 + 3 * 8 + 27 * 8
    assert ctypes.sizeof(_lib.Store) == 9 * 8
    assert ctypes.sizeof(_lib.Filter) == 8 + 12 + 8 + 4
    assert ctypes.sizeof(_lib.VitCfg) == 14 * 4


def test_weight_blob_layout_matches_engine():
    pkg = importlib.import_module("bsc-nav_b200")
    vit = importlib.import_module("bsc-nav_b200.vit")
    _lib = importlib.import_module("bsc-nav_b200._lib")
    for cfg in (pkg.tiny(pkg.CLIP_L14), pkg.tiny(pkg.DINOV2_L14_REG), pkg.tiny(pkg.CLIP_B32, dim=128, depth=1, heads=2, mlp=256)):
        hb, fb = vit.pack_weights(cfg, pkg.random_weights(cfg, 0))
        c = vit._cfg_struct(cfg)
        assert hb.numel() == _lib.lib().bscnav_vit_half_elems(ctypes.byref(c))
        assert fb.numel() == _lib.lib().bscnav_vit_float_elems(ctypes.byref(c))
