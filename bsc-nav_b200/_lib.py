"""ctypes binding of libbscnav_b200.so (the C ABI declared in include/bscnav_b200.h).

There is no CPU fallback: if the library is missing or a call fails this raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(HERE, "libbscnav_b200.so")

OK, EINVAL, ECAPACITY, ECUDA, EUNSUPPORTED = 0, -1, -2, -3, -4
_ERR = {EINVAL: "invalid argument", ECAPACITY: "capacity exceeded", ECUDA: "CUDA error", EUNSUPPORTED: "unsupported device"}
KEY_EMPTY = 0xFFFFFFFFFFFFFFFF


class BscnavError(RuntimeError):
    pass


class Geom(C.Structure):
    _fields_ = [
        ("height", C.c_int32), ("width", C.c_int32), ("gs", C.c_int32), ("minh", C.c_int32), ("maxh", C.c_int32),
        ("levels", C.c_int32), ("n_patch_h", C.c_int32), ("n_patch_w", C.c_int32), ("cache", C.c_int32),
        ("dim", C.c_int32), ("max_slots", C.c_int32), ("reserved", C.c_int32),
        ("cs", C.c_double), ("min_depth", C.c_double), ("max_depth", C.c_double),
        ("kinv", C.c_double * 9), ("kimg", C.c_double * 9), ("kpatch", C.c_double * 9),
    ]


class Store(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("occupied_ids", "first_touch", "slot_cell", "slot_keys", "row_key", "tok_h",
                                          "norm2", "feat", "counters")]


class Colour(C.Structure):
    _fields_ = [(n, C.c_void_p) for n in ("grid_rgb", "weight", "cv_key", "cv_map")]


class Filter(C.Structure):
    _fields_ = [("radius2", C.c_int64), ("cur", C.c_int32 * 3), ("floor_min", C.c_int32), ("floor_max", C.c_int32),
                ("reserved", C.c_int32)]


class VitCfg(C.Structure):
    _fields_ = [("image", C.c_int32), ("patch", C.c_int32), ("dim", C.c_int32), ("depth", C.c_int32), ("heads", C.c_int32),
                ("mlp", C.c_int32), ("n_reg", C.c_int32), ("act", C.c_int32), ("ln_pre", C.c_int32),
                ("patch_k_pad", C.c_int32), ("ln_eps", C.c_float), ("head_dim", C.c_int32), ("text_context", C.c_int32),
                ("precision", C.c_int32)]


_P = C.c_void_p
_SIGS = {
    "bscnav_abi_version": (C.c_int, []),
    "bscnav_last_error": (C.c_char_p, []),
    "bscnav_launch_count": (C.c_int64, []),
    "bscnav_store_reset": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), _P]),
    "bscnav_store_publish_counts": (C.c_int, [C.POINTER(Store), _P, _P]),
    "bscnav_store_row_norms": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), C.c_int64, _P]),
    "bscnav_fuse_scratch_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "bscnav_fuse_frames": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), _P, _P, _P, _P, C.c_int32, C.c_int32, _P, _P,
                                     C.c_int64, _P, _P]),
    "bscnav_query_scratch_bytes": (C.c_int64, [C.POINTER(Geom), C.c_int32, C.c_int32]),
    "bscnav_query_topk": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), _P, C.c_int32, C.c_int32, C.POINTER(Filter), _P,
                                    C.c_int64, _P, _P, _P, _P]),
    "bscnav_shadow_bytes": (C.c_int64, [C.POINTER(Geom)]),
    "bscnav_store_build_shadow": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), _P, _P]),
    "bscnav_query_batch_scratch_bytes": (C.c_int64, [C.POINTER(Geom), C.c_int32, C.c_int32, C.c_int64]),
    "bscnav_query_topk_batch": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), _P, _P, C.c_int32, C.c_int32, C.POINTER(Filter),
                                          C.c_int64, _P, C.c_int64, _P, _P, _P, _P]),
    "bscnav_cluster_centers": (C.c_int, [_P, _P, _P, C.c_int32, C.c_int32, C.c_double, C.c_int32, C.c_int32, _P, _P, _P, _P, _P, _P]),
    "bscnav_colour_reset": (C.c_int, [C.POINTER(Geom), C.POINTER(Colour), _P]),
    "bscnav_colour_scratch_bytes": (C.c_int64, [C.POINTER(Geom), C.c_int32, C.c_int32]),
    "bscnav_colour_frames": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), C.POINTER(Colour), _P, _P, C.c_int32, _P, _P, _P, C.c_int32,
                                       C.c_int32, _P, C.c_int64, _P]),
    "bscnav_merge_delta": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), _P, _P, _P, _P, C.c_int32, _P, C.c_int64, _P]),
    "bscnav_export_scratch_bytes": (C.c_int64, [C.POINTER(Geom)]),
    "bscnav_export_delta": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), C.c_uint32, C.c_uint32, _P, _P, _P, _P, C.c_int32,
                                      _P, _P, C.c_int64, _P]),
    "bscnav_packed_row_bytes": (C.c_int64, [C.POINTER(Geom)]),
    "bscnav_export_packed": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), C.c_uint32, C.c_uint32, _P, C.c_int32, _P, _P, C.c_int64, _P]),
    "bscnav_merge_packed_scratch_bytes": (C.c_int64, [C.c_int32, C.c_int32]),
    "bscnav_merge_packed": (C.c_int, [C.POINTER(Geom), C.POINTER(Store), _P, C.c_int32, C.c_int32, _P, C.c_int32, C.c_int32, _P,
                                      C.c_int64, _P]),
    "bscnav_vit_half_elems": (C.c_int64, [C.POINTER(VitCfg)]),
    "bscnav_vit_float_elems": (C.c_int64, [C.POINTER(VitCfg)]),
    "bscnav_vit_workspace_bytes": (C.c_int64, [C.POINTER(VitCfg), C.c_int32]),
    "bscnav_vit_create": (C.c_int, [C.POINTER(VitCfg), _P, _P, _P, C.c_int64, C.c_int32, C.POINTER(_P)]),
    "bscnav_vit_destroy": (None, [_P]),
    "bscnav_vit_range_status": (C.c_int, [_P, C.c_int32, C.POINTER(C.c_int32), _P]),
    "bscnav_vit_profile": (C.c_int, [_P, C.c_int32]),
    "bscnav_vit_profile_read": (C.c_int, [_P, _P, _P, _P]),
    "bscnav_vit_forward_frames": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, _P, _P]),
    "bscnav_vit_forward_images": (C.c_int, [_P, _P, C.c_int32, _P, _P]),
    "bscnav_vit_encode_images": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, _P, _P]),
    "bscnav_score_views": (C.c_int, [_P, _P, C.c_int32, C.c_int32, _P, _P, _P]),
    "bscnav_attention_f16_hd": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "bscnav_text_encode": (C.c_int, [_P, _P, C.c_int32, _P, C.c_int32, _P, C.c_int32, _P, _P]),
    "bscnav_gaussian_pool": (C.c_int, [_P, C.c_int32, C.c_int32, C.c_int32, _P, _P]),
    "bscnav_attention_f16": (C.c_int, [_P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "bscnav_gemm_bf16x3": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, _P]),
    "bscnav_debug_attention_ws2": (C.c_int, [C.c_int32, _P]),
    "bscnav_gemm_f16": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P]),
    "bscnav_gemm_f16_sliced": (C.c_int, [_P, _P, _P, _P, C.c_int32, C.c_int32, C.c_int32, C.c_int32, C.c_int32, _P, C.c_int64,
                                          C.POINTER(C.c_int32), _P]),
}
EXPORTS = tuple(_SIGS)

_lib = None


def lib():
    """Load the shared library (once).  Raises if it has not been built: there is no fallback path."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise BscnavError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                              "(bscnav_b200 has no CPU or PyTorch fallback)")
        l = C.CDLL(LIB_PATH)
        for name, (res, args) in _SIGS.items():
            fn = getattr(l, name)
            fn.restype, fn.argtypes = res, args
        _lib = l
    return _lib


def check(status: int, what: str = ""):
    if status >= 0:
        return status
    detail = lib().bscnav_last_error().decode() if status == ECUDA else ""
    raise BscnavError(f"{what}: {_ERR.get(status, status)} {detail}".strip())


def ptr(t):
    """Device pointer of a torch tensor (or None)."""
    return None if t is None else C.c_void_p(t.data_ptr())


def stream_ptr():
    import torch
    return C.c_void_p(torch.cuda.current_stream().cuda_stream)
