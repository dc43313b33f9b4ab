"""bscnav_b200 -- B200-native spatial-memory build + query engine behind BSC-Nav's VoxelTokenMemory API.

Host code is Python; all arithmetic on the hot path runs in libbscnav_b200.so (hand-written sm_100a CUDA,
C ABI in include/bscnav_b200.h).  There is no CPU fallback.
"""
from .vit_config import CLIP_B32, CLIP_L14, CONFIGS, DINOV2_L14_REG, OPENCLIP_H14, OPENCLIP_H14_TEXT, ViTConfig, random_weights, tiny  # noqa: F401


def __getattr__(name):
    # device-touching modules are imported lazily so that `import` works on a box without the built library
    import importlib
    lazy = {
        "SpatialStore": "store", "StoreConfig": "store", "ViTEngine": "vit", "pack_weights": "vit", "gaussian_pool": "vit",
        "gemm_f16": "vit", "gemm_f16_sliced": "vit", "gemm_bf16x3": "vit", "load_state_dict": "weights", "load_checkpoint": "weights", "interpolate_pos_embed": "weights", "attention_f16": "vit", "score_views": "vit", "ViewChecker": "agent", "match_long_memory_label": "agent", "VoxelTokenMemory": "memory",
        "Memory": "memory", "GESObjectNavRobot": "agent", "LocalizeAgent": "agent", "PoseChain": "geometry",
    }
    if name in lazy:
        return getattr(importlib.import_module(f".{lazy[name]}", __name__), name)
    raise AttributeError(name)
