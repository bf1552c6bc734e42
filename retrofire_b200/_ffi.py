"""ctypes view of include/retrofire_b200.h and loader of librf_b200.so.

The shared library is the product: if it is missing or a symbol is absent this module
raises — there is no Python/CPU fallback for the hot path.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("RF_B200_LIB") or os.path.join(HERE, "librf_b200.so")   # RF_B200_LIB: a tuning build (build.build_variant)

RF_MAX_ATTR_LANES = 8
RF_VS_UNIFORM_F32 = 32
RF_FS_UNIFORM_F32 = 8

# rf_status
RF_OK, RF_E_INVALID, RF_E_INDEX_OOB, RF_E_TARGET_OOB, RF_E_BAD_TEXTURE = 0, 1, 2, 3, 4
RF_E_UNSUPPORTED_SHADER, RF_E_CUDA, RF_E_NCCL, RF_E_NOMEM, RF_E_UNSUPPORTED = 5, 6, 7, 8, 9
STATUS_NAMES = {
    0: "RF_OK", 1: "RF_E_INVALID", 2: "RF_E_INDEX_OOB", 3: "RF_E_TARGET_OOB", 4: "RF_E_BAD_TEXTURE",
    5: "RF_E_UNSUPPORTED_SHADER", 6: "RF_E_CUDA", 7: "RF_E_NCCL", 8: "RF_E_NOMEM", 9: "RF_E_UNSUPPORTED",
}

# rf_vs_id / rf_fs_id
VS_MVP, VS_MVP_LINEARIZE, VS_SOLIDS, VS_SPRITE = 0, 1, 2, 3
(FS_COLOR3F, FS_COLOR3F_SRGB, FS_COLOR4F, FS_CHECKER, FS_TEX_CLAMP_LIT, FS_TEX_CLAMP,
 FS_TEX_REPEAT_POT, FS_SPRITE_DISC, FS_NORMAL_VIS, FS_TEX_ONCE) = range(10)

# rf_color_fmt
FMT_RGBA8888, FMT_XRGB8888, FMT_ARGB8888, FMT_BGRA8888, FMT_RGB888, FMT_RGB565, FMT_RGBA4444 = range(7)
FMT_HOST_BYTES = {0: 4, 1: 4, 2: 4, 3: 4, 4: 3, 5: 2, 6: 2}
TEXEL_RGB888, TEXEL_RGBA8888 = 0, 1

PRIM_TRIS, PRIM_EDGES = 0, 1
CULL_NONE, CULL_BACK, CULL_FRONT = 0, 1, 2
DEPTH_NONE, DEPTH_LESS, DEPTH_EQUAL, DEPTH_GREATER = 0, 1, 2, 3


class RfDraw(C.Structure):
    _fields_ = [
        ("indices", C.POINTER(C.c_uint32)),
        ("n_prims", C.c_uint32),
        ("verts", C.POINTER(C.c_float)),
        ("n_verts", C.c_uint32),
        ("vert_stride_f32", C.c_uint32),
        ("mesh", C.c_void_p),
        ("n_attr_lanes", C.c_uint32),
        ("persp_mask", C.c_uint32),
        ("vs", C.c_uint32),
        ("fs", C.c_uint32),
        ("vs_uniform", C.c_float * RF_VS_UNIFORM_F32),
        ("fs_uniform", C.c_float * RF_FS_UNIFORM_F32),
        ("texture", C.c_void_p),
        ("viewport", C.c_float * 16),
        ("face_cull", C.c_uint8),
        ("depth_test", C.c_uint8),
        ("color_write", C.c_uint8),
        ("depth_write", C.c_uint8),
        ("depth_sort", C.c_uint8),
        ("prim_kind", C.c_uint8),
        ("bbox_cull", C.c_uint8),
        ("_pad", C.c_uint8 * 1),
        ("bbox", C.c_float * 6),
    ]


class RfStats(C.Structure):
    _fields_ = [
        ("calls", C.c_uint64),
        ("prims_i", C.c_uint64), ("prims_o", C.c_uint64),
        ("verts_i", C.c_uint64), ("verts_o", C.c_uint64),
        ("frags_i", C.c_uint64), ("frags_o", C.c_uint64),
        ("time_ns", C.c_uint64),
        ("objs_i", C.c_uint64), ("objs_o", C.c_uint64),
    ]


# every symbol include/retrofire_b200.h declares: name -> (restype, argtypes)
_P = C.c_void_p
SYMBOLS = {
    "rf_abi_version": (C.c_uint32, []),
    "rf_ctx_create": (C.c_int, [C.c_int, _P, C.POINTER(_P)]),
    "rf_ctx_destroy": (None, [_P]),
    "rf_last_error": (C.c_char_p, [_P]),
    "rf_ctx_set_row_band": (C.c_int, [_P, C.c_uint32, C.c_uint32]),
    "rf_ctx_set_geometry_async": (C.c_int, [_P, C.c_int]),
    "rf_target_create": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, C.c_int, C.POINTER(_P)]),
    "rf_target_destroy": (None, [_P]),
    "rf_target_clear": (C.c_int, [_P, _P, C.POINTER(C.c_uint8), C.POINTER(C.c_float)]),
    "rf_target_upload_color": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rf_target_download_color": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rf_target_upload_depth": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rf_target_download_depth": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rf_host_alloc": (C.c_int, [C.c_size_t, C.POINTER(_P)]),
    "rf_host_free": (None, [_P]),
    "rf_target_download_color_async": (C.c_int, [_P, _P, _P, C.c_size_t]),
    "rf_target_color_devptr": (_P, [_P]),
    "rf_target_depth_devptr": (_P, [_P]),
    "rf_texture_create": (C.c_int, [_P, C.c_uint32, C.c_uint32, C.c_uint32, _P, C.c_size_t, C.POINTER(_P)]),
    "rf_texture_destroy": (None, [_P]),
    "rf_mesh_create": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P, C.c_uint32, C.c_uint32, C.POINTER(_P)]),
    "rf_mesh_destroy": (None, [_P]),
    "rf_render_many": (C.c_int, [_P, _P, _P, C.c_uint32]),
    "rf_ctx_peer_export": (C.c_int, [_P, _P, C.POINTER(_P)]),
    "rf_ctx_peer_attach": (C.c_int, [_P, C.c_uint32, C.c_uint32, _P, C.POINTER(_P)]),
    "rf_target_peer_export": (C.c_int, [_P, _P, _P, C.POINTER(_P)]),
    "rf_target_peer_attach": (C.c_int, [_P, _P, C.c_uint32, C.c_uint32, _P, C.POINTER(_P)]),
    "rf_ctx_replays": (C.c_int, [_P, C.POINTER(C.c_uint64)]),
    "rf_render": (C.c_int, [_P, _P, C.POINTER(RfDraw), C.POINTER(RfStats)]),
    "rf_render_frames": (C.c_int, [_P, C.POINTER(_P), C.c_uint32, C.POINTER(RfDraw), _P]),
    "rf_flush": (C.c_int, [_P]),
    "rf_sync": (C.c_int, [_P]),
    "rf_ctx_stats": (C.c_int, [_P, C.POINTER(RfStats), C.c_int]),
    "rf_ctx_last_pass": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint32)]),
    "rf_ctx_profile": (C.c_int, [_P, C.c_int]),
    "rf_ctx_kernel_times": (C.c_int, [_P, C.POINTER(C.c_uint64), C.POINTER(C.c_uint64)]),
    "rf_kernel_name": (C.c_char_p, [C.c_uint32]),
}
RF_N_KERNELS = 12

_lib = None


def load() -> C.CDLL:
    """Load librf_b200.so and bind every declared symbol. Raises if anything is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'` "
            "(retrofire_b200 has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)  # AttributeError if the .so does not export it
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
