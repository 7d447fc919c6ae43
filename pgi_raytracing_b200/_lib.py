"""ctypes loader of ``libpgrt_b200.so`` (the C ABI declared in ``include/pgrt.h``).

The library is the product: there is no Python or CPU fallback.  ``load()`` raises when the shared object is
missing; creating a context raises when no GPU is present.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("PGRT_LIB") or os.path.join(_HERE, "libpgrt_b200.so")   # PGRT_LIB: A/B builds of the same ABI (tools/)
CSRC = os.path.join(_HERE, "csrc")

PGRT_OK, PGRT_ERR_INVALID, PGRT_ERR_CUDA, PGRT_ERR_NO_DEVICE, PGRT_ERR_OVERFLOW = 0, 1, 2, 3, 4
INVALID_ID = 0xFFFFFFFF
MAX_INFLIGHT = 32


class Material(C.Structure):
    _fields_ = [("diffuse", C.c_float * 3), ("specular", C.c_float * 3), ("shininess", C.c_float), ("ior", C.c_float),
                ("type", C.c_int32), ("diffuse_tex", C.c_int32)]


class Light(C.Structure):
    _fields_ = [("position", C.c_float * 3), ("ambient", C.c_float * 3), ("diffuse", C.c_float * 3), ("specular", C.c_float * 3)]


class RenderParams(C.Structure):
    _fields_ = [("sampling_width", C.c_int32), ("jitter", C.c_int32), ("focal_distance", C.c_float), ("aperture", C.c_float),
                ("max_depth", C.c_int32), ("gamma_level", C.c_float), ("seed", C.c_uint32), ("camera_mode", C.c_int32),
                ("shader_mode", C.c_int32), ("scheduler", C.c_int32), ("shadow_mode", C.c_int32), ("reserved", C.c_int32 * 5)]


class BuildStats(C.Structure):
    _fields_ = [("triangles", C.c_uint32), ("nodes", C.c_uint32), ("build_ms", C.c_float), ("sort_ms", C.c_float),
                ("sah_cost", C.c_float), ("tree_ms", C.c_float), ("collapse_ms", C.c_float), ("depth", C.c_uint32),
                ("ploc_passes", C.c_uint32), ("node_bytes", C.c_uint32), ("reserved", C.c_uint32 * 2)]


class RenderStats(C.Structure):
    _fields_ = [("rays_primary", C.c_uint64), ("rays_shadow", C.c_uint64), ("rays_reflection", C.c_uint64), ("rays_refraction", C.c_uint64),
                ("nodes_visited", C.c_uint64), ("tris_tested", C.c_uint64), ("frame_ms", C.c_float), ("trace_ms", C.c_float), ("shade_ms", C.c_float), ("trace_launches", C.c_uint32),
                ("launches", C.c_uint32), ("batches", C.c_uint32), ("overflow_retries", C.c_uint32), ("max_nodes_per_ray", C.c_uint32),
                ("reserved", C.c_uint32 * 4)]


class LevelStats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("shadow_rays", C.c_uint64), ("nodes", C.c_uint64), ("tris", C.c_uint64),
                ("shadow_nodes", C.c_uint64), ("shadow_tris", C.c_uint64), ("max_nodes", C.c_uint32), ("shadow_max_nodes", C.c_uint32),
                ("trace_ms", C.c_float), ("shade_ms", C.c_float)]


# every symbol include/pgrt.h declares: name -> (restype, argtypes)
_VP, _I32, _U32, _U64, _F = C.c_void_p, C.c_int32, C.c_uint32, C.c_uint64, C.c_float
SYMBOLS = {
    "pgrt_version": (C.c_char_p, []),
    "pgrt_create": (C.c_int, [C.POINTER(_VP), C.c_int]),
    "pgrt_destroy": (None, [_VP]),
    "pgrt_last_error": (C.c_char_p, [_VP]),
    "pgrt_set_stream": (C.c_int, [_VP, _VP]),
    "pgrt_add_mesh": (C.c_int, [_VP, _VP, _VP, _VP, _U32, _I32, C.POINTER(_U32)]),
    "pgrt_set_materials": (C.c_int, [_VP, C.POINTER(Material), _I32]),
    "pgrt_set_texture": (C.c_int, [_VP, _I32, _VP, _I32, _I32, _I32, _I32]),
    "pgrt_set_envmap": (C.c_int, [_VP, _VP, _I32, _I32, _I32, _I32]),
    "pgrt_set_lights": (C.c_int, [_VP, C.POINTER(Light), _I32]),
    "pgrt_commit": (C.c_int, [_VP, C.POINTER(BuildStats)]),
    "pgrt_clear_scene": (C.c_int, [_VP]),
    "pgrt_set_camera": (C.c_int, [_VP, _I32, _I32, _F, C.POINTER(_F), C.POINTER(_F)]),
    "pgrt_default_params": (None, [C.POINTER(RenderParams)]),
    "pgrt_render": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.POINTER(RenderStats), _I32]),
    "pgrt_render_device": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.POINTER(RenderStats), _I32]),
    "pgrt_render_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _I32, _I32]),
    "pgrt_render_device_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _I32, _I32]),
    "pgrt_render_shard_device_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _I32, _I32]),
    "pgrt_render_accumulate": (C.c_int, [_VP, C.POINTER(RenderParams), C.c_int32, _VP, C.POINTER(RenderStats)]),
    "pgrt_frame_alloc": (C.c_int, [_VP, _U64, C.POINTER(_VP)]),
    "pgrt_frame_free": (C.c_int, [_VP, _VP]),
    "pgrt_frame_export": (C.c_int, [_VP, _VP, C.c_char_p]),
    "pgrt_frame_import": (C.c_int, [_VP, C.c_char_p, C.POINTER(_VP)]),
    "pgrt_frame_unmap": (C.c_int, [_VP, _VP]),
    "pgrt_host_frame_register": (C.c_int, [_VP, _VP, _U64, C.POINTER(_VP)]),
    "pgrt_host_frame_unregister": (C.c_int, [_VP, _VP]),
    "pgrt_debug_flush_l2": (C.c_int, [_VP, _I32, _U64, _U32]),
    "pgrt_debug_l2_bandwidth": (C.c_int, [_VP, _U64, _I32, C.POINTER(_F)]),
    "pgrt_debug_frame_cycles": (C.c_int, [_VP, _I32, C.POINTER(_U64)]),
    "pgrt_enable_peer_access": (C.c_int, [_VP, _I32]),
    "pgrt_render_shard_to_frame_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _I32, _I32]),
    "pgrt_render_rgba8": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.POINTER(RenderStats), _I32]),
    "pgrt_render_rgba8_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _I32, _I32]),
    "pgrt_render_shard_to_frame_rgba8_begin": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _I32, _I32]),
    "pgrt_slot_signal": (C.c_int, [_VP, _I32, _VP, _U32]),
    "pgrt_slot_signal_add": (C.c_int, [_VP, _I32, _VP]),
    "pgrt_stream_wait_value32": (C.c_int, [_VP, _VP, _VP, _U32]),
    "pgrt_stream_write_value32": (C.c_int, [_VP, _VP, _VP, _U32]),
    "pgrt_render_end": (C.c_int, [_VP, _I32, C.POINTER(RenderStats)]),
    "pgrt_slot_stream": (_VP, [_VP, _I32]),
    "pgrt_stream_wait_slot": (C.c_int, [_VP, _I32, _VP]),
    "pgrt_get_pixel": (C.c_int, [_VP, C.POINTER(RenderParams), _I32, _I32, C.POINTER(_F)]),
    "pgrt_primary_ids": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _VP]),
    "pgrt_set_shard": (C.c_int, [_VP, _I32, _I32]),
    "pgrt_shard_pixels": (_U64, [_VP]),
    "pgrt_render_shard_device": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, C.POINTER(RenderStats), _I32]),
    "pgrt_untile": (C.c_int, [_VP, _VP, _I32, _VP]),
    "pgrt_untile_on_stream": (C.c_int, [_VP, _VP, _I32, _VP, _VP]),
    "pgrt_intersect": (C.c_int, [_VP, _VP, _U64]),
    "pgrt_interpolate": (C.c_int, [_VP, _VP, _VP, _VP, _VP, _U64, _I32, _VP]),
    "pgrt_trace": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _U64, _I32, _VP]),
    "pgrt_is_illuminated": (C.c_int, [_VP, C.POINTER(RenderParams), _VP, _VP, _VP, _U64, _VP]),
    "pgrt_eval_mix_srgb": (C.c_int, [_VP, _VP, _VP, _VP, _U64, _VP]),
    "pgrt_eval_texture": (C.c_int, [_VP, _I32, _VP, _U64, _VP]),
    "pgrt_eval_envmap": (C.c_int, [_VP, _VP, _U64, _VP]),
    "pgrt_eval_gamma": (C.c_int, [_VP, _VP, _F, _U64, _VP]),
    "pgrt_eval_primary_rays": (C.c_int, [_VP, C.POINTER(RenderParams), _VP]),
    "pgrt_eval_secondary_rays": (C.c_int, [_VP, _VP, _U64, _I32, _VP]),
    "pgrt_num_triangles": (_U32, [_VP]),
    "pgrt_num_geometries": (_U32, [_VP]),
    "pgrt_last_level_stats": (C.c_int, [_VP, _I32, C.POINTER(LevelStats)]),
    "pgrt_kernel_launches": (_U64, [_VP]),
}

_lib = None


def build(force: bool = False) -> str:
    """Compile the CUDA library in-tree for sm_100a (nvcc cross-compiles without a GPU)."""
    subprocess.check_call(["make", "-C", CSRC, "-s"] + (["-B"] if force else []), stdout=subprocess.DEVNULL)
    return LIB_PATH


def load():
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise RuntimeError(f"{LIB_PATH} is missing: run `python -c 'import __graft_entry__ as g; g.build()'` "
                           "(the render loop has no CPU fallback)")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SYMBOLS.items():
        fn = getattr(lib, name)   # AttributeError here = header / library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib
