"""ctypes binding of libgh_raster.so (C ABI: include/gh_rasterizer.h).

Plain pointers and sizes only -- torch is used by the callers for device memory and streams, never
here.  The library must have been built (python -m gaussianhaircut_b200.build); a missing library is
a hard error, there is no fallback path.
"""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "lib", "libgh_raster.so")

GH_OK = 0
GH_E_INVALID_ARG = 1
GH_E_NO_COLORS = 2
GH_E_CUDA = 3
GH_E_PREFILTERED = 4

ABI_VERSION = 3

_p = C.c_void_p
_i = C.c_int
_f = C.c_float
_ll = C.c_longlong
_sz = C.c_size_t

# name -> (restype, argtypes); every symbol declared in include/gh_rasterizer.h
SIGNATURES = {
    "gh_abi_version": (_i, []),
    "gh_num_channels": (_i, []),
    "gh_last_error": (C.c_char_p, []),
    "gh_kernel_launch_count": (C.c_ulonglong, []),
    "gh_stage_timing_enable": (None, [_i]),
    "gh_stage_timing_read": (_i, [C.POINTER(C.c_double), C.POINTER(C.c_ulonglong), _i]),
    "gh_forward_workspace_sizes": (_i, [_i, _i, _i, C.POINTER(_sz), C.POINTER(_sz)]),
    "gh_binning_workspace_size": (_i, [_ll, C.POINTER(_sz)]),
    "gh_forward_preprocess": (_i, [
        _i, _i, _i, _i, _i,                  # P D M width height
        _p, _p, _p, _p, _p,                  # means3D means2D_precomp shs colors_precomp opacities
        _p, _f, _p,                          # scales scale_modifier rotations
        _p, _p,                              # cov3D_precomp conic_precomp
        _p, _p, _p,                          # viewmatrix projmatrix cam_pos
        _f, _f, _i,                          # tan_fovx tan_fovy prefiltered
        _p,                                  # radii
        _p, _p,                              # geom_buffer img_buffer
        C.POINTER(_i), C.POINTER(_i),        # num_rendered max_tile_len
        _i, _p]),                            # debug stream
    "gh_forward_render": (_i, [
        _i, _i, _i, _p, _p, _p, _p, _p, _p, _i, _i, _p, _i, _p]),
    "gh_backward": (_i, [
        _i, _i, _i, _i, _i, _i,              # P D M R width height
        _p,                                  # background
        _p, _p, _p,                          # means3D shs colors_precomp
        _p, _f, _p,                          # scales scale_modifier rotations
        _p, _p,                              # cov3D_precomp conic_precomp
        _p, _p, _p,                          # viewmatrix projmatrix campos
        _f, _f,                              # tan_fovx tan_fovy
        _p,                                  # radii
        _p, _p, _p,                          # geom binning img
        _p,                                  # dL_dpix
        _p, _p, _p, _p,                      # dL_dmean2D dL_dconic dL_dopacity dL_dcolor
        _p, _p, _p, _p, _p,                  # dL_dmean3D dL_dcov3D dL_dsh dL_dscale dL_drot
        _i, _p]),                            # debug stream
    "gh_mark_visible": (_i, [_i, _p, _p, _p, _p, _p]),
    "gh_adam_step": (_i, [_i, _p, _p, _p, _p, _p, _p, _f, _f, _f, _i, _p, _p, _p, _p]),
    "gh_image_loss_workspace_size": (_i, [_i, _i, C.POINTER(C.c_size_t)]),
    "gh_allreduce_p2p": (_i, [_p, _p, C.c_ulonglong, _i, _i, C.c_size_t, C.c_size_t, C.c_uint, _p, _p, _p]),
    "gh_image_loss": (_i, [_i, _i, _p, _p, _p, _p, _p, _f, _f, _f, _f, _p, _p, _p, _p]),
    "gh_project_workspace_size": (_i, [_i, C.POINTER(C.c_size_t)]),
    "gh_project_forward": (_i, [
        _i, _i, _i,                          # P width height
        _p, _p, _p, _p,                      # xyz scaling rotation dirs
        _p, _p,                              # features_dc features_rest
        _p, _p, _p,                          # opacity label orient_conf
        _p, _p, _p,                          # viewmatrix projmatrix campos
        _f, _f, _f, _i, C.c_uint, _f,        # tan_fovx tan_fovy scale_modifier sh_degree flags det_eps
        _p, _p, _p, _p, _p, _p,              # means2D colors opacities conic cov3D visible
        _p]),                                # stream
    "gh_project_forward_binned": (_i, [
        _i, _i, _i,
        _p, _p, _p, _p,
        _p, _p,
        _p, _p, _p,
        _p, _p, _p,
        _f, _f, _f, _i, C.c_uint, _f,
        _p, _p, _p, _p, _p, _p,              # means2D colors opacities conic cov3D visible
        _p, _p, _p, C.POINTER(C.c_int), C.POINTER(C.c_int),   # radii geom_buffer img_buffer num_rendered max_tile_len
        _p]),
    "gh_project_backward": (_i, [
        _i, _i, _i,
        _p, _p, _p, _p,
        _p, _p,
        _p, _p, _p,
        _p, _p, _p,
        _f, _f, _f, _i, C.c_uint, _f,
        _p,                                  # visible
        _p,                                  # geom_buffer
        _p, _p, _p, _p,                      # dL_dmeans2D dL_dconic dL_dcolors dL_dopacity
        _p, _p, _p, _p, _p, _p,              # d_xyz d_scaling d_rotation d_dirs d_features_dc d_features_rest
        _p, _p, _p, _p, _p,                  # d_opacity d_label d_orient_conf d_means2D d_camera
        _p, _p, _p]),                        # nan_flag workspace stream
    "gh_densify_classify": (_i, [_i, _p, _p, _p, _p, _f, _f, _f, _f, _p, _p]),
    "gh_densify_scatter": (_i, [_i, _i, _p, _p, _p, _p, _p, _p, _p, _i, _i, _i, _p, _p, _i, _i, _i, _p, _i, _p]),
    "gh_debug_export": (_i, [_i, _i, _i, _ll, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p, _p]),
}

_lib = None


class GhError(RuntimeError):
    """Raised for any non-zero status of the native library (the reference raises RuntimeError from
    AT_ERROR / std::runtime_error at the same places)."""

    def __init__(self, code: int, message: str):
        super().__init__(message)
        self.code = code


def load() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.isfile(LIB_PATH):
        raise RuntimeError(
            f"gaussianhaircut_b200: native library not found at {LIB_PATH}. "
            "Build it with `python -m gaussianhaircut_b200.build` (needs nvcc; there is no CPU fallback).")
    lib = C.CDLL(LIB_PATH)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError if the symbol is missing: fail loudly
        fn.restype = res
        fn.argtypes = args
    if lib.gh_abi_version() != ABI_VERSION:
        raise RuntimeError(f"libgh_raster.so ABI {lib.gh_abi_version()} != binding ABI {ABI_VERSION}; rebuild")
    _lib = lib
    return lib


STAGE_NAMES = ("preprocess", "tile_scan", "emit", "tile_sort", "blend_forward", "blend_backward",
               "preprocess_backward")


def stage_timing_read():
    """{stage: (total_ms, calls)} accumulated since gh_stage_timing_enable(1)."""
    lib = load()
    n = len(STAGE_NAMES)
    ms = (C.c_double * n)()
    calls = (C.c_ulonglong * n)()
    lib.gh_stage_timing_read(ms, calls, n)
    return {STAGE_NAMES[i]: (ms[i], int(calls[i])) for i in range(n)}


def check(status: int) -> None:
    if status != GH_OK:
        msg = load().gh_last_error().decode("utf-8", "replace")
        raise GhError(status, msg or f"libgh_raster error {status}")
