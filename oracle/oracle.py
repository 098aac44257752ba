"""TEST INFRASTRUCTURE ONLY -- numpy/ctypes front end of oracle/gh_oracle.c (Oracle-B).

A CPU restatement of the reference rasterizer used to CHECK the CUDA product.  Only tests/,
__graft_entry__.smoke() and bench.py's baseline leg may import this; the product package
(gaussianhaircut_b200/) never does.

    build()                         gcc -> oracle/lib/libgh_oracle.so
    forward(inputs) -> dict         K1..K6 of the reference (radii, keys, point list, ranges, image, ...)
    backward(inputs, fwd, dL) -> dict   K7..K9 (the nine gradient tensors of the reference binding)
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
SRC = os.path.join(HERE, "gh_oracle.c")
LIB = os.path.join(HERE, "lib", "libgh_oracle.so")
NUM_CHANNELS = 10
_lib = None


def build(force: bool = False) -> str:
    if not force and os.path.isfile(LIB) and os.path.getmtime(LIB) >= os.path.getmtime(SRC):
        return LIB
    os.makedirs(os.path.dirname(LIB), exist_ok=True)
    # -ffp-contract=off: only the fmaf() calls written in the source may fuse
    cmd = ["gcc", "-O2", "-ffp-contract=off", "-fno-fast-math", "-fopenmp", "-shared", "-fPIC", "-o", LIB, SRC, "-lm"]
    subprocess.run(cmd, check=True)
    return LIB


def _load():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(LIB)
        _lib.gho_preprocess.restype = C.c_int
        _lib.gho_binning.restype = C.c_int
    return _lib


def _p(a):
    return None if a is None else a.ctypes.data_as(C.c_void_p)


def _f32(a):
    if a is None:
        return None
    a = np.ascontiguousarray(np.asarray(a, dtype=np.float32))
    return a if a.size else None


def forward(means3D, opacities, colors, viewmatrix, projmatrix, tanfovx, tanfovy, W, H, bg,
            scales=None, rotations=None, cov3D_precomp=None, conic_precomp=None, scale_modifier=1.0):
    """Reference forward on the CPU.  Arrays are numpy float32; optional ones None/empty when absent."""
    lib = _load()
    means3D, opacities, colors = _f32(means3D), _f32(opacities), _f32(colors)
    scales, rotations = _f32(scales), _f32(rotations)
    cov3D_precomp, conic_precomp = _f32(cov3D_precomp), _f32(conic_precomp)
    vm, pm, bg = _f32(viewmatrix).reshape(-1), _f32(projmatrix).reshape(-1), _f32(bg)
    if colors is None:
        raise RuntimeError("For non-RGB, provide precomputed Gaussian colors!")   # rasterizer_impl.cu:244-247
    P = means3D.shape[0]
    gx, gy = (W + 15) // 16, (H + 15) // 16
    T = gx * gy
    radii = np.zeros(P, np.int32); depths = np.zeros(P, np.float32); means2D = np.zeros((P, 2), np.float32)
    conic_opacity = np.zeros((P, 4), np.float32); cov3D = np.zeros((P, 6), np.float32)
    tiles = np.zeros(P, np.uint32)
    culled = lib.gho_preprocess(
        C.c_int(P), C.c_int(W), C.c_int(H), _p(means3D), _p(opacities), _p(scales), C.c_float(scale_modifier),
        _p(rotations), _p(cov3D_precomp), _p(conic_precomp), _p(vm), _p(pm), C.c_float(tanfovx), C.c_float(tanfovy),
        _p(radii), _p(depths), _p(means2D), _p(conic_opacity), _p(cov3D), _p(tiles))
    R = int(tiles.astype(np.int64).sum())
    keys = np.zeros(R, np.uint64); plist = np.zeros(R, np.uint32); ranges = np.zeros((T, 2), np.uint32)
    rc = lib.gho_binning(C.c_int(P), C.c_int(W), C.c_int(H), _p(radii), _p(depths), _p(means2D), _p(tiles),
                         C.c_longlong(R), _p(keys), _p(plist), _p(ranges))
    if rc != 0:
        raise RuntimeError(f"gho_binning failed: {rc}")
    out = np.zeros((NUM_CHANNELS, H, W), np.float32); final_T = np.zeros(H * W, np.float32)
    n_contrib = np.zeros(H * W, np.uint32)
    n_eval, n_con = C.c_longlong(0), C.c_longlong(0)
    lib.gho_render_forward(C.c_int(W), C.c_int(H), _p(ranges), _p(plist), _p(means2D), _p(conic_opacity), _p(colors),
                           _p(bg), _p(out), _p(final_T), _p(n_contrib), C.byref(n_eval), C.byref(n_con))
    return {"num_rendered": R, "out_color": out, "radii": radii, "depths": depths, "means2D": means2D,
            "conic_opacity": conic_opacity, "cov3D": cov3D, "tiles_touched": tiles, "keys": keys,
            "point_list": plist, "ranges": ranges, "final_T": final_T, "n_contrib": n_contrib,
            "culled": int(culled), "pairs_evaluated": n_eval.value, "pairs_contributing": n_con.value}


def backward(fwd, dL_dout, means3D, colors, viewmatrix, projmatrix, tanfovx, tanfovy, W, H, bg,
             scales=None, rotations=None, cov3D_precomp=None, conic_precomp=None, scale_modifier=1.0):
    """Reference backward on the CPU; returns the 9 tensors in the reference binding's order/shapes
    (rasterize_points.cu:160-168,205)."""
    lib = _load()
    means3D, colors, dL = _f32(means3D), _f32(colors), _f32(dL_dout)
    scales, rotations = _f32(scales), _f32(rotations)
    cov3D_precomp, conic_precomp = _f32(cov3D_precomp), _f32(conic_precomp)
    vm, pm, bg = _f32(viewmatrix).reshape(-1), _f32(projmatrix).reshape(-1), _f32(bg)
    P = means3D.shape[0]
    g = {"dL_dmeans2D": np.zeros((P, 3), np.float32), "dL_dcolors": np.zeros((P, NUM_CHANNELS), np.float32),
         "dL_dopacity": np.zeros((P, 1), np.float32), "dL_dmeans3D": np.zeros((P, 3), np.float32),
         "dL_dcov3D": np.zeros((P, 6), np.float32), "dL_dconic": np.zeros((P, 2, 2), np.float32),
         "dL_dsh": np.zeros((P, 0, 3), np.float32), "dL_dscales": np.zeros((P, 3), np.float32),
         "dL_drotations": np.zeros((P, 4), np.float32)}
    lib.gho_render_backward(C.c_int(P), C.c_int(W), C.c_int(H), _p(fwd["ranges"]), _p(fwd["point_list"]),
                            _p(fwd["means2D"]), _p(fwd["conic_opacity"]), _p(colors), _p(bg), _p(fwd["final_T"]),
                            _p(fwd["n_contrib"]), _p(dL), _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]),
                            _p(g["dL_dopacity"]), _p(g["dL_dcolors"]))
    lib.gho_preprocess_backward(C.c_int(P), C.c_int(W), C.c_int(H), _p(means3D), _p(fwd["radii"]), _p(scales),
                                C.c_float(scale_modifier), _p(rotations), _p(cov3D_precomp), _p(conic_precomp),
                                _p(vm), _p(pm), C.c_float(tanfovx), C.c_float(tanfovy),
                                _p(g["dL_dmeans2D"]), _p(g["dL_dconic"]), _p(g["dL_dmeans3D"]), _p(g["dL_dcov3D"]),
                                _p(g["dL_dscales"]), _p(g["dL_drotations"]))
    return g


def mark_visible(means3D, viewmatrix):
    lib = _load()
    means3D, vm = _f32(means3D), _f32(viewmatrix).reshape(-1)
    out = np.zeros(means3D.shape[0], np.uint8)
    lib.gho_mark_visible(C.c_int(means3D.shape[0]), _p(means3D), _p(vm), _p(out))
    return out.astype(bool)
