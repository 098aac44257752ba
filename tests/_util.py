"""Shared helpers for the parity tests (test infrastructure)."""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

if os.path.join(ROOT, "oracle") not in sys.path:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402  (oracle/synth.py: seeded scenes, cameras, the caller-preamble restatement)


def ref_module():
    """Oracle-A: the reference extension built in place under oracle/_ref (GPU only)."""
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    return build_ref.load()


def ref_available() -> bool:
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import build_ref
    return build_ref.is_built()


def _align(off: int, a: int = 128) -> int:
    return (off + a - 1) // a * a


def parse_ref_buffers(P: int, W: int, H: int, R: int, geom: torch.Tensor, binning: torch.Tensor, img: torch.Tensor):
    """Unpack the reference's three opaque byte buffers with its `obtain` arithmetic
    (rasterizer_impl.h:21-27, rasterizer_impl.cu:155-194).  Only the arrays in front of the
    CUB-sized temp regions are addressable without knowing CUB's temp size -- enough for the gates."""
    out = {}
    g = geom.cpu().numpy()
    base = geom.data_ptr()
    off = _align(base) - base
    out["depths"] = g[off:off + 4 * P].view(np.float32).copy(); off += 4 * P
    off = _align(base + off) - base; off += 3 * P                       # clamped bool[3P]
    off = _align(base + off) - base; off += 4 * P                       # internal_radii
    off = _align(base + off) - base
    out["means2D"] = g[off:off + 8 * P].view(np.float32).reshape(P, 2).copy(); off += 8 * P
    off = _align(base + off) - base; off += 24 * P                      # cov3D
    off = _align(base + off) - base
    out["conic_opacity"] = g[off:off + 16 * P].view(np.float32).reshape(P, 4).copy(); off += 16 * P
    off = _align(base + off) - base; off += 12 * P                      # rgb 3P
    off = _align(base + off) - base
    out["tiles_touched"] = g[off:off + 4 * P].view(np.uint32).copy()

    b = binning.cpu().numpy()
    base = binning.data_ptr()
    off = _align(base) - base
    out["point_list"] = b[off:off + 4 * R].view(np.uint32).copy(); off += 4 * R
    off = _align(base + off) - base; off += 4 * R                       # point_list_unsorted
    off = _align(base + off) - base
    out["keys"] = b[off:off + 8 * R].view(np.uint64).copy(); off += 8 * R

    i = img.cpu().numpy()
    base = img.data_ptr()
    N = W * H
    off = _align(base) - base
    out["final_T"] = i[off:off + 4 * N].view(np.float32).copy(); off += 4 * N
    off = _align(base + off) - base
    out["n_contrib"] = i[off:off + 4 * N].view(np.uint32).copy(); off += 4 * N
    off = _align(base + off) - base
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out["ranges"] = i[off:off + 8 * T].view(np.uint32).reshape(T, 2).copy()
    return out


def settings_tuple(mod, s: dict):
    return mod.GaussianRasterizationSettings(
        image_height=s["image_height"], image_width=s["image_width"], tanfovx=s["tanfovx"], tanfovy=s["tanfovy"],
        bg=s["bg"], scale_modifier=s["scale_modifier"], viewmatrix=s["viewmatrix"], projmatrix=s["projmatrix"],
        sh_degree=s["sh_degree"], campos=s["campos"], prefiltered=s["prefiltered"], debug=s["debug"])


def native_args(inp: dict, empty_device=None):
    """The 21 positional args of `_C.rasterize_gaussians` (reference __init__.py:63-85)."""
    kw, s = inp["kwargs"], inp["settings"]
    e = torch.Tensor([])
    g = lambda k: e if kw[k] is None else kw[k]  # noqa: E731
    return (s["bg"], kw["means3D"], kw["means2D"], g("colors_precomp"), kw["opacities"], g("scales"), g("rotations"),
            s["scale_modifier"], g("cov3D_precomp"), g("conic_precomp"), s["viewmatrix"], s["projmatrix"],
            s["tanfovx"], s["tanfovy"], s["image_height"], s["image_width"], e, s["sh_degree"], s["campos"],
            s["prefiltered"], s["debug"])


def backward_args(inp: dict, radii, dL, geom, R, binning, img):
    kw, s = inp["kwargs"], inp["settings"]
    e = torch.Tensor([])
    g = lambda k: e if kw[k] is None else kw[k]  # noqa: E731
    return (s["bg"], kw["means3D"], radii, g("colors_precomp"), g("scales"), g("rotations"), s["scale_modifier"],
            g("cov3D_precomp"), g("conic_precomp"), s["viewmatrix"], s["projmatrix"], s["tanfovx"], s["tanfovy"],
            dL, e, s["sh_degree"], s["campos"], geom, R, binning, img, s["debug"])


GRAD_NAMES = ("dL_dmeans2D", "dL_dcolors", "dL_dopacity", "dL_dmeans3D", "dL_dcov3D", "dL_dconic",
              "dL_dsh", "dL_dscales", "dL_drotations")


def rel_err(a: torch.Tensor, b: torch.Tensor) -> float:
    """norm-relative error ||a-b|| / ||b|| (0 if both are zero)."""
    a = a.double().flatten(); b = b.double().flatten()
    nb = b.norm().item()
    d = (a - b).norm().item()
    if nb == 0.0:
        return 0.0 if d == 0.0 else float("inf")
    return d / nb


def make_inputs(scene_kind: str, n: int, W: int, H: int, mode: str, cam_k: int = 0, seed: int = 0,
                opacity_mode: str = "random", device=None, **cam_kw):
    if scene_kind == "strands":
        scene = synth.make_strand_scene(n, seed=seed, opacity_mode=opacity_mode)
        strand_dir = True
    else:
        scene = synth.make_blob_scene(n, seed=seed)
        strand_dir = False
    cam = synth.make_camera(cam_k, W, H, **cam_kw)
    if not strand_dir:
        # blobs: dir feature is whatever; keep the same preamble code path
        pass
    inp = synth.rasterizer_inputs(scene, cam, mode=mode, device=device)
    return inp
