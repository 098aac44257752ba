"""Python front door: the exact public surface of the reference package
`diff_gaussian_rasterization` (ext/diff_gaussian_rasterization_hair/diff_gaussian_rasterization/
__init__.py) so that `src/gaussian_renderer` imports and calls it unchanged:

    GaussianRasterizationSettings   12-field NamedTuple            (__init__.py:170-182)
    GaussianRasterizer              nn.Module, .forward/.markVisible (__init__.py:184-236)
    rasterize_gaussians             functional entry                (__init__.py:21-44)

The autograd Function packs the same positional argument tuples the reference passes to its native
module (`__init__.py:63-85`, `:114-135`) and hands them to `gaussianhaircut_b200._C`, whose functions
mirror that native module on top of the C ABI.
"""
from __future__ import annotations

from typing import NamedTuple

import torch
import torch.nn as nn

from . import _C


class GaussianRasterizationSettings(NamedTuple):
    image_height: int
    image_width: int
    tanfovx: float
    tanfovy: float
    bg: torch.Tensor
    scale_modifier: float
    viewmatrix: torch.Tensor
    projmatrix: torch.Tensor
    sh_degree: int
    campos: torch.Tensor
    prefiltered: bool
    debug: bool


# ---- multi-GPU hook (no counterpart in the reference, which is single-GPU): when a gradient arena is installed, the
# backward writes every per-Gaussian gradient into that caller-owned flat float32 buffer (e.g. the symmetric-memory
# buffer of dist.PeerAllReduce) and the tensors autograd hands to `.grad` are views of it -- so ONE collective over
# the arena reduces everything the op produced, through the public API.  Process-wide on purpose: autograd runs the
# backward on its own thread, a thread-local would not be seen there.
_ARENA = {"storage": None, "last": None}


def set_gradient_arena(storage):
    """Install (or, with None, remove) the flat float32 buffer the next backward passes write their gradients into.
    It must hold at least `_C.arena_floats(P)` elements.  Returns the previous buffer."""
    prev = _ARENA["storage"]
    _ARENA["storage"] = storage
    _ARENA["last"] = None
    return prev


def last_gradient_arena():
    """(flat, views) of the most recent backward that ran with an installed arena, else None.  `flat[:_C.trainable_floats(P)]`
    is what a data-parallel caller all-reduces in the native call shape (dist.trainable_slice)."""
    if _ARENA["last"] is None:
        return None
    flat, P = _ARENA["last"]
    return _C.alloc_grad_arena(P, flat.device, zero=False, storage=flat)


def _cpu_snapshot(args):
    return tuple(a.detach().cpu().clone() if isinstance(a, torch.Tensor) else a for a in args)


class _RasterizeGaussians(torch.autograd.Function):
    @staticmethod
    def forward(ctx, means3D, means2D_precomp, sh, colors_precomp, opacities, scales, rotations,
                cov3Ds_precomp, conics_precomp, raster_settings):
        rs = raster_settings
        # positional layout of _C.rasterize_gaussians (reference __init__.py:63-85)
        args = (rs.bg, means3D, means2D_precomp, colors_precomp, opacities, scales, rotations,
                rs.scale_modifier, cov3Ds_precomp, conics_precomp, rs.viewmatrix, rs.projmatrix,
                rs.tanfovx, rs.tanfovy, rs.image_height, rs.image_width, sh, rs.sh_degree, rs.campos,
                rs.prefiltered, rs.debug)
        if rs.debug:
            snapshot = _cpu_snapshot(args)   # replayable fixture, same format as the reference's dump
            try:
                out = _C.rasterize_gaussians(*args)
            except Exception:
                torch.save(snapshot, "snapshot_fw.dump")
                print("\nAn error occured in forward. Please forward snapshot_fw.dump for debugging.")
                raise
        else:
            out = _C.rasterize_gaussians(*args)
        num_rendered, color, radii, geomBuffer, binningBuffer, imgBuffer = out
        ctx.raster_settings = rs
        ctx.num_rendered = num_rendered
        ctx.save_for_backward(colors_precomp, means3D, scales, rotations, cov3Ds_precomp, conics_precomp,
                              radii, sh, geomBuffer, binningBuffer, imgBuffer)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, grad_out_color, _grad_radii):
        rs = ctx.raster_settings
        (colors_precomp, means3D, scales, rotations, cov3Ds_precomp, conics_precomp,
         radii, sh, geomBuffer, binningBuffer, imgBuffer) = ctx.saved_tensors
        # positional layout of _C.rasterize_gaussians_backward (reference __init__.py:114-135)
        args = (rs.bg, means3D, radii, colors_precomp, scales, rotations, rs.scale_modifier,
                cov3Ds_precomp, conics_precomp, rs.viewmatrix, rs.projmatrix, rs.tanfovx, rs.tanfovy,
                grad_out_color, sh, rs.sh_degree, rs.campos, geomBuffer, ctx.num_rendered,
                binningBuffer, imgBuffer, rs.debug)
        if rs.debug:
            snapshot = _cpu_snapshot(args)
            try:
                grads = _C.rasterize_gaussians_backward(*args)
            except Exception:
                torch.save(snapshot, "snapshot_bw.dump")
                print("\nAn error occured in backward. Writing snapshot_bw.dump for debugging.\n")
                raise
        elif _ARENA["storage"] is not None:
            flat, v, g_sh0 = _C.rasterize_gaussians_backward_arena(*args, arena_storage=_ARENA["storage"])
            P = int(means3D.size(0))
            # no second reference to the views is kept: autograd then adopts them as `.grad` instead of cloning
            _ARENA["last"] = (flat, P)
            grads = (v["means2D"], v["colors"], v["opacity"], v["means3D"], v["cov3D"], v["conic"].view(P, 2, 2), g_sh0,
                     v["scales"], v["rotations"])
            del v
        else:
            grads = _C.rasterize_gaussians_backward(*args)
        (g_means2D, g_colors, g_opacities, g_means3D, g_cov3D, g_conic, g_sh, g_scales, g_rot) = grads
        # native side keeps half the off-diagonal derivative (backward.cu:554); the public gradient is
        # w.r.t. the 3-vector conic (a, b, c)  (reference __init__.py:149-153)
        g_conic3 = torch.stack([g_conic[:, 0, 0], 2 * g_conic[:, 0, 1], g_conic[:, 1, 1]], dim=-1)
        # gradient order = input order of forward(); raster_settings gets None
        return (g_means3D, g_means2D, g_sh, g_colors, g_opacities, g_scales, g_rot, g_cov3D, g_conic3, None)


def rasterize_gaussians(means3D, means2D_precomp, sh, colors_precomp, opacities, scales, rotations,
                        cov3Ds_precomp, conics_precomp, raster_settings):
    return _RasterizeGaussians.apply(means3D, means2D_precomp, sh, colors_precomp, opacities, scales,
                                     rotations, cov3Ds_precomp, conics_precomp, raster_settings)


class GaussianRasterizer(nn.Module):
    def __init__(self, raster_settings: GaussianRasterizationSettings):
        super().__init__()
        self.raster_settings = raster_settings

    def markVisible(self, positions):
        """Boolean mask of points passing the near-plane test of this camera (no grad)."""
        with torch.no_grad():
            rs = self.raster_settings
            return _C.mark_visible(positions, rs.viewmatrix, rs.projmatrix)

    def forward(self, means3D, means2D, opacities, shs=None, colors_precomp=None, scales=None,
                rotations=None, cov3D_precomp=None, conic_precomp=None):
        rs = self.raster_settings
        if (shs is None) == (colors_precomp is None):
            raise Exception('Please provide excatly one of either SHs or precomputed colors!')
        has_sr = (scales is not None) or (rotations is not None)
        if ((scales is None or rotations is None) and cov3D_precomp is None) or (has_sr and cov3D_precomp is not None):
            raise Exception('Please provide exactly one of either scale/rotation pair or precomputed 3D covariance!')
        empty = torch.Tensor([])   # "absent": empty CPU tensor -> NULL pointer at the C ABI
        shs = empty if shs is None else shs
        colors_precomp = empty if colors_precomp is None else colors_precomp
        scales = empty if scales is None else scales
        rotations = empty if rotations is None else rotations
        cov3D_precomp = empty if cov3D_precomp is None else cov3D_precomp
        conic_precomp = empty if conic_precomp is None else conic_precomp
        return rasterize_gaussians(means3D, means2D, shs, colors_precomp, opacities, scales, rotations,
                                   cov3D_precomp, conic_precomp, rs)
