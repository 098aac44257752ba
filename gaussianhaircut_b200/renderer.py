"""Drop-in for the reference's `gaussian_renderer` module: `render()` and `render_hair()` with the same
signatures and the same returned dictionary (src/gaussian_renderer/__init__.py:23-114, :116-236), but one fused
autograd node from the models' RAW parameters to the rasterized 10-channel image:

    reference                                                    here
    ---------------------------------------------------------    -------------------------------------------------
    ~100 PyTorch kernels of model-side projection + autograd      gh_project_forward / gh_project_backward (1 launch each)
    six boolean-mask gathers (+ scatter of the radii)             none: culled Gaussians keep their index, conic = 0
    GaussianRasterizer forward / backward                         the same C-ABI rasterizer calls (rasterizer.py / _C.py)
    (P,2,2) conic gradient restack + per-tensor gradient buffers  none: the projection backward reads the blend backward's
                                                                  accumulation records in place

A trainer switches by importing `render` / `render_hair` from here instead of from `gaussian_renderer`; everything
it reads back afterwards is the same: the four maps, `viewspace_points` (+ its `.grad` for the densification
statistics), `visibility_filter`, `radii`, and `.grad` of every model parameter and of trainable camera tensors.
tests/test_gpu_projection.py checks all of that against the reference's own functions imported unmodified.
"""
from __future__ import annotations

import math
from typing import Dict, Optional

import torch
import torch.nn.functional as F

from . import _C, projection
from ._alloc import empty_rows

__all__ = ["render", "render_hair", "render_raw", "set_nan_flag"]

_EMPTY = torch.Tensor([])

# The optimizer's NaN guard (src/train_gaussians.py:174-181) without a pass over the gradients: when a flag tensor is
# installed, the projection backward ORs 1 into it if any parameter gradient is NaN; hand the same tensor to
# FusedAdam.step(nan_flag_in=...).  The caller zeroes it once per iteration (FusedAdam does, after reading it).
_NAN_FLAG = {"t": None}


def set_nan_flag(flag):
    """Install (or remove with None) the device int32[1] tensor that receives the NaN verdict of the next backward."""
    _NAN_FLAG["t"] = flag


_TAN_CACHE: Dict[int, tuple] = {}


def _tan_half(fov) -> float:
    """tan(fov / 2) as a host float.  The reference reads it back from the device on every call (`.item()`,
    __init__.py:43-44); here the value is cached per FoV tensor (identity + in-place version), so a camera whose
    field of view does not change costs one device read in total."""
    if isinstance(fov, torch.Tensor):
        key = id(fov)
        hit = _TAN_CACHE.get(key)
        if hit is not None and hit[0]() is fov and hit[1] == fov._version and not fov.requires_grad:
            return hit[2]
        val = float(torch.tan(fov.detach() * 0.5).item())
        if not fov.requires_grad:
            import weakref
            if len(_TAN_CACHE) > 4096:
                _TAN_CACHE.clear()
            try:
                _TAN_CACHE[key] = (weakref.ref(fov), fov._version, val)
            except TypeError:
                pass
        return val
    return math.tan(float(fov) * 0.5)


class _FusedRender(torch.autograd.Function):
    """(xyz, scaling, rotation, dirs, f_dc, f_rest, opacity, label, conf, viewspace, viewmatrix, projmatrix, campos, tanfov,
    static) -> (raw 10-channel image, radii).  `viewspace` is a (P,3) leaf: its storage receives the NDC means and its
    gradient the rasterizer's dL/dmeans2D (what `add_densification_stats` reads).  `static`: dict of non-tensors + the
    frozen head block of render_hair()."""

    @staticmethod
    def forward(ctx, xyz, scaling, rotation, dirs, f_dc, f_rest, opacity, label, conf, viewspace, viewmatrix, projmatrix,
                campos, tanfov, static):
        st = static
        dev = xyz.device
        head = st.get("head")
        n_head = 0 if head is None else int(head["xyz"].shape[0])
        P_own = int(xyz.shape[0])
        P = n_head + P_own
        pi = projection.pack_inputs(xyz, scaling, rotation, dirs, f_dc, f_rest, opacity, label, conf, viewmatrix, projmatrix,
                                    campos, st["tanx"], st["tany"], st["W"], st["H"], st["sh_degree"], st["mod"], st["cfg"])
        bg = st["bg"]
        if n_head == 0 and P_own > 0 and not st["debug"]:
            # one pass over the Gaussians does the projection AND the rasterizer's preprocess + tile histogram
            out, radii, geom, img, R, max_len = projection.project_forward_binned(pi, means2D_out=viewspace.detach())
            color, binning = _C.forward_render(bg, out["colors"], radii, geom, img, R, max_len, st["H"], st["W"])
            ctx.pi, ctx.st, ctx.R, ctx.n_head, ctx.hp = pi, st, R, 0, None
            ctx.bufs = (pi.xyz, out["colors"], out["conic"], out["visible"], radii, geom, binning, img)
            ctx.mark_non_differentiable(radii)
            return color, radii
        if n_head == 0:
            out = projection.project_forward(pi, means2D_out=viewspace.detach())
            means3D = pi.xyz
        else:
            # the frozen head Gaussians first, then this model's: both write row blocks of the same buffers
            out = projection.alloc_outputs(P, dev, means2D_out=viewspace.detach())
            hp = projection.pack_inputs(head["xyz"], head["scaling"], head["rotation"], None, head["f_dc"], head["f_rest"],
                                        head["opacity"], None, None, viewmatrix, projmatrix, campos, st["tanx"], st["tany"],
                                        st["W"], st["H"], st["sh_degree"], st["mod"], projection.HEAD_PRECOMP)
            sl = lambda a, b: {k: (v[a:b] if v is not None else None) for k, v in out.items()}  # noqa: E731
            projection.project_forward(hp, out=sl(0, n_head))
            projection.project_forward(pi, out=sl(n_head, P))
            means3D = torch.cat([hp.xyz, pi.xyz], dim=0)
        R, color, radii, geom, binning, img = _C.rasterize_gaussians(
            bg, means3D, _EMPTY, out["colors"], out["opacity"], _EMPTY, _EMPTY, st["mod"], _EMPTY, out["conic"],
            pi.V, pi.Pm, st["tanx"], st["tany"], st["H"], st["W"], _EMPTY, st["sh_degree"], pi.campos,
            False, st["debug"])            # prefiltered=False: culled Gaussians are still in the list (conic = 0)
        ctx.pi, ctx.st, ctx.R, ctx.n_head = pi, st, R, n_head
        ctx.hp = hp if n_head > 0 else None
        ctx.bufs = (means3D, out["colors"], out["conic"], out["visible"], radii, geom, binning, img)
        ctx.mark_non_differentiable(radii)
        return color, radii

    @staticmethod
    def backward(ctx, g_color, _g_radii):
        pi, st, n_head = ctx.pi, ctx.st, ctx.n_head
        means3D, colors, conic, visible, radii, geom, binning, img = ctx.bufs
        need = ctx.needs_input_grad
        cam_grads = bool(need[10] or need[11] or need[12] or need[13])
        if n_head == 0:
            _C.rasterize_gaussians_backward_records(st["bg"], means3D, radii, colors, conic, pi.V, pi.Pm, st["tanx"], st["tany"],
                                                    g_color, pi.campos, geom, ctx.R, binning, img, st["debug"])
            g = projection.project_backward(pi, visible, geom_buffer=geom, camera_grads=cam_grads, want_means2D_grad=True,
                                            nan_flag=_NAN_FLAG["t"])
            g_view = g["means2D"]
        else:
            g9 = _C.rasterize_gaussians_backward(st["bg"], means3D, radii, colors, _EMPTY, _EMPTY, st["mod"], _EMPTY, conic,
                                                 pi.V, pi.Pm, st["tanx"], st["tany"], g_color, _EMPTY, st["sh_degree"],
                                                 pi.campos, geom, ctx.R, binning, img, st["debug"])
            g_m2, g_col, g_op, _g3, _gc3, g_conic = g9[0], g9[1], g9[2], g9[3], g9[4], g9[5]
            g = projection.project_backward(pi, visible[n_head:], dL_dmeans2D=g_m2[n_head:], dL_dconic4=g_conic[n_head:],
                                            dL_dcolors=g_col[n_head:], dL_dopacity=g_op[n_head:], camera_grads=cam_grads,
                                            nan_flag=_NAN_FLAG["t"])
            g_view = g_m2                  # `viewspace_points.grad` covers head rows too (retain_grad on the concatenation, :134-137)
            if cam_grads:
                # the head block is frozen, but its conic / depth / view direction still depend on the camera; only its
                # screen-space mean is detached (:132), so that path contributes nothing
                gh = projection.project_backward(ctx.hp, visible[:n_head], dL_dmeans2D=torch.zeros_like(g_m2[:n_head]), dL_dconic4=g_conic[:n_head],
                                                 dL_dcolors=g_col[:n_head], dL_dopacity=g_op[:n_head], camera_grads=True)
                for k in ("viewmatrix", "projmatrix", "campos", "tanfov"):
                    g[k] = g[k] + gh[k]
        pick = lambda k, idx: g.get(k) if need[idx] else None  # noqa: E731   (autograd rejects gradients for non-Variable inputs)
        return (pick("xyz", 0), pick("scaling", 1), pick("rotation", 2), pick("dirs", 3), pick("f_dc", 4), pick("f_rest", 5),
                pick("opacity", 6), pick("label", 7), pick("conf", 8), g_view if need[9] else None,
                pick("viewmatrix", 10), pick("projmatrix", 11), pick("campos", 12), pick("tanfov", 13), None)


def _post(renders: torch.Tensor, radii: torch.Tensor, viewspace: torch.Tensor) -> Dict[str, torch.Tensor]:
    """The reference's epilogue (gaussian_renderer/__init__.py:98-113)."""
    rendered_image, rendered_mask, rendered_cov2D, rendered_orient_conf, _ = renders.split([3, 2, 3, 1, 1], dim=0)
    rendered_dir2D = F.normalize(rendered_cov2D[:2], dim=0)
    to_mirror = torch.ones_like(rendered_dir2D[[0]])
    to_mirror[rendered_dir2D[[0]] < 0] *= -1
    rendered_orient_angle = torch.acos(rendered_dir2D[[1]].clamp(-1 + 1e-3, 1 - 1e-3) * to_mirror) / math.pi
    return {"render": rendered_image, "mask": rendered_mask, "orient_angle": rendered_orient_angle,
            "orient_conf": rendered_orient_conf, "viewspace_points": viewspace, "visibility_filter": radii > 0,
            "radii": radii, "raw": renders}


def _static(viewpoint_camera, bg_color, scaling_modifier, sh_degree, debug, cfg) -> Dict[str, object]:
    return {"W": int(viewpoint_camera.image_width), "H": int(viewpoint_camera.image_height),
            "tanx": _tan_half(viewpoint_camera.FoVx), "tany": _tan_half(viewpoint_camera.FoVy),
            "bg": bg_color, "mod": float(scaling_modifier), "sh_degree": int(sh_degree), "debug": bool(debug), "cfg": cfg}


def _tanfov_tensor(viewpoint_camera):
    """A differentiable (2,) tensor when the field of view is trainable (cameras.py:95-107), else None."""
    fx, fy = viewpoint_camera.FoVx, viewpoint_camera.FoVy
    if isinstance(fx, torch.Tensor) and isinstance(fy, torch.Tensor) and (fx.requires_grad or fy.requires_grad):
        return torch.stack([torch.tan(fx * 0.5).reshape(()), torch.tan(fy * 0.5).reshape(())])
    return None


def render_raw(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier: float = 1.0):
    """`render()` without the epilogue: (raw (10,H,W) image, radii, viewspace_points).  The raw image is what
    `losses.hair_image_loss` consumes (it contains the epilogue), so a trainer built on this repository never
    materialises the four separate maps."""
    P = int(pc._xyz.shape[0])
    viewspace = empty_rows(P, (3,), torch.float32, pc._xyz.device).requires_grad_(True)
    st = _static(viewpoint_camera, bg_color, scaling_modifier, pc.active_sh_degree, getattr(pipe, "debug", False),
                 projection.GAUSSIAN_MODEL)
    renders, radii = _FusedRender.apply(
        pc._xyz, pc._scaling, pc._rotation, None, pc._features_dc, pc._features_rest, pc._opacity, pc._label, pc._orient_conf,
        viewspace, viewpoint_camera.world_view_transform, viewpoint_camera.full_proj_transform, viewpoint_camera.camera_center,
        _tanfov_tensor(viewpoint_camera), st)
    return renders, radii, viewspace


def render(viewpoint_camera, pc, pipe, bg_color: torch.Tensor, scaling_modifier: float = 1.0):
    """Same contract as the reference's `render` (src/gaussian_renderer/__init__.py:23-114); `pc` is a GaussianModel
    (src/scene/gaussian_model.py:45) or anything with its raw parameter attributes."""
    renders, radii, viewspace = render_raw(viewpoint_camera, pc, pipe, bg_color, scaling_modifier)
    return _post(renders, radii, viewspace)


def _head_block(pc) -> Optional[Dict[str, torch.Tensor]]:
    """The frozen head Gaussians that render_hair() composites behind the hair (the *_precomp attributes the strand
    trainers attach, src/train_strands.py:67-73).  Cached on the model: they never change during strand training."""
    cache = getattr(pc, "_gh_head_block", None)
    if cache is not None and cache["key"] is pc.xyz_precomp:
        return cache
    with torch.no_grad():
        shs = pc.shs_view                                           # (n, 3, 16) channel-major
        feats = shs.transpose(1, 2).contiguous()                    # (n, 16, 3) like get_features
        blk = {"key": pc.xyz_precomp, "xyz": pc.xyz_precomp.contiguous(), "scaling": pc.scaling_precomp.contiguous(),
               "rotation": pc.rotation_precomp.contiguous(), "opacity": pc.opacity_precomp.contiguous(),
               "f_dc": feats[:, :1].contiguous(), "f_rest": feats[:, 1:].contiguous()}
    pc._gh_head_block = blk
    return blk


def render_hair(viewpoint_camera, pc, pc_hair, pipe, bg_color: torch.Tensor, scaling_modifier: float = 1.0):
    """Same contract as the reference's `render_hair` (src/gaussian_renderer/__init__.py:116-236): the frozen head
    Gaussians of `pc` (label < 0.5) followed by the strand Gaussians of `pc_hair`."""
    head = _head_block(pc)
    n_head = int(head["xyz"].shape[0])
    P = n_head + int(pc_hair.get_xyz.shape[0])
    dev = pc_hair.get_xyz.device
    viewspace = empty_rows(P, (3,), torch.float32, dev).requires_grad_(True)
    st = _static(viewpoint_camera, bg_color, scaling_modifier, pc_hair.active_sh_degree, getattr(pipe, "debug", False),
                 projection.HAIR_MODEL)
    st["head"] = head if n_head > 0 else None
    renders, radii = _FusedRender.apply(
        pc_hair.get_xyz, pc_hair.get_scaling, pc_hair._rotation, pc_hair._dir, pc_hair._features_dc, pc_hair._features_rest,
        None, None, pc_hair._orient_conf, viewspace, viewpoint_camera.world_view_transform,
        viewpoint_camera.full_proj_transform, viewpoint_camera.camera_center, _tanfov_tensor(viewpoint_camera), st)
    return _post(renders, radii, viewspace)
