"""Fused projection preamble ("next" row 1 of SURVEY.md 8f): the PyTorch code the reference runs before every
rasterizer call, as one forward and one backward kernel (csrc/gh_project.cu) behind the C ABI
(`gh_project_forward` / `gh_project_backward`, include/gh_rasterizer.h).

Replaces, for one camera, on the RAW parameters a model stores:

    GaussianModel.get_conic / get_covariance_2d / get_covariance      src/scene/gaussian_model.py:230-315
    get_mean_2d :317-337, get_depths :339-342, get_direction_2d :344-393, filter_points :143-228
    the activations get_scaling / get_opacity / get_label / get_orient_conf   :106-141
    eval_sh + clamp, the 10-channel feature concatenation, the mask gathers    src/gaussian_renderer/__init__.py:29-83
    (and the strand models' copies of the same functions, src/scene/gaussian_model_latent_strands.py:109-440)

including the gradients w.r.t. the camera matrices, which the reference gets from autograd because cameras are
trainable (src/scene/cameras.py:124-150).  `renderer.py` builds `render()` / `render_hair()` on top of it.
There is no CPU path: CPU tensors raise.
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _capi
from ._alloc import empty_rows

# activation / mode codes (GhProjArgs in csrc/gh_project_math.h)
GAUSSIAN_MODEL = dict(scale_act=1, opacity_act=1, label_act=1, conf_act=1, dir_mode=0, det_eps=1e-12)   # gaussian_model.py
HAIR_MODEL = dict(scale_act=0, opacity_act=2, label_act=2, conf_act=1, dir_mode=1, det_eps=1e-7)         # gaussian_model_latent_strands.py
# the frozen head Gaussians inside render_hair(): activated scales / opacities, label = dir2D = confidence = 0
HEAD_PRECOMP = dict(scale_act=0, opacity_act=0, label_act=3, conf_act=3, dir_mode=2, det_eps=1e-12)


def encode_flags(cfg: Dict[str, object]) -> int:
    return (int(cfg["scale_act"]) & 3) | ((int(cfg["opacity_act"]) & 3) << 2) | ((int(cfg["label_act"]) & 3) << 4) | \
           ((int(cfg["conf_act"]) & 3) << 6) | ((int(cfg["dir_mode"]) & 3) << 8)


def _ptr(t: Optional[torch.Tensor]):
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _f32(t: Optional[torch.Tensor], name: str, device: torch.device, align: int = 4) -> Optional[torch.Tensor]:
    if t is None:
        return None
    if not t.is_cuda:
        raise RuntimeError(f"gaussianhaircut_b200 projection: '{name}' must be a CUDA tensor (there is no CPU path)")
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {t.dtype} for argument '{name}'")
    if t.device != device:
        raise RuntimeError(f"argument '{name}' is on {t.device}, expected {device}")
    t = t.detach().contiguous()
    if t.data_ptr() % align:
        t = t.clone()
    return t


def _stream(device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


class ProjectionInputs:
    """The (detached, contiguous) tensors of one projection call, kept between forward and backward."""
    __slots__ = ("P", "W", "H", "xyz", "scaling", "rotation", "dirs", "f_dc", "f_rest", "opacity", "label", "conf",
                 "V", "Pm", "campos", "tanx", "tany", "mod", "sh_degree", "flags", "det_eps", "device")


def pack_inputs(xyz, scaling, rotation, dirs, f_dc, f_rest, opacity, label, conf, viewmatrix, projmatrix, campos,
                tanfovx: float, tanfovy: float, width: int, height: int, sh_degree: int, scaling_modifier: float,
                cfg: Dict[str, object]) -> ProjectionInputs:
    if xyz.ndim != 2 or xyz.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    dev = xyz.device
    pi = ProjectionInputs()
    pi.device = dev
    pi.P, pi.W, pi.H = int(xyz.size(0)), int(width), int(height)
    pi.xyz = _f32(xyz, "xyz", dev)
    pi.scaling = _f32(scaling, "scaling", dev)
    pi.rotation = _f32(rotation, "rotation", dev, align=16)
    pi.dirs = _f32(dirs, "dirs", dev)
    pi.f_dc = _f32(f_dc, "features_dc", dev)
    pi.f_rest = _f32(f_rest, "features_rest", dev)
    pi.opacity = _f32(opacity, "opacity", dev)
    pi.label = _f32(label, "label", dev)
    pi.conf = _f32(conf, "orient_conf", dev)
    pi.V = _f32(viewmatrix, "viewmatrix", dev)
    pi.Pm = _f32(projmatrix, "projmatrix", dev)
    pi.campos = _f32(campos, "campos", dev)
    pi.tanx, pi.tany = float(tanfovx), float(tanfovy)
    pi.mod, pi.sh_degree = float(scaling_modifier), int(sh_degree)
    pi.flags, pi.det_eps = encode_flags(cfg), float(cfg["det_eps"])
    n = pi.P
    for name, t, per in (("scaling", pi.scaling, 3), ("rotation", pi.rotation, 4), ("features_dc", pi.f_dc, 3)):
        if t is None or t.numel() != n * per:
            raise RuntimeError(f"projection: '{name}' must have {per} floats per Gaussian")
    if pi.sh_degree > 0 and (pi.f_rest is None or pi.f_rest.numel() != n * 45):
        raise RuntimeError("projection: 'features_rest' must be (P, 15, 3) for sh_degree > 0")
    return pi


def _common_args(pi: ProjectionInputs):
    return (pi.P, pi.W, pi.H, _ptr(pi.xyz), _ptr(pi.scaling), _ptr(pi.rotation), _ptr(pi.dirs), _ptr(pi.f_dc), _ptr(pi.f_rest),
            _ptr(pi.opacity), _ptr(pi.label), _ptr(pi.conf), _ptr(pi.V), _ptr(pi.Pm), _ptr(pi.campos),
            pi.tanx, pi.tany, pi.mod, pi.sh_degree, pi.flags, pi.det_eps)


def alloc_outputs(P: int, dev: torch.device, want_cov3D: bool = False, means2D_out: Optional[torch.Tensor] = None):
    f32 = torch.float32
    return {"means2D": means2D_out if means2D_out is not None else empty_rows(P, (3,), f32, dev),
            "colors": empty_rows(P, (10,), f32, dev), "opacity": empty_rows(P, (1,), f32, dev), "conic": empty_rows(P, (3,), f32, dev),
            "visible": empty_rows(P, (), torch.uint8, dev),
            "cov3D": empty_rows(P, (6,), f32, dev) if want_cov3D else None}


def project_forward(pi: ProjectionInputs, want_cov3D: bool = False, means2D_out: Optional[torch.Tensor] = None,
                    out: Optional[Dict[str, torch.Tensor]] = None):
    """-> dict(means2D (P,3) NDC, colors (P,10), opacity (P,1), conic (P,3) [zeros where culled], visible (P) uint8,
    cov3D (P,6) or None).  `means2D_out`: write the NDC means into this (P,3) float32 tensor instead of allocating;
    `out`: write everything into these preallocated tensors (e.g. row slices of larger buffers)."""
    lib = _capi.load()
    dev, P = pi.device, pi.P
    if out is None:
        out = alloc_outputs(P, dev, want_cov3D, means2D_out)
    else:
        for k, v in out.items():
            if v is not None and (not v.is_contiguous() or v.shape[0] != P):
                raise RuntimeError(f"projection: preallocated output '{k}' must be contiguous with {P} rows")
    if P == 0:
        return out
    with torch.cuda.device(dev):
        _capi.check(lib.gh_project_forward(*_common_args(pi), _ptr(out["means2D"]), _ptr(out["colors"]), _ptr(out["opacity"]),
                                           _ptr(out["conic"]), _ptr(out["cov3D"]), _ptr(out["visible"]), _stream(dev)))
    return out


def project_forward_binned(pi: ProjectionInputs, means2D_out: Optional[torch.Tensor] = None):
    """project_forward AND the rasterizer's first phase in one pass over the Gaussians (gh_project_forward_binned):
    -> (out dict as project_forward, radii (P,) int32, geomBuffer, imgBuffer, num_rendered, max_tile_len).  Continue
    with `_C.forward_render`.  Bit-identical to project_forward followed by `_C.rasterize_gaussians` on its outputs
    with prefiltered=False (tests/test_gpu_projection.py)."""
    from . import _C
    lib = _capi.load()
    dev, P = pi.device, pi.P
    if P == 0:
        raise RuntimeError("project_forward_binned: empty model (use project_forward + rasterize_gaussians)")
    out = alloc_outputs(P, dev, False, means2D_out)
    geom, img, radii = _C.alloc_forward_workspaces(P, int(pi.W), int(pi.H), dev)
    n_rendered, max_len = C.c_int(0), C.c_int(0)
    with torch.cuda.device(dev):
        _capi.check(lib.gh_project_forward_binned(
            *_common_args(pi), _ptr(out["means2D"]), _ptr(out["colors"]), _ptr(out["opacity"]), _ptr(out["conic"]), None,
            _ptr(out["visible"]), _ptr(radii), _ptr(geom), _ptr(img), C.byref(n_rendered), C.byref(max_len), _stream(dev)))
    return out, radii, geom, img, int(n_rendered.value), int(max_len.value)


_CAM_WS: Dict[tuple, torch.Tensor] = {}


def _camera_workspace(P: int, dev: torch.device) -> torch.Tensor:
    nbytes = C.c_size_t()
    _capi.check(_capi.load().gh_project_workspace_size(P, C.byref(nbytes)))
    key = (dev.index, torch.cuda.current_stream(dev).cuda_stream)
    ws = _CAM_WS.get(key)
    if ws is None or ws.numel() < nbytes.value:
        ws = torch.zeros(nbytes.value, dtype=torch.uint8, device=dev)
        _CAM_WS[key] = ws
    return ws


# parameter-gradient segments of a flat gradient arena (floats per Gaussian), in this order; every segment is padded to a
# multiple of 4 floats so that a 4-float-granular all-reduce covers whole segments and every view stays 16-byte aligned
GRAD_SEGMENTS = (("rotation", 4), ("xyz", 3), ("scaling", 3), ("f_dc", 3), ("f_rest", 45), ("opacity", 1), ("label", 1), ("conf", 1),
                 ("dirs", 3))


def grad_arena_floats(P: int, with_dirs: bool = False) -> int:
    """float32 elements of the model-gradient arena for P Gaussians (61 per Gaussian + padding; +3 with `dirs`)."""
    return sum((P * n + 3) // 4 * 4 for k, n in GRAD_SEGMENTS if with_dirs or k != "dirs")


def carve_grad_arena(storage: torch.Tensor, P: int, with_dirs: bool = False) -> Dict[str, torch.Tensor]:
    need = grad_arena_floats(P, with_dirs)
    if storage.dtype != torch.float32 or not storage.is_contiguous() or storage.numel() < need:
        raise RuntimeError(f"projection: gradient arena must be a contiguous float32 tensor of at least {need} elements")
    flat = storage.view(-1)
    out, off = {}, 0
    shapes = {"f_dc": (P, 1, 3), "f_rest": (P, 15, 3)}
    for k, n in GRAD_SEGMENTS:
        if k == "dirs" and not with_dirs:
            continue
        out[k] = flat[off:off + P * n].view(shapes.get(k, (P, n)))
        off += (P * n + 3) // 4 * 4
    return out


# process-wide, like rasterizer.set_gradient_arena (autograd's backward thread must see it)
_GRAD_ARENA = {"storage": None, "last": None}


def set_gradient_arena(storage):
    """Install (or remove with None) a caller-owned flat float32 buffer -- e.g. dist.PeerAllReduce(...).buffer -- into
    which the next projection backward writes every parameter gradient (`carve_grad_arena` layout), so that a
    data-parallel trainer reduces the whole model gradient with ONE collective over `flat[:grad_arena_floats(P)]`."""
    prev = _GRAD_ARENA["storage"]
    _GRAD_ARENA["storage"] = storage
    _GRAD_ARENA["last"] = None
    return prev


def project_backward(pi: ProjectionInputs, visible: torch.Tensor, geom_buffer: Optional[torch.Tensor] = None,
                     dL_dmeans2D=None, dL_dconic4=None, dL_dcolors=None, dL_dopacity=None,
                     camera_grads: bool = True, want_means2D_grad: bool = False, nan_flag: Optional[torch.Tensor] = None):
    """Incoming gradients: either the rasterizer's geometry workspace after `gh_backward` (its accumulation records are
    read directly) or the four API-shaped tensors (dL_dmeans2D (P,3), dL_dconic (P,2,2) native layout, dL_dcolors
    (P,10), dL_dopacity (P,1)).  -> dict of parameter gradients (+ 'viewmatrix' (4,4), 'projmatrix' (4,4), 'campos' (3),
    'tanfov' (2) when `camera_grads`, + 'means2D' (P,3) = the incoming NDC gradient when `want_means2D_grad`).
    `nan_flag` (device int32[1], zeroed by the caller): OR-ed with 1 when a parameter gradient is NaN."""
    lib = _capi.load()
    dev, P = pi.device, pi.P
    f = dict(dtype=torch.float32, device=dev)
    if _GRAD_ARENA["storage"] is not None:
        a = carve_grad_arena(_GRAD_ARENA["storage"], P, with_dirs=pi.dirs is not None)
        _GRAD_ARENA["last"] = P
        g = {"xyz": a["xyz"], "scaling": a["scaling"], "rotation": a["rotation"], "dirs": a.get("dirs"),
             "f_dc": a["f_dc"], "f_rest": a["f_rest"],
             "opacity": a["opacity"] if pi.opacity is not None else None,
             "label": a["label"] if pi.label is not None else None,
             "conf": a["conf"] if pi.conf is not None else None}
        del a
    else:
        f32 = torch.float32
        g = {"xyz": empty_rows(P, (3,), f32, dev), "scaling": empty_rows(P, (3,), f32, dev), "rotation": empty_rows(P, (4,), f32, dev),
             "dirs": empty_rows(P, (3,), f32, dev) if pi.dirs is not None else None,
             "f_dc": empty_rows(P, (1, 3), f32, dev), "f_rest": empty_rows(P, (15, 3), f32, dev),
             "opacity": empty_rows(P, (1,), f32, dev) if pi.opacity is not None else None,
             "label": empty_rows(P, (1,), f32, dev) if pi.label is not None else None,
             "conf": empty_rows(P, (1,), f32, dev) if pi.conf is not None else None}
    g["means2D"] = empty_rows(P, (3,), torch.float32, dev) if want_means2D_grad else None
    cam = torch.zeros(37, **f) if camera_grads else None
    if P != 0:
        with torch.cuda.device(dev):
            ws = _camera_workspace(P, dev) if camera_grads else None
            conic4 = None if dL_dconic4 is None else _f32(dL_dconic4, "dL_dconic", dev, align=16)
            _capi.check(lib.gh_project_backward(
                *_common_args(pi), _ptr(visible), _ptr(geom_buffer),
                _ptr(_f32(dL_dmeans2D, "dL_dmeans2D", dev)), _ptr(conic4), _ptr(_f32(dL_dcolors, "dL_dcolors", dev, align=8)),
                _ptr(_f32(dL_dopacity, "dL_dopacity", dev)),
                _ptr(g["xyz"]), _ptr(g["scaling"]), _ptr(g["rotation"]), _ptr(g["dirs"]), _ptr(g["f_dc"]), _ptr(g["f_rest"]),
                _ptr(g["opacity"]), _ptr(g["label"]), _ptr(g["conf"]), _ptr(g["means2D"]), _ptr(cam), _ptr(nan_flag), _ptr(ws),
                _stream(dev)))
    if camera_grads:
        g["viewmatrix"], g["projmatrix"] = cam[0:16].view(4, 4), cam[16:32].view(4, 4)
        g["campos"], g["tanfov"] = cam[32:35], cam[35:37]
    return g
