"""Host-side mirror of the reference's native module `diff_gaussian_rasterization._C`.

Same three functions, same positional signatures, same return tuples and error behaviour as the
reference's torch/pybind binding (ext/diff_gaussian_rasterization_hair/ext.cpp:15-19,
rasterize_points.cu:35-123, :125-206, :208-227) -- but the work is done by libgh_raster.so through
its C ABI (include/gh_rasterizer.h).  PyTorch is used for device memory and the current stream only.
"""
from __future__ import annotations

import ctypes as C
from typing import Optional, Tuple

import torch

from . import _capi
from ._alloc import empty_rows, row_capacity

NUM_CHANNELS = 10  # reference cuda_rasterizer/config.h:15

# floats per Gaussian of each gradient the backward returns (rasterize_points.cu:160-168);
# they are views into ONE flat zero-filled arena so that a multi-GPU caller can all-reduce
# everything the op produced with a single collective and no packing copy.
# (rotations first: the native side stores them as one 16-byte vector per Gaussian; conic sits at a
# multiple of 4 floats per Gaussian for the same reason).  The first 21 floats per Gaussian are the
# gradients the optimizer consumes in the native call shape (scales+rotations, no conic): a multi-GPU
# caller only has to all-reduce `flat[:P * GRAD_FLOATS_TRAINABLE_NATIVE]`.  means2D follows: its gradient
# is only read for the densification statistics (per-view NORMS are accumulated, so it must not be
# summed over ranks as a gradient; dist.allreduce_densification_stats reduces the statistics instead).
_GRAD_LAYOUT = (("rotations", 4), ("colors", NUM_CHANNELS), ("opacity", 1), ("means3D", 3), ("scales", 3),
                ("means2D", 3), ("conic", 4), ("cov3D", 6))
_N_TRAINABLE_SEGMENTS = 5
GRAD_FLOATS_TRAINABLE_NATIVE = 4 + NUM_CHANNELS + 1 + 3 + 3
GRAD_FLOATS_PER_GAUSSIAN = sum(n for _, n in _GRAD_LAYOUT)


def _seg_floats(P: int, n: int) -> int:
    """Every arena segment is padded to a multiple of 4 floats: segments stay 16-byte aligned for any P and
    a 4-float-granular all-reduce of the trainable prefix never touches the means2D segment behind it."""
    return (P * n + 3) // 4 * 4


def arena_floats(P: int) -> int:
    """float32 elements of the gradient arena for P Gaussians (34 * P when P % 4 == 0)."""
    return sum(_seg_floats(P, n) for _, n in _GRAD_LAYOUT)


def trainable_floats(P: int) -> int:
    """Length of the arena prefix the optimizer consumes in the native call shape (21 * P when P % 4 == 0)."""
    return sum(_seg_floats(P, n) for _, n in _GRAD_LAYOUT[:_N_TRAINABLE_SEGMENTS])


def _ptr(t: Optional[torch.Tensor]):
    """Device pointer or NULL for an absent optional (empty tensor, reference __init__.py:210-222)."""
    if t is None or t.numel() == 0:
        return None
    return C.c_void_p(t.data_ptr())


def _prep(t: torch.Tensor, name: str, device: torch.device, align: int = 4) -> torch.Tensor:
    """contiguous float32 on `device`, like `.contiguous().data<float>()` in the reference binding."""
    if t.numel() == 0:
        return t
    if t.dtype != torch.float32:
        raise RuntimeError(f"expected scalar type Float but found {t.dtype} for argument '{name}'")
    if t.device != device:
        raise RuntimeError(f"argument '{name}' is on {t.device}, expected {device}")
    t = t.contiguous()
    if t.data_ptr() % align:
        t = t.clone()
    return t


def _stream(device: torch.device):
    return C.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def rasterize_gaussians(
    background: torch.Tensor, means3D: torch.Tensor, means2D_precomp: torch.Tensor,
    colors: torch.Tensor, opacity: torch.Tensor, scales: torch.Tensor, rotations: torch.Tensor,
    scale_modifier: float, cov3D_precomp: torch.Tensor, conic_precomp: torch.Tensor,
    viewmatrix: torch.Tensor, projmatrix: torch.Tensor, tan_fovx: float, tan_fovy: float,
    image_height: int, image_width: int, sh: torch.Tensor, degree: int, campos: torch.Tensor,
    prefiltered: bool, debug: bool,
) -> Tuple[int, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor, torch.Tensor]:
    """RasterizeGaussiansCUDA (rasterize_points.cu:35-123).

    Returns (num_rendered, out_color (C,H,W), radii (P,) int32, geomBuffer, binningBuffer, imgBuffer).
    """
    if means3D.ndim != 2 or means3D.size(1) != 3:
        raise RuntimeError("means3D must have dimensions (num_points, 3)")
    lib = _capi.load()
    device = means3D.device
    if device.type != "cuda":
        raise RuntimeError("gaussianhaircut_b200 rasterizer: tensors must live on a CUDA device (no CPU path)")
    P, H, W = int(means3D.size(0)), int(image_height), int(image_width)
    byte_opts = dict(dtype=torch.uint8, device=device)

    if P == 0:
        # reference: native call skipped, all-zero image (NOT background), R = 0 (rasterize_points.cu:86-87,122)
        return (0, torch.zeros((NUM_CHANNELS, H, W), dtype=torch.float32, device=device),
                torch.zeros((0,), dtype=torch.int32, device=device),
                torch.empty(0, **byte_opts), torch.empty(0, **byte_opts), torch.empty(0, **byte_opts))

    with torch.cuda.device(device):
        means3D = _prep(means3D, "means3D", device)
        colors = _prep(colors, "colors", device, align=8)
        opacity = _prep(opacity, "opacity", device)
        scales = _prep(scales, "scales", device)
        rotations = _prep(rotations, "rotations", device, align=16)
        cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device)
        conic_precomp = _prep(conic_precomp, "conic_precomp", device)
        viewmatrix = _prep(viewmatrix, "viewmatrix", device)
        projmatrix = _prep(projmatrix, "projmatrix", device)
        background = _prep(background, "background", device)
        M = int(sh.size(1)) if sh.numel() != 0 else 0

        geom_bytes, img_bytes, geom_cap = C.c_size_t(), C.c_size_t(), C.c_size_t()
        _capi.check(lib.gh_forward_workspace_sizes(P, W, H, C.byref(geom_bytes), C.byref(img_bytes)))
        # bucketed sizes for the per-Gaussian buffers (see _alloc.py): a model that grows a little keeps hitting the
        # allocator's cached blocks
        _capi.check(lib.gh_forward_workspace_sizes(row_capacity(P), W, H, C.byref(geom_cap), None))
        geomBuffer = torch.empty(geom_cap.value, **byte_opts)[:geom_bytes.value]
        imgBuffer = torch.empty(img_bytes.value, **byte_opts)
        radii = empty_rows(P, (), torch.int32, device)
        out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=device)
        stream = _stream(device)

        n_rendered, max_len = C.c_int(0), C.c_int(0)
        _capi.check(lib.gh_forward_preprocess(
            P, int(degree), M, W, H,
            _ptr(means3D), _ptr(means2D_precomp), _ptr(sh), _ptr(colors), _ptr(opacity),
            _ptr(scales), float(scale_modifier), _ptr(rotations),
            _ptr(cov3D_precomp), _ptr(conic_precomp),
            _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
            float(tan_fovx), float(tan_fovy), int(bool(prefiltered)),
            _ptr(radii), _ptr(geomBuffer), _ptr(imgBuffer),
            C.byref(n_rendered), C.byref(max_len), int(bool(debug)), stream))
        R = int(n_rendered.value)

        bin_bytes = C.c_size_t()
        _capi.check(lib.gh_binning_workspace_size(R, C.byref(bin_bytes)))
        binningBuffer = torch.empty(bin_bytes.value, **byte_opts)
        _capi.check(lib.gh_forward_render(
            P, W, H, _ptr(background), _ptr(colors), _ptr(radii),
            _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer),
            R, int(max_len.value), _ptr(out_color), int(bool(debug)), stream))
    return R, out_color, radii, geomBuffer, binningBuffer, imgBuffer


def alloc_forward_workspaces(P: int, W: int, H: int, device: torch.device):
    """(geomBuffer, imgBuffer, radii) of a forward pass -- what rasterize_gaussians allocates before its first native
    call; used by the fused projection (projection.project_forward_binned), which fills them itself."""
    lib = _capi.load()
    byte_opts = dict(dtype=torch.uint8, device=device)
    geom_bytes, img_bytes, geom_cap = C.c_size_t(), C.c_size_t(), C.c_size_t()
    _capi.check(lib.gh_forward_workspace_sizes(P, W, H, C.byref(geom_bytes), C.byref(img_bytes)))
    _capi.check(lib.gh_forward_workspace_sizes(row_capacity(P), W, H, C.byref(geom_cap), None))
    geomBuffer = torch.empty(geom_cap.value, **byte_opts)[:geom_bytes.value]
    imgBuffer = torch.empty(img_bytes.value, **byte_opts)
    radii = empty_rows(P, (), torch.int32, device)
    return geomBuffer, imgBuffer, radii


def forward_render(background: torch.Tensor, colors: torch.Tensor, radii: torch.Tensor, geomBuffer: torch.Tensor,
                   imgBuffer: torch.Tensor, num_rendered: int, max_tile_len: int, image_height: int, image_width: int,
                   debug: bool = False):
    """Second phase of the forward (emit, sort, blend) on workspaces whose first phase already ran
    (gh_forward_preprocess or gh_project_forward_binned).  -> (out_color (C,H,W), binningBuffer)."""
    lib = _capi.load()
    device = colors.device
    P, H, W = int(colors.shape[0]), int(image_height), int(image_width)
    with torch.cuda.device(device):
        background = _prep(background, "background", device)
        colors = _prep(colors, "colors", device, align=8)
        out_color = torch.empty((NUM_CHANNELS, H, W), dtype=torch.float32, device=device)
        bin_bytes = C.c_size_t()
        _capi.check(lib.gh_binning_workspace_size(int(num_rendered), C.byref(bin_bytes)))
        binningBuffer = torch.empty(bin_bytes.value, dtype=torch.uint8, device=device)
        _capi.check(lib.gh_forward_render(
            P, W, H, _ptr(background), _ptr(colors), _ptr(radii),
            _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer),
            int(num_rendered), int(max_tile_len), _ptr(out_color), int(bool(debug)), _stream(device)))
    return out_color, binningBuffer


def alloc_grad_arena(P: int, device: torch.device, zero: bool = True, storage: torch.Tensor | None = None):
    """One flat float32 buffer holding all per-Gaussian gradients + named (P, n) views.
    `zero=False` skips the fill: gh_backward writes every element itself.  `storage`: use (the head of)
    this float32 buffer instead of allocating, e.g. a symmetric-memory arena (dist.PeerAllReduce)."""
    n = arena_floats(P)
    if storage is not None:
        if storage.dtype != torch.float32 or storage.numel() < n or not storage.is_contiguous():
            raise RuntimeError("gradient arena storage must be a contiguous float32 tensor of at least _C.arena_floats(P) elements")
        flat = storage.view(-1)[:n]
        if zero:
            flat.zero_()
    else:
        alloc = torch.zeros if zero else torch.empty
        flat = alloc(n, dtype=torch.float32, device=device)
    views, off = {}, 0
    for name, k in _GRAD_LAYOUT:
        views[name] = flat[off:off + P * k].view(P, k)
        off += _seg_floats(P, k)
    return flat, views


def rasterize_gaussians_backward_arena(
    background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, conic_precomp,
    viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos,
    geomBuffer, R, binningBuffer, imageBuffer, debug, arena_storage=None,
):
    """Same work as `rasterize_gaussians_backward`, but returns (flat_arena, views, dL_dsh): every
    per-Gaussian gradient is a (P, n) view into ONE flat float32 buffer, which is what a multi-GPU
    caller all-reduces (one collective per step, no packing copy).  `arena_storage` places the arena
    in caller-owned memory (a symmetric-memory buffer for dist.PeerAllReduce)."""
    lib = _capi.load()
    device = means3D.device
    P = int(means3D.size(0))
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    M = int(sh.size(1)) if sh.numel() != 0 else 0
    flat, g = alloc_grad_arena(P, device, zero=False, storage=arena_storage)     # gh_backward writes every element
    dL_dsh = torch.zeros((P, M, 3), dtype=torch.float32, device=device)
    if P != 0:
        with torch.cuda.device(device):
            means3D = _prep(means3D, "means3D", device)
            colors = _prep(colors, "colors", device, align=8)
            scales = _prep(scales, "scales", device)
            rotations = _prep(rotations, "rotations", device, align=16)
            cov3D_precomp = _prep(cov3D_precomp, "cov3D_precomp", device)
            conic_precomp = _prep(conic_precomp, "conic_precomp", device)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            background = _prep(background, "background", device)
            dL_dout_color = _prep(dL_dout_color, "dL_dout_color", device)
            radii = radii.contiguous()
            _capi.check(lib.gh_backward(
                P, int(degree), M, int(R), W, H,
                _ptr(background), _ptr(means3D), _ptr(sh), _ptr(colors),
                _ptr(scales), float(scale_modifier), _ptr(rotations),
                _ptr(cov3D_precomp), _ptr(conic_precomp),
                _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
                float(tan_fovx), float(tan_fovy), _ptr(radii),
                _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
                _ptr(dL_dout_color),
                _ptr(g["means2D"]), _ptr(g["conic"]), _ptr(g["opacity"]), _ptr(g["colors"]),
                _ptr(g["means3D"]), _ptr(g["cov3D"]), _ptr(dL_dsh), _ptr(g["scales"]), _ptr(g["rotations"]),
                int(bool(debug)), _stream(device)))
    return flat, g, dL_dsh


def rasterize_gaussians_backward_records(background, means3D, radii, colors, conic_precomp, viewmatrix, projmatrix,
                                         tan_fovx, tan_fovy, dL_dout_color, campos, geomBuffer, R, binningBuffer,
                                         imageBuffer, debug) -> None:
    """Blend backward only, for the conic_precomp call shape: the per-Gaussian accumulation records (dL/d colour,
    2-D mean, conic, opacity) are LEFT in `geomBuffer` for `gh_project_backward` to consume directly -- no unpack
    pass, no intermediate gradient tensors (projection.project_backward(geom_buffer=...))."""
    lib = _capi.load()
    device = means3D.device
    P = int(means3D.size(0))
    if P == 0:
        return
    H, W = int(dL_dout_color.size(1)), int(dL_dout_color.size(2))
    if conic_precomp is None or conic_precomp.numel() == 0:
        raise RuntimeError("rasterize_gaussians_backward_records needs the conic_precomp call shape")
    with torch.cuda.device(device):
        dL = _prep(dL_dout_color, "dL_dout_color", device)
        _capi.check(lib.gh_backward(
            P, 0, 0, int(R), W, H,
            _ptr(background), _ptr(means3D), None, _ptr(colors),
            None, 1.0, None, None, _ptr(conic_precomp),
            _ptr(viewmatrix), _ptr(projmatrix), _ptr(campos),
            float(tan_fovx), float(tan_fovy), _ptr(radii),
            _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imageBuffer),
            _ptr(dL),
            None, None, None, None, None, None, None, None, None,
            int(bool(debug)), _stream(device)))


def rasterize_gaussians_backward(
    background: torch.Tensor, means3D: torch.Tensor, radii: torch.Tensor, colors: torch.Tensor,
    scales: torch.Tensor, rotations: torch.Tensor, scale_modifier: float,
    cov3D_precomp: torch.Tensor, conic_precomp: torch.Tensor,
    viewmatrix: torch.Tensor, projmatrix: torch.Tensor, tan_fovx: float, tan_fovy: float,
    dL_dout_color: torch.Tensor, sh: torch.Tensor, degree: int, campos: torch.Tensor,
    geomBuffer: torch.Tensor, R: int, binningBuffer: torch.Tensor, imageBuffer: torch.Tensor,
    debug: bool,
):
    """RasterizeGaussiansBackwardCUDA (rasterize_points.cu:125-206).

    Returns (dL_dmeans2D (P,3), dL_dcolors (P,C), dL_dopacity (P,1), dL_dmeans3D (P,3), dL_dcov3D (P,6),
    dL_dconic (P,2,2), dL_dsh (P,M,3), dL_dscales (P,3), dL_drotations (P,4)).
    """
    _flat, g, dL_dsh = rasterize_gaussians_backward_arena(
        background, means3D, radii, colors, scales, rotations, scale_modifier, cov3D_precomp, conic_precomp,
        viewmatrix, projmatrix, tan_fovx, tan_fovy, dL_dout_color, sh, degree, campos,
        geomBuffer, R, binningBuffer, imageBuffer, debug)
    P = int(means3D.size(0))
    return (g["means2D"], g["colors"], g["opacity"], g["means3D"], g["cov3D"],
            g["conic"].view(P, 2, 2), dL_dsh, g["scales"], g["rotations"])


def mark_visible(means3D: torch.Tensor, viewmatrix: torch.Tensor, projmatrix: torch.Tensor) -> torch.Tensor:
    """markVisible (rasterize_points.cu:208-227): bool (P,) -- near-plane test only."""
    lib = _capi.load()
    device = means3D.device
    P = int(means3D.size(0))
    present = torch.zeros((P,), dtype=torch.bool, device=device)
    if P != 0:
        with torch.cuda.device(device):
            means3D = _prep(means3D, "means3D", device)
            viewmatrix = _prep(viewmatrix, "viewmatrix", device)
            projmatrix = _prep(projmatrix, "projmatrix", device)
            _capi.check(lib.gh_mark_visible(P, _ptr(means3D), _ptr(viewmatrix), _ptr(projmatrix),
                                            _ptr(present), _stream(device)))
    return present


def debug_export(P: int, W: int, H: int, R: int, geomBuffer, binningBuffer, imgBuffer):
    """Parity-test helper: unpack the opaque workspaces into reference-shaped arrays
    (keys, point_list, ranges, final_T, n_contrib, depths, means2D, conic_opacity)."""
    lib = _capi.load()
    device = geomBuffer.device
    T = ((W + 15) // 16) * ((H + 15) // 16)
    out = {
        "keys": torch.zeros(R, dtype=torch.int64, device=device),
        "point_list": torch.zeros(R, dtype=torch.int32, device=device),
        "ranges": torch.zeros((T, 2), dtype=torch.int32, device=device),
        "final_T": torch.zeros(H * W, dtype=torch.float32, device=device),
        "n_contrib": torch.zeros(H * W, dtype=torch.int32, device=device),
        "depths": torch.zeros(P, dtype=torch.float32, device=device),
        "means2D": torch.zeros((P, 2), dtype=torch.float32, device=device),
        "conic_opacity": torch.zeros((P, 4), dtype=torch.float32, device=device),
    }
    with torch.cuda.device(device):
        _capi.check(lib.gh_debug_export(
            P, W, H, R, _ptr(geomBuffer), _ptr(binningBuffer), _ptr(imgBuffer),
            _ptr(out["keys"]), _ptr(out["point_list"]), _ptr(out["ranges"]),
            _ptr(out["final_T"]), _ptr(out["n_contrib"]),
            _ptr(out["depths"]), _ptr(out["means2D"]), _ptr(out["conic_opacity"]), _stream(device)))
    return out
