"""Fused projection preamble ("next" row 8f-1) without a GPU.

The per-Gaussian arithmetic of the CUDA kernels lives in gaussianhaircut_b200/csrc/gh_project_math.h as
host/device functions; tests/host_harness/project_host.cpp compiles that very header with g++.  Here its forward
and hand-derived backward are checked against PyTorch autograd of the restated preamble (oracle/synth.py
`project_reference`, itself pinned on the reference's Python in test_oracle_cpu.py / the pyref goldens) -- values,
every parameter gradient and the camera gradients (view matrix, projection matrix, camera centre, tan fov).
The CUDA kernels add only staging and reductions around these functions (checked on the GPU in
tests/test_gpu_projection.py)."""
import ctypes as C
import os
import subprocess
import sys

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
import synth  # noqa: E402

HARNESS_SRC = os.path.join(ROOT, "tests", "host_harness", "project_host.cpp")
HARNESS_SO = os.path.join(ROOT, "tests", "host_harness", "libproject_host.so")
MATH_H = os.path.join(ROOT, "gaussianhaircut_b200", "csrc", "gh_project_math.h")


@pytest.fixture(scope="module")
def host():
    newest = max(os.path.getmtime(HARNESS_SRC), os.path.getmtime(MATH_H))
    if not os.path.isfile(HARNESS_SO) or os.path.getmtime(HARNESS_SO) < newest:
        subprocess.run(["g++", "-O2", "-shared", "-fPIC", "-std=c++17", "-w", HARNESS_SRC, "-o", HARNESS_SO], check=True)
    return C.CDLL(HARNESS_SO)


def _flags(cfg):
    from gaussianhaircut_b200.projection import encode_flags
    return encode_flags(cfg)


def _p(t):
    return None if t is None else C.c_void_p(t.data_ptr())


def _run_host(host, raw, cam, cfg, sh_degree, mod, grads=None):
    P = raw["xyz"].shape[0]
    f = lambda k: raw[k].detach().contiguous().float() if k in raw and raw[k] is not None else None  # noqa: E731
    t = {k: f(k) for k in ("xyz", "scaling", "rotation", "dirs", "f_dc", "f_rest", "opacity", "label", "conf")}
    V = cam["world_view_transform"].detach().contiguous().float()
    Pm = cam["full_proj_transform"].detach().contiguous().float()
    cc = cam["camera_center"].detach().contiguous().float()
    common = (P, cam["image_width"], cam["image_height"], _p(t["xyz"]), _p(t["scaling"]), _p(t["rotation"]), _p(t["dirs"]),
              _p(t["f_dc"]), _p(t["f_rest"]), _p(t["opacity"]), _p(t["label"]), _p(t["conf"]), _p(V), _p(Pm), _p(cc),
              C.c_float(float(cam["tanfovx"])), C.c_float(float(cam["tanfovy"])), C.c_float(mod), sh_degree,
              C.c_uint(_flags(cfg)), C.c_float(cfg["det_eps"]))
    out = {"means2D": torch.zeros(P, 3), "colors": torch.zeros(P, 10), "opacity": torch.zeros(P, 1), "conic": torch.zeros(P, 3),
           "cov3D": torch.zeros(P, 6), "mask": torch.zeros(P, dtype=torch.uint8)}
    host.gh_host_project_forward(*common, _p(out["means2D"]), _p(out["colors"]), _p(out["opacity"]), _p(out["conic"]),
                                 _p(out["cov3D"]), _p(out["mask"]))
    if grads is None:
        return out
    g = {k: v.detach().contiguous().float() for k, v in grads.items()}
    d = {"xyz": torch.zeros(P, 3), "scaling": torch.zeros(P, 3), "rotation": torch.zeros(P, 4), "dirs": torch.zeros(P, 3),
         "f_dc": torch.zeros(P, 1, 3), "f_rest": torch.zeros(P, 15, 3), "opacity": torch.zeros(P, 1), "label": torch.zeros(P, 1),
         "conf": torch.zeros(P, 1)}
    cam29 = np.zeros(29, dtype=np.float64)
    host.gh_host_project_backward(*common, _p(out["mask"]), _p(g["means2D"]), _p(g["conic"]), _p(g["colors"]), _p(g["opacity"]),
                                  _p(d["xyz"]), _p(d["scaling"]), _p(d["rotation"]), _p(d["dirs"]), _p(d["f_dc"]), _p(d["f_rest"]),
                                  _p(d["opacity"]), _p(d["label"]), _p(d["conf"]), cam29.ctypes.data_as(C.c_void_p))
    dV = np.zeros((4, 4)); dPm = np.zeros((4, 4))
    dV[:, :3] = cam29[:12].reshape(4, 3)
    dPm[:, [0, 1, 3]] = cam29[12:24].reshape(4, 3)
    d.update(viewmatrix=torch.from_numpy(dV).float(), projmatrix=torch.from_numpy(dPm).float(),
             campos=torch.from_numpy(cam29[24:27]).float(), tan=torch.from_numpy(cam29[27:29]).float())
    return out, d


def _rel(a, b):
    if b is None:                     # autograd: no dependency at all
        b = torch.zeros_like(a)
    a, b = a.double().flatten(), b.double().flatten()
    nb = b.norm().item()
    return (a - b).norm().item() / nb if nb > 0 else (a - b).norm().item()


CASES = [
    # flavour, scene, sh_degree, modifier, camera tweak
    ("gaussian_model", "strands", 3, 1.0, None),
    ("gaussian_model", "blobs", 3, 1.0, None),
    ("gaussian_model", "blobs", 1, 0.7, None),
    ("gaussian_model", "blobs", 0, 1.0, None),
    ("gaussian_model", "blobs", 3, 1.0, "narrow"),       # narrow field of view: the +-1.3 tan(fov) clamp is active
    ("hair", "strands", 3, 1.0, None),
    ("hair", "strands", 2, 1.0, "narrow"),
]


@pytest.mark.parametrize("flavour,kind,deg,mod,tweak", CASES)
def test_projection_math_matches_autograd(host, flavour, kind, deg, mod, tweak):
    torch.manual_seed(0)
    scene = synth.make_strand_scene(6, seed=3) if kind == "strands" else synth.make_blob_scene(500, seed=5)
    if flavour == "hair" and kind != "strands":
        pytest.skip("hair flavour uses strand scenes")
    W, H = 200, 120
    cam = dict(synth.make_camera(7, W, H, focal_factor=(6.0 if tweak == "narrow" else 1.2)))
    cfg = dict(synth.PROJECT_GAUSSIAN_MODEL if flavour == "gaussian_model" else synth.PROJECT_HAIR_MODEL)
    raw = synth.raw_params_from_scene(scene, flavour)
    if kind == "blobs":
        raw["rotation"] = raw["rotation"] * (0.5 + torch.rand(raw["rotation"].shape[0], 1))      # un-normalised quaternions
    leaves = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    camg = dict(cam)
    for k in ("world_view_transform", "full_proj_transform", "camera_center"):
        camg[k] = cam[k].clone().requires_grad_(True)
    camg["tanfovx"] = torch.tensor(cam["tanfovx"], dtype=torch.float32, requires_grad=True)
    camg["tanfovy"] = torch.tensor(cam["tanfovy"], dtype=torch.float32, requires_grad=True)
    ref = synth.project_reference(leaves, camg, cfg, sh_degree=deg, scaling_modifier=mod)
    mask = ref["mask"]
    assert 0 < int(mask.sum())
    if tweak == "narrow":
        assert int(mask.sum()) < mask.numel()            # some Gaussians are culled: their rows must come back as zeros

    out = _run_host(host, raw, cam, cfg, deg, mod)
    assert torch.equal(out["mask"].bool(), mask)
    assert _rel(out["means2D"], ref["means2D"]) <= 1e-5
    assert _rel(out["conic"][mask], ref["conic"][mask]) <= 1e-4 and float(out["conic"][~mask].abs().sum()) == 0.0
    assert _rel(out["colors"], ref["colors"]) <= 1e-5
    assert _rel(out["opacity"], ref["opacity"]) <= 1e-6
    assert _rel(out["cov3D"], ref["cov3D"]) <= 1e-5

    # upstream gradients as the rasterizer would deliver them: only for visible Gaussians (the reference gathers
    # with points_mask before the rasterizer, so culled rows receive nothing)
    g = torch.Generator().manual_seed(1)
    gin = {"means2D": torch.randn(ref["means2D"].shape, generator=g), "conic": torch.randn(ref["conic"].shape, generator=g) * 1e-3,
           "colors": torch.randn(ref["colors"].shape, generator=g), "opacity": torch.randn(ref["opacity"].shape, generator=g)}
    gin["means2D"][:, 2] = 0.0                           # the rasterizer never writes a z gradient
    m = mask[:, None].float()
    loss = sum((ref[k] * gin[k] * m).sum() for k in gin)
    loss.backward()
    _, d = _run_host(host, raw, cam, cfg, deg, mod, grads=gin)
    names = {"xyz": "xyz", "scaling": "scaling", "rotation": "rotation", "f_dc": "f_dc", "f_rest": "f_rest", "conf": "conf"}
    if flavour == "gaussian_model":
        names.update(opacity="opacity", label="label")
    else:
        names.update(dirs="dirs")
    for k, src in names.items():
        gref = leaves[src].grad
        assert gref is not None, k
        assert _rel(d[k].reshape(gref.shape), gref) <= 2e-4, f"{k}: {_rel(d[k].reshape(gref.shape), gref)}"
        assert float(d[k].reshape(gref.shape)[~mask].abs().sum()) == 0.0, f"{k}: culled rows must be zero"
    assert _rel(d["viewmatrix"], camg["world_view_transform"].grad) <= 2e-4
    assert _rel(d["projmatrix"], camg["full_proj_transform"].grad) <= 2e-4
    assert _rel(d["campos"], camg["camera_center"].grad) <= 2e-4
    gt = torch.stack([camg["tanfovx"].grad, camg["tanfovy"].grad])
    assert _rel(d["tan"], gt) <= 2e-4, f"tan fov: {d['tan']} vs {gt}"


# ------------------------------------------------------------------ pinned on the reference's own Python (goldens)
GOLDEN = {k: os.path.join(ROOT, "tests", "golden", f"pyref_project_{k}.npz") for k in ("gm", "hair")}


def _golden_inputs(kind):
    d = np.load(GOLDEN[kind])
    cam = dict(synth.make_camera(int(d["cam_k"]), int(d["W"]), int(d["H"])))
    if kind == "gm":
        scene = synth.make_blob_scene(int(d["n"]), seed=int(d["seed"]))
        raw = synth.raw_params_from_scene(scene, "gaussian_model")
        raw["rotation"] = raw["rotation"] * (0.5 + torch.rand(int(d["n"]), 1, generator=torch.Generator().manual_seed(int(d["rot_seed"]))))
        cfg = dict(synth.PROJECT_GAUSSIAN_MODEL)
    else:
        scene = synth.make_strand_scene(int(d["strands"]), seed=int(d["seed"]), opacity_mode="ones")
        raw = {"xyz": scene["xyz"].clone(), "dirs": scene["dir"].clone(), "f_dc": scene["f_dc"].clone(), "f_rest": scene["f_rest"].clone(),
               "conf": torch.log(scene["orient_conf"])}
        cfg = dict(synth.PROJECT_HAIR_MODEL)
    return d, cam, raw, cfg


def _hair_derived(dirs, scale):
    """scaling / rotation of the strand models as functions of the segment vector
    (gaussian_model_latent_strands.py:109-115, :489-498)."""
    scaling = torch.cat([dirs.norm(dim=-1, keepdim=True) * 0.5, torch.full_like(dirs[:, :2], scale)], dim=-1)
    ex = torch.cat([torch.ones_like(dirs[:, :1]), torch.zeros_like(dirs[:, :2])], dim=-1)
    return scaling, synth.parallel_transport(ex, dirs)


@pytest.mark.parametrize("kind", ["gm", "hair"])
def test_project_reference_matches_the_reference_python(kind):
    """oracle/synth.py `project_reference` against outputs AND autograd gradients of the reference's own model
    classes (tests/golden/make_golden_pyref_project.py)."""
    d, cam, raw, cfg = _golden_inputs(kind)
    leaves = {k: v.clone().requires_grad_(True) for k, v in raw.items()}
    camg = dict(cam)
    for k in ("world_view_transform", "full_proj_transform", "camera_center"):
        camg[k] = cam[k].clone().requires_grad_(True)
    inp = dict(leaves)
    if kind == "hair":
        inp["scaling"], inp["rotation"] = _hair_derived(leaves["dirs"], float(d["scale"]))
    ref = synth.project_reference(inp, camg, cfg)
    mask = torch.from_numpy(d["mask"])
    assert torch.equal(ref["mask"], mask)
    for k, tol in (("means2D", 1e-5), ("conic", 2e-4), ("colors", 2e-5), ("opacity", 1e-6)):
        assert _rel(ref[k].detach(), torch.from_numpy(d[k])) <= tol, k
    m = mask[:, None].float()
    terms = [(ref[k] * torch.from_numpy(d["W_" + k]) * m).sum() for k in ("conic", "means2D", "colors")]
    if ref["opacity"].requires_grad:
        terms.append((ref["opacity"] * torch.from_numpy(d["W_opacity"]) * m).sum())
    sum(terms).backward()
    for k, t in leaves.items():
        assert _rel(t.grad, torch.from_numpy(d["g_" + k])) <= 5e-4, f"{k}: {_rel(t.grad, torch.from_numpy(d['g_' + k]))}"
    assert _rel(camg["world_view_transform"].grad, torch.from_numpy(d["g_viewmatrix"])) <= 5e-4
    assert _rel(camg["full_proj_transform"].grad, torch.from_numpy(d["g_projmatrix"])) <= 5e-4
    assert _rel(camg["camera_center"].grad, torch.from_numpy(d["g_campos"])) <= 5e-4


@pytest.mark.parametrize("kind", ["gm", "hair"])
def test_product_projection_math_matches_the_reference_python(host, kind):
    """The product's per-Gaussian arithmetic (gh_project_math.h, compiled for the host) directly against the
    reference's own Python: values and every gradient incl. the camera matrices."""
    d, cam, raw, cfg = _golden_inputs(kind)
    if kind == "hair":
        dl = raw["dirs"].clone().requires_grad_(True)
        scaling, rotation = _hair_derived(dl, float(d["scale"]))
        raw = dict(raw, scaling=scaling.detach(), rotation=rotation.detach())
    gin = {k: torch.from_numpy(d["W_" + k]).clone() for k in ("means2D", "conic", "colors", "opacity")}
    out, g = _run_host(host, raw, cam, cfg, 3, 1.0, grads=gin)
    mask = torch.from_numpy(d["mask"])
    assert torch.equal(out["mask"].bool(), mask)
    for k, tol in (("means2D", 1e-5), ("colors", 2e-5), ("opacity", 1e-6)):
        assert _rel(out[k], torch.from_numpy(d[k])) <= tol, k
    assert _rel(out["conic"][mask], torch.from_numpy(d["conic"])[mask]) <= 2e-4
    names = ["xyz", "f_dc", "f_rest", "conf"] + (["scaling", "rotation", "opacity", "label"] if kind == "gm" else [])
    for k in names:
        gref = torch.from_numpy(d["g_" + k])
        assert _rel(g[k].reshape(gref.shape), gref) <= 5e-4, f"{k}: {_rel(g[k].reshape(gref.shape), gref)}"
    if kind == "hair":       # scaling and rotation are functions of the segment vector: chain through them
        (scaling * g["scaling"]).sum().add((rotation * g["rotation"]).sum()).backward()
        total = g["dirs"] + dl.grad
        assert _rel(total, torch.from_numpy(d["g_dirs"])) <= 5e-4, f"dirs: {_rel(total, torch.from_numpy(d['g_dirs']))}"
    assert _rel(g["viewmatrix"], torch.from_numpy(d["g_viewmatrix"])) <= 5e-4
    assert _rel(g["projmatrix"], torch.from_numpy(d["g_projmatrix"])) <= 5e-4
    assert _rel(g["campos"], torch.from_numpy(d["g_campos"])) <= 5e-4
