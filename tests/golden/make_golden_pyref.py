"""Golden vectors from the reference's own PYTHON restatement of stage 1, generated in the build
container (the Python reference cannot travel to the GPU box, so the vectors are committed).

The reference authors restate the rasterizer's per-Gaussian preprocess in PyTorch, with the CUDA
source pasted in comments (src/scene/gaussian_model.py:143-337, `filter_points`, `get_covariance`,
`get_covariance_2d`, `get_conic`, `get_mean_2d`, `get_depths`; src/utils/sh_utils.py:57-112 `eval_sh`).
This script imports those modules UNMODIFIED from /root/reference/src and runs them on the CPU:

  * absent third-party modules that gaussian_model.py imports at module scope but that stage 1 never
    calls (`plyfile`, `simple_knn._C`) are stubbed in sys.modules;
  * the helpers hard-code device="cuda" (general_utils.py:66,84,112; gaussian_model.py:234,385), so
    torch's factory functions are wrapped to map that to the CPU -- arithmetic is untouched.

    python tests/golden/make_golden_pyref.py        (writes tests/golden/pyref_stage1.npz)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
REF_SRC = "/root/reference/src"
sys.path.insert(0, ROOT)


def _stub(name, **attrs):
    m = types.ModuleType(name)
    for k, v in attrs.items():
        setattr(m, k, v)
    sys.modules[name] = m
    return m


def _cpu_factories():
    for fn_name in ("zeros", "ones", "arange", "empty", "tensor", "eye", "full", "zeros_like", "ones_like"):
        orig = getattr(torch, fn_name)

        def wrapped(*a, __orig=orig, **kw):
            if str(kw.get("device", "")) .startswith("cuda"):
                kw["device"] = "cpu"
            return __orig(*a, **kw)
        setattr(torch, fn_name, wrapped)


def main():
    _stub("plyfile", PlyData=object, PlyElement=object)
    _stub("simple_knn")
    _stub("simple_knn._C", distCUDA2=lambda *a, **k: None)
    _cpu_factories()
    sys.path.insert(0, REF_SRC)
    spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(REF_SRC, "scene", "gaussian_model.py"))
    gm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gm)
    from utils.sh_utils import eval_sh          # reference module
    sys.path.insert(0, os.path.join(ROOT, "oracle"))
    import synth       # only for the seeded scene + camera matrices

    scene = synth.make_strand_scene(20, seed=7)
    W, H = 200, 120
    cam_d = synth.make_camera(9, W, H)
    cam = types.SimpleNamespace(
        image_width=W, image_height=H, FoVx=torch.tensor(cam_d["FoVx"]), FoVy=torch.tensor(cam_d["FoVy"]),
        world_view_transform=cam_d["world_view_transform"], full_proj_transform=cam_d["full_proj_transform"],
        camera_center=cam_d["camera_center"])

    pc = gm.GaussianModel(3)
    pc._xyz = scene["xyz"].clone()
    pc._scaling = torch.log(scene["scaling"])
    pc._rotation = scene["rotation"].clone()          # already unit norm; build_rotation re-normalises
    pc._opacity = torch.logit(scene["opacity"].clamp(1e-6, 1 - 1e-6))
    with torch.no_grad():
        conic = pc.get_conic(cam)                     # also fills pc.cov2d, pc.cov
        mean2d = pc.get_mean_2d(cam)                  # fills pc.xyz_proj
        mask = pc.filter_points(cam)
        depths = pc.get_depths(cam)
        cov6 = pc.cov
        shs_view = torch.cat([scene["f_dc"], scene["f_rest"]], dim=1).transpose(1, 2)
        d = scene["xyz"] - cam.camera_center[None]
        d = d / d.norm(dim=1, keepdim=True)
        rgb = eval_sh(3, shs_view, d)
    out = {
        "W": np.array(W), "H": np.array(H), "cam_k": np.array(9), "seed": np.array(7), "strands": np.array(20),
        "tanfovx": np.array(cam_d["tanfovx"]), "tanfovy": np.array(cam_d["tanfovy"]),
        "viewmatrix": cam_d["world_view_transform"].numpy(), "projmatrix": cam_d["full_proj_transform"].numpy(),
        "xyz": scene["xyz"].numpy(), "scaling": pc.get_scaling.numpy(), "rotation": scene["rotation"].numpy(),
        "conic": conic.numpy(), "cov2d": pc.cov2d.numpy(), "cov3D": cov6.numpy(), "mean2d_ndc": mean2d.numpy(),
        "points_mask": mask.numpy(), "depths": depths.numpy(), "sh_rgb": rgb.numpy(),
        "f_dc": scene["f_dc"].numpy(), "f_rest": scene["f_rest"].numpy(), "campos": cam_d["camera_center"].numpy(),
    }
    path = os.path.join(HERE, "pyref_stage1.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; visible", int(mask.sum()), "of", mask.numel())


if __name__ == "__main__":
    main()
