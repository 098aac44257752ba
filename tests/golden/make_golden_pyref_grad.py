"""Gradient golden vectors for the caller preamble (groundwork for a fused projection
op): the reference's own PyTorch stage 1 (src/scene/gaussian_model.py `get_conic`, `get_mean_2d`,
`get_depths`; src/utils/sh_utils.py `eval_sh`), imported unmodified as in make_golden_pyref.py, is
differentiated with autograd on the CPU for a seeded scalar loss

    L = sum(conic * Wc) + sum(mean2d * Wm) + sum(depths * Wd) + sum(eval_sh * Wr)

with respect to the Gaussian parameters AND the camera matrices (cameras are trainable in the reference).

    python tests/golden/make_golden_pyref_grad.py     (writes tests/golden/pyref_stage1_grad.npz)
"""
import importlib.util
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_pyref as base  # noqa: E402  (stubs + cpu factory wrappers + paths)


def main():
    base._stub("plyfile", PlyData=object, PlyElement=object)
    base._stub("simple_knn")
    base._stub("simple_knn._C", distCUDA2=lambda *a, **k: None)
    base._cpu_factories()
    sys.path.insert(0, base.REF_SRC)
    spec = importlib.util.spec_from_file_location("ref_gaussian_model", os.path.join(base.REF_SRC, "scene", "gaussian_model.py"))
    gm = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(gm)
    from utils.sh_utils import eval_sh
    sys.path.insert(0, os.path.join(base.ROOT, "oracle"))
    import synth

    strands, seed, cam_k, W, H = 6, 5, 13, 160, 96
    scene = synth.make_strand_scene(strands, seed=seed)
    cam_d = synth.make_camera(cam_k, W, H)
    leaf = lambda t: t.clone().requires_grad_(True)   # noqa: E731
    vm, pm, cc = leaf(cam_d["world_view_transform"]), leaf(cam_d["full_proj_transform"]), leaf(cam_d["camera_center"])
    cam = types.SimpleNamespace(image_width=W, image_height=H, FoVx=torch.tensor(cam_d["FoVx"]), FoVy=torch.tensor(cam_d["FoVy"]),
                                world_view_transform=vm, full_proj_transform=pm, camera_center=cc)
    pc = gm.GaussianModel(3)
    pc._xyz = leaf(scene["xyz"])
    pc._scaling = leaf(torch.log(scene["scaling"]))
    pc._rotation = leaf(scene["rotation"])
    pc._opacity = torch.logit(scene["opacity"].clamp(1e-6, 1 - 1e-6))
    f_dc, f_rest = leaf(scene["f_dc"]), leaf(scene["f_rest"])

    conic = pc.get_conic(cam)
    mean2d = pc.get_mean_2d(cam)
    depths = pc.get_depths(cam)
    shs_view = torch.cat([f_dc, f_rest], dim=1).transpose(1, 2)
    d = pc._xyz - cc[None]
    d = d / d.norm(dim=1, keepdim=True)
    rgb = eval_sh(3, shs_view, d)

    g = torch.Generator().manual_seed(99)
    Wc, Wm, Wd, Wr = (torch.rand(t.shape, generator=g) for t in (conic, mean2d, depths, rgb))
    loss = (conic * Wc).sum() * 1e-3 + (mean2d * Wm).sum() + (depths * Wd).sum() + (rgb * Wr).sum()
    loss.backward()
    out = {
        "strands": np.array(strands), "seed": np.array(seed), "cam_k": np.array(cam_k), "W": np.array(W), "H": np.array(H),
        "weight_seed": np.array(99), "loss": loss.detach().numpy(),
        "g_xyz": pc._xyz.grad.numpy(), "g_log_scaling": pc._scaling.grad.numpy(), "g_rotation": pc._rotation.grad.numpy(),
        "g_f_dc": f_dc.grad.numpy(), "g_f_rest": f_rest.grad.numpy(),
        "g_viewmatrix": vm.grad.numpy(), "g_projmatrix": pm.grad.numpy(), "g_campos": cc.grad.numpy(),
    }
    path = os.path.join(HERE, "pyref_stage1_grad.npz")
    np.savez_compressed(path, **out)
    print("wrote", path, os.path.getsize(path), "bytes; loss", float(loss), {k: float(np.abs(v).max()) for k, v in out.items() if k.startswith("g_")})


if __name__ == "__main__":
    main()
