"""Golden vectors for the fused projection preamble from the reference's own Python.

The reference's model classes are imported UNMODIFIED (stubs + CPU factory wrappers of make_golden_pyref.py) and
asked for exactly what `render()` / `render_hair()` ask them (src/gaussian_renderer/__init__.py:29-83, :122-186):
conic, NDC mean, depth, direction feature, activations, SH colour, prefilter mask -- from RAW parameters that are
autograd leaves, plus the gradients of a fixed random-weight loss w.r.t. every parameter and the camera matrices.

    python tests/golden/make_golden_pyref_project.py     (writes tests/golden/pyref_project_{gm,hair}.npz)
"""
import os
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import make_golden_pyref as base  # noqa: E402


def main():
    base._cpu_factories()
    sys.path.insert(0, os.path.join(base.ROOT, "oracle"))
    import ref_python
    import synth
    ref_python.install_stubs()
    sys.path.insert(0, base.REF_SRC)
    from scene.gaussian_model import GaussianModel
    from scene.gaussian_model_latent_strands import GaussianModelHair
    from utils.sh_utils import eval_sh
    from utils.general_utils import parallel_transport

    W, H, cam_k = 200, 120, 7
    cam_d = synth.make_camera(cam_k, W, H)
    leaf = lambda t: t.detach().clone().requires_grad_(True)   # noqa: E731

    def camera():
        vm, pm, cc = leaf(cam_d["world_view_transform"]), leaf(cam_d["full_proj_transform"]), leaf(cam_d["camera_center"])
        return types.SimpleNamespace(image_width=W, image_height=H, FoVx=torch.tensor(cam_d["FoVx"]), FoVy=torch.tensor(cam_d["FoVy"]),
                                     world_view_transform=vm, full_proj_transform=pm, camera_center=cc)

    def outputs(pc, cam, sh_deg, opacity, label):
        conic = pc.get_conic(cam)
        m2 = pc.get_mean_2d(cam)
        shs_view = pc.get_features.transpose(1, 2).view(-1, 3, 16)
        d = pc.get_xyz - cam.camera_center.repeat(pc.get_features.shape[0], 1)
        d = d / d.norm(dim=1, keepdim=True)
        rgb = torch.clamp_min(eval_sh(sh_deg, shs_view, d) + 0.5, 0.0)
        dir2 = pc.get_direction_2d(cam)
        colors = torch.cat([rgb, label, torch.ones_like(label), dir2, pc.get_orient_conf, pc.get_depths(cam)], dim=-1)
        mask = pc.filter_points(cam)
        return {"conic": conic, "means2D": m2, "colors": colors, "opacity": opacity, "mask": mask}

    def finish(name, out, leaves, cam, extra):
        g = torch.Generator().manual_seed(99)
        Wt = {k: torch.rand(out[k].shape, generator=g) for k in ("conic", "means2D", "colors", "opacity")}
        Wt["conic"] *= 1e-3
        Wt["means2D"][:, 2] = 0.0
        m = out["mask"][:, None].float()
        terms = [(out[k] * Wt[k] * m).sum() for k in ("conic", "means2D", "colors")]
        if out["opacity"].requires_grad:
            terms.append((out["opacity"] * Wt["opacity"] * m).sum())
        loss = sum(terms)
        loss.backward()
        res = {k: v.detach().numpy() for k, v in out.items()}
        res.update({"W_" + k: v.numpy() for k, v in Wt.items()})
        res.update({"g_" + k: v.grad.numpy() for k, v in leaves.items()})
        res.update(g_viewmatrix=cam.world_view_transform.grad.numpy(), g_projmatrix=cam.full_proj_transform.grad.numpy(),
                   g_campos=cam.camera_center.grad.numpy(), loss=loss.detach().numpy(), W=np.array(W), H=np.array(H), cam_k=np.array(cam_k))
        res.update(extra)
        path = os.path.join(HERE, f"pyref_project_{name}.npz")
        np.savez_compressed(path, **res)
        print("wrote", path, os.path.getsize(path), "bytes; visible", int(out["mask"].sum()), "of", out["mask"].numel())

    # ---- GaussianModel (train_gaussians.py): anisotropic blobs with raw, un-normalised quaternions
    scene = synth.make_blob_scene(300, seed=5)
    raw = synth.raw_params_from_scene(scene, "gaussian_model")
    raw["rotation"] = raw["rotation"] * (0.5 + torch.rand(300, 1, generator=torch.Generator().manual_seed(3)))
    pc = GaussianModel(3)
    leaves = {k: leaf(v) for k, v in raw.items()}
    pc._xyz, pc._scaling, pc._rotation = leaves["xyz"], leaves["scaling"], leaves["rotation"]
    pc._features_dc, pc._features_rest = leaves["f_dc"], leaves["f_rest"]
    pc._opacity, pc._label, pc._orient_conf = leaves["opacity"], leaves["label"], leaves["conf"]
    pc.active_sh_degree = 3
    cam = camera()
    out = outputs(pc, cam, 3, pc.get_opacity, pc.get_label)
    finish("gm", out, leaves, cam, {"seed": np.array(5), "n": np.array(300), "rot_seed": np.array(3)})

    # ---- GaussianModelHair (train_strands.py / train_latent_strands.py): strand segments
    scene = synth.make_strand_scene(4, seed=8, opacity_mode="ones")
    hair = GaussianModelHair.__new__(GaussianModelHair)
    hair.setup_functions()
    hair.active_sh_degree = hair.max_sh_degree = 3
    leaves = {"xyz": leaf(scene["xyz"]), "dirs": leaf(scene["dir"]), "f_dc": leaf(scene["f_dc"]), "f_rest": leaf(scene["f_rest"]),
              "conf": leaf(torch.log(scene["orient_conf"]))}
    hair._xyz, hair._dir = leaves["xyz"], leaves["dirs"]
    hair._features_dc, hair._features_rest, hair._orient_conf = leaves["f_dc"], leaves["f_rest"], leaves["conf"]
    hair.scale = float(scene["scaling"][0, 1]) * torch.ones(1)
    ex = torch.cat([torch.ones_like(hair._xyz[:, :1]), torch.zeros_like(hair._xyz[:, :2])], dim=-1)
    hair._rotation = parallel_transport(a=ex, b=hair._dir).view(-1, 4)
    cam = camera()
    out = outputs(hair, cam, 3, hair.get_opacity, hair.get_label)
    finish("hair", out, leaves, cam, {"seed": np.array(8), "strands": np.array(4), "scale": np.array(float(scene["scaling"][0, 1]))})


if __name__ == "__main__":
    main()
