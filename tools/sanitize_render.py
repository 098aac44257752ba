"""One small fused render + losses + backward + Adam + densify pass for compute-sanitizer (projection, densify and
the records-mode backward kernels).   compute-sanitizer --tool memcheck python tools/sanitize_render.py"""
import os, sys, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
import synth, ref_python
from gaussianhaircut_b200 import renderer, losses as ghl, densify
from gaussianhaircut_b200.optim import FusedAdam
dev = torch.device("cuda:0")
W, H = 250, 187
scene = synth.make_strand_scene(37, seed=2)
raw = synth.raw_params_from_scene(scene, "gaussian_model")
names = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_label", "_scaling", "_rotation", "_orient_conf")
gnames = ("xyz", "f_dc", "f_rest", "opacity", "label", "scaling", "rotation", "orient_conf")
keys = ("xyz", "f_dc", "f_rest", "opacity", "label", "scaling", "rotation", "conf")
pc = types.SimpleNamespace(active_sh_degree=3, max_sh_degree=3, percent_dense=0.01)
for n, k in zip(names, keys):
    setattr(pc, n, torch.nn.Parameter(raw[k].to(dev).contiguous()))
pc.optimizer = FusedAdam([{"params": [getattr(pc, n)], "lr": 1e-4, "name": gn} for n, gn in zip(names, gnames)], eps=1e-15)
P = pc._xyz.shape[0]
pc.xyz_gradient_accum = torch.zeros(P, 1, device=dev); pc.denom = torch.zeros(P, 1, device=dev); pc.max_radii2D = torch.zeros(P, device=dev)
cam = ref_python.make_camera(synth.make_camera(5, W, H), dev, trainable=True)
bg = torch.tensor(synth.BG_DEFAULT, device=dev)
g = torch.Generator().manual_seed(0)
gt = (torch.rand(3, H, W, generator=g).to(dev), torch.rand(2, H, W, generator=g).to(dev), torch.rand(1, H, W, generator=g).to(dev), torch.rand(1, H, W, generator=g).to(dev))
flag = torch.zeros(1, dtype=torch.int32, device=dev)
renderer.set_nan_flag(flag)
for it in range(3):
    pkg = renderer.render(cam, pc, types.SimpleNamespace(debug=False), bg)
    loss, _ = ghl.hair_image_loss(pkg["raw"], *gt, 0.8, 0.2, 0.1, 0.1)
    loss.backward()
    densify.update_max_radii(pc, pkg["radii"])
    densify.add_densification_stats(pc, pkg["viewspace_points"], pkg["visibility_filter"])
    pc.optimizer.step(nan_flag_in=flag); pc.optimizer.zero_grad()
    c = densify.densify_and_prune(pc, 1e-9, 0.005, 0.1, 20)
torch.cuda.synchronize()
print("sanitize_render ok", c, float(loss))
