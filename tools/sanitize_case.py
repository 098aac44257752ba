"""Small forward+backward through the C ABI for compute-sanitizer runs (memcheck / racecheck / initcheck)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import torch
import _util
import gaussianhaircut_b200._C as mine
dev = torch.device("cuda:0")
for scene, n, W, H, mode in (("strands", 40, 100, 70, "native"), ("blobs", 1500, 96, 64, "render"), ("blobs", 9000, 48, 32, "native")):
    inp = _util.make_inputs(scene, n, W, H, mode, device=dev)
    r = mine.rasterize_gaussians(*_util.native_args(inp))
    dL = _util.synth.upstream_gradient(W, H, 0).to(dev)
    g = mine.rasterize_gaussians_backward(*_util.backward_args(inp, r[2], dL, r[3], r[0], r[4], r[5]))
    torch.cuda.synchronize()
    print(scene, n, mode, "R", r[0], float(r[1].sum()), float(g[1].abs().sum()))
