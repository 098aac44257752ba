"""One forward+backward at 2M Gaussians / 1080p for compute-sanitizer (mixed in-CTA and long-list tiles)."""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gaussianhaircut_b200 import _C
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import synth
dev = torch.device("cuda:0")
W, H = 1920, 1080
scene = synth.make_strand_scene(int(sys.argv[1]) if len(sys.argv) > 1 else 20000, seed=0)
inp = synth.rasterizer_inputs(scene, synth.make_camera(0, W, H), mode="native", device=dev)
kw, s = inp["kwargs"], inp["settings"]
e = torch.Tensor([])
g = lambda k: e if kw[k] is None else kw[k]
fw = (s["bg"], kw["means3D"], kw["means2D"], g("colors_precomp"), kw["opacities"], g("scales"), g("rotations"),
      s["scale_modifier"], g("cov3D_precomp"), g("conic_precomp"), s["viewmatrix"], s["projmatrix"],
      s["tanfovx"], s["tanfovy"], s["image_height"], s["image_width"], e, s["sh_degree"], s["campos"], s["prefiltered"], False)
R, color, radii, geom, binning, img = _C.rasterize_gaussians(*fw)
torch.cuda.synchronize(); print("forward ok R", R, flush=True)
dL = synth.upstream_gradient(W, H, 0).to(dev)
bw = (s["bg"], kw["means3D"], radii, g("colors_precomp"), g("scales"), g("rotations"), s["scale_modifier"],
      g("cov3D_precomp"), g("conic_precomp"), s["viewmatrix"], s["projmatrix"], s["tanfovx"], s["tanfovy"],
      dL, e, s["sh_degree"], s["campos"], geom, R, binning, img, False)
grads = _C.rasterize_gaussians_backward(*bw)
torch.cuda.synchronize(); print("backward ok", float(grads[1].abs().sum()), flush=True)
