"""Diagnostic: where does the public-API (e2e) step spend its time?  (not part of the product)"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "oracle"))
import synth
impl = sys.argv[1] if len(sys.argv) > 1 else "mine"
if impl == "mine":
    import gaussianhaircut_b200 as mod
else:
    import build_ref; mod = build_ref.load()
dev = torch.device("cuda:0")
W, H = 1920, 1080
scene = synth.make_strand_scene(5000, seed=0)
cam = synth.make_camera(0, W, H)
inp = synth.rasterizer_inputs(scene, cam, mode="native", device=dev)
s = inp["settings"]
Wt_host = synth.upstream_gradient(W, H, 0).pin_memory()
kw = {k: (v.detach().clone().requires_grad_(True) if isinstance(v, torch.Tensor) else v) for k, v in inp["kwargs"].items()}
settings = mod.GaussianRasterizationSettings(**{k: s[k] for k in mod.GaussianRasterizationSettings._fields})
def sync(): torch.cuda.synchronize()
acc = {}
def lap(name, t0):
    sync(); t = time.perf_counter(); acc[name] = acc.get(name, 0.0) + (t - t0); return t
for it in range(12):
    if it == 2: acc.clear()
    sync(); t = time.perf_counter()
    Wt = Wt_host.to(dev, non_blocking=True); t = lap("h2d", t)
    for v in kw.values():
        if isinstance(v, torch.Tensor): v.grad = None
    color, radii = mod.GaussianRasterizer(settings)(**kw); t = lap("forward", t)
    loss = (color * Wt).sum(); t = lap("loss", t)
    loss.backward(); t = lap("backward", t)
    x = loss.item(); t = lap("item", t)
print(impl, {k: round(v / 10 * 1e3, 3) for k, v in acc.items()}, "ms per step")
