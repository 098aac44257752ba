"""gh_image_loss at 1920x1080 on synthetic maps (for ncu / timing): python tools/loss_case.py [reps] [variant .so]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
from gaussianhaircut_b200 import _capi, losses as ghl
if len(sys.argv) > 2:                               # development only: time a variant build of the library
    _capi.LIB_PATH = os.path.abspath(sys.argv[2])
dev = torch.device("cuda:0")
W, H = 1920, 1080
g = torch.Generator().manual_seed(3)
r = lambda *s: torch.rand(*s, generator=g).to(dev)
render = torch.cat([r(3, H, W), r(2, H, W), r(3, H, W) * 2 - 1, r(1, H, W) * 0.9 + 0.05, r(1, H, W) * 3])
gts = (r(3, H, W), (r(2, H, W) > 0.3).float(), r(1, H, W), r(1, H, W))
ws = torch.empty(ghl.workspace_elems(W, H), dtype=torch.float64, device=dev)
reps = int(sys.argv[1]) if len(sys.argv) > 1 else 20
for _ in range(3):
    ghl.image_loss_forward_backward(render, *gts, 0.8, 0.2, 0.1, 0.1, workspace=ws)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(reps):
    losses, dL = ghl.image_loss_forward_backward(render, *gts, 0.8, 0.2, 0.1, 0.1, workspace=ws)
e1.record(); torch.cuda.synchronize()
print(f"gh_image_loss {W}x{H}: {e0.elapsed_time(e1) / reps * 1000:.1f} us/call", losses.tolist()[:5])
