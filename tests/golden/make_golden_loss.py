"""Golden vectors for the fused image loss ('next' row 4): the reference's OWN loss functions
(src/utils/loss_utils.py: l1_loss, ssim, or_loss, imported unmodified from /root/reference) evaluated on
seeded inputs on the CPU, with autograd gradients w.r.t. the (10,H,W) render.  The inline composition
(src/train_gaussians.py:126-140) and the dir->angle post-processing (src/gaussian_renderer/__init__.py:98-105)
are not importable functions in the reference; they are restated in oracle/loss_oracle.py and exercised
here with the reference's loss functions plugged in.

    python tests/golden/make_golden_loss.py        (build container only: needs /root/reference)
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "oracle"))
sys.path.insert(0, "/root/reference/src")
import loss_oracle  # noqa: E402
from utils import loss_utils as ref  # noqa: E402  (the reference module itself)

LAMBDAS = (0.8, 0.2, 0.4, 0.1)   # dl1, dssim, dmask, dorient (arbitrary, all terms active)

for name, (W, H, seed, zero_w) in {"loss_a": (45, 37, 1, False), "loss_b": (64, 48, 2, False), "loss_zero_weight": (20, 18, 3, True)}.items():
    renders, gt_image, gt_mask, gt_angle, gt_conf = loss_oracle.synthetic_case(W, H, seed, zero_weights=zero_w)
    renders.requires_grad_(True)
    loss, parts = loss_oracle.training_loss(renders, gt_image, gt_mask, gt_angle, gt_conf, *LAMBDAS,
                                            fns=(ref.l1_loss, ref.ssim, ref.or_loss))
    loss.backward()
    np.savez_compressed(os.path.join(os.path.dirname(__file__), name + ".npz"),
                        W=W, H=H, seed=seed, zero_weights=zero_w, lambdas=np.array(LAMBDAS, np.float32),
                        loss=loss.detach().numpy(), Ll1=parts["Ll1"].detach().numpy(), Lssim=parts["Lssim"].detach().numpy(),
                        Lmask=parts["Lmask"].detach().numpy(), Lorient=parts["Lorient"].detach().numpy(),
                        dL_drender=renders.grad.numpy())
    print(name, float(loss), {k: float(v) for k, v in parts.items()})
