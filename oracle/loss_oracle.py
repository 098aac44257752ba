"""TEST INFRASTRUCTURE - plain PyTorch restatement of the reference's appearance-stage loss, the oracle for
`gh_image_loss` ('next' row 4).  Only tests/ and bench.py's reference arm may import this file.

Pinned: tests/golden/loss_*.npz were produced by tests/golden/make_golden_loss.py, which imports the
reference's own src/utils/loss_utils.py (l1_loss, ssim, or_loss) unmodified; tests/test_oracle_cpu.py checks
this restatement against them (values and autograd gradients).

Each function cites the reference lines it follows (SRC = /root/reference/src)."""
import math

import torch
import torch.nn.functional as F


def l1_loss(network_output, gt, weight=None, mask=None):
    # SRC/utils/loss_utils.py:19-26
    loss = (network_output - gt).abs()
    if mask is not None:
        loss = loss * mask
    if weight is not None:
        return (loss * weight).sum() / weight.sum()
    return loss.mean()


def or_loss(network_output, gt, confs=None, weight=None, mask=None):
    # SRC/utils/loss_utils.py:31-48
    weight = torch.ones_like(gt[:1]) if weight is None else weight
    d = network_output - gt
    loss = torch.minimum(d.abs(), torch.minimum((d - 1).abs(), (d + 1).abs())) * math.pi
    if confs is not None:
        loss = loss * confs - (confs + 1e-7).log()
    if mask is not None:
        loss = loss * mask
    return (loss * weight).sum() / weight.sum()


def _window(channel, dtype, device, window_size=11, sigma=1.5):
    # SRC/utils/loss_utils.py:73-81 (float32 taps, float32 normalisation, outer product)
    g = torch.tensor([math.exp(-(x - window_size // 2) ** 2 / float(2 * sigma ** 2)) for x in range(window_size)])
    g = (g / g.sum()).unsqueeze(1)
    w2 = g.mm(g.t()).float().unsqueeze(0).unsqueeze(0)
    return w2.expand(channel, 1, window_size, window_size).contiguous().to(device=device, dtype=dtype)


def ssim(img1, img2, window_size=11):
    # SRC/utils/loss_utils.py:83-121 (size_average=True)
    channel = img1.size(-3)
    window = _window(channel, img1.dtype, img1.device, window_size)
    pad = window_size // 2
    mu1 = F.conv2d(img1, window, padding=pad, groups=channel)
    mu2 = F.conv2d(img2, window, padding=pad, groups=channel)
    mu1_sq, mu2_sq, mu1_mu2 = mu1.pow(2), mu2.pow(2), mu1 * mu2
    sigma1_sq = F.conv2d(img1 * img1, window, padding=pad, groups=channel) - mu1_sq
    sigma2_sq = F.conv2d(img2 * img2, window, padding=pad, groups=channel) - mu2_sq
    sigma12 = F.conv2d(img1 * img2, window, padding=pad, groups=channel) - mu1_mu2
    C1, C2 = 0.01 ** 2, 0.03 ** 2
    ssim_map = ((2 * mu1_mu2 + C1) * (2 * sigma12 + C2)) / ((mu1_sq + mu2_sq + C1) * (sigma1_sq + sigma2_sq + C2))
    return ssim_map.mean()


def split_render(renders):
    # SRC/gaussian_renderer/__init__.py:98-105
    image, mask, cov2d, orient_conf, _ = renders.split([3, 2, 3, 1, 1], dim=0)
    dir2d = F.normalize(cov2d[:2], dim=0)
    to_mirror = torch.ones_like(dir2d[[0]])
    to_mirror[dir2d[[0]] < 0] *= -1
    orient_angle = torch.acos(dir2d[[1]].clamp(-1 + 1e-3, 1 - 1e-3) * to_mirror) / math.pi
    return image, mask, orient_angle, orient_conf


def training_loss(renders, gt_image, gt_mask, gt_orient_angle, gt_orient_conf,
                  lambda_dl1, lambda_dssim, lambda_dmask, lambda_dorient, fns=None):
    """SRC/train_gaussians.py:126-140.  `fns` lets the golden generator plug in the reference's own
    (l1_loss, ssim, or_loss)."""
    f_l1, f_ssim, f_or = fns if fns is not None else (l1_loss, ssim, or_loss)
    image, mask, orient_angle, orient_conf = split_render(renders)
    Ll1 = f_l1(image, gt_image, mask=gt_mask[1:].detach())
    Lssim = 1.0 - f_ssim(image * gt_mask[1:], gt_image * gt_mask[1:])
    Lmask = f_l1(mask, gt_mask)
    orient_weight = torch.ones_like(gt_mask[:1]) * gt_orient_conf
    Lorient = f_or(orient_angle, gt_orient_angle, orient_conf, weight=orient_weight, mask=gt_mask[:1])
    if torch.isnan(Lorient).any():
        Lorient = torch.zeros_like(Ll1)
    loss = Ll1 * lambda_dl1 + Lssim * lambda_dssim + Lmask * lambda_dmask + Lorient * lambda_dorient
    return loss, {"Ll1": Ll1, "Lssim": Lssim, "Lmask": Lmask, "Lorient": Lorient}


def synthetic_case(W, H, seed=0, device="cpu", zero_weights=False):
    """Seeded inputs shaped like one training view: a plausible render and its supervision maps."""
    g = torch.Generator().manual_seed(seed)
    r = lambda *s: torch.rand(*s, generator=g)
    renders = torch.empty(10, H, W)
    renders[0:3] = r(3, H, W)
    renders[3:5] = r(2, H, W)
    renders[5:8] = r(3, H, W) * 2 - 1
    renders[8:9] = r(1, H, W) * 0.9 + 0.05
    renders[9:10] = r(1, H, W) * 3
    gt_image = r(3, H, W)
    gt_mask = (r(2, H, W) > 0.3).float() * (0.5 + 0.5 * r(2, H, W))
    gt_angle = r(1, H, W)
    gt_conf = torch.zeros(1, H, W) if zero_weights else r(1, H, W)
    return [t.to(device) for t in (renders, gt_image, gt_mask, gt_angle, gt_conf)]
