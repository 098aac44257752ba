"""The reference's own callers of the hot path, UNMODIFIED, running on this repository's rasterizer.

`north_star`: "drops in behind src/gaussian_renderer's GaussianRasterizer / rasterize_gaussians() surface so
train_gaussians.py and train_strands.py run unmodified".  The trainers' entry into the path is
`gaussian_renderer.render()` (src/gaussian_renderer/__init__.py:23, called at src/train_gaussians.py:110) and
`render_hair()` (:116, called at src/train_strands.py:103): both are imported here exactly as the reference
ships them (oracle/ref_python.py; sources from /root/reference/src or the copy staged by oracle/build_ref.py
under oracle/_ref/src), driven with real `GaussianModel` / `GaussianModelHair` objects, once on top of the
product (`diff_gaussian_rasterization` = this repository) and once on top of the reference's own extension
(oracle/_ref).  Everything the trainers read back must agree within the north-star tolerance: the four
rendered maps, `radii`, `visibility_filter`, `viewspace_points.grad` (densification statistics) and the
gradient of every model parameter and of the (trainable) camera matrices.
"""
import os
import sys

import pytest
import torch

import _util
from _util import rel_err

pytestmark = pytest.mark.gpu

REL_TOL = 1e-4
sys.path.insert(0, os.path.join(_util.ROOT, "oracle"))
import ref_python  # noqa: E402


def _need_reference():
    if not ref_python.available():
        pytest.skip("reference Python sources not staged (python oracle/build_ref.py in the build container)")
    if not _util.ref_available():
        pytest.skip("oracle/_ref not built")


def _loss(pkg, W):
    return sum((pkg[k] * W[k]).sum() for k in ("render", "mask", "orient_angle", "orient_conf"))


def _weights(H, Wd, device, seed):
    g = torch.Generator().manual_seed(seed)
    return {k: torch.rand(c, H, Wd, generator=g).to(device) for k, c in
            (("render", 3), ("mask", 2), ("orient_angle", 1), ("orient_conf", 1))}


def _compare_pkg(a, b):
    assert torch.equal(a["radii"], b["radii"]), "radii differ"
    assert torch.equal(a["visibility_filter"], b["visibility_filter"])
    for k in ("render", "mask", "orient_conf"):
        assert rel_err(a[k], b[k]) <= REL_TOL, f"{k}: {rel_err(a[k], b[k])}"
    # acos of a normalised direction: compare where the direction is defined (hair pixels)
    assert rel_err(a["orient_angle"], b["orient_angle"]) <= 1e-3, f"orient_angle: {rel_err(a['orient_angle'], b['orient_angle'])}"


@pytest.mark.parametrize("strands,W,H", [(300, 512, 384), (5000, 1920, 1080)], ids=["30k-512x384", "500k-1080p"])
def test_reference_render_runs_unmodified_on_the_product(cuda_device, strands, W, H):
    _need_reference()
    synth = _util.synth
    scene = synth.make_strand_scene(strands, seed=3)
    cam_d = synth.make_camera(11, W, H)
    bg = torch.tensor(synth.BG_DEFAULT, device=cuda_device)
    Wt = _weights(H, W, cuda_device, 5)
    res = {}
    for which in ("mine", "ref"):
        gr = ref_python.load_renderer(which)
        pc = ref_python.make_gaussian_model(scene, cuda_device)
        cam = ref_python.make_camera(cam_d, cuda_device, trainable=True)
        pkg = gr.render(cam, pc, ref_python.pipe(), bg)                 # the reference's function, as shipped
        _loss(pkg, Wt).backward()
        torch.cuda.synchronize()
        res[which] = (pkg, pc, cam)
    (pa, ma, ca), (pb, mb, cb) = res["mine"], res["ref"]
    assert type(ma).__module__ == "scene.gaussian_model" and int(pa["visibility_filter"].sum()) > 0.5 * pa["radii"].numel()
    _compare_pkg(pa, pb)
    for name in ref_python.MODEL_PARAMS:
        ga, gb = getattr(ma, name).grad, getattr(mb, name).grad
        assert ga is not None and gb is not None, name
        assert rel_err(ga, gb) <= REL_TOL, f"{name}: {rel_err(ga, gb)}"
    for name in ("world_view_transform", "full_proj_transform", "camera_center"):
        ga, gb = getattr(ca, name).grad, getattr(cb, name).grad
        assert rel_err(ga, gb) <= REL_TOL, f"camera {name}: {rel_err(ga, gb)}"
    # what add_densification_stats consumes (gaussian_model.py:739-741)
    assert rel_err(pa["viewspace_points"].grad, pb["viewspace_points"].grad) <= REL_TOL


def test_reference_render_hair_runs_unmodified_on_the_product(cuda_device):
    _need_reference()
    synth = _util.synth
    W, H = 512, 512
    head = synth.make_blob_scene(20000, seed=2, spread=0.08, max_scale=0.004)
    hair = synth.make_strand_scene(500, seed=4, opacity_mode="ones")
    cam_d = synth.make_camera(7, W, H)
    bg = torch.tensor(synth.BG_DEFAULT, device=cuda_device)
    Wt = _weights(H, W, cuda_device, 9)
    res = {}
    for which in ("mine", "ref"):
        gr = ref_python.load_renderer(which)
        pc, pc_hair = ref_python.make_hair_models(head, hair, cuda_device)
        cam = ref_python.make_camera(cam_d, cuda_device, trainable=True)
        pkg = gr.render_hair(cam, pc, pc_hair, ref_python.pipe(), bg)   # the reference's function, as shipped
        _loss(pkg, Wt).backward()
        torch.cuda.synchronize()
        res[which] = (pkg, pc_hair, cam)
    (pa, ha, ca), (pb, hb, cb) = res["mine"], res["ref"]
    assert type(ha).__module__ == "scene.gaussian_model_latent_strands"
    assert int(pa["visibility_filter"].sum()) > 0.5 * pa["radii"].numel()
    _compare_pkg(pa, pb)
    for name in ref_python.HAIR_PARAMS:
        ga, gb = getattr(ha, name).grad, getattr(hb, name).grad
        assert ga is not None and gb is not None, name
        assert rel_err(ga, gb) <= REL_TOL, f"{name}: {rel_err(ga, gb)}"
    for name in ("world_view_transform", "full_proj_transform", "camera_center"):
        assert rel_err(getattr(ca, name).grad, getattr(cb, name).grad) <= REL_TOL, name
    assert rel_err(pa["viewspace_points"].grad, pb["viewspace_points"].grad) <= REL_TOL


def test_reference_render_debug_flag_through_the_caller(cuda_device):
    """pipe.debug=True reaches raster_settings.debug (gaussian_renderer/__init__.py:50) -> per-stage checks."""
    _need_reference()
    synth = _util.synth
    scene = synth.make_strand_scene(50, seed=1)
    cam_d = synth.make_camera(2, 160, 96)
    bg = torch.tensor(synth.BG_DEFAULT, device=cuda_device)
    gr = ref_python.load_renderer("mine")
    outs = []
    for dbg in (False, True):
        pc = ref_python.make_gaussian_model(scene, cuda_device)
        cam = ref_python.make_camera(cam_d, cuda_device)
        outs.append(gr.render(cam, pc, ref_python.pipe(debug=dbg), bg))
    assert torch.equal(outs[0]["render"], outs[1]["render"]) and torch.equal(outs[0]["radii"], outs[1]["radii"])
