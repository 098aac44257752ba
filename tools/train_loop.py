"""The appearance-stage training loop of the reference (src/train_gaussians.py:96-181) on synthetic data, with adaptive
density control active -- BASELINE.json configs[3] (500k Gaussians, cameras sharded over the GPUs, one gradient
all-reduce per step) and configs[4] (2M Gaussians, densify / prune every 100 iterations, at 1 and 8 GPUs).

    python tools/train_loop.py --strands 20000 --iters 300 [--impl mine|reference]
    python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 tools/train_loop.py --strands 20000 --iters 300

Per iteration (one camera per rank):
  mine       renderer.render_raw (fused projection + rasterizer) -> gh_image_loss -> backward (model gradients written
             into ONE flat arena) -> [N>1: one gh_allreduce_p2p over the arena] -> FusedAdam (NaN guard on the device)
             -> densification statistics; every `--densify-every` iterations densify.densify_and_prune (+ the SUM / SUM /
             MAX reduction of the statistics first at N>1; every rank then takes identical decisions with identical
             random draws because all ranks seed the generator identically)
  reference  the reference's own render(), loss_utils, GaussianModel.densify_and_prune and torch.optim.Adam with the
             trainer's NaN host syncs, imported unmodified (oracle/ref_python.py); single GPU only (the reference has no
             multi-GPU path).
Prints one JSON line (rank 0): iterations/s, Gaussian-iterations/s (sum over iterations of the model size / time),
the model size trajectory.  Time = CUDA events around the whole loop, max over ranks."""
import argparse, json, os, sys, time, types
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "oracle"))


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--strands", type=int, default=20000)
    ap.add_argument("--iters", type=int, default=300)
    ap.add_argument("--warmup", type=int, default=10)
    ap.add_argument("--densify-every", type=int, default=100)
    ap.add_argument("--impl", default="mine", choices=["mine", "reference"])
    ap.add_argument("--width", type=int, default=1920)
    ap.add_argument("--height", type=int, default=1080)
    ap.add_argument("--cameras", type=int, default=64)
    ap.add_argument("--no-densify", action="store_true")
    ap.add_argument("--profile", action="store_true", help="synchronise between the phases of an iteration and report their times")
    ap.add_argument("--arena-growth", type=float, default=3.0, help="N>1: capacity of the gradient arena as a multiple of the initial model size")
    args = ap.parse_args()
    import torch
    import torch.distributed as dist
    import synth, ref_python
    world = int(os.environ.get("WORLD_SIZE", "1")); rank = int(os.environ.get("RANK", "0")); lrank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference" and rank != 0:
        return
    dev = torch.device("cuda", lrank); torch.cuda.set_device(dev)
    use_dist = world > 1 and args.impl == "mine"
    if use_dist:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    W, H = args.width, args.height
    scene = synth.make_strand_scene(args.strands, seed=0)
    g = torch.Generator().manual_seed(1)
    scene["scaling"] = scene["scaling"] * torch.tensor([1.0, 1.0, 1.5])
    # a scene extent for which the synthetic strand segments straddle the clone / split threshold
    extent = 0.1
    P0 = scene["xyz"].shape[0]
    ncam = max(args.cameras, world)
    cams = [ref_python.make_camera(synth.make_camera(k % 64, W, H), dev) for k in range(min(ncam, 64))]
    gen = torch.Generator().manual_seed(11)
    n_gt = 8                                        # distinct supervision sets, resident on the device (data_device = cuda)
    gts = []
    for k in range(n_gt):
        gts.append((torch.rand(3, H, W, generator=gen).to(dev),
                    ((torch.rand(2, H, W, generator=gen) > 0.3).float() * (0.5 + 0.5 * torch.rand(2, H, W, generator=gen))).to(dev),
                    torch.rand(1, H, W, generator=gen).to(dev), torch.rand(1, H, W, generator=gen).to(dev)))
    lambdas = (0.8, 0.2, 0.1, 0.1)
    bg = torch.tensor(synth.BG_DEFAULT, device=dev)
    names = ("_xyz", "_features_dc", "_features_rest", "_opacity", "_label", "_scaling", "_rotation", "_orient_conf")
    gnames = ("xyz", "f_dc", "f_rest", "opacity", "label", "scaling", "rotation", "orient_conf")
    keys = ("xyz", "f_dc", "f_rest", "opacity", "label", "scaling", "rotation", "conf")
    lrs = (1.6e-6, 2.5e-4, 2.5e-4 / 20, 5e-3, 2.5e-4, 5e-4, 1e-4, 1e-4)
    sizes = []
    phase = {}
    timing = {"on": False}

    def mark(name, t0):
        if args.profile:
            torch.cuda.synchronize()
            t1 = time.time()
            if timing["on"]:
                phase[name] = phase.get(name, 0.0) + (t1 - t0)
            return t1
        return t0

    if args.impl == "mine":
        from gaussianhaircut_b200 import renderer, losses as ghl, densify, projection, dist as ghdist
        from gaussianhaircut_b200.optim import FusedAdam
        raw = synth.raw_params_from_scene(scene, "gaussian_model")
        pc = types.SimpleNamespace(active_sh_degree=3, max_sh_degree=3, percent_dense=0.01)
        for n, k in zip(names, keys):
            setattr(pc, n, torch.nn.Parameter(raw[k].to(dev).contiguous()))
        pc.optimizer = FusedAdam([{"params": [getattr(pc, n)], "lr": lr, "name": gn} for n, gn, lr in zip(names, gnames, lrs)], eps=1e-15)
        pc.xyz_gradient_accum = torch.zeros(P0, 1, device=dev); pc.denom = torch.zeros(P0, 1, device=dev); pc.max_radii2D = torch.zeros(P0, device=dev)
        ws = torch.empty(ghl.workspace_elems(W, H), dtype=torch.float64, device=dev)
        nan_flag = torch.zeros(1, dtype=torch.int32, device=dev)
        if not use_dist:
            renderer.set_nan_flag(nan_flag)        # single GPU: the NaN guard rides on the projection backward
        par = None
        if use_dist:
            cap = projection.grad_arena_floats(int(P0 * args.arena_growth) + 1024)          # room for the model to grow
            par = ghdist.PeerAllReduce(cap, dev)
            projection.set_gradient_arena(par.buffer)
        pipe = types.SimpleNamespace(debug=False)
        order = ("xyz", "f_dc", "f_rest", "opacity", "label", "scaling", "rotation", "conf")
        if not args.no_densify:
            densify.reserve_pools(pc, int(P0 * 1.1))      # like the arena: sized once for the run, no cudaMalloc in the loop

        def iteration(it):
            cam = cams[(it * world + rank) % len(cams)]
            gt = gts[(it * world + rank) % n_gt]
            t0 = mark("_", time.time())
            renders, radii, viewspace = renderer.render_raw(cam, pc, pipe, bg)
            t0 = mark("render", t0)
            losses8, dL = ghl.image_loss_forward_backward(renders.detach(), *gt, *lambdas, workspace=ws)
            t0 = mark("loss", t0)
            renders.backward(dL)
            t0 = mark("backward", t0)
            P = pc._xyz.shape[0]
            grads = None
            skip = ()
            nf = nan_flag
            if par is not None:
                par.all_reduce(n_floats=projection.grad_arena_floats(P))
                a = projection.carve_grad_arena(par.buffer, P)
                grads = [a[k] for k in order]
                skip = (par.error_flag,)
                nf = par.nan_flag                       # the verdict on the SUM, identical on every rank
            with torch.no_grad():
                densify.update_max_radii(pc, radii)
                densify.add_densification_stats(pc, viewspace, radii > 0)
            t0 = mark("allreduce+stats", t0)
            pc.optimizer.step(grads=grads, skip_flags=skip, nan_flag_in=nf)
            pc.optimizer.zero_grad(set_to_none=True)
            t0 = mark("optimizer", t0)
            if not args.no_densify and (it + 1) % args.densify_every == 0:
                if par is not None:
                    ghdist.allreduce_densification_stats(pc.xyz_gradient_accum, pc.denom, pc.max_radii2D)
                torch.manual_seed(1000 + it); torch.cuda.manual_seed(1000 + it)      # the shared counter: same draws on every rank
                c = densify.densify_and_prune(pc, 2e-4, 0.005, extent, 20 if it > 150 else None)
                sizes.append(c["total"])
                if par is not None and projection.grad_arena_floats(c["total"]) > par.buffer.numel():
                    raise SystemExit("model outgrew the gradient arena (--arena-growth)")
                t0 = mark("densify", t0)
            return losses8
    else:
        if not ref_python.available():
            print(json.dumps({"impl": "reference", "unavailable": "reference Python sources not staged under oracle/_ref/src"})); return
        gr = ref_python.load_renderer("ref")
        import utils.loss_utils as lu
        pc = ref_python.make_gaussian_model(scene, dev)
        pc.spatial_lr_scale = 1.0
        targs = types.SimpleNamespace(percent_dense=0.01, position_lr_init=lrs[0], position_lr_final=lrs[0] * 0.01, position_lr_delay_mult=0.01,
                                      position_lr_max_steps=30000, feature_lr=lrs[1], opacity_lr=lrs[3], label_lr=lrs[4], scaling_lr=lrs[5],
                                      rotation_lr=lrs[6], train_orient_conf=True, orient_conf_lr=lrs[7])
        pc.training_setup(targs)
        pipe = ref_python.pipe()

        def iteration(it):
            cam = cams[it % len(cams)]
            gt_image, gt_mask, gt_angle, gt_conf = gts[it % n_gt]
            pkg = gr.render(cam, pc, pipe, bg)
            Ll1 = lu.l1_loss(pkg["render"], gt_image, mask=gt_mask[1:].detach())
            Lssim = 1.0 - lu.ssim(pkg["render"] * gt_mask[1:], gt_image * gt_mask[1:])
            Lmask = lu.l1_loss(pkg["mask"], gt_mask)
            Lor = lu.or_loss(pkg["orient_angle"], gt_angle, pkg["orient_conf"], weight=torch.ones_like(gt_mask[:1]) * gt_conf, mask=gt_mask[:1])
            if torch.isnan(Lor).any():
                Lor = torch.zeros_like(Ll1)
            loss = Ll1 * lambdas[0] + Lssim * lambdas[1] + Lmask * lambdas[2] + Lor * lambdas[3]
            loss.backward()
            with torch.no_grad():
                vis, radii = pkg["visibility_filter"], pkg["radii"]
                pc.max_radii2D[vis] = torch.max(pc.max_radii2D[vis], radii[vis])
                pc.add_densification_stats(pkg["viewspace_points"], vis)
                if not args.no_densify and (it + 1) % args.densify_every == 0:
                    torch.manual_seed(1000 + it); torch.cuda.manual_seed(1000 + it)
                    pc.densify_and_prune(2e-4, 0.005, extent, 20 if it > 150 else None)
                    sizes.append(int(pc._xyz.shape[0]))
                for p in [pc._xyz, pc._features_dc, pc._features_rest, pc._opacity, pc._label, pc._scaling, pc._rotation]:
                    if p.grad is not None and p.grad.isnan().any():
                        pc.optimizer.zero_grad(set_to_none=True)
                pc.optimizer.step()
                pc.optimizer.zero_grad(set_to_none=True)
            return loss

    def sync():
        if use_dist:
            dist.barrier()
        torch.cuda.synchronize()

    for it in range(args.warmup):
        iteration(-1000 + it)                   # negative counters never hit a densification boundary
    if use_dist and args.impl == "mine" and not args.no_densify:
        # NCCL sets up the connections of an algorithm / protocol the first time a collective of that size class runs
        # (0.4 s at 2 GPUs, 1.5 s at 8, measured): a one-time cost of a 30 000-iteration run that would otherwise be
        # charged to the first densification of this 300-iteration window
        from gaussianhaircut_b200 import dist as _ghd
        _ghd.allreduce_densification_stats(torch.zeros_like(pc.xyz_gradient_accum), torch.zeros_like(pc.denom), torch.zeros_like(pc.max_radii2D))
    sizes.clear()
    timing["on"] = True
    sync()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    gauss_iters = 0
    e0.record()
    for it in range(args.iters):
        gauss_iters += int(pc._xyz.shape[0])
        out = iteration(it)
    e1.record(); sync()
    ms = e0.elapsed_time(e1)
    if use_dist:
        t = torch.tensor([ms], device=dev); dist.all_reduce(t, op=dist.ReduceOp.MAX); ms = float(t.item())
        # replicas stayed identical: same model size and same parameters on every rank
        chk = torch.stack([pc._xyz.detach().double().sum(), torch.tensor(float(pc._xyz.shape[0]), device=dev, dtype=torch.float64)])
        every = [torch.zeros_like(chk) for _ in range(world)]
        dist.all_gather(every, chk)
        same = all(torch.equal(every[0], e) for e in every[1:])
    else:
        same = True
    if rank == 0:
        n_eff = world if args.impl == "mine" else 1
        line = {"workload": f"train_gaussians.py loop, strands({args.strands}) = {P0} Gaussians initially, {W}x{H}, "
                            f"densify/prune every {args.densify_every} iterations" + (" (off)" if args.no_densify else ""),
                "impl": args.impl, "n_gpus": n_eff, "iterations": args.iters, "ms_per_iteration": ms / args.iters,
                "iterations_per_s": args.iters / (ms * 1e-3),
                "gaussian_views_per_s": gauss_iters * n_eff / (ms * 1e-3), "model_size_after_each_densification": sizes,
                "final_model_size": int(pc._xyz.shape[0]), "replicas_identical": same,
                "loss": float(out[0] if args.impl == "mine" else out.detach()), "data": "synthetic"}
        if args.profile:
            line["phase_ms_per_iteration"] = {k: 1e3 * v / args.iters for k, v in phase.items() if k != "_"}
        print(json.dumps(line), flush=True)
    if use_dist:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
