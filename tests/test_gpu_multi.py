"""Multi-GPU (needs >= 2 GPUs on the box; skipped otherwise): gh_allreduce_p2p -- the peer-memory
gradient all-reduce -- against NCCL's all_reduce on the same data, through both data paths
(NVLS multimem when the allocation has a multicast mapping, plain peer loads/stores otherwise)."""
import os
import sys

import pytest
import torch

pytestmark = pytest.mark.gpu

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _worker(rank, world, port, result_path):
    import torch.distributed as dist
    sys.path.insert(0, ROOT)
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dev = torch.device(f"cuda:{rank}")
    torch.cuda.set_device(dev)
    dist.init_process_group("nccl", rank=rank, world_size=world, device_id=dev)
    from gaussianhaircut_b200 import dist as gd, _C
    P = 50_001                                   # odd on purpose: arena segments are padded to 4 floats
    n = _C.arena_floats(P)
    assert _C.trainable_floats(P) % 4 == 0
    msgs = []
    for use_mc in (True, False):
        par = gd.PeerAllReduce(n, dev, use_multicast=use_mc)
        path = "multimem" if par.multicast else "peer ld/st"
        g = torch.Generator(device="cpu").manual_seed(100 + rank)
        for it, (off, cnt) in enumerate([(0, None), (0, _C.trainable_floats(P)), (1024, 4096), (0, None)]):
            x = torch.randn(par.buffer.numel(), generator=g).to(dev) * (10.0 ** (it - 1))
            par.buffer.copy_(x)
            ref = x.clone()
            hi = par.buffer.numel() if cnt is None else off + (cnt + 3) // 4 * 4
            part = ref[off:hi].clone()
            dist.all_reduce(part)
            ref[off:hi] = part
            par.all_reduce(n_floats=cnt, offset_floats=off)
            torch.cuda.synchronize(dev)
            err = (par.buffer - ref).abs().max().item() / max(1e-30, ref.abs().max().item())
            assert par.ok(), f"rank {rank}: a peer did not arrive ({path})"
            assert err <= 1e-6, f"rank {rank} {path} case {it}: rel err {err}"
            assert int(par.nan_flag.item()) == 0
        # every rank holds bit-identical sums (each slice is reduced once, by its owner)
        gathered = [torch.empty_like(par.buffer) for _ in range(world)]
        dist.all_gather(gathered, par.buffer)
        assert all(torch.equal(gathered[0], t) for t in gathered[1:]), f"{path}: ranks disagree"
        # the NaN guard rides on the reduction: a NaN on ONE rank raises the flag on EVERY rank, outside the range it does not
        par.buffer.fill_(1.0)
        if rank == world - 1:
            par.buffer[777] = float("nan")
        par.all_reduce(n_floats=4096, offset_floats=0)
        torch.cuda.synchronize(dev)
        assert int(par.nan_flag.item()) == 1, f"rank {rank} {path}: NaN not reported"
        par.buffer.fill_(1.0)
        if rank == world - 1:
            par.buffer[8000] = float("nan")
        par.all_reduce(n_floats=4096, offset_floats=0)
        torch.cuda.synchronize(dev)
        assert int(par.nan_flag.item()) == 0 and par.ok()
        msgs.append(path)
    # a peer that never arrives: bounded wait, reduction SKIPPED (no partial sums), sticky error flag
    os.environ["GH_ALLREDUCE_TIMEOUT_MS"] = "300"
    par = gd.PeerAllReduce(1 << 16, dev, use_multicast=False)
    par.buffer.fill_(float(rank + 1))
    torch.cuda.synchronize(dev)
    dist.barrier()
    if rank == 0:
        par.all_reduce()                          # rank 1 does not call
        torch.cuda.synchronize(dev)
        assert not par.ok() and int(par.error_flag.item()) != 0
        assert bool((par.buffer == 1.0).all()), "a failed all-reduce must not touch the arena"
        with pytest.raises(RuntimeError, match="did not reach the barrier"):
            par.all_reduce(check=True)            # sticky
    dist.barrier()
    del os.environ["GH_ALLREDUCE_TIMEOUT_MS"]
    if rank == 0:
        open(result_path, "w").write("ok " + ",".join(msgs))
    dist.destroy_process_group()


def test_peer_allreduce_matches_nccl(tmp_path):
    if not torch.cuda.is_available() or torch.cuda.device_count() < 2:
        pytest.skip("needs >= 2 GPUs")
    import torch.multiprocessing as mp
    world = 2
    out = str(tmp_path / "result.txt")
    mp.spawn(_worker, args=(world, 29533, out), nprocs=world, join=True)
    assert open(out).read().startswith("ok")
