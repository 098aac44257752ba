"""All-reduce of the gradient arena alone: NCCL vs gh_allreduce_p2p (peer ld/st, NVLS multimem).
   python -m torch.distributed.run --nproc-per-node N --master-addr 127.0.0.1 tools/allreduce_case.py [P]"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch, torch.distributed as dist
from gaussianhaircut_b200 import dist as gd
rank, world = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"])
dev = torch.device(f"cuda:{int(os.environ['LOCAL_RANK'])}")
torch.cuda.set_device(dev)
dist.init_process_group("nccl", device_id=dev)
P = int(sys.argv[1]) if len(sys.argv) > 1 else 500_000
n = 24 * P
reps = 30

def timed(fn):
    for _ in range(5):
        fn()
    dist.barrier(); torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record(); torch.cuda.synchronize()
    t = torch.tensor([e0.elapsed_time(e1) / reps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return t.item() * 1000

x = torch.randn(n, device=dev)
res = {"nccl": timed(lambda: dist.all_reduce(x))}
for name, mc in (("peer ld/st", False), ("multimem", True)):
    par = gd.PeerAllReduce(34 * P, dev, use_multicast=mc)
    if mc and not par.multicast:
        continue
    par.buffer.normal_()
    res[name] = timed(lambda: par.all_reduce(n_floats=n))
    assert par.ok()
if rank == 0:
    mb = n * 4 / 1e6
    print(f"all-reduce of {mb:.0f} MB over {world} GPUs: " + ", ".join(f"{k} {v:.1f} us ({2 * (world - 1) / world * mb / v * 1e3:.0f} GB/s bus)" for k, v in res.items()))
dist.destroy_process_group()
