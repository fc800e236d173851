"""gic_allreduce (peer memory) against NCCL all_reduce on the c2 gradient sizes (G 51.6 MB, D 6.4 MB, both 58 MB), alone on
the machine: CUDA events, max over ranks.  torchrun --nproc-per-node N profiles/allreduce_bench.py"""
import os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import torch
import torch.distributed as dist
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
from gic_b200 import parallel
rank, world = dist.get_rank(), dist.get_world_size()
sizes = {"D 6.4 MB": 1_600_000, "G 51.6 MB": 12_900_000, "G+D 58 MB": 14_500_000}
comm = parallel.PeerComm(sum((n * 4 + 255) & ~255 for n in sizes.values()) + 4096, dev)
for name, n in sizes.items():
    t = comm.alloc(n); t.normal_()
    x = torch.randn(n, device=dev)
    sq = torch.zeros(1, device=dev)
    res = {}
    for tag, fn in (("peer", lambda: comm.allreduce_(t, 0, sq)), ("nccl", lambda: dist.all_reduce(x))):
        for _ in range(5):
            fn()
        torch.cuda.synchronize(); dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(20):
            fn()
        e1.record(); torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / 20], device=dev)
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        res[tag] = float(ms)
    if rank == 0:
        bus = lambda ms: 2.0 * (world - 1) / world * n * 4 / (ms * 1e-3) / 1e9
        print(f"{name:10s} world {world}: peer {res['peer'] * 1e3:7.1f} us ({bus(res['peer']):6.0f} GB/s bus)   nccl {res['nccl'] * 1e3:7.1f} us ({bus(res['nccl']):6.0f} GB/s bus)", flush=True)
if rank == 0:
    print("peer wait expired:", comm.error(), flush=True)
torch.cuda.synchronize(); dist.barrier()
import threading
tm = threading.Timer(15.0, lambda: os._exit(0)); tm.daemon = True; tm.start()
dist.destroy_process_group(); tm.cancel()
