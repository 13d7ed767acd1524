"""Latency of the peer-arena allreduce and of torch/NCCL's small allreduce (tools only; run under torchrun)."""
import os
import sys
import torch
import torch.distributed as dist
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from tfel_b200 import dist as tdist  # noqa: E402
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
ctx = tdist.init_context(local)
us = ctx.p2p_bench(2000) if ctx.p2p else float("nan")
t = torch.zeros(2, dtype=torch.float64, device="cuda")
for _ in range(20):
    dist.all_reduce(t)
torch.cuda.synchronize()
e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(2000):
    dist.all_reduce(t)
e1.record(); e1.synchronize()
if dist.get_rank() == 0:
    print("ranks %d: p2p arena allreduce %.2f us, NCCL allreduce (eager, torch) %.2f us, p2p enabled %s"
          % (dist.get_world_size(), us, e0.elapsed_time(e1) * 1e3 / 2000, ctx.p2p))
ctx.close()
dist.destroy_process_group()
