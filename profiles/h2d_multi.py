"""Aggregate pinned host->device bandwidth with every rank copying the config-2 batch (49 MB) at the same time
(torchrun, one process per GPU): what the host side of an N-GPU box can feed.  Rank 0 prints one JSON line."""
import json, os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
import torch.distributed as dist
from silent_speech_decoder_b200 import dist as sdist
rank, world, local = sdist.init_from_env()
bound = sdist.bind_host_thread_to_gpu(local) if world > 1 else False
dev = torch.device("cuda", local)
h = torch.randn(256, 1000, 48).pin_memory()
d = torch.empty_like(h, device=dev)
for _ in range(3):
    d.copy_(h, non_blocking=True)
torch.cuda.synchronize()
if world > 1:
    dist.barrier()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(40):
    d.copy_(h, non_blocking=True)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 40
t = torch.tensor([ms], dtype=torch.float64, device=dev)
if world > 1:
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
if rank == 0:
    gb = h.numel() * 4 / 1e9
    print(json.dumps({"ranks": world, "bound_to_gpu_numa": bool(bound), "ms_per_copy_max_over_ranks": float(t), "per_gpu_GBps": gb / (float(t) * 1e-3),
                      "aggregate_GBps": world * gb / (float(t) * 1e-3), "cpus": os.cpu_count()}))
if world > 1:
    dist.barrier(); dist.destroy_process_group()
