"""Utterance-sharded data parallelism for the CTC path (one process per GPU).

Every utterance's lattice, argmax-collapse and beam search is independent, so ranks never
exchange log-probs or gradients.  The only collectives are
  * loss:   one all-reduce (SUM) of two scalars -- sum_b nll_b / max(S_b,1) and the batch size --
            so that each rank can report the global 'mean' loss; gradients are scaled by the
            GLOBAL batch size inside the kernel (``mean_denominator``), so they need no exchange;
  * decode: one gather of the ragged int32 transcripts (flattened ids + counts) to rank 0.
Backend: NCCL over NVLink on GPUs; gloo in the CPU tests of this host logic.
"""

from __future__ import annotations

import os
from typing import Callable, List, Optional, Sequence, Tuple

import torch
import torch.distributed as dist


def init_from_env(backend: Optional[str] = None) -> Tuple[int, int, int]:
    """Reads RANK / WORLD_SIZE / LOCAL_RANK / MASTER_* (torchrun). Returns (rank, world, local_rank)."""
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world > 1 and not dist.is_initialized():
        if backend is None:
            backend = "nccl" if torch.cuda.is_available() else "gloo"
        if backend == "nccl":
            torch.cuda.set_device(local)
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        os.environ.setdefault("MASTER_PORT", "29511")
        dist.init_process_group(backend=backend, rank=rank, world_size=world,
                                device_id=torch.device("cuda", local) if backend == "nccl" else None)
    elif torch.cuda.is_available():
        torch.cuda.set_device(local)
    return rank, world, local


def bind_host_thread_to_gpu(local_rank: int) -> bool:
    """Pin the calling thread to the CPUs nearest ``cuda:local_rank`` (NVML's affinity mask), so the pinned host
    buffers a rank allocates afterwards are first-touched on that GPU's NUMA node.  With eight ranks copying
    log-probs to eight GPUs at once the host side of the copy is what they share.  Best effort: False when NVML
    or the affinity call is unavailable."""
    try:
        import pynvml

        pynvml.nvmlInit()
        vis = os.environ.get("CUDA_VISIBLE_DEVICES")
        idx = int(vis.split(",")[local_rank]) if vis and all(x.strip().isdigit() for x in vis.split(",")) else local_rank
        pynvml.nvmlDeviceSetCpuAffinity(pynvml.nvmlDeviceGetHandleByIndex(idx))
        return True
    except Exception:
        return False


def shard_by_cost(lengths: Sequence[int], world: int) -> List[List[int]]:
    """Deal utterances to ranks longest-first, round-robin (decode cost grows with T)."""
    order = sorted(range(len(lengths)), key=lambda i: (-int(lengths[i]), i))
    shards: List[List[int]] = [[] for _ in range(world)]
    for k, i in enumerate(order):
        # snake order keeps the per-rank frame totals within one utterance of each other
        r = k % (2 * world)
        shards[r if r < world else 2 * world - 1 - r].append(i)
    return shards


def gather_ragged(rows: List[List[int]], indices: List[int], total: int, group=None) -> Optional[List[List[int]]]:
    """Rank 0 receives every rank's (indices, rows) and returns them in original order."""
    if not dist.is_initialized() or dist.get_world_size(group) == 1:
        out: List[List[int]] = [[] for _ in range(total)]
        for i, r in zip(indices, rows):
            out[i] = r
        return out
    world, rank = dist.get_world_size(group), dist.get_rank(group)
    dev = torch.device("cuda", torch.cuda.current_device()) if dist.get_backend(group) == "nccl" else torch.device("cpu")
    counts = torch.tensor([len(r) for r in rows], dtype=torch.int32)
    flat = torch.tensor([x for r in rows for x in r], dtype=torch.int32)
    idx = torch.tensor(indices, dtype=torch.int32)
    meta = torch.tensor([len(rows), int(flat.numel())], dtype=torch.int64, device=dev)
    metas = [torch.zeros_like(meta) for _ in range(world)]
    dist.all_gather(metas, meta, group=group)
    n_max = max(int(m[0]) for m in metas)
    f_max = max(int(m[1]) for m in metas)
    pack = torch.zeros(2 * n_max + f_max, dtype=torch.int32, device=dev)
    pack[: len(rows)] = idx.to(dev)
    pack[n_max : n_max + len(rows)] = counts.to(dev)
    pack[2 * n_max : 2 * n_max + flat.numel()] = flat.to(dev)
    bufs = [torch.zeros_like(pack) for _ in range(world)] if rank == 0 else None
    dist.gather(pack, bufs, dst=0, group=group)
    if rank != 0:
        return None
    out = [[] for _ in range(total)]
    for m, buf in zip(metas, bufs):
        n = int(m[0])
        buf = buf.cpu()
        ids, cnt = buf[:n].tolist(), buf[n_max : n_max + n].tolist()
        flat_r = buf[2 * n_max : 2 * n_max + int(m[1])].tolist()
        lo = 0
        for i, c in zip(ids, cnt):
            out[i] = flat_r[lo : lo + c]
            lo += c
    return out


def sharded_decode(decode_ids: Callable[[torch.Tensor, torch.Tensor], List[List[int]]], log_probs: torch.Tensor,
                   lengths: torch.Tensor, group=None) -> Optional[List[List[int]]]:
    """Every rank holds the same (B,T,V) host tensor, decodes its share, rank 0 gets all rows."""
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    mine = shard_by_cost(lengths.tolist(), world)[rank]
    sel = torch.tensor(mine, dtype=torch.long)
    rows = decode_ids(log_probs[sel], lengths[sel]) if mine else []
    return gather_ragged(rows, mine, int(lengths.numel()), group)


def _loss_parts(nll: torch.Tensor, target_lengths: torch.Tensor) -> torch.Tensor:
    # everything stays on the stream: no host-side scalar is copied in (that would synchronise the step)
    if nll.is_cuda and nll.dim() == 1:
        from . import ops

        return ops.loss_parts(nll, target_lengths)  # one launch for both scalars
    s = (nll / target_lengths.to(nll.device).clamp_min(1)).sum(dtype=torch.float64)  # host tensors (gloo tests)
    return torch.stack([s, torch.full_like(s, float(nll.numel()))])


def global_mean_loss(nll: torch.Tensor, target_lengths: torch.Tensor, group=None) -> Tuple[torch.Tensor, int]:
    """'mean' reduction over the GLOBAL batch from local per-utterance nll. One all-reduce of 2 scalars."""
    part = _loss_parts(nll, target_lengths)
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group)
    return (part[0] / part[1]).to(nll.dtype), int(round(float(part[1])))


def global_mean_loss_async(nll: torch.Tensor, target_lengths: torch.Tensor, group=None) -> Callable[[], torch.Tensor]:
    """Same reduction, issued on the collective's own stream: returns a function that waits for it and gives the
    mean.  The gradient never needs the reduced value (it is scaled by the global batch size in-kernel,
    ``mean_denominator``), so a training step calls this right after the loss launch, carries on with the backward
    pass of the encoder, and reads the number when it logs it."""
    part = _loss_parts(nll, target_lengths)
    work = None
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        work = dist.all_reduce(part, op=dist.ReduceOp.SUM, group=group, async_op=True)

    def result() -> torch.Tensor:
        if work is not None:
            work.wait()
        return (part[0] / part[1]).to(nll.dtype)

    return result
