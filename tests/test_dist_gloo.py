"""World-size-2 gloo tests of the multi-GPU host logic (sharding, ragged gather, global mean)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from oracle import ctc_loss as ol
from oracle import greedy as og
from silent_speech_decoder_b200 import dist as sdist
from silent_speech_decoder_b200 import synth


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, q):
    os.environ.update(RANK=str(rank), WORLD_SIZE=str(world), LOCAL_RANK=str(rank), MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    r, w, _ = sdist.init_from_env(backend="gloo")
    assert (r, w) == (rank, world)
    # decode: every rank decodes its shard with a stand-in decoder (the oracle), rank 0 gets all rows
    lp = synth.peaked(11, 60, 5)
    ln = synth.ragged_lengths(11, 60, 6)
    rows = sdist.sharded_decode(lambda x, n: og.greedy_ids(x.numpy(), n.tolist(), 1), lp, ln)
    # loss: each rank owns a slice of the batch; the global mean needs only 2 scalars
    tg, tl = synth.targets(11, 8, 7)
    lo, hi = (0, 6) if rank == 0 else (6, 11)
    nll, _, _ = ol.torch_ctc(lp[lo:hi], ln[lo:hi], tg[lo:hi], tl[lo:hi], 1, True, torch.float64)
    mean, n = sdist.global_mean_loss(nll, tl[lo:hi])
    later = sdist.global_mean_loss_async(nll, tl[lo:hi])  # same number through the overlapped form
    assert float(later()) == float(mean)
    q.put((rank, rows, float(mean), n))
    dist.barrier()
    dist.destroy_process_group()


def test_two_rank_decode_gather_and_loss_mean():
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    got = {}
    for _ in range(2):
        rank, rows, mean, n = q.get(timeout=120)
        got[rank] = (rows, mean, n)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    lp = synth.peaked(11, 60, 5)
    ln = synth.ragged_lengths(11, 60, 6)
    tg, tl = synth.targets(11, 8, 7)
    assert got[1][0] is None
    assert got[0][0] == og.greedy_ids(lp.numpy(), ln.tolist(), 1)
    _, mean64, _ = ol.torch_ctc(lp, ln, tg, tl, 1, True, torch.float64)
    for r in (0, 1):
        assert abs(got[r][1] - float(mean64)) < 1e-12 and got[r][2] == 11


def test_shard_by_cost_balances_frames():
    ln = synth.ragged_lengths(1000, 500, 3).tolist()
    for world in (1, 2, 4, 8):
        shards = sdist.shard_by_cost(ln, world)
        assert sorted(i for s in shards for i in s) == list(range(1000))
        tot = [sum(ln[i] for i in s) for s in shards]
        assert max(tot) - min(tot) <= 500
        assert max(len(s) for s in shards) - min(len(s) for s in shards) <= 1


def test_gather_ragged_single_process():
    assert sdist.gather_ragged([[1, 2], []], [1, 0], 3) == [[], [1, 2], []]
