"""N>1 host logic on CPU with the gloo backend, world_size 2 (no GPU needed)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from rl_collision_avoidance_b200.parallel import (allreduce_moments, average_gradients, broadcast_parameters,
                                                 normalize_from_moments, plan_minibatches, shard_worlds, agree_to_stop)


def test_shard_worlds_covers_everything_once():
    for total, ws in ((171, 1), (171, 2), (171, 8), (328, 8), (7, 8)):
        seen = []
        for r in range(ws):
            off, cnt = shard_worlds(total, r, ws)
            seen += list(range(off, off + cnt))
        assert seen == list(range(total))
    with pytest.raises(ValueError):
        shard_worlds(10, 3, 2)


def _free_port():
    s = socket.socket()
    s.bind(('127.0.0.1', 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world_size, port, out):
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        rs = np.random.RandomState(0)
        full = torch.from_numpy(rs.standard_normal(1000) * 2.0 + 0.5)
        off, cnt = shard_worlds(10, rank, world_size)          # 10 "worlds" of 100 advantages each
        local = full[off * 100:(off + cnt) * 100].float()
        mom = torch.tensor([local.double().sum(), (local.double() ** 2).sum(), float(local.numel())], dtype=torch.float64)
        allreduce_moments(mom)
        norm = normalize_from_moments(local, mom)
        ref = ((full - full.mean()) / full.std(unbiased=False))[off * 100:(off + cnt) * 100].float()
        ok_norm = bool((norm - ref).abs().max() < 1e-6)
        g = torch.full((16,), float(rank + 1))
        average_gradients(g)
        ok_grad = bool((g == 3.0).all())                       # 1 + 2, the 1/world_size goes into Adam's grad_scale
        p = torch.full((8,), float(rank))
        broadcast_parameters(p, src=0)
        ok_bcast = bool((p == 0).all())
        out.put((rank, ok_norm, ok_grad, ok_bcast))
    finally:
        dist.destroy_process_group()


def test_world_size_2_gloo_collectives():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = [q.get(timeout=120) for _ in procs]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    assert sorted(r[0] for r in res) == [0, 1]
    assert all(r[1] and r[2] and r[3] for r in res), res


def test_plan_minibatches_single_process():
    steps, sizes, weights = plan_minibatches(41, 16, drop_last=True)
    assert (steps, sizes, weights) == (2, [16, 16], [1.0, 1.0])
    steps, sizes, weights = plan_minibatches(41, 16, drop_last=False)
    assert (steps, sizes, weights) == (3, [16, 16, 9], [1.0, 1.0, 1.0])
    assert plan_minibatches(5, 16, drop_last=True)[0] == 0


def _sched_worker(rank, world_size, port, out):
    """Ranks with DIFFERENT row counts (stage 2: per-rank filter_index; model/ppo.py:212-223) must issue the same
    number of gradient all-reduces, and the weighted average must equal the gradient of the global minibatch."""
    os.environ['MASTER_ADDR'] = '127.0.0.1'
    os.environ['MASTER_PORT'] = str(port)
    dist.init_process_group('gloo', rank=rank, world_size=world_size)
    try:
        n_local = [70, 41][rank]
        rs = np.random.RandomState(3)
        rows_all = [torch.from_numpy(rs.standard_normal((70, 5))), torch.from_numpy(rs.standard_normal((41, 5)))]
        rows = rows_all[rank]
        res = {}
        for drop_last in (True, False):
            steps, sizes, weights = plan_minibatches(n_local, 16, drop_last, distributed=True)
            reduces = 0
            ok = True
            for i in range(steps):
                nb = sizes[i]
                # "gradient" of a mean loss over this rank's rows of the step, times the data-parallel weight
                g = rows[i * 16:i * 16 + nb].mean(0) * weights[i] if nb else torch.zeros(5, dtype=torch.float64)
                average_gradients(g)
                reduces += 1
                g = g / world_size
                union = torch.cat([r[i * 16:i * 16 + min(16, max(0, len(r) - i * 16))] for r in rows_all])
                ok = ok and bool((g - union.mean(0)).abs().max() < 1e-12)
            res[drop_last] = (steps, reduces, ok, sizes)
        stop = agree_to_stop(rank == 1)            # only rank 1 wants to stop: both must
        out.put((rank, res, stop))
    finally:
        dist.destroy_process_group()


def test_world_size_2_different_row_counts_same_collective_count():
    ctx = mp.get_context('spawn')
    q = ctx.Queue()
    port = _free_port()
    procs = [ctx.Process(target=_sched_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    res = sorted([q.get(timeout=120) for _ in procs])
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    (r0, a, stop0), (r1, b, stop1) = res
    assert stop0 and stop1
    assert a[True][:3] == (2, 2, True) and b[True][:3] == (2, 2, True)           # min(70 // 16, 41 // 16) full batches
    assert a[True][3] == [16, 16] and b[True][3] == [16, 16]
    assert a[False][:3] == (5, 5, True) and b[False][:3] == (5, 5, True)         # max(ceil(70/16), ceil(41/16))
    assert a[False][3] == [16, 16, 16, 16, 6] and b[False][3] == [16, 16, 9, 0, 0]


def test_grad_buckets_cover_the_flat_buffer_once():
    from rl_collision_avoidance_b200.parallel import grad_buckets
    # offsets as rlca_policy_param_offset lays them out: tensor starts padded to 32 floats, state_dict order
    sizes = [2, 480, 32, 3072, 32, 256 * 4096, 256, 128 * 260, 128, 128, 1, 128, 1,
             480, 32, 3072, 32, 256 * 4096, 256, 128 * 260, 128, 128, 1]
    off = [0]
    for n in sizes:
        off.append(off[-1] + (n + 31) // 32 * 32)
    early, late = grad_buckets(off)
    spans = sorted(early + late)
    assert spans[0][0] == 0 and spans[-1][1] == off[-1]
    assert all(a[1] == b[0] for a, b in zip(spans[:-1], spans[1:]))           # contiguous, no overlap
    assert sum(b - a for a, b in early) > 0.95 * off[-1]                     # the early ranges are the bulk
    for i in (1, 2, 3, 4, 13, 14, 15, 16):                                   # every conv tensor is in a late range
        assert any(a <= off[i] and off[i + 1] <= b for a, b in late)
