"""Multi-rank host logic on CPU: world_size-2 gloo process group (the N>1 path of bench.py minus CUDA)."""
import os
import socket

import numpy as np
import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from pyflyt_b200.sharding import reduce_step_stats, shard_range


def test_shard_range_partitions_exactly():
    for n in (1, 7, 64, 65536, 65537):
        for w in (1, 2, 3, 8):
            spans = [shard_range(n, r, w) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_range(8, 2, 2)


def _worker(rank, world, port, out):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port), RANK=str(rank), WORLD_SIZE=str(world))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        lo, hi = shard_range(1000, rank, world)
        # each rank "times" its shard; the job time is the max over ranks, the work is the sum
        t_max = reduce_step_stats([10.0 + rank], "max")[0]
        n_sum = reduce_step_stats([hi - lo], "sum")[0]
        # oracle shards agree with the unsharded run because noise/actions are keyed by GLOBAL env id
        from engines import OracleEngine, build_model

        m = build_model("quadx", "cf2x")
        n = 64
        g_lo, g_hi = shard_range(n, rank, world)
        rng = np.random.default_rng(0)
        start = np.column_stack([rng.uniform(-1, 1, n), rng.uniform(-1, 1, n), rng.uniform(5, 6, n)])
        noise = rng.normal(4.0, 1.0, (40, n))
        sp = rng.uniform([-1, -1, -1, 0.2], [1, 1, 1, 0.6], (n, 4))
        e = OracleEngine(m, None, g_hi - g_lo, start[g_lo:g_hi], np.zeros((g_hi - g_lo, 3)))
        e.reset()
        e.set_mode(0)
        e.set_setpoints(sp[g_lo:g_hi])
        e.aviary_step(noise[:, g_lo:g_hi], n_steps=20)
        mine = torch.from_numpy(e.state()[:, 3].copy())
        gathered = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(gathered, mine)  # equal shard sizes here (64 / 2)
        if rank == 0:
            full = OracleEngine(m, None, n, start, np.zeros((n, 3)))
            full.reset()
            full.set_mode(0)
            full.set_setpoints(sp)
            full.aviary_step(noise, n_steps=20)
            ok = bool(np.array_equal(torch.cat(gathered).numpy(), full.state()[:, 3]))
            out.put((t_max, n_sum, ok))
    finally:
        dist.destroy_process_group()


def test_two_rank_gloo_shards_reduce_and_match_unsharded():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in procs:
        p.start()
    for p in procs:
        p.join(120)
        assert p.exitcode == 0
    t_max, n_sum, ok = q.get(timeout=10)
    assert t_max == 11.0 and n_sum == 1000.0 and ok
