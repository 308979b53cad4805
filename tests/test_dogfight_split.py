"""The dogfight with an arena's two aircraft on different ranks (BASELINE configs[4], "NCCL all-gather"): host-side
partition logic on CPU (gloo, world_size 2), the split kernels vs the fused arena kernel on one GPU, and the NCCL run
on two GPUs (skipped where the box has one)."""
import os
import subprocess
import sys

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))


def test_split_agent_range_and_spawns():
    from pyflyt_b200.pz_envs import spawn_poses, split_agent_range

    assert split_agent_range(8192, 0, 8) == (0, 2048) and split_agent_range(8192, 7, 8) == (14336, 16384)
    with pytest.raises(ValueError):
        split_agent_range(3, 0, 4)
    pos, orn = spawn_poses(64, 10.0, 50.0, seed=1)
    assert pos.shape == (128, 3) and orn.shape == (128, 3)
    # members of an arena start opposite each other on the circle (pi / team_size apart), heading roughly outwards
    a0, a1 = np.arctan2(pos[:64, 1], pos[:64, 0]), np.arctan2(pos[64:, 1], pos[64:, 0])
    assert np.allclose(np.abs(np.angle(np.exp(1j * (a1 - a0)))), np.pi, atol=1e-9)
    r = np.hypot(pos[:, 0], pos[:, 1])
    assert r.min() >= 10.0 and r.max() <= 50.0 and pos[:, 2].min() >= 10.0 and pos[:, 2].max() <= 50.0
    assert np.all(np.angle(np.exp(1j * (orn[:64, 2] - a0))) >= -1e-9) and np.all(np.angle(np.exp(1j * (orn[:64, 2] - a0))) <= np.pi / 8 + 1e-9)


def _gloo_worker(rank, world, port, q):
    import torch
    import torch.distributed as dist

    from pyflyt_b200.pz_envs import split_agent_range

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    num_arenas = 6
    lo, hi = split_agent_range(num_arenas, rank, world)
    payload = torch.arange(lo, hi, dtype=torch.float32)[:, None].repeat(1, 20)  # row content = its global agent id
    table = torch.zeros(2 * num_arenas, 20)
    dist.all_gather_into_tensor(table, payload)
    # the opponent lookup of k_df_split_combat: pid = (1 - member) * num_arenas + arena
    ok = True
    for gid in range(lo, hi):
        member, arena = divmod(gid, num_arenas)
        pid = (1 - member) * num_arenas + arena
        ok &= float(table[pid, 0]) == pid and float(table[gid, 0]) == gid
    q.put((rank, ok))
    dist.destroy_process_group()


def test_gather_layout_gloo_world2():
    import torch.multiprocessing as mp

    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29641 + os.getpid() % 200
    procs = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in procs]
    res = sorted(q.get(timeout=120) for _ in procs)
    [p.join(60) for p in procs]
    assert res == [(0, True), (1, True)]


@pytest.mark.gpu
def test_split_matches_fused_arena_kernel():
    """Same spawns / actions / noise through the fused (shuffle) kernel and the split (payload table) kernels."""
    import torch

    sys.path.insert(0, HERE)
    from dist_dogfight_split import run, scenario
    from engines import build_model, dogfight_config, make_cuda_engine

    num_arenas, steps = 2048, 25
    s_obs, s_rew, s_term, env = run(num_arenas, steps, "cuda:0")
    assert env.collectives == 1 + 4 * steps
    pos, orn, nz0, acts, nz = scenario(num_arenas, steps)
    # fused layout: agents of arena g at rows 2g, 2g+1; split layout: member-major
    perm = (np.arange(2)[None, :] * num_arenas + np.arange(num_arenas)[:, None]).reshape(-1)
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    cfg = dogfight_config(1, False, lethal_distance=150.0, lethal_angle=1.0, damage_per_hit=0.05)
    fused = make_cuda_engine(build_model("fixedwing", "acrowing"), cfg, 2 * num_arenas, f64(pos[perm].astype(np.float32)), f64(orn[perm].astype(np.float32)))
    o = fused.env_reset(f64(nz0[:, perm]))
    assert np.abs(o - s_obs[0].cpu().numpy()[perm]).max() < 1e-4
    for k in range(steps):
        ob, r, te, tr, _ = fused.env_step(f64(acts[k][perm]), f64(nz[k][:, perm]))
        so, sr, st = s_obs[k + 1].cpu().numpy()[perm], s_rew[k].cpu().numpy()[perm], s_term[k].cpu().numpy()[perm]
        bad = (te != st) | (np.abs(r - sr) > 1e-2 + 1e-4 * np.abs(r))
        assert bad.mean() < 1e-3, (k, int(bad.sum()))
        assert np.abs(ob[~bad] - so[~bad]).max() < 1e-3, k
    assert int(s_term[-1].sum()) > 0
    assert torch.isfinite(s_obs).all()


@pytest.mark.gpu
def test_split_matches_oracle():
    """The split kernels (k_df_split_physics -> payload table -> k_df_split_combat) against the fp64 CPU oracle DIRECTLY (not
    through the fused kernel): same spawns / actions / injected noise, tolerances of tests/test_dogfight.py."""
    sys.path.insert(0, HERE)
    from dist_dogfight_split import run, scenario
    from engines import OracleEngine, build_model, dogfight_config

    num_arenas, steps = 2048, 30
    s_obs, s_rew, s_term, env = run(num_arenas, steps, "cuda:0")
    pos, orn, nz0, acts, nz = scenario(num_arenas, steps)
    perm = (np.arange(2)[None, :] * num_arenas + np.arange(num_arenas)[:, None]).reshape(-1)  # oracle: arena-major; split: member-major
    f64 = lambda a: np.ascontiguousarray(a, dtype=np.float64)  # noqa: E731
    cfg = dogfight_config(1, False, lethal_distance=150.0, lethal_angle=1.0, damage_per_hit=0.05)
    orc = OracleEngine(build_model("fixedwing", "acrowing"), cfg, 2 * num_arenas, f64(pos[perm].astype(np.float32)), f64(orn[perm].astype(np.float32)))
    o = orc.env_reset(f64(nz0[:, perm]))
    assert np.abs(o - s_obs[0].cpu().numpy()[perm]).max() < 2e-3
    hits = 0
    for k in range(steps):
        ob, r, te, tr, _ = orc.env_step(f64(acts[k][perm]), f64(nz[k][:, perm]))
        so, sr, st = s_obs[k + 1].cpu().numpy()[perm], s_rew[k].cpu().numpy()[perm], s_term[k].cpu().numpy()[perm]
        # a hit / range decision within fp32 rounding of its threshold may flip in a handful of arenas
        bad = (te != st) | (np.abs(r - sr) > 0.05 + 1e-3 * np.abs(r))
        assert bad.mean() < 2e-3, (k, int(bad.sum()))
        assert np.abs(ob[~bad] - so[~bad]).max() < 2e-2, k
        hits += int((ob[:, 18] < 1.0).sum())
    assert hits > 0 and int(s_term[-1].sum()) > 0


@pytest.mark.gpu
@pytest.mark.parametrize("exchange", ["nccl", "peer", "peer-signal"])
def test_split_two_ranks(exchange):
    """Two ranks over NVLink: NCCL all-gather between the kernels, or the exchange fused into the physics kernel (peer
    stores into symmetric memory + barrier); both must equal the single-rank run bit for bit."""
    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs (gpurun --gpus 2)")
    cmd = [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node", "2", "--master-addr", "127.0.0.1",
           "--master-port", {"nccl": "29533", "peer": "29534", "peer-signal": "29535"}[exchange], os.path.join(HERE, "dist_dogfight_split.py")]
    out = subprocess.run(cmd, capture_output=True, text=True, timeout=600, env=dict(os.environ, PFB_SPLIT_EXCHANGE=exchange))
    assert out.returncode == 0 and f"SPLIT_OK world=2 exchange={exchange}" in out.stdout, out.stdout[-2000:] + out.stderr[-2000:]


@pytest.mark.gpu
def test_split_peer_exchange_single_rank_equals_nccl_path():
    """world = 1: the peer-store path (own table as the only peer, double-buffered) against the payload + copy path."""
    import torch

    sys.path.insert(0, HERE)
    from dist_dogfight_split import run

    a = run(1024, 20, "cuda:0", exchange="nccl")
    for ex in ("peer", "peer-signal"):
        b = run(1024, 20, "cuda:0", exchange=ex)
        assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]), ex
