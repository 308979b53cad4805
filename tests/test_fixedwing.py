"""Fixedwing (lifting-surface aero path, BASELINE configs[2]): oracle vs the unmodified reference, the
kernel body on the host, and — on the GPU box — the CUDA path, all on the same golden fixtures."""
import glob
import os

import numpy as np
import pytest

from engines import GOLDEN, HostSimEngine, OracleEngine, build_model, load_golden, make_cuda_engine, replay_vehicle, replay_waypoints, waypoints_config

AVIARY = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "fixedwing_*.npz")))
ENVS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "fwwp_*.npz")))


def test_fixtures_present():
    assert len(AVIARY) >= 6 and len(ENVS) >= 3


@pytest.mark.parametrize("name", AVIARY)
def test_oracle_reproduces_reference(name):
    err = replay_vehicle(OracleEngine, load_golden(name))
    assert err["contact_mismatch"] == 0
    for k in ("pos", "euler", "angvel", "linvel", "aux"):
        assert err[k] < 1e-10, (name, k, err[k])


@pytest.mark.parametrize("name", ENVS)
def test_oracle_reproduces_reference_waypoints_env(name):
    err = replay_waypoints(OracleEngine, load_golden(name))
    assert err["flag_mismatch"] == 0 and err["obs"] < 1e-10 and err["reward"] < 1e-10, err


@pytest.mark.parametrize("name", AVIARY)
def test_kernel_body_on_host(name):
    err = replay_vehicle(HostSimEngine, load_golden(name))
    assert err["contact_mismatch"] == 0
    assert err["pos"] < 2e-4 and err["euler"] < 1e-4 and err["linvel"] < 1e-3, (name, err)


class _FullBlockHostEngine(HostSimEngine):
    """the kernel body with the one-basic-block surface code (fixedwing_substep<FULL>): what the step kernels run when the model has
    all five surfaces and there is no wind"""

    name = "hostsim-full"
    full_block = True


@pytest.mark.parametrize("name", [n for n in AVIARY if "wind" not in n])
def test_full_block_kernel_body_on_host(name):
    err = replay_vehicle(_FullBlockHostEngine, load_golden(name))
    assert err["contact_mismatch"] == 0
    assert err["pos"] < 2e-4 and err["euler"] < 1e-4 and err["linvel"] < 1e-3, (name, err)


@pytest.mark.parametrize("vehicle", ["fixedwing", "acrowing"])
def test_full_block_equals_generic_surface_code_on_host(vehicle):
    """FULL only removes the launch-uniform per-surface tests at compile time: same arithmetic, and with g++ the two instantiations
    agree bit for bit (64 aircraft, 240 Aviary steps through stalls and recoveries)."""
    n, steps = 64, 240
    rng = np.random.default_rng(11)
    f = lambda a: a.astype(np.float32).astype(np.float64)  # noqa: E731
    model = build_model("fixedwing", vehicle)
    start = f(np.column_stack([rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.uniform(40, 60, n)]))
    orn = f(rng.uniform(-0.3, 0.3, (n, 3)))
    noise = f(rng.normal(1.0, 1.0, (steps * 2, n)))
    eng = [HostSimEngine(model, None, n, start, orn), _FullBlockHostEngine(model, None, n, start, orn)]
    for e in eng:
        e.reset()
        e.set_mode(0)
    for i in range(0, steps, 30):
        sp = f(np.column_stack([rng.uniform(-0.8, 0.8, (n, 3)), rng.uniform(0.2, 1.0, n)]))
        for e in eng:
            e.set_setpoints(sp)
            e.aviary_step(noise[2 * i : 2 * i + 60], n_steps=30)
    a, b = eng[0].state(), eng[1].state()
    assert np.array_equal(a, b), np.abs(a - b).max()
    assert np.array_equal(eng[0].aux(), eng[1].aux())


@pytest.mark.gpu
@pytest.mark.parametrize("name", AVIARY)
def test_cuda_matches_reference(name):
    err = replay_vehicle(make_cuda_engine, load_golden(name), every=2)
    assert err["contact_mismatch"] == 0
    assert err["pos"] < 2e-4 and err["euler"] < 1e-4 and err["linvel"] < 1e-3, (name, err)


@pytest.mark.gpu
@pytest.mark.parametrize("name", ENVS)
def test_cuda_waypoints_env_matches_reference(name):
    err = replay_waypoints(make_cuda_engine, load_golden(name))
    assert err["flag_mismatch"] == 0
    assert err["obs"] < 2e-3 and err["reward"] < 2e-3, err  # target deltas are O(100 m) fp32 numbers


@pytest.mark.gpu
def test_cuda_batch_16384_matches_oracle():
    """BASELINE configs[2]: 16 384 aircraft, seeded inputs through oracle and CUDA (Aviary level, mode 0)."""
    n, steps = 16384, 120
    rng = np.random.default_rng(3)
    f = lambda a: a.astype(np.float32).astype(np.float64)  # noqa: E731
    model = build_model("fixedwing", "fixedwing")
    start = f(np.column_stack([rng.uniform(-5, 5, n), rng.uniform(-5, 5, n), rng.uniform(40, 60, n)]))
    orn = f(rng.uniform(-0.2, 0.2, (n, 3)))
    noise = f(rng.normal(1.0, 1.0, (steps * 2, n)))
    eng = [OracleEngine(model, None, n, start, orn), make_cuda_engine(model, None, n, start, orn)]
    for e in eng:
        e.reset()
        e.set_mode(0)
    for i in range(0, steps, 30):
        sp = f(np.column_stack([rng.uniform(-0.5, 0.5, (n, 3)), rng.uniform(0.3, 1.0, n)]))
        for e in eng:
            e.set_setpoints(sp)
            e.aviary_step(noise[2 * i : 2 * i + 60], n_steps=30)
    a, b = eng[0].state(), eng[1].state()
    assert np.abs(a[:, 3] - b[:, 3]).max() < 5e-4
    assert np.abs(a[:, 0] - b[:, 0]).max() < 1e-3
    assert np.abs(eng[0].aux() - eng[1].aux()).max() < 1e-5


@pytest.mark.gpu
def test_cuda_waypoints_autoreset_and_determinism():
    import torch

    from pyflyt_b200.gym_envs.fixedwing_waypoints_env import FixedwingWaypointsVecEnv

    def run():
        env = FixedwingWaypointsVecEnv(num_envs=16384, seed=7, goal_reach_distance=30.0)
        env.reset()
        reached = 0
        for _ in range(60):
            env.rollout(1)
            reached = max(reached, int(env._info()["num_targets_reached"].max()))
        torch.cuda.synchronize()
        out = (env.aviary.obs.clone(), env.aviary.reward.clone(), env.aviary.state_tensor.clone())
        env.close()
        return out, reached

    (a, ra), (b, rb) = run(), run()
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert torch.isfinite(a[0]).all() and ra >= 1 and ra == rb


@pytest.mark.gpu
def test_waypoints_spare_reset_equals_inline_reset():
    """Fixedwing-Waypoints autoreset copies the env's spare (state + targets of the next episode, rebuilt on a side
    stream); it must equal integrating every warm-up inside the step launch bit for bit."""
    import torch

    from pyflyt_b200.gym_envs import FixedwingWaypointsVecEnv

    outs = []
    for inline in (False, True):
        env = FixedwingWaypointsVecEnv(num_envs=4096, seed=7, inline_reset=inline, max_duration_seconds=0.5, goal_reach_distance=30.0)
        env.reset()
        resets, trace = 0, []
        for k in range(80):
            env.rollout(1)
            resets += int((env.aviary.term | env.aviary.trunc).sum())
            trace.append(env.aviary.obs.sum().item())
            if k == 30:
                env.aviary.start_pos[::2, 2] += 5.0  # stale spares must be ignored
        torch.cuda.synchronize()
        outs.append((env.aviary.obs.clone(), env.aviary.reward.clone(), env.aviary.state_tensor.clone(), resets, trace))
        env.close()
    a, b = outs
    assert a[3] > 4096 and a[3] == b[3] and a[4] == b[4]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
