"""Rocket (gimbal + variable mass, BASELINE configs[3]): oracle vs the unmodified reference, the kernel body
on the host, and the CUDA path, on the same golden fixtures."""
import glob
import os

import numpy as np
import pytest

from engines import GOLDEN, HostSimEngine, OracleEngine, build_model, landing_config, load_golden, make_cuda_engine, replay_landing, replay_vehicle

AVIARY = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "rocket_*.npz")))
ENVS = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "landing_*.npz")))
# fp32 tolerances: the drop fixtures fall ~250 m at the 100 m/s clamp; ang vel / euler in rad
TOL = dict(pos=2e-3, euler=2e-4, linvel=2e-3, angvel=2e-4, aux=1e-5)


def test_fixtures_present():
    assert len(AVIARY) >= 5 and len(ENVS) >= 4


@pytest.mark.parametrize("name", AVIARY)
def test_oracle_reproduces_reference(name):
    err = replay_vehicle(OracleEngine, load_golden(name))
    assert err["contact_mismatch"] == 0
    for k in ("pos", "euler", "angvel", "linvel", "aux"):
        assert err[k] < 1e-9, (name, k, err[k])


@pytest.mark.parametrize("name", ENVS)
def test_oracle_reproduces_reference_landing_env(name):
    err = replay_landing(OracleEngine, load_golden(name))
    assert err["flag_mismatch"] == 0 and err["obs"] < 1e-9 and err["reward"] < 1e-9, err


@pytest.mark.parametrize("name", AVIARY)
def test_kernel_body_on_host(name):
    err = replay_vehicle(HostSimEngine, load_golden(name))
    assert err["contact_mismatch"] == 0
    for k, tol in TOL.items():
        assert err[k] < tol, (name, k, err[k])


@pytest.mark.gpu
@pytest.mark.parametrize("name", AVIARY)
def test_cuda_matches_reference(name):
    err = replay_vehicle(make_cuda_engine, load_golden(name), every=2)
    assert err["contact_mismatch"] == 0
    for k, tol in TOL.items():
        assert err[k] < tol, (name, k, err[k])


@pytest.mark.gpu
@pytest.mark.parametrize("name", ENVS)
def test_cuda_landing_env_matches_reference(name):
    err = replay_landing(make_cuda_engine, load_golden(name))
    assert err["flag_mismatch"] == 0
    assert err["obs"] < 5e-3 and err["reward"] < 5e-3, err  # positions are O(400 m) fp32 numbers


@pytest.mark.gpu
def test_cuda_batch_16384_matches_oracle():
    """BASELINE configs[3]: 16 384 rockets in a randomised accelerated drop, oracle vs CUDA through the env."""
    n, steps = 16384, 40
    rng = np.random.default_rng(4)
    f = lambda a: a.astype(np.float32).astype(np.float64)  # noqa: E731
    model = build_model("rocket", "rocket", starting_fuel_ratio=0.05)
    env = landing_config("quaternion", False, False, True)
    start = f(np.column_stack([rng.uniform(-20, 20, n), rng.uniform(-20, 20, n), rng.uniform(400, 450, n)]))
    orn = f(rng.uniform(-0.3, 0.3, (n, 3)))
    orc, cud = OracleEngine(model, env, n, start, orn), make_cuda_engine(model, env, n, start, orn)
    nz0 = f(rng.normal(1.0, 1.0, (20, n)))
    o0, o1 = orc.env_reset(nz0), cud.env_reset(nz0)
    assert np.abs(o0 - o1).max() < 2e-3
    for k in range(steps):
        act = f(rng.uniform([-1, -1, -1, 0, 0, -1, -1], [1, 1, 1, 1, 1, 1, 1], (n, 7)))
        nz = f(rng.normal(1.0, 1.0, (6, n)))
        ob0, r0, te0, tr0, in0 = orc.env_step(act, nz)
        ob1, r1, te1, tr1, in1 = cud.env_step(act, nz)
        assert np.array_equal(te0, te1) and np.array_equal(tr0, tr1) and np.array_equal(in0, in1), k
        assert np.abs(ob0 - ob1).max() < 5e-3 and np.abs(r0 - r1).max() < 5e-3, k


@pytest.mark.gpu
def test_cuda_landing_autoreset_and_determinism():
    import torch

    from pyflyt_b200.gym_envs.rocket_landing_env import RocketLandingVecEnv

    def run():
        env = RocketLandingVecEnv(num_envs=16384, seed=3)
        obs, _ = env.reset()
        z0 = obs[:, 12].clone()
        done_total = 0
        for _ in range(250):
            env.rollout(1)
            done_total += int((env.aviary.term.bool() | env.aviary.trunc.bool()).sum())
        torch.cuda.synchronize()
        out = (env.aviary.obs.clone(), env.aviary.reward.clone(), env.aviary.state_tensor.clone())
        env.close()
        return out, done_total, z0

    (a, da, z0), (b, db, _) = run(), run()
    for x, y in zip(a, b):
        assert torch.equal(x, y)
    assert torch.isfinite(a[0]).all() and da == db and da > 100  # crashes happen; random throttle slows most drops
    assert float(z0.min()) > 380.0 and float(z0.max()) < 455.0  # randomize_drop: U(0.8, 0.9) * ceiling minus the warm-up fall


@pytest.mark.gpu
@pytest.mark.parametrize("randomize", [True, False])
def test_landing_spare_reset_equals_inline_reset(randomize):
    """Rocket-Landing autoreset copies the env's spare (rebuilt on a side stream; a randomised drop is keyed by the
    episode number); it must equal integrating every warm-up inside the step launch bit for bit."""
    import torch

    from pyflyt_b200.gym_envs import RocketLandingVecEnv

    outs = []
    for inline in (False, True):
        env = RocketLandingVecEnv(num_envs=4096, seed=7, inline_reset=inline, max_duration_seconds=0.4, randomize_drop=randomize)
        env.reset()
        resets, trace = 0, []
        for k in range(70):
            env.rollout(1)
            resets += int((env.aviary.term | env.aviary.trunc).sum())
            trace.append(env.aviary.obs.sum().item())
        torch.cuda.synchronize()
        outs.append((env.aviary.obs.clone(), env.aviary.reward.clone(), env.aviary.state_tensor.clone(), resets, trace))
        env.close()
    a, b = outs
    assert a[3] > 4096 and a[3] == b[3] and a[4] == b[4]
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2])
