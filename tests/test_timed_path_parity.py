"""The instantiation bench.py times — Philox noise generated in-kernel, NEXT_STEP autoreset with spare post-reset states —
pinned to the fp64 CPU oracle.  The device streams are stateless (tests/philox_replay.py regenerates them from
(seed, env id, call number, tag)), so the oracle is driven in lock-step with exactly the noise the kernel drew, the actions
it was given and the same reset schedule, at BASELINE.json's full batch (65 536 envs) for 200 env steps (~1e5 autoresets)."""
import numpy as np
import pytest

from engines import OracleEngine, build_model, hover_config
from philox_replay import Streams

pytestmark = pytest.mark.gpu


def _f(a):
    return np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)


@pytest.mark.parametrize("n,steps,randact", [(65536, 200, False), (8192, 120, True)])
def test_hover_philox_autoreset_matches_oracle(n, steps, randact):
    import torch

    from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

    seed = 20240924
    env = QuadXHoverVecEnv(num_envs=n, seed=seed)  # autoreset=True, Philox noise: k_hover_step<0, false, false, true, false>
    av = env.aviary
    dump = torch.zeros((6, n), dtype=torch.float32, device=av.device)
    av.set_noise_dump(dump)
    streams = Streams(seed, n, noise_loc=4.0)
    model = build_model("quadx", "cf2x")
    orc = OracleEngine(model, hover_config(0, "quaternion", False, 3.0), n, np.tile([[0.0, 0.0, 1.0]], (n, 1)), np.zeros((n, 3)))

    obs_g, _ = env.reset()
    obs_o = orc.o.env_reset(noise=streams.user_reset_noise(0).astype(np.float64))
    assert np.abs(obs_g.double().cpu().numpy() - obs_o).max() < 1e-5

    rng = np.random.default_rng(1)
    episode = np.ones(n, dtype=np.int64)       # number of the spare an env takes at its next autoreset
    done_prev = np.zeros(n, dtype=bool)
    live = np.ones(n, dtype=bool)              # envs still compared (an env whose termination decision flipped is dropped)
    worst_obs = worst_rew = worst_noise = 0.0
    n_resets = n_done = n_flip = 0
    for k in range(steps):
        if randact:
            env.rollout(1)
            act = av.setpoints.double().cpu().numpy()  # the kernel writes the actions it drew back
            assert np.array_equal(act.astype(np.float32), streams.actions(k))
        else:
            act = _f(rng.uniform([-np.pi, -np.pi, -np.pi, 0.0], [np.pi, np.pi, np.pi, 0.8], (n, 4)))
            env.step(torch.as_tensor(act, dtype=torch.float32, device=av.device))
        og, rg = av.obs.double().cpu().numpy(), av.reward.double().cpu().numpy()
        teg, trg, ig = av.term.cpu().numpy().astype(bool), av.trunc.cpu().numpy().astype(bool), av.info_bits.cpu().numpy()
        # ---- oracle, same step: everyone steps on the replayed noise, then the envs that finished on the previous call are
        #      reset instead (NEXT_STEP): first observation of episode `episode[i]`, reward 0, flags cleared
        nz = streams.step_noise(k)
        oo, ro, teo, tro, io = orc.o.env_step(act, nz.astype(np.float64))
        teo, tro = teo.astype(bool), tro.astype(bool)
        if done_prev.any():
            rz = np.zeros((20, n))
            idx = np.nonzero(done_prev)[0]
            rz[:, idx] = streams.autoreset_noise(episode[idx], envs=idx)
            obs_r = orc.o.env_reset(mask=done_prev.astype(np.uint8), noise=rz)
            oo[done_prev], ro[done_prev], teo[done_prev], tro[done_prev], io[done_prev] = obs_r[done_prev], 0.0, False, False, 0
            episode[idx] += 1
            n_resets += len(idx)
        # ---- the kernel's own draws (dumped) vs the host replay, for the envs that ran all three Aviary steps
        full = ~done_prev & ~(teg | trg)
        worst_noise = max(worst_noise, float(np.abs(dump.cpu().numpy()[:, full] - nz[:, full]).max()))
        # ---- outcomes
        flip = live & ((teg != teo) | (trg != tro))
        if flip.any():  # a termination decided within rounding of its threshold: drop the env from the comparison
            n_flip += int(flip.sum())
            live &= ~flip
        cmp = live
        assert np.array_equal(ig[cmp] & 3, io[cmp] & 3), k
        worst_obs = max(worst_obs, float(np.abs(og[cmp] - oo[cmp]).max()))
        worst_rew = max(worst_rew, float(np.abs(rg[cmp] - ro[cmp]).max()))
        done_prev = teg | trg
        n_done += int(done_prev.sum())
        # (the oracle follows the kernel's reset schedule, which is its own for every env that is still compared)
    print(f"\n[timed-path parity] {n} envs x {steps} steps: {n_done} episodes finished, {n_resets} autoresets, flips {n_flip}; "
          f"max |obs| {worst_obs:.2e}, max |reward| {worst_rew:.2e}, max |noise dump - replay| {worst_noise:.2e}")
    assert n_resets > n  # every env was reset more than once on average
    assert worst_noise < 5e-4  # the device Box-Muller uses __logf / __sincosf / sqrt.approx: ~2e-4 absolute on N(4, 1) draws
    assert n_flip <= max(2, n // 4096)
    assert worst_obs < 1e-4 and worst_rew < 1e-4
    env.close()


def test_fixedwing_waypoints_philox_autoreset_matches_oracle():
    """The same pin for BASELINE configs[2]: k_fwwp_step with Philox noise (N(1, 1): one motor), device-drawn waypoints and
    NEXT_STEP autoreset through spare post-reset states, 16 384 aircraft x 150 env steps, against the oracle driven with the
    replayed noise AND the replayed waypoint draws (waypoint_handler.py:65-83)."""
    import torch

    from engines import waypoints_config
    from pyflyt_b200.gym_envs.fixedwing_waypoints_env import FixedwingWaypointsVecEnv

    n, steps, seed, T, dome = 16384, 150, 777, 4, 100.0
    env = FixedwingWaypointsVecEnv(num_envs=n, seed=seed, goal_reach_distance=25.0, max_duration_seconds=3.0)  # episodes of <= 90 steps
    av = env.aviary
    streams = Streams(seed, n, noise_loc=1.0)
    model = build_model("fixedwing", "fixedwing")
    cfg = waypoints_config("quaternion", False, T, 25.0, dome, max_duration=3.0)
    orc = OracleEngine(model, cfg, n, np.tile([[0.0, 0.0, 10.0]], (n, 1)), np.zeros((n, 3)))

    obs_g, _ = env.reset()
    tg = streams.waypoint_targets(0x80000000, T, dome)
    obs_o = orc.o.env_reset(noise=streams.user_reset_noise(0).astype(np.float64), targets=tg.astype(np.float64).reshape(n, -1))
    assert np.abs(obs_g.double().cpu().numpy() - obs_o).max() < 2e-3

    rng = np.random.default_rng(2)
    episode = np.ones(n, dtype=np.int64)
    done_prev = np.zeros(n, dtype=bool)
    live = np.ones(n, dtype=bool)
    worst_obs = worst_rew = 0.0
    n_resets = n_flip = reached = 0
    for k in range(steps):
        act = _f(rng.uniform(-1.0, 1.0, (n, 4)) * [0.5, 0.3, 0.3, 1.0])
        env.step(torch.as_tensor(act, dtype=torch.float32, device=av.device))
        og, rg = av.obs.double().cpu().numpy(), av.reward.double().cpu().numpy()
        teg, trg, ig = av.term.cpu().numpy().astype(bool), av.trunc.cpu().numpy().astype(bool), av.info_bits.cpu().numpy()
        oo, ro, teo, tro, io = orc.o.env_step(act, streams.step_noise(k, 4).astype(np.float64))
        teo, tro = teo.astype(bool), tro.astype(bool)
        if done_prev.any():
            idx = np.nonzero(done_prev)[0]
            rz = np.zeros((20, n))
            rz[:, idx] = streams.autoreset_noise(episode[idx], envs=idx)
            tgr = np.zeros((n, T, 3))
            tgr[idx] = streams.waypoint_targets(episode[idx], T, dome, envs=idx)
            obs_r = orc.o.env_reset(mask=done_prev.astype(np.uint8), noise=rz, targets=tgr.reshape(n, -1))
            oo[done_prev], ro[done_prev], teo[done_prev], tro[done_prev], io[done_prev] = obs_r[done_prev], 0.0, False, False, 0
            episode[idx] += 1
            n_resets += len(idx)
        flip = live & ((teg != teo) | (trg != tro) | ((ig >> 3) != (io >> 3)))  # termination / target-reached decisions at a threshold
        n_flip += int(flip.sum())
        live &= ~flip
        worst_obs = max(worst_obs, float(np.abs(og[live] - oo[live]).max()))
        worst_rew = max(worst_rew, float(np.abs(rg[live] - ro[live]).max()))
        reached = max(reached, int((ig >> 3).max()))
        done_prev = teg | trg
    print(f"\n[timed-path parity, fixedwing-waypoints] {n} envs x {steps} steps: {n_resets} autoresets, flips {n_flip}, max targets reached {reached}; "
          f"max |obs| {worst_obs:.2e}, max |reward| {worst_rew:.2e}")
    assert n_resets > n and reached >= 1
    assert n_flip <= n // 1000
    assert worst_obs < 5e-3 and worst_rew < 5e-3  # target deltas are O(100 m) fp32 numbers
    env.close()
