"""MAQuadXHover (pz_envs/quadx_envs/ma_quadx_hover_env.py): the oracle against the unmodified reference
PettingZoo env (fixtures from tools/gen_golden.py mahover), then the CUDA per-agent kernel and the arena bookkeeping."""
import glob
import os

import numpy as np
import pytest

from engines import GOLDEN, OracleEngine, load_golden, make_cuda_engine, replay_ma_hover

FIX = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "mahover_*.npz")))


def test_fixtures_present():
    assert len(FIX) >= 4


@pytest.mark.parametrize("name", FIX)
def test_oracle_reproduces_reference(name):
    err = replay_ma_hover(OracleEngine, load_golden(name))
    assert err["flag_mismatch"] == 0, err
    assert err["obs"] < 1e-7 and err["reward"] < 1e-6, err


@pytest.mark.gpu
@pytest.mark.parametrize("name", FIX)
def test_cuda_matches_reference(name):
    err = replay_ma_hover(make_cuda_engine, load_golden(name))
    assert err["flag_mismatch"] == 0, err
    if "mode6" in name:  # the reference's z-velocity PID limit-cycles on cf2x: pin the position envelope
        assert err["pos"] < 1e-3 and err["obs"] < 0.1 and err["reward"] < 5e-2, err
    else:
        assert err["obs"] < 2e-4 and err["reward"] < 2e-3, err


@pytest.mark.gpu
def test_arena_env_follows_the_reference_episode():
    """The PettingZoo bookkeeping of MAQuadXHoverVecEnv (culling, zero actions for culled agents) against a fixture's first
    episode, with the reference's own noise draws."""
    import torch

    from pyflyt_b200.pz_envs import MAQuadXHoverVecEnv

    g = load_golden("mahover_mode0")
    A = int(g["n_agents"])
    env = MAQuadXHoverVecEnv(num_arenas=1, start_pos=g["start_pos"], start_orn=g["start_orn"], autoreset=False)
    dev = env.device
    t = lambda a: torch.as_tensor(np.ascontiguousarray(a, dtype=np.float32), device=dev)  # noqa: E731
    obs, _ = env.reset(noise=t(g["ep0_reset_noise"].reshape(-1, A)))
    assert np.abs(obs.double().cpu().numpy() - g["ep0_reset_obs"]).max() < 2e-4
    for i in range(len(g["ep0_actions"])):
        alive_ref = g["ep0_alive"][i]
        assert np.array_equal(env.alive.cpu().numpy(), alive_ref), i
        ob, r, te, tr, _ = env.step(t(g["ep0_actions"][i]), noise=t(g["ep0_noise"][i].reshape(-1, A)))
        ob, r = ob.double().cpu().numpy(), r.double().cpu().numpy()
        assert np.abs(ob[alive_ref] - g["ep0_obs"][i][alive_ref]).max() < 2e-4, i
        assert np.abs(r[alive_ref] - g["ep0_reward"][i][alive_ref]).max() < 2e-3, i
        assert np.array_equal(te.cpu().numpy()[alive_ref], g["ep0_term"][i][alive_ref]) and (r[~alive_ref] == 0).all()
    assert not bool(env.alive.any())  # the fixture's episode ends with every agent done
    env.close()


@pytest.mark.gpu
def test_arena_env_autoreset_and_determinism():
    import torch

    from pyflyt_b200.pz_envs import MAQuadXHoverVecEnv

    def run():
        env = MAQuadXHoverVecEnv(num_arenas=2048, seed=4, flight_dome_size=3.0)
        obs, _ = env.reset()
        assert obs.shape == (8192, 24)
        g = torch.Generator(device=env.device).manual_seed(0)
        lo = torch.tensor([-3.14, -3.14, -3.14, 0.0], device=env.device)
        hi = torch.tensor([3.14, 3.14, 3.14, 0.8], device=env.device)
        episodes = 0
        for _ in range(120):
            act = lo + (hi - lo) * torch.rand((8192, 4), device=env.device, generator=g)
            obs, rew, term, trunc, info = env.step(act)
            dead = ~info["alive"]
            assert bool((rew[dead & ~term] == 0).all())
            # an arena is either running (someone alive) or was just reset (everyone alive again, step counter 0)
            per = info["alive"].view(2048, 4)
            assert bool(per.any(dim=1).all())
            episodes += int((env.aviary.step_counts.view(2048, 4)[:, 0] == 0).sum())
        torch.cuda.synchronize()
        out = (obs.clone(), rew.clone(), env.aviary.state_tensor.clone(), episodes)
        env.close()
        return out

    a, b = run(), run()
    assert torch.equal(a[0], b[0]) and torch.equal(a[1], b[1]) and torch.equal(a[2], b[2]) and a[3] == b[3]
    assert a[3] > 2048 and torch.isfinite(a[0]).all()
