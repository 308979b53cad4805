"""debug: Rocket-Landing timed path at ceiling 120 — history of |gpu - oracle| for the worst env of the failing test"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from engines import OracleEngine, build_model, landing_config  # noqa: E402
from philox_replay import Streams  # noqa: E402

from pyflyt_b200.gym_envs.rocket_landing_env import RocketLandingVecEnv  # noqa: E402

n, steps, seed, ceiling, W = 16384, 166, 4242, 120.0, 15038
f = lambda a: np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)  # noqa: E731
env = RocketLandingVecEnv(num_envs=n, seed=seed, ceiling=ceiling, max_duration_seconds=30.0)
av = env.aviary
streams = Streams(seed, n, noise_loc=1.0)
model = build_model("rocket", "rocket", starting_fuel_ratio=0.05)
cfg = landing_config("quaternion", False, False, True, ceiling=ceiling, max_duration=30.0, contact_response=True)
sp, so = streams.drop_poses(0x80000000, ceiling, 200.0)
sp, so = sp.astype(np.float64), so.astype(np.float64)
orc = OracleEngine(model, cfg, n, sp, so)
env.reset()
orc.o.env_reset(noise=streams.user_reset_noise(0).astype(np.float64))
rng = np.random.default_rng(3)
episode = np.ones(n, dtype=np.int64)
done_prev = np.zeros(n, dtype=bool)
np.set_printoptions(precision=6, suppress=True, linewidth=250)
for k in range(steps):
    act = f(rng.uniform([-1, -1, -1, 0, 0, -1, -1], [1, 1, 1, 1, 1, 1, 1], (n, 7)))
    env.step(torch.as_tensor(act, dtype=torch.float32, device=av.device))
    og = av.obs.double().cpu().numpy()
    teg, trg, ig = av.term.cpu().numpy().astype(bool), av.trunc.cpu().numpy().astype(bool), av.info_bits.cpu().numpy()
    oo, ro, teo, tro, io = orc.o.env_step(act, streams.step_noise(k, 3).astype(np.float64))
    if done_prev.any():
        idx = np.nonzero(done_prev)[0]
        rz = np.zeros((20, n))
        rz[:, idx] = streams.autoreset_noise(episode[idx], envs=idx)
        p_, o_ = streams.drop_poses(episode[idx], ceiling, 200.0, envs=idx)
        sp[idx], so[idx] = p_, o_
        orc.o.set_start(sp, so)
        obs_r = orc.o.env_reset(mask=done_prev.astype(np.uint8), noise=rz)
        oo[done_prev] = obs_r[done_prev]
        episode[idx] += 1
    if k >= 120:
        d = np.abs(og[W] - oo[W])
        print(f"step {k} reset_now={bool(done_prev[W])} term={bool(teg[W])} info={int(ig[W])} max diff {d.max():.3e} col {int(d.argmax())} | angvel {og[W][0:3]} quat-diff {d[3:7].max():.2e} "
              f"vel gpu {og[W][7:10]} orc {oo[W][7:10]} pos {og[W][10:13]} posdiff {d[10:13].max():.2e} aux gpu {og[W][20:29]} auxdiff {d[20:29].max():.2e}")
    done_prev = teg | trg
env.close()
