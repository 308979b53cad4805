"""debug: Rocket-Landing timed path, low ceiling: where do GPU and oracle differ on crash steps?"""
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

n, steps, seed, ceiling = 4096, 60, 4242, 120.0
f = lambda a: np.ascontiguousarray(a, dtype=np.float32).astype(np.float64)  # noqa: E731
for cr in (True, False):
    env = RocketLandingVecEnv(num_envs=n, seed=seed, ceiling=ceiling, max_duration_seconds=30.0, autoreset=False, contact_response=cr)
    av = env.aviary
    streams = Streams(seed, n, noise_loc=1.0)
    model = build_model("rocket", "rocket", starting_fuel_ratio=0.05)
    cfg = landing_config("quaternion", False, False, True, ceiling=ceiling, max_duration=30.0, contact_response=cr)
    sp, so = streams.drop_poses(0x80000000, ceiling, 200.0)
    orc = OracleEngine(model, cfg, n, sp.astype(np.float64), so.astype(np.float64))
    og0, _ = env.reset()
    oo0 = orc.o.env_reset(noise=streams.user_reset_noise(0).astype(np.float64))
    print("contact_response", cr, "reset diff", np.abs(og0.double().cpu().numpy() - oo0).max())
    rng = np.random.default_rng(3)
    done = np.zeros(n, dtype=bool)
    prev_g, prev_o = og0.double().cpu().numpy(), oo0
    for k in range(steps):
        act = f(rng.uniform([-1, -1, -1, 0, 0, -1, -1], [1, 1, 1, 1, 1, 1, 1], (n, 7)))
        env.step(torch.as_tensor(act, dtype=torch.float32, device=av.device))
        og, rg = av.obs.double().cpu().numpy(), av.reward.double().cpu().numpy()
        teg, ig = av.term.cpu().numpy().astype(bool), av.info_bits.cpu().numpy()
        oo, ro, teo, tro, io = orc.o.env_step(act, streams.step_noise(k, 3).astype(np.float64))
        live = ~done
        d = np.abs(og - oo)
        d[~live] = 0
        crash = live & teg
        if crash.any():
            w = np.unravel_index(np.argmax(d * crash[:, None]), d.shape)
            nc = ~crash & live
            print(f"step {k}: {int(crash.sum())} crashes (oracle {int((live & teo.astype(bool)).sum())}), worst crash-step obs diff {d[crash].max():.3e} at env {w[0]} col {w[1]}; "
                  f"non-crash worst {d[nc].max() if nc.any() else 0:.3e}; reward diff crash {np.abs(rg - ro)[crash].max():.3e}")
            if d[crash].max() > 1e-2:
                i = w[0]
                np.set_printoptions(precision=5, suppress=True, linewidth=200)
                print("  prev gpu", prev_g[i][:13]); print("  prev orc", prev_o[i][:13])
                print("  now  gpu", og[i][:13], rg[i], ig[i]); print("  now  orc", oo[i][:13], ro[i], io[i])
        done |= teg | teo.astype(bool)
        prev_g, prev_o = og, oo
    env.close()
