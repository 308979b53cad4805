"""Experiment: cost structure of k_hover_step (kernel time from the library's own CUDA events around the launch)."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

dev = torch.device("cuda:0")
n = 65536
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

def timeit(env, K=100, do_flush=True):
    av = env.aviary
    for _ in range(10):
        env.rollout(1)
    av.profile_begin(K)
    done = 0
    for k in range(K):
        if do_flush:
            flush.fill_(float(k))
        env.rollout(1)
        if k % 10 == 0:
            done += int((av.term | av.trunc).sum())
    torch.cuda.synchronize()
    ms = sorted(av.profile_read(K))
    av.profile_begin(0)
    return ms[len(ms) // 2] * 1e3, ms[0] * 1e3, done / (K / 10)

high = np.tile(np.array([[0.0, 0.0, 1000.0]]), (n, 1))
for label, kw in (("resets (default env)", dict()),
                  ("no terminations, autoreset on", dict(start_pos=high, flight_dome_size=1e9, max_duration_seconds=1e6)),
                  ("no terminations, autoreset off", dict(start_pos=high, flight_dome_size=1e9, max_duration_seconds=1e6, autoreset=False))):
    env = QuadXHoverVecEnv(num_envs=n, seed=0, device=dev, **kw)
    env.reset()
    for fl in (True, False):
        med, mn, done = timeit(env, do_flush=fl)
        print(f"{label:32s} flush={fl!s:5s} kernel median {med:6.2f} us  min {mn:6.2f} us   done/step ~{done:.0f}", flush=True)
    env.close()
