"""Experiment: distribution of the cold (L2-flushed) k_hover_step time, spare-copy vs inline resets, write- vs read-flush."""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

dev = torch.device("cuda:0")
n = 65536
buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

def run(env, K, flush):
    av = env.aviary
    for _ in range(30):
        env.rollout(1)
    av.profile_begin(K)
    for k in range(K):
        if flush == "write":
            buf.fill_(float(k))
        elif flush == "read":
            buf.sum()
        env.rollout(1)
    torch.cuda.synchronize()
    ms = sorted(av.profile_read(K))
    av.profile_begin(0)
    q = lambda f: ms[int(f * (len(ms) - 1))] * 1e3
    return f"min {q(0):5.1f}  p10 {q(.1):5.1f}  p50 {q(.5):5.1f}  p90 {q(.9):5.1f}  max {q(1):5.1f} us"

for inline in (0, 2, 1):
    env = QuadXHoverVecEnv(num_envs=n, seed=0, device=dev, inline_reset=inline)
    env.reset()
    for flush in ("write", "read", "none"):
        print(f"inline_reset={inline!s:5s} flush={flush:5s} {run(env, 200, flush)}", flush=True)
    env.close()
