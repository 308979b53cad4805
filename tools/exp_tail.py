"""Experiment: how much of k_hover_step's duration is the autoreset tail (10 warm-up Aviary steps per finished env)?"""
import os, sys
import torch
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

dev = torch.device("cuda:0")
n = 65536
flush = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)

def timeit(env, actions, K=100, do_flush=True):
    av = env.aviary
    for _ in range(10):
        av.env_step(actions=actions)
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(K)]
    done = 0
    for k in range(K):
        if do_flush:
            flush.fill_(float(k))
        ev[k][0].record()
        av.env_step(actions=actions)
        ev[k][1].record()
        done += int((av.term | av.trunc).sum()) if k % 10 == 0 else 0
    torch.cuda.synchronize()
    ms = sorted(a.elapsed_time(b) for a, b in ev)
    return ms[len(ms) // 2] * 1e3, done / (K / 10)

for label, autoreset in (("autoreset", True), ("no-autoreset", False)):
    env = QuadXHoverVecEnv(num_envs=n, seed=0, device=dev, autoreset=autoreset)
    env.reset()
    lo = torch.tensor([-3.14159265] * 3 + [0.0], device=dev); hi = torch.tensor([3.14159265] * 3 + [0.8], device=dev)
    rand = lo + (hi - lo) * torch.rand((n, 4), device=dev)
    calm = torch.zeros((n, 4), device=dev); calm[:, 3] = 0.3
    for aname, act in (("random", rand), ("calm", calm)):
        env.reset()
        for fl in (True, False):
            us, done = timeit(env, act, do_flush=fl)
            print(f"{label:13s} actions={aname:6s} flush={fl!s:5s} median {us:7.2f} us/step   done/step ~{done:.0f}", flush=True)
    env.close()
