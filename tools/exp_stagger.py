"""Experiment: cold QuadX-Hover step launch (65 536 envs) with (a) a staggered start of the tiles (PFB_HOVER_STAGGER=ns,mod) and
(b) the L2 flushed by WRITING 256 MiB (bench.py's protocol: the L2 is then full of dirty lines, every miss of the step evicts one)
vs by READING 256 MiB (clean lines).  One JSON line per setting: p10 / p50 / p90 of the library's event pair around the launch."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def q(ms, f):
    ms = sorted(ms)
    return round(ms[int(f * (len(ms) - 1))] * 1e3, 2)


def main():
    from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

    dev = torch.device("cuda:0")
    n, K = 65536, 300
    buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
    g = torch.Generator(device=dev).manual_seed(1)
    lo = torch.tensor([-3.14159265, -3.14159265, -3.14159265, 0.0], device=dev)
    hi = torch.tensor([3.14159265, 3.14159265, 3.14159265, 0.8], device=dev)
    actions = lo + (hi - lo) * torch.rand((16, n, 4), device=dev, generator=g)
    settings = [("0,1", "write"), ("0,1", "read"), ("0,1", "none"), ("1000,2", "write"), ("2000,2", "write"), ("3000,2", "write"), ("4000,2", "write"),
                ("1000,3", "write"), ("1500,3", "write"), ("1000,4", "write"), ("500,8", "write"), ("2000,2", "read"), ("2000,2", "none")]
    for stag, flush in settings:
        os.environ["PFB_HOVER_STAGGER"] = stag
        env = QuadXHoverVecEnv(num_envs=n, seed=0, device=dev)
        env.reset()
        av = env.aviary
        for k in range(40):
            av.env_step(actions=actions[k % 16])
        av.profile_begin(K)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        torch.cuda.synchronize()
        e0.record()
        acc = torch.zeros((), device=dev)
        for k in range(K):
            if flush == "write":
                buf.fill_(float(k))
            elif flush == "read":
                acc += buf.sum()
            av.env_step(actions=actions[k % 16])
        e1.record()
        torch.cuda.synchronize()
        ms = av.profile_read(K)
        av.profile_begin(0)
        chk = float(av.obs.double().sum())
        env.close()
        print(json.dumps({"stagger": stag, "flush": flush, "p10": q(ms, 0.1), "p50": q(ms, 0.5), "p90": q(ms, 0.9), "obs_checksum": chk}), flush=True)


if __name__ == "__main__":
    main()
