"""Experiment: QuadX-Hover step time with M independent batches of 65 536 envs stepped round-robin, back to back (one event pair
around K launches, no flush kernel, no per-step events).  With (M - 1) x ~23 MB touched between two steps of the same batch above the
126 MB L2, every launch finds its inputs in DRAM ("inputs larger than L2").  One JSON line per M."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

    dev = torch.device("cuda:0")
    n, K = 65536, 240
    g = torch.Generator(device=dev).manual_seed(1)
    lo = torch.tensor([-3.14159265, -3.14159265, -3.14159265, 0.0], device=dev)
    hi = torch.tensor([3.14159265, 3.14159265, 3.14159265, 0.8], device=dev)
    for M in ([int(x) for x in os.environ["PFB_ROTATION_M"].split(",")] if os.environ.get("PFB_ROTATION_M") else (1, 2, 4, 6, 8, 12, 16, 24)):
        envs = [QuadXHoverVecEnv(num_envs=n, seed=0, device=dev, env_offset=j * n) for j in range(M)]
        acts = [lo + (hi - lo) * torch.rand((4, n, 4), device=dev, generator=g) for _ in range(M)]
        for e in envs:
            e.reset()
        for k in range(40 * M):
            envs[k % M].aviary.env_step(actions=acts[k % M][(k // M) % 4])
        best = []
        for rep in range(5):
            torch.cuda.synchronize()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for k in range(K):
                envs[k % M].aviary.env_step(actions=acts[k % M][(k // M) % 4])
            e1.record()
            torch.cuda.synchronize()
            best.append(e0.elapsed_time(e1) * 1e3 / K)
        best.sort()
        print(json.dumps({"lib": os.environ.get("PYFLYT_B200_LIB", "default")[-40:], "batches": M, "us_per_step_median": round(best[2], 3), "us_per_step_min": round(best[0], 3), "us_per_step_max": round(best[-1], 3),
                          "env_steps_per_s": n / (best[2] * 1e-6)}), flush=True)
        for e in envs:
            e.close()


if __name__ == "__main__":
    main()
