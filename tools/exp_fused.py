"""Experiment: QuadX-Hover synthetic rollout, fused launches (pfb_env_rollout(n >= 4): up to 16 env steps per launch + the spare top-up)
vs one launch per step, 65 536 envs, 12 rotating batches (inputs larger than the L2) and one batch.  One JSON line per setting."""
import json
import os
import sys

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

    dev = torch.device("cuda:0")
    n, K = 65536, 320
    for M in ([int(x) for x in os.environ["PFB_FUSED_M"].split(",")] if os.environ.get("PFB_FUSED_M") else (1, 12)):
        envs = [QuadXHoverVecEnv(num_envs=n, seed=0, device=dev, env_offset=j * n) for j in range(M)]
        for e in envs:
            e.reset()
        for chunk in ([int(x) for x in os.environ["PFB_FUSED_CHUNKS"].split(",")] if os.environ.get("PFB_FUSED_CHUNKS") else (1, 4, 8, 16, 32)):
            for k in range(2 * M):
                envs[k % M].rollout(max(chunk, 16))
            res = []
            for rep in range(5):
                torch.cuda.synchronize()
                e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                e0.record()
                for k in range(K // chunk):
                    envs[k % M].rollout(chunk)
                e1.record()
                torch.cuda.synchronize()
                res.append(e0.elapsed_time(e1) * 1e3 / (K // chunk * chunk))
            res.sort()
            print(json.dumps({"lib": os.environ.get("PYFLYT_B200_LIB", "default")[-36:], "batches": M, "steps_per_call": chunk, "us_per_step_median": round(res[2], 3), "us_per_step_min": round(res[0], 3),
                              "env_steps_per_s": n / (res[2] * 1e-6), "launches": envs[0].aviary.launch_count}), flush=True)
        for e in envs:
            e.close()


if __name__ == "__main__":
    main()
