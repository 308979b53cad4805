"""Experiment: end-to-end QuadX-Hover step through pfb_env_step_mapped as a function of the dynamic shared memory the step launch
requests (PFB_MAPPED_DYN_SMEM: caps the CTAs resident per SM, i.e. the number of waves).  One JSON line per setting."""
import json
import os
import sys
import time

import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))


def main():
    from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

    n, K = 65536, 200
    dev = torch.device("cuda", 0)
    for smem in [0, 8 * 1024, 12 * 1024, 19 * 1024, 27 * 1024, 40 * 1024]:
        os.environ["PFB_MAPPED_DYN_SMEM"] = str(smem)
        env = QuadXHoverVecEnv(num_envs=n, seed=1)
        av = env.aviary
        env.reset()
        g = torch.Generator(device="cpu").manual_seed(0)
        lo, hi = torch.tensor([-3.14159, -3.14159, -3.14159, 0.0]), torch.tensor([3.14159, 3.14159, 3.14159, 0.8])
        act_h = [(lo + (hi - lo) * torch.rand((n, 4), generator=g)).pin_memory() for _ in range(4)]
        slab_h = torch.empty(av.out_slab_bytes(n, env.obs_dim), dtype=torch.uint8).pin_memory()
        obs_h, rew_h, te_h, tr_h = av.slab_views(slab_h, n, env.obs_dim)
        best = 1e9
        for rep in range(3):
            for k in range(5):
                av.env_step_mapped(act_h[k % 4], obs_h, rew_h, te_h, tr_h)
            torch.cuda.synchronize(dev)
            t0 = time.perf_counter()
            for k in range(K):
                av.env_step_mapped(act_h[k % 4], obs_h, rew_h, te_h, tr_h)
                torch.cuda.synchronize(dev)
            best = min(best, (time.perf_counter() - t0) / K)
        chk = float(obs_h.double().sum())
        print(json.dumps({"dyn_smem": smem, "us_per_step": best * 1e6, "env_steps_per_s": n / best, "obs_checksum": chk}), flush=True)
        env.close()


if __name__ == "__main__":
    main()
