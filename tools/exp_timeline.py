"""Experiment (needs the -DPFB_TIMELINE build: PYFLYT_B200_LIB=pyflyt_b200/lib/variants/tl/libpyflyt_b200.so): per-warp %globaltimer stamps
of cold QuadX-Hover step launches -> where the launch time goes (entry ramp, input wait, integration, epilogue; builder CTAs)."""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv  # noqa: E402

dev = torch.device("cuda:0")
n = 65536
env = QuadXHoverVecEnv(num_envs=n, seed=0, device=dev)
env.reset()
av = env.aviary
buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
tl = torch.zeros((6, n), dtype=torch.float32, device=dev)
av.set_noise_dump(tl)
g = torch.Generator(device=dev).manual_seed(1)
lo = torch.tensor([-3.14159265, -3.14159265, -3.14159265, 0.0], device=dev)
hi = torch.tensor([3.14159265, 3.14159265, 3.14159265, 0.8], device=dev)
actions = lo + (hi - lo) * torch.rand((16, n, 4), device=dev, generator=g)
for k in range(60):
    av.env_step(actions=actions[k % 16])
rows = []
for rep in range(12):
    buf.fill_(float(rep))
    tl.zero_()
    torch.cuda.synchronize()
    av.env_step(actions=actions[rep % 16])
    torch.cuda.synchronize()
    grid = 2048 + 2 * 148
    t = tl.view(torch.int64).reshape(-1)[: grid * 4].reshape(grid, 4).cpu().numpy().astype(np.float64)
    t0 = t[:, 0].min()
    t = (t - t0) / 1e3  # us since the first warp entered
    b, s = t[:296], t[296:]
    b0, b1 = b[:148], b[148:]  # phase 0 (start the next spare) / phase 1 (finish the one started by the previous launch)
    b0, b1 = b0[b0[:, 1] > 0], b1[b1[:, 1] > 0]
    b = b[b[:, 1] > 0]  # builder CTAs that had work
    def q(x):
        return [round(float(np.percentile(x, p)), 2) for p in (0, 10, 50, 90, 100)]
    rows.append({"step_entry": q(s[:, 0]), "step_inputs": q(s[:, 1]), "step_loop_done": q(s[:, 2]), "step_exit": q(s[:, 3]),
                 "builders_with_work": int(len(b)), "b_entry": q(b[:, 0]) if len(b) else None, "b_loaded": q(b[:, 1]) if len(b) else None,
                 "b_chain_done": q(b[:, 2]) if len(b) else None, "b_exit": q(b[:, 3]) if len(b) else None,
                 "phase0": {"n": int(len(b0)), "loaded": q(b0[:, 1]), "chain_done": q(b0[:, 2]), "exit": q(b0[:, 3])} if len(b0) else None,
                 "phase1": {"n": int(len(b1)), "loaded": q(b1[:, 1]), "chain_done": q(b1[:, 2]), "exit": q(b1[:, 3])} if len(b1) else None,
                 "last_exit": round(float(t[:, 3].max()), 2)})
for r in rows[2:]:
    print(json.dumps(r))
env.close()
