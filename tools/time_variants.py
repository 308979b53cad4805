#!/usr/bin/env python
"""Times the QuadX-Hover step of one build of libpyflyt_b200.so (PYFLYT_B200_LIB selects it) and checks it against the
oracle; run once per variant by tools/run_variants.sh.  Prints one JSON line."""
import json, os, sys
import numpy as np
import torch

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
from pyflyt_b200.gym_envs.quadx_hover_env import QuadXHoverVecEnv

dev = torch.device("cuda:0")
n = 65536
buf = torch.empty(256 * 1024 * 1024 // 4, dtype=torch.float32, device=dev)
g = torch.Generator(device=dev).manual_seed(1)
lo = torch.tensor([-3.14159265, -3.14159265, -3.14159265, 0.0], device=dev)
hi = torch.tensor([3.14159265, 3.14159265, 3.14159265, 0.8], device=dev)
actions = lo + (hi - lo) * torch.rand((16, n, 4), device=dev, generator=g)


def q(ms, f):
    ms = sorted(ms)
    return ms[int(f * (len(ms) - 1))] * 1e3


def run(inline, flush, K=200):
    env = QuadXHoverVecEnv(num_envs=n, seed=0, device=dev, inline_reset=inline)
    env.reset()
    av = env.aviary
    for k in range(40):
        av.env_step(actions=actions[k % 16])
    av.profile_begin(K)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    torch.cuda.synchronize()
    e0.record()
    for k in range(K):
        if flush:
            buf.fill_(float(k))
        av.env_step(actions=actions[k % 16])
    e1.record()
    torch.cuda.synchronize()
    ms = av.profile_read(K)
    av.profile_begin(0)
    env.close()
    return {"p10": q(ms, 0.1), "p50": q(ms, 0.5), "p90": q(ms, 0.9), "loop_us_per_step": e0.elapsed_time(e1) * 1e3 / K}


def parity():
    from engines import CudaEngine, OracleEngine, build_model, hover_config
    m = 2048
    rng = np.random.default_rng(0)
    f = lambda a: a.astype(np.float32).astype(np.float64)
    model = build_model("quadx", "cf2x")
    env = hover_config(0, "quaternion", False, 3.0)
    start, orn = np.tile([[0.0, 0.0, 1.0]], (m, 1)), np.zeros((m, 3))
    orc, cud = OracleEngine(model, env, m, start, orn), CudaEngine(model, env, m, start, orn)
    nz = f(rng.normal(4.0, 1.0, (20, m)))
    o0, o1 = orc.env_reset(nz), cud.env_reset(nz)
    worst = float(np.abs(o0 - o1).max())
    for _ in range(20):
        act = f(rng.uniform([-1, -1, -1, 0.0], [1, 1, 1, 0.8], (m, 4)))
        nz = f(rng.normal(4.0, 1.0, (6, m)))
        a, b = orc.env_step(act, nz), cud.env_step(act, nz)
        worst = max(worst, float(np.abs(a[0] - b[0]).max()), float(np.abs(a[1] - b[1]).max()))
        if not (np.array_equal(a[2], b[2]) and np.array_equal(a[3], b[3])):
            return {"worst": worst, "flags": "MISMATCH"}
    return {"worst": worst, "flags": "ok"}


out = {"lib": os.environ.get("PYFLYT_B200_LIB", "default"), "cold_side": run(0, True), "cold_same": run(2, True), "warm_same": run(2, False),
       "warm_side": run(0, False), "parity": parity()}
print(json.dumps(out), flush=True)
