#!/usr/bin/env python
"""Precision-policy study (CPU only): the kernel body of pyflyt_b200/csrc compiled for the host with
different fp32/fp64 choices, flown next to the fp64 oracle on the SURVEY §8(d) config-1 scenario
(mode 0, start z = 50 m, 1000 env-steps = 3000 Aviary steps = 6000 substeps, rate commands x0.3 held
for 10 env-steps, injected motor noise).  Prints max / median |dpos| per variant; DESIGN.md quotes it.
"""
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "tests"))
from engines import HostSimEngine, OracleEngine, build_model  # noqa: E402

VARIANTS = {
    "q64 x64 v64 R64": "-DPFB_R_DOUBLE=1",
    "q64 x64 v64 R32 (shipped)": "",
    "q64 x64 v32 R32": "-DPFB_R_DOUBLE=0 -DPFB_V_DOUBLE=0",
    "q64 x32 v32 R32": "-DPFB_R_DOUBLE=0 -DPFB_V_DOUBLE=0 -DPFB_X_DOUBLE=0",
    "q32 x64 v64 R32": "-DPFB_R_DOUBLE=0 -DPFB_Q_DOUBLE=0",
    "all fp32": "-DPFB_R_DOUBLE=0 -DPFB_V_DOUBLE=0 -DPFB_X_DOUBLE=0 -DPFB_Q_DOUBLE=0",
}


def main(n=64, steps=3000, mode=0, seed=0):
    model = build_model("quadx", "cf2x")
    rng = np.random.default_rng(seed)
    start = np.tile(np.array([[0.0, 0.0, 50.0]]), (n, 1))
    noise = rng.normal(4.0, 1.0, size=(steps * 2, n))
    acts = np.zeros((steps, n, 4))
    for k in range(0, steps, 30):
        a = rng.uniform([-np.pi, -np.pi, -np.pi, 0.0], [np.pi, np.pi, np.pi, 0.8], size=(n, 4))
        a[:, :3] *= 0.3
        acts[k : k + 30] = a
    # identical inputs for both: actions and noise draws are fp32-representable numbers
    noise = noise.astype(np.float32).astype(np.float64)
    acts = acts.astype(np.float32).astype(np.float64)
    orc = OracleEngine(model, None, n, start, np.zeros((n, 3)))
    orc.reset()
    orc.set_mode(mode)
    ref = np.zeros((steps // 30, n, 3))
    for i in range(steps):
        orc.set_setpoints(acts[i])
        orc.aviary_step(noise[2 * i : 2 * i + 2])
        if i % 30 == 29:
            ref[i // 30] = orc.state()[:, 3]
    travelled = np.linalg.norm(orc.state()[:, 3] - start, axis=1)
    print(f"oracle: n={n} steps={steps} final |x - x0| median {np.median(travelled):.1f} m, max speed ~ {np.abs(orc.state()[:,2]).max():.1f} m/s")
    for name, flags in VARIANTS.items():
        hs = HostSimEngine(model, None, n, start, np.zeros((n, 3)), flags=flags)
        hs.reset()
        hs.set_mode(mode)
        err = np.zeros((steps // 30, n))
        for i in range(steps):
            hs.set_setpoints(acts[i])
            hs.aviary_step(noise[2 * i : 2 * i + 2])
            if i % 30 == 29:
                err[i // 30] = np.abs(hs.state()[:, 3] - ref[i // 30]).max(axis=1)
        worst = err.max(axis=0)
        print(f"{name:28s} max|dpos| {worst.max():.2e}  median {np.median(worst):.2e}  p90 {np.quantile(worst, 0.9):.2e}   (at 333 env-steps: {err[:33].max():.2e})")


if __name__ == "__main__":
    main()
