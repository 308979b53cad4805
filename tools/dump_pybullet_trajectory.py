#!/usr/bin/env python
"""Where REAL PyBullet + PyFlyt are importable (not in the build image), dump the same Aviary-level
fixtures as tools/gen_golden.py from the true engine, so the restated Bullet step (oracle/fakebullet,
SURVEY.md §A.3) can be pinned by anyone with `pip install pybullet PyFlyt`.

    python tools/dump_pybullet_trajectory.py out_dir

The output .npz files have the layout of tests/golden/quadx_*.npz; point tests/engines.GOLDEN at out_dir
(env PFB_GOLDEN_DIR) and run `pytest tests/test_oracle_golden.py` to compare the oracle with PyBullet.
"""
import os
import sys

import numpy as np


def main(out_dir):
    try:
        import pybullet  # noqa: F401

        if "fakebullet" in os.path.realpath(pybullet.__file__):
            raise ImportError("the fake engine is on sys.path")
        from PyFlyt.core import Aviary
    except ImportError as e:
        print(f"real pybullet / PyFlyt not importable here ({e}); nothing to do")
        return 0
    os.makedirs(out_dir, exist_ok=True)

    class Recorder:
        def __init__(self, seed):
            self._rng, self.log = np.random.default_rng(seed), []

        def normal(self, *a, **k):
            v = self._rng.normal(*a, **k)
            self.log.append(float(v))
            return v

        def __getattr__(self, n):
            return getattr(self._rng, n)

    def fly(name, mode, start_pos, start_orn, sched, n_steps, seed):
        rng = Recorder(seed)
        env = Aviary(start_pos=np.array([start_pos], float), start_orn=np.array([start_orn], float), drone_type="quadx", np_random=rng)
        env.set_mode(mode)
        sp0 = np.array(env.drones[0].setpoint, float)
        st, aux, pwm, con, sps = [], [], [], [], []
        for i in range(n_steps):
            if i in sched:
                env.set_setpoint(0, np.array(sched[i], float))
            sps.append(np.array(env.drones[0].setpoint, float))
            env.step()
            d = env.drones[0]
            st.append(np.array(d.state)); aux.append(np.array(d.aux_state)); pwm.append(np.array(d.pwm))
            con.append(bool(np.any(env.contact_array[env.planeId])))
        np.savez_compressed(os.path.join(out_dir, name + ".npz"), kind="quadx_aviary", mode=mode, drone_model="cf2x",
                            start_pos=np.array(start_pos, float), start_orn=np.array(start_orn, float), setpoint_after_set_mode=sp0,
                            setpoints=np.array(sps), noise=np.array(rng.log), state=np.array(st), aux=np.array(aux), pwm=np.array(pwm),
                            contact=np.array(con))
        env.disconnect()

    fly("quadx_mode7_hold", 7, [0, 0, 1], [0, 0, 0], {}, 1000, 1)
    fly("quadx_mode7_setpoints", 7, [0, 0, 1], [0, 0, 0], {0: [1, 0, 0, 1], 500: [0, 0, np.pi / 4, 2]}, 1000, 2)
    fly("quadx_cf2x_mode0", 0, [0.3, -0.2, 5.0], [0.1, -0.15, 0.7], {0: [0.3, -0.2, 0.1, 0.45], 150: [-0.5, 0.4, -0.3, 0.3]}, 300, 10)
    print("wrote fixtures to", out_dir)
    return 0


if __name__ == "__main__":
    sys.exit(main(sys.argv[1] if len(sys.argv) > 1 else "pybullet_golden"))
