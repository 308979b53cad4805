"""Golden DISTRIBUTIONS of the reference's env-level random draws, sampled from the UNMODIFIED reference code (imported from
/root/reference on the fake Bullet, oracle/ref_in_loop.py) and stored as quantile tables:

  * WaypointHandler.reset                   gym_envs/utils/waypoint_handler.py:53-90   (Fixedwing-Waypoints and QuadX-Waypoints settings)
  * MAFixedwingDogfightEnv._get_start_pos_orn   pz_envs/fixedwing_envs/ma_fixedwing_dogfight_env.py:177-217

tests/test_draw_distributions.py compares the device streams (replayed on the host by tests/philox_replay.py, which is pinned bit for
bit to the kernels) with these tables: the kernels draw from Philox instead of numpy's PCG64 / MT19937, so the draws cannot be
equal — their distributions must be.   python tools/gen_draw_fixtures.py  ->  tests/golden/draw_quantiles.npz
"""
import os
import sys
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle import ref_in_loop  # noqa: E402

Q = np.linspace(0.0, 1.0, 2001)


def main():
    ref_in_loop.install()
    from PyFlyt.gym_envs.utils.waypoint_handler import WaypointHandler
    from PyFlyt.pz_envs.fixedwing_envs.ma_fixedwing_dogfight_env import MAFixedwingDogfightEnv

    out = {"q": Q}
    rng = np.random.default_rng(12345)
    for tag, dome, min_height, yaw in (("fw", 100.0, 0.5, False), ("qx", 5.0, 0.1, True)):
        wh = WaypointHandler(enable_render=False, num_targets=4, use_yaw_targets=yaw, goal_reach_distance=1.0, goal_reach_angle=0.1,
                             flight_dome_size=dome, min_height=min_height, np_random=rng)
        tg, yw = [], []
        for _ in range(60000):
            wh.reset(p=None, np_random=rng)
            tg.append(wh.targets.copy())
            if yaw:
                yw.append(wh.yaw_targets.copy())
        tg = np.concatenate(tg)
        for k, name in enumerate("xyz"):
            out[f"wp_{tag}_{name}"] = np.quantile(tg[:, k], Q)
        out[f"wp_{tag}_r"] = np.quantile(np.linalg.norm(tg, axis=1), Q)
        if yaw:
            out[f"wp_{tag}_yaw"] = np.quantile(np.concatenate(yw), Q)
    # dogfight spawns: the method only reads these attributes of self
    me = types.SimpleNamespace(team_size=1, spawn_min_radius=10.0, spawn_max_radius=50.0, spawn_min_height=20.0, spawn_max_height=50.0,
                               num_possible_agents=2)
    pos, orn = [], []
    for seed in range(120000):
        p_, o_ = MAFixedwingDogfightEnv._get_start_pos_orn(me, seed)
        pos.append(p_)
        orn.append(o_)
    pos, orn = np.stack(pos), np.stack(orn)  # [M, 2, 3]
    out["df_radius"] = np.quantile(np.hypot(pos[..., 0], pos[..., 1]).reshape(-1), Q)
    out["df_height"] = np.quantile(pos[..., 2].reshape(-1), Q)
    ang = np.arctan2(pos[..., 1], pos[..., 0])
    out["df_angle0"] = np.quantile(ang[:, 0], Q)                                             # agent 0's bearing: uniform on the circle
    out["df_opposite"] = np.quantile(np.abs(np.angle(np.exp(1j * (ang[:, 1] - ang[:, 0])))), Q)  # agents of an arena: pi apart
    out["df_heading_jitter"] = np.quantile(np.angle(np.exp(1j * (orn[..., 2] - ang))).reshape(-1), Q)  # yaw - bearing in [0, pi/8)
    path = os.path.join(ROOT, "tests", "golden", "draw_quantiles.npz")
    np.savez_compressed(path, **{k: v.astype(np.float32) if k != "q" else v for k, v in out.items()})
    print("wrote", path, os.path.getsize(path), "bytes;", sorted(out))


if __name__ == "__main__":
    main()
