"""The env-level random draws of the kernels (waypoints, dogfight spawns, randomised rocket drops) come from Philox instead of
numpy's generators, so they cannot equal the reference's draws number for number; their DISTRIBUTIONS must.  The host replay of
the device streams (tests/philox_replay.py, pinned bit for bit to the kernels: tests/test_philox_replay.py and the timed-path GPU
tests) is compared with quantile tables sampled from the UNMODIFIED reference code (tools/gen_draw_fixtures.py ->
tests/golden/draw_quantiles.npz: WaypointHandler.reset waypoint_handler.py:53-90, _get_start_pos_orn
ma_fixedwing_dogfight_env.py:177-217) by a two-sample Kolmogorov-Smirnov distance, and with the closed-form uniform ranges of
rocket_base_env.py:192-199."""
import os

import numpy as np

from philox_replay import Streams, philox4x32_10, unit_open

GOLD = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", "draw_quantiles.npz"))
Q = GOLD["q"]


def ks_to_table(sample, name, atom=None):
    """sup |F_sample - F_reference| with the reference CDF given as a quantile table.  `atom`: a value carrying probability mass
    (the z floor of the waypoints): there the CDFs are compared by their right limits only"""
    table = GOLD[name].astype(np.float64)
    s = np.sort(np.asarray(sample, dtype=np.float64))
    f_ref = np.interp(s, table, Q)
    hi = np.arange(1, len(s) + 1) / len(s)
    lo = np.arange(0, len(s)) / len(s)
    if atom is not None:
        at = s <= atom + 1e-6
        f_atom = Q[np.searchsorted(table, atom + 1e-6, side="right") - 1]  # reference mass at and below the atom
        f_ref = np.where(at, f_atom, f_ref)
        m = int(at.sum())
        hi = np.where(at, m / len(s), hi)
        lo = np.where(at, m / len(s), lo)
    return float(max(np.abs(hi - f_ref).max(), np.abs(lo - f_ref).max()))


# two samples of ~2e5: the 0.1 % critical value of the KS statistic is 1.95 * sqrt(2 / 2e5) = 0.006; the quantile table's
# own resolution (2001 points, float32) adds ~1e-3
KS_MAX = 0.01


def test_waypoint_draws_match_the_reference_distribution():
    n, T = 60000, 4
    for tag, dome, mh, yaw in (("fw", 100.0, 0.5, False), ("qx", 5.0, 0.1, True)):
        tg = Streams(31337, n).waypoint_targets(3, T, dome, min_height=mh, yaw=yaw).astype(np.float64).reshape(n * T, -1)
        assert tg[:, 2].min() >= mh - 1e-6  # the z floor (waypoint_handler.py:81-83)
        for k, c in enumerate("xyz"):
            assert ks_to_table(tg[:, k], f"wp_{tag}_{c}", atom=mh if c == "z" else None) < KS_MAX, (tag, c)
        assert ks_to_table(np.linalg.norm(tg[:, :3], axis=1), f"wp_{tag}_r") < KS_MAX, tag
        if yaw:
            assert ks_to_table(tg[:, 3], f"wp_{tag}_yaw") < KS_MAX
    # independence of the targets of one env: neighbours are uncorrelated
    tg = Streams(31337, n).waypoint_targets(3, T, 100.0, min_height=0.5).astype(np.float64)
    for a in range(3):
        assert abs(np.corrcoef(tg[:, 0, a], tg[:, 1, a])[0, 1]) < 0.02
    # different seq (episode number) or env id -> a different draw
    t2 = Streams(31337, n).waypoint_targets(4, T, 100.0, min_height=0.5)
    assert np.mean(np.all(np.isclose(tg, t2), axis=(1, 2))) == 0.0


def dogfight_spawns(seed, num_arenas, seq, rmin=10.0, rmax=50.0, A=2):
    """pfb_dogfight.cu df_reset_agent (random_spawn): one base angle per arena (stream of the arena's first agent, tag 6, word 0),
    radius / height / heading jitter per agent (tag 6 | 1)"""
    n = num_arenas * A
    st = Streams(seed, n)
    li = np.arange(n) % A
    first = np.arange(n) - li
    seq = np.broadcast_to(np.asarray(seq, dtype=np.uint32), (n,))  # per agent (arena-uniform): the episode number of an autoreset
    a = philox4x32_10(st.env_lo[first], st.env_hi[first], seq, np.uint32(6 << 24), st.k0, st.k1)
    b = philox4x32_10(st.env_lo, st.env_hi, seq, np.uint32((6 << 24) | 1), st.k0, st.k1)
    two_pi = np.float32(6.28318530717958647692)
    rad = (two_pi / np.float32(A)) * li.astype(np.float32) + two_pi * unit_open(a[0])
    radius = np.float32(rmin) + np.float32(rmax - rmin) * unit_open(b[0])
    height = np.float32(rmin) + np.float32(rmax - rmin) * unit_open(b[1])  # (sic) the radius range: ma_fixedwing_dogfight_env.py:199-203
    yaw = rad + unit_open(b[2]) * np.float32(0.39269908169872414)
    pos = np.stack([radius * np.cos(rad), radius * np.sin(rad), height], axis=1).astype(np.float64)
    return pos.reshape(num_arenas, A, 3), yaw.astype(np.float64).reshape(num_arenas, A)


def test_dogfight_spawns_match_the_reference_distribution():
    pos, yaw = dogfight_spawns(777, 120000, 1)
    ang = np.arctan2(pos[..., 1], pos[..., 0])
    assert ks_to_table(np.hypot(pos[..., 0], pos[..., 1]).reshape(-1), "df_radius") < KS_MAX
    assert ks_to_table(pos[..., 2].reshape(-1), "df_height") < KS_MAX
    assert ks_to_table(ang[:, 0], "df_angle0") < KS_MAX
    opposite = np.abs(np.angle(np.exp(1j * (ang[:, 1] - ang[:, 0]))))
    assert np.abs(opposite - np.pi).max() < 1e-5 and abs(float(GOLD["df_opposite"][1000]) - np.pi) < 1e-5  # pi apart, in both
    assert ks_to_table(np.angle(np.exp(1j * (yaw - ang))).reshape(-1), "df_heading_jitter") < KS_MAX


def test_rocket_drops_match_the_reference_ranges():
    """rocket_base_env.py:192-199: xy ~ U(-0.1 max_displacement, +), z ~ U(0.8, 0.9) ceiling, euler ~ U(-0.3, 0.3)^3"""
    n = 200000
    pos, orn = Streams(99, n).drop_poses(2, 500.0, 200.0)
    pos, orn = pos.astype(np.float64), orn.astype(np.float64)

    def ks_uniform(x, lo, hi):
        s = np.sort(x)
        f = (s - lo) / (hi - lo)
        return float(max(np.abs(np.arange(1, n + 1) / n - f).max(), np.abs(np.arange(0, n) / n - f).max()))

    for k in range(2):
        assert ks_uniform(pos[:, k], -20.0, 20.0) < 0.005
    assert ks_uniform(pos[:, 2], 400.0, 450.0) < 0.005
    for k in range(3):
        assert ks_uniform(orn[:, k], -0.3, 0.3) < 0.005
    c = np.corrcoef(np.column_stack([pos, orn]).T)
    assert np.abs(c - np.eye(6)).max() < 0.01  # the six coordinates are independent
