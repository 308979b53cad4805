"""Independent pins for the restated Bullet free-body step (SURVEY.md §7 "Hard parts", §A.3): both the Python engine the
UNMODIFIED reference flies on (oracle/fakebullet) and the C oracle (oracle/pfb_oracle.c) are checked against physics that
does not come from this repository:

* a torque-free composite body with an off-origin centre of mass and products of inertia conserves the world-frame angular
  momentum about its COM, its kinetic energy and its linear momentum.  Bullet integrates with explicit / semi-implicit
  Euler, so the discrete map conserves them only to first order in dt: the drift must be small AND halve when dt halves
  (a sign error in the gyroscopic term or in the COM coupling of the 6x6 Newton-Euler system gives an O(1), dt-independent
  drift);
* a constant torque about a principal axis of the central inertia tensor spins the body up as w(t) = tau t / lambda along
  that axis — exactly, for Euler steps, because w x I w vanishes — while the COM stays at rest;
* the two implementations of the same discrete map agree to round-off.
"""
import ctypes as C
import math
import os
import sys

import numpy as np
import pytest

ROOT = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..")
sys.path.insert(0, os.path.join(ROOT, "oracle", "fakebullet"))

from engines import build_model  # noqa: E402

URDF = """<?xml version="1.0"?>
<robot name="lopsided">
  <link name="base">
    <inertial><origin xyz="0.05 -0.02 0.01" rpy="0.1 -0.2 0.3"/><mass value="1.7"/>
      <inertia ixx="0.11" iyy="0.23" izz="0.31" ixy="0.01" ixz="-0.02" iyz="0.015"/></inertial>
  </link>
  <link name="arm"><inertial><origin xyz="0.1 0 0" rpy="0 0.4 0"/><mass value="0.6"/>
      <inertia ixx="0.02" iyy="0.05" izz="0.04" ixy="0" ixz="0.004" iyz="0"/></inertial></link>
  <link name="tail"><inertial><origin xyz="0 0 0"/><mass value="0.35"/>
      <inertia ixx="0" iyy="0" izz="0" ixy="0" ixz="0" iyz="0"/></inertial></link>
  <joint name="j0" type="fixed"><parent link="base"/><child link="arm"/><origin xyz="-0.6 0.25 0.1" rpy="0.2 0 -0.5"/></joint>
  <joint name="j1" type="fixed"><parent link="base"/><child link="tail"/><origin xyz="-1.1 0 0.3"/></joint>
</robot>
"""


def _central(M, c, I_O):
    return I_O - M * (np.dot(c, c) * np.eye(3) - np.outer(c, c))


def _invariants(M, c, I_O, pos, R, v, w):
    """(L about the COM, kinetic energy, linear momentum, COM position), world frame; v, w = base origin velocities."""
    Ic = R @ _central(M, c, I_O) @ R.T
    vc = v + np.cross(w, R @ c)
    return Ic @ w, 0.5 * M * vc @ vc + 0.5 * w @ Ic @ w, M * vc, pos + R @ c


# ---------------------------------------------------------------------------------------------------------------------
# the Python engine under the reference
# ---------------------------------------------------------------------------------------------------------------------
def _fake_world(tmp_path, hz):
    import pybullet as fb  # oracle/fakebullet/pybullet.py

    path = tmp_path / "lopsided.urdf"
    path.write_text(URDF)
    w = fb.World()
    w.setTimeStep(1.0 / hz)
    uid = w.loadURDF(str(path), basePosition=(0.3, -0.2, 5.0), baseOrientation=fb.getQuaternionFromEuler((0.3, -0.5, 1.1)))
    w.changeDynamics(uid, -1, linearDamping=0.0, angularDamping=0.0)
    return fb, w, uid


def _fake_run(tmp_path, hz, seconds, w0, v0, torque_base=None):
    fb, w, uid = _fake_world(tmp_path, hz)
    b = w.bodies[uid]
    w.resetBaseVelocity(uid, v0, w0)
    M, c, I_O = b.composite()
    out = []
    for _ in range(int(round(seconds * hz))):
        if torque_base is not None:
            w.applyExternalTorque(uid, -1, tuple(torque_base), fb.LINK_FRAME)
        w.stepSimulation()
        out.append(_invariants(M, c, I_O, b.pos.copy(), b.R(), b.v.copy(), b.w.copy()))
    return (M, c, I_O), b, out


def test_fakebullet_torque_free_tumble_conserves_to_first_order(tmp_path):
    w0, v0 = (0.7, -0.4, 0.9), (1.0, 2.0, -0.5)
    drift = {}
    for hz in (240, 480, 960):
        (M, c, I_O), b, tr = _fake_run(tmp_path, hz, 4.0, w0, v0)
        L0, E0, P0, _ = tr[0]
        L1, E1, P1, _ = tr[-1]
        drift[hz] = (np.linalg.norm(L1 - L0) / np.linalg.norm(L0), abs(E1 - E0) / E0, np.linalg.norm(P1 - P0) / np.linalg.norm(P0))
    assert np.linalg.norm(c) > 0.1  # the COM really is off the base origin
    for k in range(3):
        assert drift[240][k] < 2e-2, drift
        for a, b_ in ((240, 480), (480, 960)):  # first-order convergence: halving dt halves the drift
            assert 1.6 < drift[a][k] / drift[b_][k] < 2.4, (k, drift)


def test_fakebullet_principal_axis_spin_up_closed_form(tmp_path):
    fb, w, uid = _fake_world(tmp_path, 240)
    M, c, I_O = w.bodies[uid].composite()
    lam, vec = np.linalg.eigh(_central(M, c, I_O))
    for ax in range(3):
        u, tau = vec[:, ax], 0.37
        (M, c, I_O), b, tr = _fake_run(tmp_path, 240, 2.0, (0, 0, 0), (0, 0, 0), torque_base=tau * u)
        n = len(tr)
        w_body = b.R().T @ b.w
        assert np.abs(w_body - u * tau * n / 240.0 / lam[ax]).max() < 1e-11
        com = np.array([t[3] for t in tr])
        assert np.abs(com - com[0]).max() < 2e-3  # first order in dt (the base origin swings around the fixed COM)
        # no force -> the COM stays at rest: its velocity is the O(dt) residue of Euler steps on a base origin that swings
        # around it at |w||c| (checked relative to that speed)
        swing = np.linalg.norm(b.w) * np.linalg.norm(c)
        assert np.abs(np.array([t[2] for t in tr]) / M).max() < 2e-2 * swing


# ---------------------------------------------------------------------------------------------------------------------
# the C oracle, on the vehicles' own composite bodies (fixed-wing: COM 0.45 m behind the base, Ixz != 0; rocket: COM offset)
# ---------------------------------------------------------------------------------------------------------------------
def _bare(kind, name, hz):
    """The vehicle's rigid body with every force source switched off."""
    kw = dict(starting_fuel_ratio=1.0) if kind == "rocket" else {}
    m = build_model(kind, name, None, hz, hz // 2, **kw)
    m.gravity = 0.0
    m.n_surfaces = 0
    m.n_bodies = 0
    m.drag_const[:] = [0.0, 0.0, 0.0]
    for k in range(4):
        m.thrust_coef[k] = 0.0
        m.torque_coef[k] = 0.0
    m.fuel_max_rate = 0.0  # the booster never ignites with a zero setpoint; no fuel burn either way
    m.n_shapes = 0
    return m


def _oracle_run(m, seconds, w0, v0, torque_base=None, euler0=(0.3, -0.5, 1.1)):
    from oracle.oracle import Oracle, lib

    o = Oracle(m, None, n=1, start_pos=np.array([[0.3, -0.2, 500.0]]), start_orn=np.array([euler0]))
    o.reset()
    if int(m.kind) != 2:
        o.set_mode(-1 if int(m.kind) == 0 else 0)
    o.set_base_velocity(np.array([v0], dtype=np.float64), np.array([w0], dtype=np.float64))
    if torque_base is not None:
        t = np.ascontiguousarray(torque_base, dtype=np.float64)
        L = lib()
        L.orc_debug_set_wrench.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p]
        L.orc_debug_set_wrench(o._h, None, t.ctypes.data_as(C.c_void_p))
    n_aviary = int(round(seconds * m.physics_hz)) // o.updates_per_step
    out = []
    for _ in range(n_aviary):
        o.aviary_step(1, np.ones((o.updates_per_step, 1)))
        pos, quat, v, w = o.raw()
        x, y, z, q = quat[0]
        R = np.array([[1 - 2 * (y * y + z * z), 2 * (x * y - q * z), 2 * (x * z + q * y)],
                      [2 * (x * y + q * z), 1 - 2 * (x * x + z * z), 2 * (y * z - q * x)],
                      [2 * (x * z - q * y), 2 * (y * z + q * x), 1 - 2 * (x * x + y * y)]])
        out.append((pos[0].copy(), R, v[0].copy(), w[0].copy()))
    return o, out


def _body_of(m):
    if int(m.kind) == 2:  # rocket: dry composite + full fuel tank (boosters.py:214-231, fuel ratio 1)
        Mf = m.fuel_total_mass
        M = m.dry_mass + Mf
        r = np.array(m.fuel_pos[:])
        mc = np.array(m.dry_first_moment[:]) + Mf * r
        I = np.array(m.dry_inertia[:]).reshape(3, 3) + np.diag(m.fuel_max_inertia[:]) + Mf * (r @ r * np.eye(3) - np.outer(r, r))
        return M, mc / M, I
    return m.mass, np.array(m.com[:]), np.array(m.inertia[:]).reshape(3, 3)


@pytest.mark.parametrize("kind,name", [("fixedwing", "fixedwing"), ("fixedwing", "acrowing"), ("rocket", "rocket")])
def test_oracle_torque_free_tumble_conserves_to_first_order(kind, name):
    w0, v0 = (0.5, -0.3, 0.7), (3.0, -1.0, 0.5)
    drift = {}
    for hz in (240, 480, 960):
        m = _bare(kind, name, hz)
        M, c, I_O = _body_of(m)
        _, tr = _oracle_run(m, 4.0, w0, v0)
        inv = [_invariants(M, c, I_O, *t) for t in (tr[0], tr[-1])]
        drift[hz] = (np.linalg.norm(inv[1][0] - inv[0][0]) / np.linalg.norm(inv[0][0]), abs(inv[1][1] - inv[0][1]) / inv[0][1],
                     np.linalg.norm(inv[1][2] - inv[0][2]) / np.linalg.norm(inv[0][2]))
    assert np.linalg.norm(c) > 0.1 or kind == "rocket"
    for k in range(3):
        assert drift[240][k] < 2e-2, (name, drift)
        for a, b_ in ((240, 480), (480, 960)):
            assert 1.6 < drift[a][k] / drift[b_][k] < 2.4, (name, k, drift)


@pytest.mark.parametrize("kind,name", [("fixedwing", "fixedwing"), ("fixedwing", "acrowing"), ("rocket", "rocket")])
def test_oracle_principal_axis_spin_up_closed_form(kind, name):
    m = _bare(kind, name, 240)
    M, c, I_O = _body_of(m)
    lam, vec = np.linalg.eigh(_central(M, c, I_O))
    for ax in range(3):
        u, tau = vec[:, ax], 0.05 * lam[ax]
        _, tr = _oracle_run(m, 2.0, (0, 0, 0), (0, 0, 0), torque_base=tau * u)
        pos, R, v, w = tr[-1]
        steps = len(tr) * 2
        assert np.abs(R.T @ w - u * tau * steps / 240.0 / lam[ax]).max() < 1e-10 * max(1.0, tau / lam[ax])
        com = np.array([t[0] + t[1] @ c for t in tr])
        assert np.abs(com - com[0]).max() < 5e-3
        vc = np.array([t[2] + np.cross(t[3], t[1] @ c) for t in tr])
        assert np.abs(vc).max() < 2e-2 * max(np.linalg.norm(w) * np.linalg.norm(c), 1e-9) + 1e-12


def test_oracle_and_fakebullet_are_the_same_map(tmp_path):
    """The lopsided synthetic body through both engines: identical trajectories to round-off (tumble with gravity on)."""
    fb, w, uid = _fake_world(tmp_path, 240)
    b = w.bodies[uid]
    w.setGravity(0.0, 0.0, -9.81)
    M, c, I_O = b.composite()
    m = _bare("fixedwing", "fixedwing", 240)
    m.gravity = -9.81
    m.mass = M
    for k in range(3):
        m.com[k] = c[k]
    for k in range(9):
        m.inertia[k] = I_O.reshape(-1)[k]
    w0, v0 = (0.7, -0.4, 0.9), (1.0, 2.0, -0.5)
    w.resetBaseVelocity(uid, v0, w0)
    # same initial pose: the oracle's start pose is the base inertial frame
    from oracle.oracle import Oracle

    e0 = fb.getEulerFromQuaternion(tuple(b.quat))
    o = Oracle(m, None, n=1, start_pos=b.pos[None].copy(), start_orn=np.array([e0]))
    o.reset()
    o.set_mode(0)
    o.set_base_velocity(np.array([v0], dtype=np.float64), np.array([w0], dtype=np.float64))
    worst = 0.0
    for _ in range(480):
        o.aviary_step(1, np.ones((2, 1)))
        w.stepSimulation()
        w.stepSimulation()
        pos, quat, v, ww = o.raw()
        worst = max(worst, np.abs(pos[0] - b.pos).max(), np.abs(v[0] - b.v).max(), np.abs(ww[0] - b.w).max(),
                    min(np.abs(quat[0] - b.quat).max(), np.abs(quat[0] + b.quat).max()))
    assert worst < 1e-9, worst
