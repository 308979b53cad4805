"""TEST INFRASTRUCTURE — a restatement of the part of PyBullet that PyFlyt's core calls.

This module is installed as ``pybullet`` in ``sys.modules`` (see ``oracle/ref_in_loop.py``)
so that the UNMODIFIED reference package under ``/root/reference/PyFlyt`` can be imported
and flown in a container that has no PyBullet wheel.  It is *our* fp64 restatement of
Bullet3's ``btMultiBody`` free-base step for bodies whose joints are all fixed
 — **parity unpinned**: no reference test pins numbers at this boundary and
the real engine (third-party, un-pinned dependency ``pybullet``, pyproject.toml:18 of the
reference) cannot be installed here.

Only ``tests/``, ``tools/`` generators, ``__graft_entry__.smoke()`` and ``bench.py``'s CPU-baseline
legs may import anything under ``oracle/``.  The product package never does.

Call sites in the reference that this file serves (file:line under /root/reference):
  PyFlyt/core/aviary.py:104,207,211,225-242,320,516,523
  PyFlyt/core/abstractions/base_drone.py:115-122,303-304
  PyFlyt/core/drones/quadx.py:228,509-510,517-526
  PyFlyt/core/abstractions/motors.py:152-155, boring_bodies.py:86-88,121-127,
  lifting_surfaces.py:81-83,315-324, boosters.py:195-212, drones/fixedwing.py:200-201,
  gym_envs/rocket_envs/rocket_base_env.py:228, rocket_landing_env.py:119
"""

from __future__ import annotations

import math
import os
import re
import xml.etree.ElementTree as ET

import numpy as np

# connection modes / flags (values follow pybullet's public constants)
GUI = 1
DIRECT = 2
WORLD_FRAME = 2
LINK_FRAME = 1
URDF_USE_INERTIA_FROM_FILE = 2

GRAVITY_DEFAULT = np.zeros(3)
MAX_COORDINATE_VELOCITY = 100.0  # btMultiBody::m_maxCoordinateVelocity
ANGULAR_MOTION_THRESHOLD = 0.5 * (math.pi / 2.0)  # btTransformUtil
#: relative contact breaking threshold factor (gContactBreakingThreshold); a manifold point
#: exists while distance < factor * angular-motion-disc of the smaller shape.
CONTACT_BREAKING_FACTOR = 0.02


# --------------------------------------------------------------------------------------
# pure-math helpers (module-level in pybullet as well)
# --------------------------------------------------------------------------------------
def getQuaternionFromEuler(eulerAngles, physicsClientId=0):
    """ZYX (yaw-pitch-roll) euler → (x, y, z, w); btQuaternion::setEulerZYX."""
    roll, pitch, yaw = (float(a) for a in eulerAngles)
    hy, hp, hr = yaw * 0.5, pitch * 0.5, roll * 0.5
    cy, sy = math.cos(hy), math.sin(hy)
    cp, sp = math.cos(hp), math.sin(hp)
    cr, sr = math.cos(hr), math.sin(hr)
    return (
        sr * cp * cy - cr * sp * sy,
        cr * sp * cy + sr * cp * sy,
        cr * cp * sy - sr * sp * cy,
        cr * cp * cy + sr * sp * sy,
    )


def getEulerFromQuaternion(quaternion, physicsClientId=0):
    """(x, y, z, w) → (roll, pitch, yaw); btQuaternion::getEulerZYX with the gimbal-lock branch."""
    x, y, z, w = (float(a) for a in quaternion)
    sqx, sqy, sqz, sqw = x * x, y * y, z * z, w * w
    sarg = -2.0 * (x * z - w * y)
    if sarg <= -0.99999:
        pitch = -0.5 * math.pi
        roll = 0.0
        yaw = 2.0 * math.atan2(x, -y)
    elif sarg >= 0.99999:
        pitch = 0.5 * math.pi
        roll = 0.0
        yaw = 2.0 * math.atan2(-x, y)
    else:
        pitch = math.asin(sarg)
        roll = math.atan2(2.0 * (y * z + w * x), sqw - sqx - sqy + sqz)
        yaw = math.atan2(2.0 * (x * y + w * z), sqw + sqx - sqy - sqz)
    return (roll, pitch, yaw)


def _mat_from_quat(q) -> np.ndarray:
    x, y, z, w = (float(a) for a in q)
    d = x * x + y * y + z * z + w * w
    s = 2.0 / d
    xs, ys, zs = x * s, y * s, z * s
    wx, wy, wz = w * xs, w * ys, w * zs
    xx, xy, xz = x * xs, x * ys, x * zs
    yy, yz, zz = y * ys, y * zs, z * zs
    return np.array(
        [
            [1.0 - (yy + zz), xy - wz, xz + wy],
            [xy + wz, 1.0 - (xx + zz), yz - wx],
            [xz - wy, yz + wx, 1.0 - (xx + yy)],
        ]
    )


def getMatrixFromQuaternion(quaternion, physicsClientId=0):
    """Row-major body→world rotation, 9-tuple (btMatrix3x3::setRotation)."""
    return tuple(_mat_from_quat(quaternion).reshape(-1))


def _quat_mul(a, b) -> np.ndarray:
    ax, ay, az, aw = a
    bx, by, bz, bw = b
    return np.array(
        [
            aw * bx + ax * bw + ay * bz - az * by,
            aw * by + ay * bw + az * bx - ax * bz,
            aw * bz + az * bw + ax * by - ay * bx,
            aw * bw - ax * bx - ay * by - az * bz,
        ]
    )


def _quat_from_mat(m: np.ndarray) -> np.ndarray:
    tr = m[0, 0] + m[1, 1] + m[2, 2]
    if tr > 0.0:
        s = math.sqrt(tr + 1.0)
        w = 0.5 * s
        s = 0.5 / s
        return np.array([(m[2, 1] - m[1, 2]) * s, (m[0, 2] - m[2, 0]) * s, (m[1, 0] - m[0, 1]) * s, w])
    i = int(np.argmax([m[0, 0], m[1, 1], m[2, 2]]))
    j, k = (i + 1) % 3, (i + 2) % 3
    s = math.sqrt(m[i, i] - m[j, j] - m[k, k] + 1.0)
    q = np.zeros(4)
    q[i] = 0.5 * s
    s = 0.5 / s
    q[3] = (m[k, j] - m[j, k]) * s
    q[j] = (m[j, i] + m[i, j]) * s
    q[k] = (m[k, i] + m[i, k]) * s
    return q


def _rpy_mat(rpy) -> np.ndarray:
    return _mat_from_quat(getQuaternionFromEuler(rpy))


# --------------------------------------------------------------------------------------
# URDF → composite body description
# --------------------------------------------------------------------------------------
def _floats(s: str | None, n: int, default=0.0):
    if s is None:
        return [default] * n
    return [float(v) for v in s.split()]


class _Link:
    __slots__ = ("name", "mass", "inertia_local", "r", "R", "frame_r", "frame_R", "shapes")

    def __init__(self):
        self.name = ""
        self.mass = 0.0
        self.inertia_local = np.zeros((3, 3))  # about the link COM, in the link inertial frame axes
        self.r = np.zeros(3)  # link COM (inertial frame origin) in the base inertial frame
        self.R = np.eye(3)  # link inertial frame axes in the base inertial frame
        self.frame_r = np.zeros(3)  # URDF link frame origin in the base inertial frame
        self.frame_R = np.eye(3)
        self.shapes = []  # collision primitives: (kind, dims, r_in_base, R_in_base)


def parse_urdf(path: str, global_scaling: float = 1.0) -> list[_Link]:
    """Parses a URDF whose joints are all ``fixed`` into a flat list [base, link0, link1, ...].

    Link index i (PyBullet) == i-th ``<joint>`` in file order; the child of that joint.
    The file is truncated at the first ``</robot>`` (rocket.urdf has a duplicated close tag).
    """
    with open(path, "r", encoding="utf-8") as f:
        text = f.read()
    end = text.find("</robot>")
    if end >= 0:
        text = text[: end + len("</robot>")]
    text = re.sub(r"<!--.*?-->", "", text, flags=re.S)
    root = ET.fromstring(text)

    raw_links = {}
    for le in root.findall("link"):
        d = dict(name=le.get("name"), mass=0.0, I=np.zeros((3, 3)), ixyz=np.zeros(3), irpy=np.zeros(3), cols=[])
        ine = le.find("inertial")
        if ine is not None:
            o = ine.find("origin")
            if o is not None:
                d["ixyz"] = np.array(_floats(o.get("xyz"), 3)) * global_scaling
                d["irpy"] = np.array(_floats(o.get("rpy"), 3))
            m = ine.find("mass")
            if m is not None:
                d["mass"] = float(m.get("value"))
            it = ine.find("inertia")
            if it is not None:
                g = lambda k: float(it.get(k, 0.0))  # noqa: E731
                d["I"] = np.array(
                    [
                        [g("ixx"), g("ixy"), g("ixz")],
                        [g("ixy"), g("iyy"), g("iyz")],
                        [g("ixz"), g("iyz"), g("izz")],
                    ]
                )
        for ce in le.findall("collision"):
            o = ce.find("origin")
            cxyz = np.array(_floats(o.get("xyz") if o is not None else None, 3)) * global_scaling
            crpy = np.array(_floats(o.get("rpy") if o is not None else None, 3))
            geo = ce.find("geometry")
            if geo is None:
                continue
            if geo.find("box") is not None:
                dims = np.array(_floats(geo.find("box").get("size"), 3)) * global_scaling
                d["cols"].append(("box", dims, cxyz, crpy))
            elif geo.find("cylinder") is not None:
                c = geo.find("cylinder")
                dims = np.array([float(c.get("radius")), float(c.get("length"))]) * global_scaling
                d["cols"].append(("cylinder", dims, cxyz, crpy))
            elif geo.find("sphere") is not None:
                dims = np.array([float(geo.find("sphere").get("radius"))]) * global_scaling
                d["cols"].append(("sphere", dims, cxyz, crpy))
            elif geo.find("plane") is not None:
                d["cols"].append(("plane", np.zeros(1), cxyz, crpy))
        raw_links[d["name"]] = d

    joints = []
    children = set()
    for je in root.findall("joint"):
        jt = je.get("type")
        if jt != "fixed":
            raise NotImplementedError(f"fake bullet only handles fixed joints, got {jt!r} in {path}")
        o = je.find("origin")
        joints.append(
            dict(
                parent=je.find("parent").get("link"),
                child=je.find("child").get("link"),
                xyz=np.array(_floats(o.get("xyz") if o is not None else None, 3)) * global_scaling,
                rpy=np.array(_floats(o.get("rpy") if o is not None else None, 3)),
            )
        )
        children.add(joints[-1]["child"])
    roots = [n for n in raw_links if n not in children]
    assert len(roots) == 1, f"expected exactly one root link in {path}, got {roots}"
    base_name = roots[0]

    # link-frame poses relative to the BASE LINK FRAME
    frame_pose = {base_name: (np.zeros(3), np.eye(3))}
    pending = list(joints)
    while pending:
        progressed = False
        for j in list(pending):
            if j["parent"] in frame_pose:
                pr, pR = frame_pose[j["parent"]]
                frame_pose[j["child"]] = (pr + pR @ j["xyz"], pR @ _rpy_mat(j["rpy"]))
                pending.remove(j)
                progressed = True
        assert progressed, "URDF joint tree is not connected"

    base_raw = raw_links[base_name]
    bR = _rpy_mat(base_raw["irpy"])
    br = base_raw["ixyz"]

    def to_base_inertial(r, R):
        # express a pose given in the base LINK frame in the base INERTIAL frame
        return bR.T @ (r - br), bR.T @ R

    out = []
    for name in [base_name] + [j["child"] for j in joints]:
        raw = raw_links[name]
        fr, fR = frame_pose[name]
        lk = _Link()
        lk.name = name
        lk.mass = raw["mass"]
        lk.inertia_local = raw["I"].copy()
        lk.frame_r, lk.frame_R = to_base_inertial(fr, fR)
        lk.r, lk.R = to_base_inertial(fr + fR @ raw["ixyz"], fR @ _rpy_mat(raw["irpy"]))
        for kind, dims, cxyz, crpy in raw["cols"]:
            cr, cR = to_base_inertial(fr + fR @ cxyz, fR @ _rpy_mat(crpy))
            lk.shapes.append((kind, dims, cr, cR))
        out.append(lk)
    return out


class _Body:
    def __init__(self, uid: int, links: list[_Link], fixed_base: bool, path: str):
        self.uid = uid
        self.links = links  # [base, link0, ...]
        self.fixed_base = fixed_base
        self.path = path
        self.pos = np.zeros(3)  # base COM, world
        self.quat = np.array([0.0, 0.0, 0.0, 1.0])
        self.v = np.zeros(3)  # world
        self.w = np.zeros(3)  # world
        self.F = np.zeros(3)  # accumulated external force (world)
        self.T = np.zeros(3)  # accumulated external torque about the base COM (world)
        self.lin_damping = 0.04  # btMultiBody defaults; PyFlyt zeroes them (base_drone.py:301-304)
        self.ang_damping = 0.04
        self.is_plane = any(s[0] == "plane" for lk in links for s in lk.shapes)

    @property
    def num_joints(self) -> int:
        return len(self.links) - 1

    def R(self) -> np.ndarray:
        return _mat_from_quat(self.quat)

    def composite(self):
        """Total mass, COM offset c (base frame) and inertia about the base origin (base frame)."""
        M = 0.0
        mc = np.zeros(3)
        I_O = np.zeros((3, 3))
        for lk in self.links:
            m = lk.mass
            M += m
            mc += m * lk.r
            I_O += lk.R @ lk.inertia_local @ lk.R.T
            I_O += m * (np.dot(lk.r, lk.r) * np.eye(3) - np.outer(lk.r, lk.r))
        c = mc / M if M > 0.0 else np.zeros(3)
        return M, c, I_O

    def lowest_point_and_threshold(self):
        """Lowest world-z over all collision primitives and the relative breaking threshold."""
        Rb = self.R()
        best = None
        for lk in self.links:
            for kind, dims, cr, cR in lk.shapes:
                centre = self.pos + Rb @ cr
                Rw = Rb @ cR
                if kind == "box":
                    half = 0.5 * dims
                    extent = abs(Rw[2, 0]) * half[0] + abs(Rw[2, 1]) * half[1] + abs(Rw[2, 2]) * half[2]
                    disc = float(np.linalg.norm(half))
                elif kind == "cylinder":
                    rad, length = dims
                    az = Rw[2, 2]
                    extent = 0.5 * length * abs(az) + rad * math.sqrt(max(0.0, 1.0 - az * az))
                    disc = math.hypot(rad, 0.5 * length)
                elif kind == "sphere":
                    extent = dims[0]
                    disc = dims[0]
                else:
                    continue
                z = centre[2] - extent
                thr = CONTACT_BREAKING_FACTOR * disc
                if best is None or (z - thr) < (best[0] - best[1]):
                    best = (z, thr)
        return best


# --------------------------------------------------------------------------------------
# the world (one per BulletClient)
# --------------------------------------------------------------------------------------
class World:
    def __init__(self):
        self.search_paths: list[str] = []
        self.resetSimulation()

    # ---- world management ------------------------------------------------------------
    def resetSimulation(self, *a, **k):
        self.bodies: dict[int, _Body] = {}
        self._next_uid = 0
        self.gravity = GRAVITY_DEFAULT.copy()
        self.dt = 1.0 / 240.0
        self.contacts: list[tuple] = []

    def setGravity(self, gx, gy, gz, *a, **k):
        self.gravity = np.array([gx, gy, gz], dtype=np.float64)

    def setTimeStep(self, timeStep, *a, **k):
        self.dt = float(timeStep)

    def setAdditionalSearchPath(self, path, *a, **k):
        self.search_paths.append(path)

    def disconnect(self, *a, **k):
        self.bodies = {}

    # no-op GUI hooks
    def addUserDebugText(self, *a, **k):
        return 0

    def resetDebugVisualizerCamera(self, *a, **k):
        return None

    def configureDebugVisualizer(self, *a, **k):
        return None

    def changeVisualShape(self, *a, **k):
        return None

    def getDebugVisualizerCamera(self, *a, **k):
        return (0, 0, tuple([0.0] * 16), tuple([0.0] * 16)) + (None,) * 8

    # ---- loading ---------------------------------------------------------------------
    def _resolve(self, fileName: str) -> str:
        if os.path.isabs(fileName) and os.path.exists(fileName):
            return fileName
        for sp in [os.getcwd()] + self.search_paths:
            cand = os.path.join(sp, fileName)
            if os.path.exists(cand):
                return cand
        raise FileNotFoundError(f"fake bullet cannot find {fileName!r} in {self.search_paths}")

    def loadURDF(
        self,
        fileName,
        basePosition=(0.0, 0.0, 0.0),
        baseOrientation=(0.0, 0.0, 0.0, 1.0),
        useFixedBase=False,
        flags=0,
        globalScaling=1.0,
        **k,
    ):
        path = self._resolve(fileName)
        links = parse_urdf(path, float(globalScaling))
        uid = self._next_uid
        self._next_uid += 1
        body = _Body(uid, links, bool(useFixedBase) or sum(l.mass for l in links) == 0.0, path)
        body.quat = np.array(baseOrientation, dtype=np.float64)
        # basePosition places the base LINK frame; the stored pose is the base inertial frame
        body.pos = np.array(basePosition, dtype=np.float64) - body.R() @ links[0].frame_r
        self.bodies[uid] = body
        return uid

    def removeBody(self, bodyUniqueId, *a, **k):
        self.bodies.pop(bodyUniqueId, None)

    def getNumBodies(self, *a, **k):
        return len(self.bodies)

    def getBodyUniqueId(self, serialIndex, *a, **k):
        return list(self.bodies.keys())[serialIndex]

    def getBodyInfo(self, bodyUniqueId, *a, **k):
        b = self.bodies[bodyUniqueId]
        return (b.links[0].name.encode(), os.path.basename(b.path).encode())

    def getNumJoints(self, bodyUniqueId, *a, **k):
        return self.bodies[bodyUniqueId].num_joints

    def getJointInfo(self, bodyUniqueId, jointIndex, *a, **k):
        b = self.bodies[bodyUniqueId]
        info = [None] * 17
        info[0] = jointIndex
        info[12] = b.links[jointIndex + 1].name.encode()
        return tuple(info)

    # ---- state access ----------------------------------------------------------------
    def getBasePositionAndOrientation(self, bodyUniqueId, *a, **k):
        b = self.bodies[bodyUniqueId]
        return tuple(b.pos), tuple(b.quat)

    def getBaseVelocity(self, bodyUniqueId, *a, **k):
        b = self.bodies[bodyUniqueId]
        return tuple(b.v), tuple(b.w)

    def resetBasePositionAndOrientation(self, bodyUniqueId, posObj, ornObj, *a, **k):
        b = self.bodies[bodyUniqueId]
        b.pos = np.array(posObj, dtype=np.float64)
        b.quat = np.array(ornObj, dtype=np.float64)
        # PyBullet zeroes the base velocity on a pose reset
        b.v = np.zeros(3)
        b.w = np.zeros(3)

    def resetBaseVelocity(self, objectUniqueId, linearVelocity=None, angularVelocity=None, *a, **k):
        b = self.bodies[objectUniqueId]
        if linearVelocity is not None:
            b.v = np.array(linearVelocity, dtype=np.float64)
        if angularVelocity is not None:
            b.w = np.array(angularVelocity, dtype=np.float64)

    def changeDynamics(
        self,
        bodyUniqueId,
        linkIndex,
        mass=None,
        localInertiaDiagonal=None,
        linearDamping=None,
        angularDamping=None,
        **k,
    ):
        b = self.bodies[bodyUniqueId]
        lk = b.links[linkIndex + 1]
        if mass is not None:
            lk.mass = float(mass)
        if localInertiaDiagonal is not None:
            lk.inertia_local = np.diag(np.asarray(localInertiaDiagonal, dtype=np.float64))
        if linearDamping is not None:
            b.lin_damping = float(linearDamping)
        if angularDamping is not None:
            b.ang_damping = float(angularDamping)

    def getDynamicsInfo(self, bodyUniqueId, linkIndex, *a, **k):
        lk = self.bodies[bodyUniqueId].links[linkIndex + 1]
        return (lk.mass, 0.5, tuple(np.diag(lk.inertia_local)), tuple(lk.r), (0, 0, 0, 1), 0, 0, 0, -1, -1, 2, 0.001)

    def _link_state(self, b: _Body, linkIndex: int):
        lk = b.links[linkIndex + 1]
        Rb = b.R()
        com = b.pos + Rb @ lk.r
        com_q = _quat_from_mat(Rb @ lk.R)
        frame = b.pos + Rb @ lk.frame_r
        frame_q = _quat_from_mat(Rb @ lk.frame_R)
        # local inertial offset expressed in the URDF link frame
        loc_r = lk.frame_R.T @ (lk.r - lk.frame_r)
        loc_q = _quat_from_mat(lk.frame_R.T @ lk.R)
        lin = b.v + np.cross(b.w, Rb @ lk.r)
        return (tuple(com), tuple(com_q), tuple(loc_r), tuple(loc_q), tuple(frame), tuple(frame_q), tuple(lin), tuple(b.w))

    def getLinkState(self, bodyUniqueId, linkIndex, computeLinkVelocity=0, computeForwardKinematics=0, *a, **k):
        st = self._link_state(self.bodies[bodyUniqueId], int(linkIndex))
        return st if computeLinkVelocity else st[:6]

    def getLinkStates(self, bodyUniqueId, linkIndices, computeLinkVelocity=0, computeForwardKinematics=0, *a, **k):
        b = self.bodies[bodyUniqueId]
        out = []
        for li in linkIndices:
            st = self._link_state(b, int(li))
            out.append(st if computeLinkVelocity else st[:6])
        return tuple(out)

    # ---- forces ----------------------------------------------------------------------
    def _link_rot_and_arm(self, b: _Body, linkIndex: int):
        Rb = b.R()
        lk = b.links[linkIndex + 1]
        return Rb @ lk.R, Rb @ lk.r

    def applyExternalForce(self, objectUniqueId, linkIndex, forceObj, posObj, flags, *a, **k):
        b = self.bodies[objectUniqueId]
        Rl, arm = self._link_rot_and_arm(b, int(linkIndex))
        f = np.asarray(forceObj, dtype=np.float64)
        p = np.asarray(posObj, dtype=np.float64)
        if flags == LINK_FRAME:
            fw = Rl @ f
            rel = Rl @ p
        else:
            fw = f
            rel = p - (b.pos + arm)
        b.F += fw
        b.T += np.cross(arm + rel, fw)

    def applyExternalTorque(self, objectUniqueId, linkIndex, torqueObj, flags, *a, **k):
        b = self.bodies[objectUniqueId]
        Rl, _ = self._link_rot_and_arm(b, int(linkIndex))
        t = np.asarray(torqueObj, dtype=np.float64)
        b.T += Rl @ t if flags == LINK_FRAME else t

    # ---- stepping --------------------------------------------------------------------
    def _detect_contacts(self):
        """Ground/pad contact FLAGS only (no response).  Uses the poses at the START of the
        step: Bullet runs collision detection before it integrates transforms."""
        self.contacts = []
        statics = [b for b in self.bodies.values() if b.fixed_base]
        for b in self.bodies.values():
            if b.fixed_base:
                continue
            low = b.lowest_point_and_threshold()
            if low is None:
                continue
            z, thr = low
            for s in statics:
                if s.is_plane:
                    top = s.pos[2]
                    hit = True
                else:
                    # static primitive (landing pad): cylinder top face, inside its radius
                    top, hit = None, False
                    for lk in s.links:
                        for kind, dims, cr, cR in lk.shapes:
                            if kind == "cylinder":
                                centre = s.pos + s.R() @ cr
                                ttop = centre[2] + 0.5 * dims[1]
                                if math.hypot(b.pos[0] - centre[0], b.pos[1] - centre[1]) <= dims[0]:
                                    top, hit = ttop, True
                    if not hit:
                        continue
                if z - top < thr:
                    self.contacts.append((0, s.uid, b.uid, -1, -1))
                    self.contacts.append((0, b.uid, s.uid, -1, -1))

    def getContactPoints(self, bodyA=None, bodyB=None, *a, **k):
        out = self.contacts
        if bodyA is not None:
            out = [c for c in out if c[1] == bodyA or c[2] == bodyA]
        if bodyB is not None:
            out = [c for c in out if c[1] == bodyB or c[2] == bodyB]
        return tuple(out)

    # ---- contact RESPONSE (opt-in) ------------------------------------------------
    # A restatement of a Bullet-like sequential-impulse contact for a free body landing on a horizontal surface (ground
    # plane z = 0, or the top of the landing pad under the base).  **Unpinned**: Bullet's btMultiBodyConstraintSolver is
    # not reproduced (no manifold reduction, no warm starting, no split impulse); what is kept is its structure — velocities
    # are predicted from the applied forces, contacts are found on the pose at the START of the step, impulses act on the
    # predicted velocities, then the pose is integrated — with: candidate points = the 8 corners of every box and 4 + 4 rim
    # points of every cylinder; per point, in order, CONTACT_ITERATIONS sweeps of a non-accumulated normal impulse that
    # removes the approach velocity and pushes out with Baumgarte bias ERP * depth / dt (restitution 0), then Coulomb
    # friction (mu = CONTACT_FRICTION) against the tangential velocity.  Same arithmetic in oracle/pfb_oracle.c and in the
    # CUDA rocket kernel.  Off by default: the fixtures recorded before it existed only carry the contact FLAG.
    CONTACT_ITERATIONS = 8
    CONTACT_ERP = 0.2
    CONTACT_SLOP = 0.001
    CONTACT_FRICTION = 0.5
    contact_response = False

    @staticmethod
    def contact_points(links):
        """candidate contact points in the base inertial frame, fixed order: shapes in link order, 8 points each"""
        pts = []
        for lk in links:
            for kind, dims, cr, cR in lk.shapes:
                if kind == "box":
                    h = 0.5 * np.asarray(dims)
                    loc = [(sx * h[0], sy * h[1], sz * h[2]) for sz in (-1, 1) for sy in (-1, 1) for sx in (-1, 1)]
                elif kind == "cylinder":
                    rad, half = dims[0], 0.5 * dims[1]
                    loc = [(rad * cx, rad * cy, sz * half) for sz in (-1, 1) for cx, cy in ((1, 0), (0, 1), (-1, 0), (0, -1))]
                else:
                    continue
                for q in loc:
                    pts.append(cr + cR @ np.asarray(q, dtype=np.float64))
        return pts

    def _surface_height(self, b):
        """height of the horizontal surface under the body: the top of a static cylinder (landing pad) if the base is over it"""
        top = 0.0
        for s in self.bodies.values():
            if not s.fixed_base or s.is_plane:
                continue
            for lk in s.links:
                for kind, dims, cr, cR in lk.shapes:
                    if kind == "cylinder":
                        centre = s.pos + s.R() @ cr
                        if math.hypot(b.pos[0] - centre[0], b.pos[1] - centre[1]) <= dims[0]:
                            top = max(top, centre[2] + 0.5 * dims[1])
        return top

    def _solve_contacts(self, b, Rb, M, c, I_O):
        dt = self.dt
        top = self._surface_height(b)
        cw = Rb @ c
        Ic = I_O - M * (np.dot(c, c) * np.eye(3) - np.outer(c, c))
        Iinv = Rb @ np.linalg.inv(Ic) @ Rb.T  # world-frame inverse inertia about the COM
        vc = b.v + np.cross(b.w, cw)          # COM velocity
        w = b.w.copy()
        pts = [Rb @ p for p in self.contact_points(b.links)]
        n = np.array([0.0, 0.0, 1.0])
        touched = False
        for _ in range(self.CONTACT_ITERATIONS):
            for pw in pts:
                depth = top - (b.pos[2] + pw[2])
                if depth <= 0.0:
                    continue
                touched = True
                r = pw - cw
                u = vc + np.cross(w, r)
                rn = np.cross(r, n)
                kn = 1.0 / M + np.dot(rn, Iinv @ rn)
                bias = self.CONTACT_ERP * max(depth - self.CONTACT_SLOP, 0.0) / dt
                jn = max(0.0, (bias - u[2]) / kn)
                if jn > 0.0:
                    vc = vc + (jn / M) * n
                    w = w + Iinv @ (jn * rn)
                    u = vc + np.cross(w, r)
                    ut = np.array([u[0], u[1], 0.0])
                    sp = math.sqrt(ut[0] * ut[0] + ut[1] * ut[1])
                    if sp > 1e-12:
                        t = ut / sp
                        rt = np.cross(r, t)
                        kt = 1.0 / M + np.dot(rt, Iinv @ rt)
                        jt = min(sp / kt, self.CONTACT_FRICTION * jn)
                        vc = vc - (jt / M) * t
                        w = w - Iinv @ (jt * rt)
        if touched:
            b.w = w
            b.v = vc - np.cross(w, cw)

    def stepSimulation(self, *a, **k):
        dt = self.dt
        self._detect_contacts()
        for b in self.bodies.values():
            if b.fixed_base:
                b.F[:] = 0.0
                b.T[:] = 0.0
                continue
            Rb = b.R()
            M, c, I_O = b.composite()
            # gravity on every link (acts at each link COM)
            Fw = b.F + M * self.gravity
            Tw = b.T + np.cross(Rb @ (M * c), self.gravity)
            # body-frame Newton-Euler about the base origin O
            F = Rb.T @ Fw
            T = Rb.T @ Tw
            w = Rb.T @ b.w
            # optional Bullet damping (zero for every PyFlyt body)
            if b.lin_damping != 0.0 or b.ang_damping != 0.0:
                vb = Rb.T @ b.v
                F = F - b.lin_damping * M * vb
                T = T - b.ang_damping * (I_O @ w)
            cx = np.array([[0.0, -c[2], c[1]], [c[2], 0.0, -c[0]], [-c[1], c[0], 0.0]])
            A = np.zeros((6, 6))
            A[:3, :3] = M * np.eye(3)
            A[:3, 3:] = -M * cx
            A[3:, :3] = M * cx
            A[3:, 3:] = I_O
            rhs = np.concatenate([F - M * np.cross(w, np.cross(w, c)), T - np.cross(w, I_O @ w)])
            sol = np.linalg.solve(A, rhs)
            a_O, wdot = sol[:3], sol[3:]
            b.w = np.clip(b.w + (Rb @ wdot) * dt, -MAX_COORDINATE_VELOCITY, MAX_COORDINATE_VELOCITY)
            b.v = np.clip(b.v + (Rb @ a_O) * dt, -MAX_COORDINATE_VELOCITY, MAX_COORDINATE_VELOCITY)
            if self.contact_response:
                self._solve_contacts(b, Rb, M, c, I_O)
            # semi-implicit Euler with the NEW velocities
            b.pos = b.pos + b.v * dt
            ang = float(np.linalg.norm(b.w))
            if ang * dt > ANGULAR_MOTION_THRESHOLD:
                ang = ANGULAR_MOTION_THRESHOLD / dt
            if ang < 0.001:
                axis = b.w * (0.5 * dt - (dt * dt * dt) * 0.020833333333 * ang * ang)
            else:
                axis = b.w * (math.sin(0.5 * ang * dt) / ang)
            dq = np.array([axis[0], axis[1], axis[2], math.cos(ang * dt * 0.5)])
            q = _quat_mul(dq, b.quat)
            b.quat = q / math.sqrt(float(np.dot(q, q)))
            b.F[:] = 0.0
            b.T[:] = 0.0
