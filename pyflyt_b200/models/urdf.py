"""Host-side vehicle description: URDF (fixed-joint trees only) → flat link table.

Mirrors what the reference obtains from ``p.loadURDF(..., flags=URDF_USE_INERTIA_FROM_FILE)``
(PyFlyt/core/abstractions/base_drone.py:116-122): every child link is kept (not merged), link index
``i`` is the i-th ``<joint>`` in file order, index ``-1`` is the base, and forces applied in
``LINK_FRAME`` at ``[0,0,0]`` act at the link's inertial frame (COM).  Everything is expressed in the
base link's inertial frame, which is the frame PyBullet reports the base pose in.
"""

from __future__ import annotations

import dataclasses
import math
import re
import xml.etree.ElementTree as ET

import numpy as np


def rpy_to_matrix(rpy) -> np.ndarray:
    """URDF fixed-axis roll-pitch-yaw → rotation matrix (Rz(yaw) Ry(pitch) Rx(roll))."""
    r, p, y = (float(v) for v in rpy)
    cr, sr, cp, sp, cy, sy = math.cos(r), math.sin(r), math.cos(p), math.sin(p), math.cos(y), math.sin(y)
    return np.array(
        [
            [cy * cp, cy * sp * sr - sy * cr, cy * sp * cr + sy * sr],
            [sy * cp, sy * sp * sr + cy * cr, sy * sp * cr - cy * sr],
            [-sp, cp * sr, cp * cr],
        ]
    )


@dataclasses.dataclass
class Shape:
    kind: str  # "box" | "cylinder" | "sphere"
    dims: list  # box: full sizes xyz; cylinder: [radius, length]; sphere: [radius]
    at: list  # centre, base inertial frame
    rot: list  # 3x3 row-major rotation, base inertial frame


@dataclasses.dataclass
class Link:
    index: int  # -1 for the base
    name: str
    mass: float
    com: list  # inertial-frame origin in the base inertial frame
    inertia: list  # 3x3 about the link COM, expressed in base-frame axes
    shapes: list

    def to_dict(self) -> dict:
        d = dataclasses.asdict(self)
        return d

    @staticmethod
    def from_dict(d: dict) -> "Link":
        shapes = [Shape(**s) for s in d.get("shapes", [])]
        return Link(int(d["index"]), d["name"], float(d["mass"]), list(d["com"]), [list(r) for r in d["inertia"]], shapes)


def _vec(text, n=3, scale=1.0):
    if text is None:
        return np.zeros(n)
    return np.array([float(t) for t in text.split()], dtype=np.float64) * scale


def load_urdf_links(path: str, scale: float = 1.0) -> list[Link]:
    """Returns ``[base, link0, link1, ...]``; raises on any non-fixed joint."""
    with open(path, "r", encoding="utf-8") as fh:
        xml = fh.read()
    # rocket.urdf carries a stray second </robot>; a strict parser needs the tail cut off
    cut = xml.find("</robot>")
    if cut != -1:
        xml = xml[: cut + 8]
    xml = re.sub(r"<!--.*?-->", "", xml, flags=re.DOTALL)
    robot = ET.fromstring(xml)

    link_nodes = {n.get("name"): n for n in robot.findall("link")}
    joints = []
    for j in robot.findall("joint"):
        if j.get("type") != "fixed":
            raise ValueError(
                f"{path}: joint {j.get('name')!r} is {j.get('type')!r}; the batched stepper models a "
                "single free rigid body, so every joint must be 'fixed'"
            )
        origin = j.find("origin")
        joints.append(
            (
                j.find("parent").get("link"),
                j.find("child").get("link"),
                _vec(origin.get("xyz") if origin is not None else None, 3, scale),
                _vec(origin.get("rpy") if origin is not None else None),
            )
        )
    child_names = {c for _, c, _, _ in joints}
    base_candidates = [n for n in link_nodes if n not in child_names]
    if len(base_candidates) != 1:
        raise ValueError(f"{path}: expected one root link, found {base_candidates}")
    base = base_candidates[0]

    # URDF link frames relative to the base link frame (walk the tree)
    pose = {base: (np.zeros(3), np.eye(3))}
    todo = list(joints)
    while todo:
        remaining = []
        for parent, child, xyz, rpy in todo:
            if parent in pose:
                pt, pr = pose[parent]
                pose[child] = (pt + pr @ xyz, pr @ rpy_to_matrix(rpy))
            else:
                remaining.append((parent, child, xyz, rpy))
        if len(remaining) == len(todo):
            raise ValueError(f"{path}: joint tree is disconnected")
        todo = remaining

    def inertial_of(node):
        ine = node.find("inertial")
        xyz, rpy, mass, tensor = np.zeros(3), np.zeros(3), 0.0, np.zeros((3, 3))
        if ine is not None:
            o = ine.find("origin")
            if o is not None:
                xyz, rpy = _vec(o.get("xyz"), 3, scale), _vec(o.get("rpy"))
            m = ine.find("mass")
            if m is not None:
                mass = float(m.get("value"))
            it = ine.find("inertia")
            if it is not None:
                a = {k: float(it.get(k, 0.0)) for k in ("ixx", "ixy", "ixz", "iyy", "iyz", "izz")}
                tensor = np.array(
                    [[a["ixx"], a["ixy"], a["ixz"]], [a["ixy"], a["iyy"], a["iyz"]], [a["ixz"], a["iyz"], a["izz"]]]
                )
        return xyz, rpy, mass, tensor

    bxyz, brpy, _, _ = inertial_of(link_nodes[base])
    base_rot = rpy_to_matrix(brpy)

    def rebase(t, r):
        return base_rot.T @ (t - bxyz), base_rot.T @ r

    links = []
    for index, name in enumerate([base] + [c for _, c, _, _ in joints]):
        node = link_nodes[name]
        ft, fr = pose[name]
        ixyz, irpy, mass, tensor = inertial_of(node)
        com, axes = rebase(ft + fr @ ixyz, fr @ rpy_to_matrix(irpy))
        shapes = []
        for col in node.findall("collision"):
            o = col.find("origin")
            ct, cr = rebase(
                ft + fr @ _vec(o.get("xyz") if o is not None else None, 3, scale),
                fr @ rpy_to_matrix(_vec(o.get("rpy") if o is not None else None)),
            )
            geo = col.find("geometry")
            if geo is None:
                continue
            if geo.find("box") is not None:
                kind, dims = "box", _vec(geo.find("box").get("size"), 3, scale)
            elif geo.find("cylinder") is not None:
                c = geo.find("cylinder")
                kind, dims = "cylinder", np.array([float(c.get("radius")), float(c.get("length"))]) * scale
            elif geo.find("sphere") is not None:
                kind, dims = "sphere", np.array([float(geo.find("sphere").get("radius"))]) * scale
            else:
                continue  # meshes/planes carry no analytic ground test
            shapes.append(Shape(kind, dims.tolist(), ct.tolist(), cr.tolist()))
        links.append(Link(index - 1, name, mass, com.tolist(), (axes @ tensor @ axes.T).tolist(), shapes))
    return links


def composite_rigid_body(links: list[Link], mass_override: dict | None = None, inertia_override: dict | None = None):
    """Total mass ``M``, COM offset ``c`` and inertia ``I_O`` about the base origin (base axes).

    ``mass_override`` / ``inertia_override`` map link index → new mass / new 3x3 COM inertia
    (the rocket's fuel tank: PyFlyt/core/abstractions/boosters.py:207-212)."""
    M, first, I_O = 0.0, np.zeros(3), np.zeros((3, 3))
    for lk in links:
        m = lk.mass if not mass_override or lk.index not in mass_override else mass_override[lk.index]
        Ic = np.asarray(lk.inertia) if not inertia_override or lk.index not in inertia_override else np.asarray(inertia_override[lk.index])
        r = np.asarray(lk.com)
        M += m
        first += m * r
        I_O += Ic + m * (float(r @ r) * np.eye(3) - np.outer(r, r))
    c = first / M if M > 0 else np.zeros(3)
    return M, c, I_O
