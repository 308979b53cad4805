"""Vehicle tables: the host-side URDF/YAML -> PfbModel path (SURVEY.md §A.2 cross-check)."""
import ctypes
import os

import numpy as np
import pytest

from pyflyt_b200.models import build_model, load_vehicle
from pyflyt_b200.models.urdf import composite_rigid_body, load_urdf_links


def test_cf2x_table():
    m = build_model("quadx")
    assert m.mass == pytest.approx(0.027)
    assert list(m.com) == [0.0, 0.0, 0.0]
    assert np.allclose(np.array(m.inertia).reshape(3, 3), np.diag([1.4e-5, 1.4e-5, 2.17e-5]))
    assert [list(p) for p in m.motor_pos] == [[0.028, -0.028, 0.0], [-0.028, 0.028, 0.0], [0.028, 0.028, 0.0], [-0.028, -0.028, 0.0]]
    assert m.max_rpm[0] == pytest.approx(np.sqrt(2.0 / (4 * 3.16e-10)))
    assert m.drag_const[0] == pytest.approx(0.5 * 1.225 * 3.0 * 4e-4)
    assert list(m.torque_coef) == [-7.94e-12, -7.94e-12, 7.94e-12, 7.94e-12]
    assert m.motor_dt_over_tau[0] == pytest.approx((1 / 240) / 0.01)
    assert list(m.pid[0][0]) == [4.0e-2, 4.0e-2, 8.0e-2]
    assert m.pid[4][0][0] == 2.0 and m.pid[4][1][0] == 0.5 and m.pid[4][2][0] == 0.05  # z_vel
    assert m.n_shapes == 1 and list(m.shapes[0].dims) == [0.045, 0.045, 0.01]


@pytest.mark.parametrize(
    "kind,name,mass,com,diag",
    [
        ("quadx", "primitive_drone", 1.0, (0, 0, 0), (0.01, 0.01, 0.016)),
        ("fixedwing", "fixedwing", 2.35, (-0.45319, 0, 0.00319), (0.36212, 0.61012, 0.97)),
        ("fixedwing", "acrowing", 2.35, (-0.39574, 0, 0.00532), (0.36412, 0.49738, 0.85525)),
        ("rocket", "rocket", 549.1, (0, 0, -0.17044), (2431.88, 2431.88, 9.3945)),
    ],
)
def test_composite_bodies_match_survey(kind, name, mass, com, diag):
    m = build_model(kind, name)
    assert m.mass == pytest.approx(mass, rel=1e-6)
    assert np.allclose(list(m.com), com, atol=1e-5)
    assert np.allclose(np.diag(np.array(m.inertia).reshape(3, 3)), diag, rtol=2e-5)


def test_rocket_dry_plus_fuel_is_full():
    m = build_model("rocket", starting_fuel_ratio=0.05)
    assert m.dry_mass + m.fuel_total_mass == pytest.approx(m.mass)
    assert m.starting_fuel_ratio == 0.05
    assert m.n_surfaces == 4 and [list(s.pos) for s in m.surfaces[:2]] == [[0.0, 0.0, 0.0], [0.0, 0.0, -2.0]]  # rocket.py:113 quirk


def test_bad_inputs_raise_like_the_reference():
    with pytest.raises(ValueError):
        build_model("quadx", physics_hz=240, control_hz=7)  # base_drone.py:94-97
    with pytest.raises(ValueError):
        build_model("blimp")
    with pytest.raises(FileNotFoundError):
        build_model("quadx", "no_such_drone")


def test_model_dir_layout_roundtrip(tmp_path):
    """A user model directory in the reference's layout gives the same table as the built-in one."""
    import yaml

    links, params = load_vehicle("cf2x")
    d = tmp_path / "mydrone"
    d.mkdir()
    urdf = ['<?xml version="1.0" ?>', '<robot name="x">']
    for lk in links:
        I = np.array(lk.inertia)
        urdf.append(
            f'<link name="{lk.name}"><inertial><origin rpy="0 0 0" xyz="{lk.com[0]} {lk.com[1]} {lk.com[2]}"/>'
            f'<mass value="{lk.mass}"/><inertia ixx="{I[0,0]}" ixy="0" ixz="0" iyy="{I[1,1]}" iyz="0" izz="{I[2,2]}"/></inertial>'
            + ('<collision><geometry><box size="0.09 0.09 0.02"/></geometry></collision>' if lk.index == -1 else "")
            + "</link>"
        )
    for lk in links[1:]:
        urdf.append(f'<joint name="j{lk.index}" type="fixed"><parent link="{links[0].name}"/><child link="{lk.name}"/></joint>')
    urdf.append("</robot>")
    (d / "mydrone.urdf").write_text("\n".join(urdf))
    (d / "mydrone.yaml").write_text(yaml.safe_dump(params))
    a, b = build_model("quadx", "cf2x"), build_model("quadx", "mydrone", model_dir=str(tmp_path))
    assert bytes(a) == bytes(b)


def test_non_fixed_joint_is_rejected(tmp_path):
    p = tmp_path / "bad.urdf"
    p.write_text('<robot name="b"><link name="a"/><link name="c"/><joint name="j" type="revolute"><parent link="a"/><child link="c"/></joint></robot>')
    with pytest.raises(ValueError):
        load_urdf_links(str(p))
