"""Analytic wind fields: the UNMODIFIED reference is flown with our AnalyticWind registered through
its own ``Aviary.register_wind_field_function`` (tools/gen_golden.py wind_fixtures -> tests/golden/wind_*.npz); the C oracle
and the CUDA kernels evaluate the same field in place of that Python callback."""
import numpy as np
import pytest

from engines import OracleEngine, load_golden, make_cuda_engine, replay_aviary, replay_hover, replay_vehicle

VEHICLES = ["wind_fixedwing_log", "wind_rocket_exp", "wind_rocket_constant"]


def test_analytic_wind_is_a_reference_wind_field_function():
    from pyflyt_b200.core.wind import AnalyticWind

    p = np.array([[0.0, 0.0, 0.5], [3.0, -2.0, 10.0], [1.0, 1.0, -4.0], [0.0, 0.0, 40.0]])
    w = AnalyticWind("exp", base=(0, 0, 1), z_ref=1.0)(0.0, p)
    assert w.shape == (4, 3) and np.allclose(w[:, 2], np.exp(p[:, 2])) and not w[:, :2].any()  # tests/test_core.py:275-278 of the reference
    w = AnalyticWind("power", base=(5.0, 0, 0), z_ref=10.0, alpha=1 / 7)(0.0, p)
    assert np.allclose(w[:, 0], 5.0 * (np.maximum(p[:, 2], 0) / 10.0) ** (1 / 7))
    w = AnalyticWind("log", base=(0, 2.0, 0), z_ref=10.0, z0=0.03)(0.0, p)
    assert w[2, 1] == 0.0 and np.isclose(w[1, 1], 2.0) and np.isclose(w[3, 1], 2.0 * np.log(40 / 0.03) / np.log(10 / 0.03))
    assert np.allclose(AnalyticWind("constant", base=(1, 2, 3))(7.0, p), [[1, 2, 3]] * 4)
    with pytest.raises(ValueError):
        AnalyticWind("thermal")


def test_oracle_reproduces_reference_in_wind():
    err = replay_aviary(OracleEngine, load_golden("wind_quadx_power"))
    assert err["contact_mismatch"] == 0
    for k in ("pos", "euler", "angvel", "linvel", "aux"):
        assert err[k] < 1e-9, (k, err[k])
    for name in VEHICLES:
        err = replay_vehicle(OracleEngine, load_golden(name))
        assert err["contact_mismatch"] == 0
        for k in ("pos", "euler", "angvel", "linvel", "aux"):
            assert err[k] < 1e-9, (name, k, err[k])
    err = replay_hover(OracleEngine, load_golden("wind_hover_quat"))
    assert err["flag_mismatch"] == 0 and err["obs"] < 1e-9 and err["reward"] < 1e-9, err


def test_wind_matters():
    """the same fixture replayed in still air misses the reference by metres: the field is not a no-op"""
    g = dict(np.load(load_golden("wind_fixedwing_log").fid.name))
    class G(dict):
        files = [k for k in g if not k.startswith("wind_")]
    still = G({k: g[k] for k in G.files})
    err = replay_vehicle(OracleEngine, still)
    assert err["pos"] > 1.0


@pytest.mark.gpu
def test_cuda_matches_reference_in_wind():
    err = replay_aviary(make_cuda_engine, load_golden("wind_quadx_power"))
    assert err["contact_mismatch"] == 0 and err["pos"] < 1e-4 and err["euler"] < 1e-4 and err["linvel"] < 1e-3, err
    for name in VEHICLES:
        err = replay_vehicle(make_cuda_engine, load_golden(name), every=2)
        assert err["contact_mismatch"] == 0
        assert err["pos"] < 5e-4 and err["euler"] < 2e-4 and err["linvel"] < 2e-3, (name, err)
    err = replay_hover(make_cuda_engine, load_golden("wind_hover_quat"))
    assert err["flag_mismatch"] == 0 and err["obs"] < 1e-4 and err["reward"] < 1e-4, err


@pytest.mark.gpu
def test_wind_validation_and_still_air_reset():
    from pyflyt_b200 import _lib
    from pyflyt_b200.core.aviary import BatchedAviary
    from pyflyt_b200.core.wind import AnalyticWind

    av = BatchedAviary(np.array([[0.0, 0.0, 5.0]]), np.zeros((1, 3)))
    with pytest.raises(TypeError):
        av.register_wind_field(lambda t, p: p * 0.0)
    w = AnalyticWind("log", base=(1, 0, 0), z_ref=10.0, z0=0.03)
    w.z0 = 20.0  # invalid after construction: the library checks as well
    with pytest.raises(_lib.PfbError):
        av.register_wind_field(w)
    av.register_wind_field(AnalyticWind("constant", base=(4.0, 0.0, 0.0)))
    av.set_mode(-1)
    av.step(60)
    drift = float(av.all_states[0, 3, 0])
    av.register_wind_field(None)
    av.reset()
    av.set_mode(-1)
    av.step(60)
    assert drift > 1e-3 and abs(float(av.all_states[0, 3, 0])) < 1e-6  # pushed downwind, then still air again
