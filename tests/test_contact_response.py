"""Ground / pad contact RESPONSE for Rocket-Landing.  The UNMODIFIED reference env is flown onto the pad
by a scripted ignition law on oracle/fakebullet with its (restated, unpinned) sequential-impulse contact switched on
(tools/gen_golden.py touchdown_fixtures): gentle touchdowns rest on the legs and end in ``env_complete``
(rocket_landing_env.py:231-263), a hard one is a fatal collision.  The C oracle and the CUDA kernel run the same arithmetic."""
import numpy as np
import pytest

from engines import OracleEngine, load_golden, make_cuda_engine, replay_landing

FIXTURES = {"landing_touchdown": 4, "landing_touchdown_soft_euler": 4, "landing_touchdown_hard": 2}  # last info: 4 complete, 2 fatal


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_fixture_outcomes(name):
    g = load_golden(name)
    assert int(g["info"][-1]) == FIXTURES[name]
    if FIXTURES[name] == 4:
        assert bool(g["trunc"][-1]) and not bool(g["term"][-1])
        rest = g["obs"][:, -1] > 0  # landing_pad_contact
        assert rest.sum() >= 20  # the rocket sat on the pad for many env steps before the env called it complete
        assert np.ptp(g["obs"][rest][5:, 12 if str(g["angle_representation"]) == "quaternion" else 11]) < 5e-3  # z at rest


@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_oracle_reproduces_reference_touchdown(name):
    err = replay_landing(OracleEngine, load_golden(name))
    assert err["flag_mismatch"] == 0, err
    assert err["obs"] < 1e-8 and err["reward"] < 1e-7, err


@pytest.mark.gpu
@pytest.mark.parametrize("name", sorted(FIXTURES))
def test_cuda_reproduces_reference_touchdown(name):
    err = replay_landing(make_cuda_engine, load_golden(name))
    assert err["flag_mismatch"] == 0, err
    assert err["obs"] < 5e-3 and err["reward"] < 5e-2, err
