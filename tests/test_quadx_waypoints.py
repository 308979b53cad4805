"""QuadX-Waypoints (gym_envs/quadx_envs/quadx_waypoints_env.py): the oracle against the unmodified
reference env (fixtures from tools/gen_golden.py qxwp), then the CUDA kernel against the fixtures and the oracle."""
import glob
import os

import numpy as np
import pytest

from engines import GOLDEN, OracleEngine, build_model, load_golden, make_cuda_engine, quadx_waypoints_config, replay_waypoints

FIX = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "qxwp_*.npz")))


def test_fixtures_present():
    assert len(FIX) >= 4
    reached = sum(int((load_golden(n)["info"] >> 3).max()) for n in FIX)
    assert reached >= 5  # the position-mode fixtures reach their waypoints


@pytest.mark.parametrize("name", FIX)
def test_oracle_reproduces_reference(name):
    err = replay_waypoints(OracleEngine, load_golden(name))
    assert err["flag_mismatch"] == 0, err
    assert err["obs"] < 1e-7 and err["reward"] < 1e-6, err
