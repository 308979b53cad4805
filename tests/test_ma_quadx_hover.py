"""MAQuadXHover (pz_envs/quadx_envs/ma_quadx_hover_env.py): the oracle against the unmodified reference
PettingZoo env (fixtures from tools/gen_golden.py mahover), then the CUDA per-agent kernel and the arena bookkeeping."""
import glob
import os

import numpy as np
import pytest

from engines import GOLDEN, OracleEngine, load_golden, make_cuda_engine, replay_ma_hover

FIX = sorted(os.path.basename(p)[:-4] for p in glob.glob(os.path.join(GOLDEN, "mahover_*.npz")))


def test_fixtures_present():
    assert len(FIX) >= 4


@pytest.mark.parametrize("name", FIX)
def test_oracle_reproduces_reference(name):
    err = replay_ma_hover(OracleEngine, load_golden(name))
    assert err["flag_mismatch"] == 0, err
    assert err["obs"] < 1e-7 and err["reward"] < 1e-6, err
