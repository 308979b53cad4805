"""CPU checks of the NumPy restatement of the device random streams (tests/philox_replay.py)."""
import ctypes as C

import numpy as np

from philox_replay import Streams, box_muller, philox4x32_10, unit_open


def test_philox_known_answers():
    """Random123 known-answer vectors for Philox4x32-10 (kat_vectors of the reference distribution)."""
    kat = [
        ((0, 0, 0, 0), (0, 0), (0x6627E8D5, 0xE169C58D, 0xBC57AC4C, 0x9B00DBD8)),
        ((0xFFFFFFFF,) * 4, (0xFFFFFFFF, 0xFFFFFFFF), (0x408F276D, 0x41C83B0E, 0xA20BC7C6, 0x6D5451FD)),
        ((0x243F6A88, 0x85A308D3, 0x13198A2E, 0x03707344), (0xA4093822, 0x299F31D0), (0xD16CFE09, 0x94FDCCEB, 0x5001E420, 0x24126EA1)),
    ]
    for ctr, key, want in kat:
        got = philox4x32_10(*[np.array([c]) for c in ctr], key[0], key[1])
        assert tuple(int(g[0]) for g in got) == want


def test_matches_the_kernel_header_compiled_for_the_host():
    """pfb_common.cuh is host-compilable: the C++ philox / unit_open the kernels use == the NumPy restatement, bit for bit;
    the Box-Muller pair agrees to float32 round-off (libm here, fast intrinsics on the device)."""
    from engines import hostsim_lib

    L = hostsim_lib()
    rng = np.random.default_rng(5)
    ctr = rng.integers(0, 2**32, size=(1000, 4), dtype=np.uint64).astype(np.uint32)
    k0, k1 = 0x12345678, 0x9ABCDEF0
    out = np.zeros((1000, 4), dtype=np.uint32)
    nrm = np.zeros((1000, 4), dtype=np.float32)
    L.hs_philox(ctr.ctypes.data_as(C.c_void_p), C.c_uint32(k0), C.c_uint32(k1), out.ctypes.data_as(C.c_void_p), nrm.ctypes.data_as(C.c_void_p), C.c_int64(1000))
    got = philox4x32_10(ctr[:, 0], ctr[:, 1], ctr[:, 2], ctr[:, 3], k0, k1)
    assert np.array_equal(np.stack(got, axis=1), out)
    n0, n1 = box_muller(out[:, 0], out[:, 1])
    n2, n3 = box_muller(out[:, 2], out[:, 3])
    assert np.abs(np.stack([n0, n1, n2, n3], axis=1) - nrm).max() < 5e-6


def test_unit_open_range_and_exactness():
    u = np.array([0, 255, 256, 0xFFFFFFFF], dtype=np.uint32)
    x = unit_open(u)
    assert x.dtype == np.float32 and x[0] == np.float32(2.0**-24) and x[1] == x[0] and x[3] == np.float32(1.0)


def test_normals_are_standard_normal():
    """Moments and a Kolmogorov-Smirnov test of 2.6 M draws of the motor-noise stream: N(4, 1) (motors.py:134-138 quirk)."""
    from scipy import stats

    s = Streams(seed=7, n_envs=65536)
    z = np.concatenate([s.step_noise(k).ravel() for k in range(7)]).astype(np.float64) - 4.0
    assert abs(z.mean()) < 3e-3 and abs(z.std() - 1.0) < 3e-3
    assert abs(stats.skew(z)) < 5e-3 and abs(stats.kurtosis(z)) < 1e-2
    assert stats.kstest(z[:200000], "norm").pvalue > 1e-3
    # independent across envs and across the two halves of a Philox call
    a = s.step_noise(0)
    assert abs(np.corrcoef(a[0], a[1])[0, 1]) < 0.02 and abs(np.corrcoef(a[0, :-1], a[0, 1:])[0, 1]) < 0.02


def test_streams_do_not_depend_on_sharding():
    full = Streams(seed=3, n_envs=4096)
    hi = Streams(seed=3, n_envs=2048, env_offset=2048)
    assert np.array_equal(full.step_noise(5)[:, 2048:], hi.step_noise(5))
    assert np.array_equal(full.actions(9)[2048:], hi.actions(9))
    assert np.array_equal(full.autoreset_noise(np.arange(4096) % 7 + 1)[:, 2048:], hi.autoreset_noise(np.arange(2048, 4096) % 7 + 1))
