"""TEST INFRASTRUCTURE — ctypes binding of oracle/pfb_oracle.c (the fp64 CPU restatement).

Only tests/, __graft_entry__.smoke() and bench.py's CPU-baseline legs may import this module.
"""

from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.realpath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "libpfb_oracle.so")
_lib = None


def build(force: bool = False) -> str:
    src = os.path.join(_HERE, "pfb_oracle.c")
    hdr = os.path.join(_HERE, "..", "include", "pyflyt_b200.h")
    stale = (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < max(os.path.getmtime(src), os.path.getmtime(hdr))
    )
    if stale:
        subprocess.run(["make", "-C", _HERE, "-B" if force else "-s"], check=True, capture_output=True)
    return _LIB_PATH


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        L = C.CDLL(_LIB_PATH)
        vp, i64, u64, dp, u8p = C.c_void_p, C.c_int64, C.c_uint64, C.POINTER(C.c_double), C.POINTER(C.c_uint8)
        L.orc_create.restype = vp
        L.orc_create.argtypes = [vp, vp, i64, u64]
        L.orc_destroy.argtypes = [vp]
        L.orc_set_start.argtypes = [vp, dp, dp]
        L.orc_df_get_actions.argtypes = [vp, dp]
        L.orc_df_set_actions.argtypes = [vp, C.POINTER(C.c_uint8), dp]
        L.orc_set_wind.argtypes = [vp, vp]
        L.orc_reset.argtypes = [vp, u8p]
        L.orc_set_mode.argtypes = [vp, C.c_int]
        L.orc_set_setpoints.argtypes = [vp, dp, C.c_int]
        L.orc_get_setpoints.argtypes = [vp, dp, C.c_int]
        L.orc_aviary_step.argtypes = [vp, C.c_int, dp]
        L.orc_get_raw.argtypes = [vp, dp, dp, dp, dp]
        L.orc_set_base_velocity.argtypes = [vp, dp, dp]
        L.orc_update_state.argtypes = [vp]
        L.orc_get_state.argtypes = [vp, dp]
        L.orc_get_aux.argtypes = [vp, dp, C.c_int]
        L.orc_get_pwm.argtypes = [vp, dp]
        L.orc_get_contact.argtypes = [vp, u8p]
        L.orc_env_reset.argtypes = [vp, u8p, dp, dp]
        L.orc_env_reset_targets.argtypes = [vp, u8p, dp, dp, dp]
        L.orc_obs_dim.argtypes = [vp]
        L.orc_env_step.argtypes = [vp, dp, dp, dp, dp, u8p, u8p, u8p]
        L.orc_env_rollout.restype = i64
        L.orc_env_rollout.argtypes = [vp, C.c_int]
        L.orc_set_num_threads.argtypes = [C.c_int]
        _lib = L
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_double))


def _u8(a):
    return None if a is None else a.ctypes.data_as(C.POINTER(C.c_uint8))


class Oracle:
    """N independent single-drone worlds stepped on the CPU in fp64."""

    def __init__(self, model, env_config=None, n: int = 1, seed: int = 0, start_pos=None, start_orn=None):
        L = lib()
        assert L.orc_sizeof_model() == C.sizeof(model), "PfbModel layout mismatch between Python and C"
        if env_config is not None:
            assert L.orc_sizeof_env_config() == C.sizeof(env_config)
        self.model, self.env_config, self.n = model, env_config, int(n)
        self._h = L.orc_create(C.byref(model), C.byref(env_config) if env_config is not None else None, self.n, seed)
        if not self._h:
            raise RuntimeError("orc_create failed (ABI mismatch?)")
        self.updates_per_step = L.orc_updates_per_step(C.c_void_p(self._h))
        sp = np.zeros((self.n, 3)) if start_pos is None else np.ascontiguousarray(np.broadcast_to(start_pos, (self.n, 3)), dtype=np.float64)
        so = np.zeros((self.n, 3)) if start_orn is None else np.ascontiguousarray(np.broadcast_to(start_orn, (self.n, 3)), dtype=np.float64)
        L.orc_set_start(self._h, _dp(sp), _dp(so))
        self.obs_dim = L.orc_obs_dim(C.c_void_p(self._h)) if env_config is not None else 21

    def __del__(self):
        if getattr(self, "_h", None):
            lib().orc_destroy(self._h)
            self._h = None

    # --- Aviary level
    def reset(self, mask=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        lib().orc_reset(self._h, _u8(m))

    def set_mode(self, mode: int):
        lib().orc_set_mode(self._h, int(mode))

    def set_wind(self, wind):
        """wind: pyflyt_b200.core.wind.AnalyticWind or None"""
        if wind is None:
            lib().orc_set_wind(self._h, None)
        else:
            w = wind.as_struct()
            lib().orc_set_wind(self._h, C.byref(w))

    def set_setpoints(self, sp):
        sp = np.ascontiguousarray(sp, dtype=np.float64)
        lib().orc_set_setpoints(self._h, _dp(sp), sp.shape[1])

    def get_setpoints(self, dim=4):
        out = np.zeros((self.n, dim))
        lib().orc_get_setpoints(self._h, _dp(out), dim)
        return out

    def aviary_step(self, n_steps: int = 1, noise=None):
        nz = None
        if noise is not None:
            nz = np.ascontiguousarray(noise, dtype=np.float64)
            assert nz.shape == (n_steps * self.updates_per_step, self.n), nz.shape
        lib().orc_aviary_step(self._h, n_steps, _dp(nz))

    def set_base_velocity(self, lin, ang):
        lin = np.ascontiguousarray(np.broadcast_to(lin, (self.n, 3)), dtype=np.float64)
        ang = np.ascontiguousarray(np.broadcast_to(ang, (self.n, 3)), dtype=np.float64)
        lib().orc_set_base_velocity(self._h, _dp(lin), _dp(ang))

    def update_state(self):
        lib().orc_update_state(self._h)

    def raw(self):
        pos, quat, v, w = np.zeros((self.n, 3)), np.zeros((self.n, 4)), np.zeros((self.n, 3)), np.zeros((self.n, 3))
        lib().orc_get_raw(self._h, _dp(pos), _dp(quat), _dp(v), _dp(w))
        return pos, quat, v, w

    def state(self):
        out = np.zeros((self.n, 12))
        lib().orc_get_state(self._h, _dp(out))
        return out.reshape(self.n, 4, 3)

    def aux_state(self, dim=4):
        out = np.zeros((self.n, dim))
        lib().orc_get_aux(self._h, _dp(out), dim)
        return out

    def pwm(self):
        out = np.zeros((self.n, 4))
        lib().orc_get_pwm(self._h, _dp(out))
        return out

    def contact(self):
        out = np.zeros(self.n, dtype=np.uint8)
        lib().orc_get_contact(self._h, _u8(out))
        return out

    # --- gym env level
    def env_reset(self, mask=None, noise=None, targets=None):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        obs = np.zeros((self.n, self.obs_dim))
        if targets is not None:
            tg = np.ascontiguousarray(targets, dtype=np.float64).reshape(self.n, -1)
            lib().orc_env_reset_targets(self._h, _u8(m), _dp(nz), _dp(tg), _dp(obs))
        else:
            lib().orc_env_reset(self._h, _u8(m), _dp(nz), _dp(obs))
        return obs

    def set_start(self, start_pos, start_orn):
        sp = np.ascontiguousarray(np.broadcast_to(start_pos, (self.n, 3)), dtype=np.float64)
        so = np.ascontiguousarray(np.broadcast_to(start_orn, (self.n, 3)), dtype=np.float64)
        lib().orc_set_start(self._h, _dp(sp), _dp(so))

    def df_get_actions(self):
        """MAFixedwingDogfight: [N][8] past_action | current_action of every agent"""
        out = np.zeros((self.n, 8))
        lib().orc_df_get_actions(self._h, _dp(out))
        return out

    def df_set_actions(self, mask, values):
        m = None if mask is None else np.ascontiguousarray(mask, dtype=np.uint8)
        v = np.ascontiguousarray(values, dtype=np.float64).reshape(self.n, 8)
        lib().orc_df_set_actions(self._h, _u8(m), _dp(v))

    def env_step(self, actions, noise=None):
        a = np.ascontiguousarray(actions, dtype=np.float64)
        nz = None if noise is None else np.ascontiguousarray(noise, dtype=np.float64)
        obs = np.zeros((self.n, self.obs_dim))
        rew = np.zeros(self.n)
        term, trunc, info = (np.zeros(self.n, dtype=np.uint8) for _ in range(3))
        lib().orc_env_step(self._h, _dp(a), _dp(nz), _dp(obs), _dp(rew), _u8(term), _u8(trunc), _u8(info))
        return obs, rew, term, trunc, info

    def env_rollout(self, n_steps: int) -> int:
        return int(lib().orc_env_rollout(self._h, int(n_steps)))
