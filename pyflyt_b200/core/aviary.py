"""``BatchedAviary`` — N independent single-drone worlds stepped in lock-step on one B200.

Mirrors the surface of the reference's ``Aviary`` that its environments use
(/root/reference/PyFlyt/core/aviary.py): ctor kwargs ``start_pos``/``start_orn``/``drone_type``/
``drone_options``/``seed``/``physics_hz`` (:69-83), ``reset`` (:218), ``step`` (:480), ``state(i)`` (:335),
``aux_state(i)`` (:353), ``all_states`` (:372), ``set_mode`` (:440), ``set_setpoint`` (:460),
``set_all_setpoints`` (:470), ``contact_array`` (:322), counters (:227-229, :528-531) — with one
difference in meaning: drone ``i`` lives in its OWN world (the reference puts them in one Bullet world),
so ``contact_array[i]`` is "drone i touched the floor during the last step".

All state is held in caller-visible ``torch`` tensors; the CUDA library (libpyflyt_b200.so) only sees
raw device pointers.  There is no CPU path.
"""

from __future__ import annotations

import ctypes as C
from typing import Any, Sequence

import numpy as np
import torch

from .. import _lib
from ..models import PfbEnvConfig, build_model


class AviaryInitException(Exception):
    """Same role as PyFlyt.core.aviary.AviaryInitException (aviary.py:21-44)."""

    def __init__(self, message: str = "AviaryInitException"):
        self.message = message
        super().__init__(self.message)

    def __str__(self) -> str:
        return f"Aviary Error: {self.message}"


def _stream_ptr(device) -> int:
    return torch.cuda.current_stream(device).cuda_stream


class BatchedAviary:
    def __init__(
        self,
        start_pos,
        start_orn,
        drone_type: str = "quadx",
        drone_options: dict[str, Any] | None = None,
        physics_hz: int = 240,
        seed: None | int = None,
        device: str | torch.device = "cuda:0",
        env_config: PfbEnvConfig | None = None,
        env_offset: int = 0,
    ):
        start_pos = np.asarray(start_pos, dtype=np.float32)
        start_orn = np.asarray(start_orn, dtype=np.float32)
        # shape checks with the reference's messages (aviary.py:120-131)
        if start_pos.ndim != 2 or start_pos.shape[-1] != 3:
            raise AviaryInitException(f"start_pos must be shape (n, 3), currently {start_pos.shape}.")
        if start_orn.shape != start_pos.shape:
            raise AviaryInitException(f"start_orn must be same shape as start_pos, currently {start_orn.shape}.")
        if isinstance(drone_type, (tuple, list)):
            if len(set(drone_type)) != 1:
                raise AviaryInitException("the batched stepper runs one vehicle kind per batch; build one BatchedAviary per kind.")
            drone_type = drone_type[0]
        if drone_type not in ("quadx", "fixedwing", "rocket"):
            raise AviaryInitException(f"Can't find `drone_type` {drone_type} amongst known types ['quadx', 'fixedwing', 'rocket'].")
        if not torch.cuda.is_available():
            raise _lib.PfbError("pyflyt_b200 needs a CUDA device (B200, sm_100a); there is no CPU fallback.")
        opts = dict(drone_options or {})
        control_hz = int(opts.pop("control_hz", 120))
        self.device = torch.device(device)
        self.num_drones = int(start_pos.shape[0])
        self.drone_type = drone_type
        self.physics_hz = int(physics_hz)
        self.physics_period = 1.0 / physics_hz
        self.model = build_model(drone_type, opts.pop("drone_model", None), opts.pop("model_dir", None), physics_hz, control_hz, **opts)
        self.env_config = env_config
        self.updates_per_step = int(physics_hz / control_hz)  # aviary.py:288-289 (single control rate)
        self.step_period = 1.0 / control_hz
        self.seed = 0 if seed is None else int(seed)

        L = _lib.lib()
        self._h = C.c_void_p()
        dev_index = self.device.index if self.device.index is not None else torch.cuda.current_device()
        _lib.check(L.pfb_create(C.byref(self.model), C.byref(env_config) if env_config is not None else None, self.num_drones, dev_index, self.seed, C.byref(self._h)))
        _lib.check(L.pfb_set_env_offset(self._h, int(env_offset)))
        n, dev = self.num_drones, self.device
        f32 = dict(dtype=torch.float32, device=dev)
        self.obs_dim = L.pfb_obs_dim(self._h)
        self.setpoint_dim = L.pfb_setpoint_dim(self._h)
        self.aux_dim = L.pfb_aux_dim(self._h)
        # persistent state: fp32 SoA, field-major [F][N] or warp-tiled [N/32][F/4][32][4] (include/pyflyt_b200.h)
        self.state_rows = int(L.pfb_state_rows(self._h))
        self.tiled = int(L.pfb_state_layout(self._h)) == 1
        if self.tiled:
            self.state_tensor = torch.zeros((int(L.pfb_state_floats(self._h)) // (self.state_rows * 32), self.state_rows // 4, 32, 4), **f32)
        else:
            self.state_tensor = torch.zeros((self.state_rows, n), **f32)
        self.istate_tensor = torch.zeros((L.pfb_istate_rows(self._h), n), dtype=torch.int32, device=dev)
        self.setpoints = torch.zeros((n, self.setpoint_dim), **f32)
        self.start_pos = torch.from_numpy(start_pos).to(dev).contiguous()
        self.start_orn = torch.from_numpy(start_orn).to(dev).contiguous()
        # obs | reward | term | trunc live in ONE slab: pfb_env_step_host then returns them with a single D2H copy
        self._out_slab = torch.zeros(self.out_slab_bytes(n, self.obs_dim), dtype=torch.uint8, device=dev)
        self.obs, self.reward, self.term, self.trunc = self.slab_views(self._out_slab, n, self.obs_dim)
        self.info_bits = torch.zeros((n,), dtype=torch.uint8, device=dev)
        self.final_obs = torch.zeros((n, self.obs_dim), **f32)
        self._drone_state = torch.zeros((n, 12), **f32)
        self._aux_state = torch.zeros((n, self.aux_dim), **f32)
        self._contact = torch.zeros((n,), dtype=torch.uint8, device=dev)
        b = _lib.PfbBuffers()
        b.state, b.istate = self.state_tensor.data_ptr(), self.istate_tensor.data_ptr()
        b.setpoint, b.start_pos, b.start_orn = self.setpoints.data_ptr(), self.start_pos.data_ptr(), self.start_orn.data_ptr()
        b.obs, b.reward, b.term, b.trunc = self.obs.data_ptr(), self.reward.data_ptr(), self.term.data_ptr(), self.trunc.data_ptr()
        b.info, b.final_obs = self.info_bits.data_ptr(), self.final_obs.data_ptr()
        b.drone_state, b.aux_state, b.contact = self._drone_state.data_ptr(), self._aux_state.data_ptr(), self._contact.data_ptr()
        b.reset_targets = None
        self._reset_targets = None
        self._buffers = b
        _lib.check(L.pfb_bind(self._h, C.byref(b)))
        self._state_fresh = False
        self.reset()

    # ------------------------------------------------------------------ lifecycle
    def __del__(self):
        h = getattr(self, "_h", None)
        if h:
            try:
                _lib.lib().pfb_destroy(h)
            except Exception:
                pass
            self._h = None

    def disconnect(self) -> None:
        self.__del__()

    def _s(self):
        return C.c_void_p(_stream_ptr(self.device))

    # ------------------------------------------------------------------ Aviary surface
    def reset(self) -> None:
        """aviary.py:218-312: every drone back to its start pose, mode 0, zero setpoint."""
        _lib.check(_lib.lib().pfb_reset(self._h, None, self._s()))
        self.physics_steps = 0
        self.aviary_steps = 0
        self.elapsed_time = 0.0
        self._state_fresh = False

    def set_mode(self, flight_modes: int | Sequence[int]) -> None:
        """aviary.py:440-458; one mode per batch (the kernel is specialised on it at compile time)."""
        if isinstance(flight_modes, (list, tuple)):
            if len(flight_modes) != self.num_drones:
                raise AssertionError(f"Expected {self.num_drones} flight_modes, got {len(flight_modes)}.")
            if len(set(flight_modes)) != 1:
                raise ValueError("the batched stepper needs one flight mode per batch")
            flight_modes = flight_modes[0]
        mode = int(flight_modes)
        lo, hi = (-1, 7) if self.drone_type == "quadx" else ((-1, 0) if self.drone_type == "fixedwing" else (0, 0))
        if mode < lo or mode > hi:
            # quadx.py:259-262, fixedwing.py:216-219, base_drone.py:252-255
            raise ValueError(f"`mode` must be between {lo} and {hi} or be registered in self.registered_controllers.keys()=dict_keys([]), got {mode}.")
        _lib.check(_lib.lib().pfb_set_mode(self._h, mode, self._s()))
        self._state_fresh = False

    def set_setpoint(self, index: int, setpoint) -> None:
        self.setpoints[index] = torch.as_tensor(setpoint, dtype=torch.float32, device=self.device)

    def set_all_setpoints(self, setpoints) -> None:
        self.setpoints.copy_(torch.as_tensor(setpoints, dtype=torch.float32, device=self.device))

    def step(self, n_steps: int = 1, noise: torch.Tensor | None = None) -> None:
        """``n_steps`` x Aviary.step() (aviary.py:480-531).  ``noise``: optional device tensor
        [n_steps*updates_per_step, N] of raw ``np_random.normal`` draws (parity tests)."""
        ptr = None
        if noise is not None:
            assert noise.dtype == torch.float32 and noise.is_cuda and noise.is_contiguous()
            assert tuple(noise.shape) == (n_steps * self.updates_per_step, self.num_drones), tuple(noise.shape)
            ptr = C.c_void_p(noise.data_ptr())
        _lib.check(_lib.lib().pfb_aviary_step(self._h, int(n_steps), ptr, self._s()))
        self.physics_steps += n_steps * self.updates_per_step
        self.aviary_steps += n_steps
        self.elapsed_time = self.physics_steps / self.physics_hz
        self._state_fresh = False

    def set_base_velocity(self, lin_vel: torch.Tensor, ang_vel: torch.Tensor) -> None:
        """``p.resetBaseVelocity`` for every drone (used by rocket_base_env.py:228): [N, 3] world-frame tensors."""
        lin = torch.as_tensor(lin_vel, dtype=torch.float32, device=self.device).reshape(self.num_drones, 3).contiguous()
        ang = torch.as_tensor(ang_vel, dtype=torch.float32, device=self.device).reshape(self.num_drones, 3).contiguous()
        _lib.check(_lib.lib().pfb_set_base_velocity(self._h, C.c_void_p(lin.data_ptr()), C.c_void_p(ang.data_ptr()), self._s()))
        self._state_fresh = False

    def _refresh(self):
        if not self._state_fresh:
            _lib.check(_lib.lib().pfb_observe_state(self._h, self._s()))
            self._state_fresh = True

    @property
    def all_states(self) -> torch.Tensor:
        """(N, 4, 3): ang_vel (body), euler, lin_vel (body), position — aviary.py:372-393."""
        self._refresh()
        return self._drone_state.view(self.num_drones, 4, 3)

    @property
    def all_aux_states(self) -> torch.Tensor:
        self._refresh()
        return self._aux_state

    def state(self, index: int) -> torch.Tensor:
        return self.all_states[index]

    def aux_state(self, index: int) -> torch.Tensor:
        return self.all_aux_states[index]

    @property
    def contact_array(self) -> torch.Tensor:
        """(N,) bool: ground contact during the last step (aviary.py:322, 523-525, per world)."""
        self._refresh()
        return self._contact.bool()

    # ------------------------------------------------------------------ raw state access (tests, debugging)
    def state_row(self, row: int) -> torch.Tensor:
        """[N] fp32 copy-free view (field-major) or gathered copy (warp-tiled) of state row ``row``."""
        if not self.tiled:
            return self.state_tensor[row]
        return self.state_tensor[:, row // 4, :, row % 4].reshape(-1)[: self.num_drones]

    def state_row_int(self, row: int) -> torch.Tensor:
        """[N] int32: a row that holds integer bits (warp-tiled layout: 17 = step_count, 18 = flags)."""
        return self.state_row(row).contiguous().view(torch.int32)

    @property
    def precise_positions(self) -> torch.Tensor:
        """(N, 3) float64 world positions as the kernels carry them: hi + lo fp32 words of the state tensor."""
        lo = {"quadx": 25, "fixedwing": 19, "rocket": 22}[self.drone_type]
        return torch.stack([self.state_row(k).double() + self.state_row(lo + k).double() for k in range(3)], dim=1)

    @property
    def step_counts(self) -> torch.Tensor:
        """[N] int32 env step counters."""
        return self.state_row_int(17) if self.tiled else self.istate_tensor[0]

    def register_wind_field(self, wind) -> None:
        """``Aviary.register_wind_field_function`` (aviary.py:324-334) for an analytic field: ``wind`` is a
        :class:`pyflyt_b200.core.wind.AnalyticWind` (the same object is a valid wind-field function for the reference) or ``None``
        for still air.  Arbitrary Python callbacks cannot run inside the step kernel (out of scope)."""
        from .wind import AnalyticWind, PfbWind

        if wind is None:
            _lib.check(_lib.lib().pfb_set_wind(self._h, None))
            self.wind_field = None
            return
        if not isinstance(wind, AnalyticWind):
            raise TypeError("the batched stepper evaluates the wind inside the CUDA kernels: pass a pyflyt_b200.core.wind.AnalyticWind")
        L = _lib.lib()
        assert L.pfb_sizeof_wind() == C.sizeof(PfbWind)
        w = wind.as_struct()
        _lib.check(L.pfb_set_wind(self._h, C.byref(w)))
        self.wind_field = wind

    register_wind_field_function = register_wind_field

    def reseed(self, seed: int) -> None:
        """``env.reset(seed=s)``: re-key the random streams and rewind every call counter, so that the same seed replays the same
        episodes (the reference re-creates ``np_random``, aviary.py:108-117)."""
        self.seed = int(seed)
        _lib.check(_lib.lib().pfb_reseed(self._h, self.seed, self._s()))

    def set_noise_dump(self, buf: torch.Tensor | None) -> None:
        """Test aid: ``buf`` [env_step_ratio * updates_per_step, N] fp32 receives every motor-noise draw of the following
        QuadX-Hover ``env_step`` calls (None = off)."""
        self._noise_dump = buf
        _lib.check(_lib.lib().pfb_set_noise_dump(self._h, None if buf is None else C.c_void_p(buf.data_ptr())))

    @property
    def launch_count(self) -> int:
        return int(_lib.lib().pfb_launch_count(self._h))

    # ------------------------------------------------------------------ fused env surface
    def env_reset(self, mask: torch.Tensor | None = None, noise: torch.Tensor | None = None, targets: torch.Tensor | None = None) -> torch.Tensor:
        """env.reset() for all / masked envs.  ``targets`` [N, 3*num_targets] installs explicit waypoints
        (parity tests); by default they are drawn on device like ``WaypointHandler.reset``."""
        if targets is not None:
            self._reset_targets = torch.as_tensor(targets, dtype=torch.float32, device=self.device).reshape(self.num_drones, -1).contiguous()
            self._buffers.reset_targets = self._reset_targets.data_ptr()
            _lib.check(_lib.lib().pfb_bind(self._h, C.byref(self._buffers)))
        elif self._reset_targets is not None:
            self._reset_targets = None
            self._buffers.reset_targets = None
            _lib.check(_lib.lib().pfb_bind(self._h, C.byref(self._buffers)))
        m = None if mask is None else C.c_void_p(mask.to(torch.uint8).contiguous().data_ptr())
        nz = None if noise is None else C.c_void_p(noise.data_ptr())
        _lib.check(_lib.lib().pfb_env_reset(self._h, m, nz, self._s()))
        self._state_fresh = False
        return self.obs

    def env_step(self, actions: torch.Tensor | None = None, noise: torch.Tensor | None = None) -> None:
        """One fused env.step() for every env; ``actions`` [N, S] fp32 on this device (None = ``self.setpoints``)."""
        act = None
        if actions is not None:
            assert actions.is_cuda and actions.dtype == torch.float32 and actions.is_contiguous()
            assert tuple(actions.shape) == (self.num_drones, self.setpoint_dim), tuple(actions.shape)
            act = C.c_void_p(actions.data_ptr())
        nz = None if noise is None else C.c_void_p(noise.data_ptr())
        _lib.check(_lib.lib().pfb_env_step(self._h, act, nz, self._s()))
        self._state_fresh = False

    @staticmethod
    def out_slab_bytes(n: int, obs_dim: int) -> int:
        return n * obs_dim * 4 + n * 4 + n + n

    @staticmethod
    def slab_views(slab: torch.Tensor, n: int, obs_dim: int):
        """(obs [n, O] f32, reward [n] f32, term [n] u8, trunc [n] u8) views of a contiguous uint8 slab (device or pinned host)."""
        o = n * obs_dim * 4
        obs = slab[:o].view(torch.float32).view(n, obs_dim)
        reward = slab[o : o + 4 * n].view(torch.float32)
        term = slab[o + 4 * n : o + 5 * n]
        trunc = slab[o + 5 * n : o + 6 * n]
        return obs, reward, term, trunc

    def dogfight_physics(self, payload: torch.Tensor, actions: torch.Tensor | None = None, noise: torch.Tensor | None = None,
                         first: bool = False, do_reset: bool = False, aviary_index: int = 0) -> None:
        """Split dogfight, half 1: integrate one Aviary step (or reset + warm-up) and publish ``payload`` [N, 20]."""
        act = None if actions is None else C.c_void_p(actions.data_ptr())
        nz = None if noise is None else C.c_void_p(noise.data_ptr())
        _lib.check(_lib.lib().pfb_dogfight_physics(self._h, act, nz, C.c_void_p(payload.data_ptr()), int(first), int(do_reset), int(aviary_index), self._s()))
        self._state_fresh = False

    def dogfight_physics_peer(self, peer_tables: torch.Tensor, world: int, slot_offset_floats: int, actions: torch.Tensor | None = None,
                              noise: torch.Tensor | None = None, first: bool = False, do_reset: bool = False, aviary_index: int = 0,
                              peer_flags: torch.Tensor | None = None, rank: int = 0, epoch: int = 0) -> None:
        """Split dogfight, half 1 with the exchange fused in: every payload is stored straight into all ranks' tables
        (``peer_tables``: int64 device tensor of ``world`` peer-mapped base pointers)."""
        act = None if actions is None else C.c_void_p(actions.data_ptr())
        nz = None if noise is None else C.c_void_p(noise.data_ptr())
        fl = None if peer_flags is None else C.c_void_p(peer_flags.data_ptr())
        _lib.check(_lib.lib().pfb_dogfight_physics_peer(self._h, act, nz, C.c_void_p(peer_tables.data_ptr()), int(world), int(slot_offset_floats),
                                                        fl, int(rank), int(epoch), int(first), int(do_reset), int(aviary_index), self._s()))
        self._state_fresh = False

    def dogfight_combat_wait(self, table: torch.Tensor, first_global_agent: int, num_arenas: int, last: int, flags: torch.Tensor,
                             world: int, epoch: int) -> None:
        """Split dogfight, half 2, waiting in-kernel until every rank's physics kernel has raised its flag to ``epoch``."""
        _lib.check(_lib.lib().pfb_dogfight_combat_wait(self._h, C.c_void_p(table.data_ptr()), int(first_global_agent), int(num_arenas), int(last),
                                                       C.c_void_p(flags.data_ptr()), int(world), int(epoch), self._s()))
        self._state_fresh = False

    def dogfight_split_step(self, actions: torch.Tensor, peer_tables: torch.Tensor, peer_flags: torch.Tensor, tables: torch.Tensor,
                            flags: torch.Tensor, world: int, rank: int, epoch0: int, first_global_agent: int, num_arenas: int) -> None:
        """A whole env step of the split dogfight (fused exchange, in-kernel signalling) in one library call."""
        _lib.check(_lib.lib().pfb_dogfight_split_step(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(peer_tables.data_ptr()),
                                                      C.c_void_p(peer_flags.data_ptr()), C.c_void_p(tables.data_ptr()), C.c_void_p(flags.data_ptr()),
                                                      int(world), int(rank), int(epoch0), int(first_global_agent), int(num_arenas), self._s()))
        self._state_fresh = False

    def dogfight_combat(self, table: torch.Tensor, first_global_agent: int, num_arenas: int, last: int) -> None:
        """Split dogfight, half 2: combat state from the all-gathered payload ``table`` [2 * num_arenas, 20]."""
        _lib.check(_lib.lib().pfb_dogfight_combat(self._h, C.c_void_p(table.data_ptr()), int(first_global_agent), int(num_arenas), int(last), self._s()))
        self._state_fresh = False

    def profile_begin(self, capacity: int) -> None:
        _lib.check(_lib.lib().pfb_profile_begin(self._h, int(capacity)))

    def profile_read(self, capacity: int) -> list[float]:
        buf = (C.c_float * capacity)()
        n = _lib.lib().pfb_profile_read(self._h, buf, capacity)
        if n < 0:
            _lib.check(n)
        return [float(buf[i]) for i in range(n)]

    def env_rollout(self, n_steps: int) -> None:
        _lib.check(_lib.lib().pfb_env_rollout(self._h, int(n_steps), self._s()))
        self._state_fresh = False

    def env_step_mapped(self, actions: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, term: torch.Tensor, trunc: torch.Tensor) -> None:
        """Zero-copy end-to-end step: the kernel reads ``actions`` from and writes the results into PINNED host tensors."""
        for t in (actions, obs, reward, term, trunc):
            assert not t.is_cuda and t.is_contiguous() and t.is_pinned()
        _lib.check(_lib.lib().pfb_env_step_mapped(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(obs.data_ptr()), C.c_void_p(reward.data_ptr()), C.c_void_p(term.data_ptr()), C.c_void_p(trunc.data_ptr()), self._s()))
        self._state_fresh = False

    def env_step_host(self, actions: torch.Tensor, obs: torch.Tensor, reward: torch.Tensor, term: torch.Tensor, trunc: torch.Tensor) -> None:
        """Pinned-host in, pinned-host out (the end-to-end path bench.py times)."""
        for t in (actions, obs, reward, term, trunc):
            assert not t.is_cuda and t.is_contiguous()
        _lib.check(_lib.lib().pfb_env_step_host(self._h, C.c_void_p(actions.data_ptr()), C.c_void_p(obs.data_ptr()), C.c_void_p(reward.data_ptr()), C.c_void_p(term.data_ptr()), C.c_void_p(trunc.data_ptr()), self._s()))
        self._state_fresh = False
