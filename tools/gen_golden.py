#!/usr/bin/env python
"""Generates tests/golden/*.npz by flying the UNMODIFIED reference (/root/reference/PyFlyt) on the
restated Bullet engine (oracle/fakebullet).  Run in the build container only; the fixtures travel.

Each fixture stores the scenario inputs (setpoints / actions, start pose), the raw
``np_random.normal`` draws the reference consumed (one per component per physics step)
and the reference's outputs per Aviary step / env step.  The C oracle (oracle/pfb_oracle.c) and the
CUDA path are both replayed against these with the same injected draws.

Scenarios mirror the reference's own tests: tests/test_core.py:13-31 (mode-7 hold),
:65-93 (two set-points), tests/test_gym_envs.py:92-112 (env determinism contract).
"""
import json
import os
import sys

import numpy as np

ROOT = os.path.join(os.path.dirname(os.path.realpath(__file__)), "..")
sys.path.insert(0, ROOT)
from oracle import ref_in_loop as ril  # noqa: E402

ril.install()
from PyFlyt.core import Aviary  # noqa: E402

OUT = os.path.join(ROOT, "tests", "golden")


def fly_quadx(name, mode, drone_model, start_pos, start_orn, setpoint_schedule, n_steps, seed, wind=None):
    """Aviary-level QuadX flight; setpoint_schedule: {step_index: setpoint(4)} applied before step."""
    rng = ril.ScriptedNoise(seed)
    env = Aviary(
        start_pos=np.array([start_pos], dtype=np.float64),
        start_orn=np.array([start_orn], dtype=np.float64),
        drone_type="quadx",
        drone_options=dict(drone_model=drone_model),
        np_random=rng,
    )
    env.set_mode(mode)
    if wind is not None:
        env.register_wind_field_function(wind)  # the UNMODIFIED reference evaluates our AnalyticWind as a plain wind-field function
        env.drones[0].update_state()  # the cached body velocity now sees the wind (update_state runs after every step anyway)
    sp_after_mode = np.array(env.drones[0].setpoint, dtype=np.float64)
    states, auxs, pwms, contacts, raws, sps = [], [], [], [], [], []
    for i in range(n_steps):
        if i in setpoint_schedule:
            env.set_setpoint(0, np.array(setpoint_schedule[i], dtype=np.float64))
        sps.append(np.array(env.drones[0].setpoint, dtype=np.float64))
        env.step()
        d = env.drones[0]
        states.append(np.array(d.state))
        auxs.append(np.array(d.aux_state))
        pwms.append(np.array(d.pwm))
        contacts.append(bool(np.any(env.contact_array[env.planeId])))
        pos, quat = env.getBasePositionAndOrientation(d.Id)
        v, w = env.getBaseVelocity(d.Id)
        raws.append(np.concatenate([pos, quat, v, w]))
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        kind="quadx_aviary",
        mode=mode,
        drone_model=drone_model,
        start_pos=np.array(start_pos, dtype=np.float64),
        start_orn=np.array(start_orn, dtype=np.float64),
        setpoint_after_set_mode=sp_after_mode,
        setpoints=np.array(sps),
        noise=np.array(rng.normal_log),
        state=np.array(states),
        aux=np.array(auxs),
        pwm=np.array(pwms),
        contact=np.array(contacts),
        raw=np.array(raws),
        **wind_fields(wind),
    )
    print(name, "final pos", states[-1][3], "draws", len(rng.normal_log))


def wind_fields(wind):
    """npz entries describing an AnalyticWind (absent = still air)"""
    if wind is None:
        return {}
    return dict(wind_kind=wind.kind, wind_base=wind.base, wind_z_ref=wind.z_ref, wind_alpha=wind.alpha, wind_z0=wind.z0)


def fly_vehicle(name, drone_type, drone_model, mode, start_pos, start_orn, setpoint_schedule, n_steps, seed, drone_options=None, pre_hook=None, wind=None):
    """Aviary-level flight of a fixedwing / rocket; setpoint_schedule: {step: setpoint} applied before the step."""
    rng = ril.ScriptedNoise(seed)
    opts = dict(drone_model=drone_model, **(drone_options or {}))
    env = Aviary(
        start_pos=np.array([start_pos], dtype=np.float64),
        start_orn=np.array([start_orn], dtype=np.float64),
        drone_type=drone_type,
        drone_options=opts,
        np_random=rng,
    )
    env.set_mode(mode)
    if wind is not None:
        env.register_wind_field_function(wind)
        env.drones[0].update_state()  # the cached surface / body velocities now see the wind (update_state runs after every step)
    if pre_hook is not None:
        pre_hook(env)
        env.drones[0].update_state()  # resetBaseVelocity alone leaves the drone's cached velocities stale
    d = env.drones[0]
    v0, w0 = env.getBaseVelocity(d.Id)
    sp_dim = len(np.atleast_1d(d.setpoint))
    states, auxs, contacts, raws, sps = [], [], [], [], []
    for i in range(n_steps):
        if i in setpoint_schedule:
            env.set_setpoint(0, np.array(setpoint_schedule[i], dtype=np.float64))
        sps.append(np.array(d.setpoint, dtype=np.float64))
        env.step()
        states.append(np.array(d.state))
        auxs.append(np.array(d.aux_state, dtype=np.float64))
        contacts.append(bool(np.any(env.contact_array[env.planeId])))
        pos, quat = env.getBasePositionAndOrientation(d.Id)
        v, w = env.getBaseVelocity(d.Id)
        raws.append(np.concatenate([pos, quat, v, w]))
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        kind=f"{drone_type}_aviary",
        mode=mode,
        drone_type=drone_type,
        drone_model=drone_model,
        drone_options=json.dumps({k: (list(v) if hasattr(v, "__len__") else v) for k, v in (drone_options or {}).items()}),
        start_pos=np.array(start_pos, dtype=np.float64),
        start_orn=np.array(start_orn, dtype=np.float64),
        setpoint_dim=sp_dim,
        start_lin_vel=np.array(v0, dtype=np.float64),
        start_ang_vel=np.array(w0, dtype=np.float64),
        has_pre_hook=pre_hook is not None,
        setpoints=np.array(sps),
        noise=np.array(rng.normal_log),
        state=np.array(states),
        aux=np.array(auxs),
        contact=np.array(contacts),
        raw=np.array(raws),
        **wind_fields(wind),
    )
    print(name, "final pos", states[-1][3], "draws", len(rng.normal_log))


def fly_waypoints(name, seed, n_steps, action_seed, angle_representation="quaternion", sparse=False, num_targets=4,
                  goal_reach_distance=2.0, dome=100.0, action_scale=1.0):
    """FixedwingWaypointsEnv (fixedwing_waypoints_env.py) with scripted actions and user-loop resets."""
    from PyFlyt.gym_envs.fixedwing_envs.fixedwing_waypoints_env import FixedwingWaypointsEnv

    env = FixedwingWaypointsEnv(sparse_reward=sparse, num_targets=num_targets, goal_reach_distance=goal_reach_distance,
                                flight_dome_size=dome, angle_representation=angle_representation)
    rng = ril.ScriptedNoise(seed)
    env._np_random = rng

    def flat(state):
        d = np.asarray(state["target_deltas"], dtype=np.float64).reshape(-1)
        pad = np.zeros(3 * num_targets)
        pad[: len(d)] = d
        return np.concatenate([state["attitude"], pad])

    state0, _ = env.reset()
    targets = [np.array(env.waypoints.targets, dtype=np.float64)]
    arng = np.random.default_rng(action_seed)
    obs, rew, term, trunc, info, acts, episode_start, resets_obs = [], [], [], [], [], [], [], []
    noise_splits = [len(rng.normal_log)]
    for i in range(n_steps):
        a = arng.uniform(-1.0, 1.0, 4) * action_scale
        a[3] = arng.uniform(0.2, 1.0)
        o, r, te, tr, inf = env.step(a)
        acts.append(a); obs.append(flat(o)); rew.append(r); term.append(te); trunc.append(tr)
        info.append(int(inf["out_of_bounds"]) | (int(inf["collision"]) << 1) | (int(inf["env_complete"]) << 2) | (int(inf["num_targets_reached"]) << 3))
        noise_splits.append(len(rng.normal_log))
        if te or tr:
            o2, _ = env.reset()
            targets.append(np.array(env.waypoints.targets, dtype=np.float64))
            resets_obs.append(flat(o2))
            episode_start.append(i + 1)
            noise_splits.append(len(rng.normal_log))
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), kind="fixedwing_waypoints", sparse=sparse, dome=dome, num_targets=num_targets,
        goal_reach_distance=goal_reach_distance, angle_representation=angle_representation, reset_obs=flat(state0),
        targets=np.array(targets), actions=np.array(acts), obs=np.array(obs), reward=np.array(rew), term=np.array(term),
        trunc=np.array(trunc), info=np.array(info), noise=np.array(rng.normal_log), noise_splits=np.array(noise_splits),
        episode_start=np.array(episode_start, dtype=np.int64),
        after_reset_obs=np.array(resets_obs) if resets_obs else np.zeros((0, 23 + 3 * num_targets)),
    )
    print(name, "steps", n_steps, "episodes", len(episode_start) + 1, "max targets reached", max(v >> 3 for v in info), "draws", len(rng.normal_log))


def fly_qx_waypoints(name, seed, n_steps, action_seed, flight_mode=0, angle_representation="quaternion", sparse=False, num_targets=4,
                     use_yaw_targets=False, goal_reach_distance=0.2, goal_reach_angle=0.1, dome=5.0, chase=False, action_scale=1.0):
    """QuadXWaypointsEnv (quadx_waypoints_env.py) with scripted actions and user-loop resets.  ``chase``: in a position
    flight mode (7: x, y, yaw, z) the action is the next target (+ jitter), so that waypoints ARE reached."""
    from PyFlyt.gym_envs.quadx_envs.quadx_waypoints_env import QuadXWaypointsEnv

    env = QuadXWaypointsEnv(sparse_reward=sparse, num_targets=num_targets, use_yaw_targets=use_yaw_targets,
                            goal_reach_distance=goal_reach_distance, goal_reach_angle=goal_reach_angle, flight_mode=flight_mode,
                            flight_dome_size=dome, angle_representation=angle_representation)
    rng = ril.ScriptedNoise(seed)
    env._np_random = rng
    T = 4 if use_yaw_targets else 3

    def flat(state):
        d = np.asarray(state["target_deltas"], dtype=np.float64).reshape(-1)
        pad = np.zeros(T * num_targets)
        pad[: len(d)] = d
        return np.concatenate([state["attitude"], pad])

    def cur_targets():
        t = np.array(env.waypoints.targets, dtype=np.float64).reshape(-1, 3)
        if use_yaw_targets:
            t = np.concatenate([t, np.array(env.waypoints.yaw_targets, dtype=np.float64)[:, None]], axis=-1)
        return t

    state0, _ = env.reset()
    targets = [cur_targets()]
    arng = np.random.default_rng(action_seed)
    obs, rew, term, trunc, info, acts, episode_start, resets_obs = [], [], [], [], [], [], [], []
    noise_splits = [len(rng.normal_log)]
    for i in range(n_steps):
        if chase and len(env.waypoints.targets) > 0:
            t = np.asarray(env.waypoints.targets[0], dtype=np.float64)
            yaw = float(env.waypoints.yaw_targets[0]) if use_yaw_targets else 0.0
            a = np.array([t[0], t[1], yaw, t[2]]) + arng.normal(0.0, 0.02, 4)
        else:
            a = arng.uniform([-np.pi, -np.pi, -np.pi, 0.0], [np.pi, np.pi, np.pi, 0.8]) * np.array([action_scale] * 3 + [1.0])
        o, r, te, tr, inf = env.step(a)
        acts.append(a); obs.append(flat(o)); rew.append(r); term.append(te); trunc.append(tr)
        info.append(int(inf["out_of_bounds"]) | (int(inf["collision"]) << 1) | (int(inf["env_complete"]) << 2) | (int(inf["num_targets_reached"]) << 3))
        noise_splits.append(len(rng.normal_log))
        if te or tr:
            o2, _ = env.reset()
            targets.append(cur_targets())
            resets_obs.append(flat(o2))
            episode_start.append(i + 1)
            noise_splits.append(len(rng.normal_log))
    att = 21 if angle_representation == "quaternion" else 20
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), kind="quadx_waypoints", sparse=sparse, dome=dome, num_targets=num_targets,
        use_yaw_targets=use_yaw_targets, goal_reach_distance=goal_reach_distance, goal_reach_angle=goal_reach_angle, flight_mode=flight_mode,
        angle_representation=angle_representation, reset_obs=flat(state0), targets=np.array(targets), actions=np.array(acts),
        obs=np.array(obs), reward=np.array(rew), term=np.array(term), trunc=np.array(trunc), info=np.array(info),
        noise=np.array(rng.normal_log), noise_splits=np.array(noise_splits), episode_start=np.array(episode_start, dtype=np.int64),
        after_reset_obs=np.array(resets_obs) if resets_obs else np.zeros((0, att + T * num_targets)),
    )
    print(name, "steps", n_steps, "episodes", len(episode_start) + 1, "max targets reached", max(v >> 3 for v in info), "draws", len(rng.normal_log))


def fly_landing(name, seed, n_steps, action_seed, options, angle_representation="quaternion", sparse=False, ignite_p=0.7):
    """RocketLandingEnv (rocket_landing_env.py) with scripted actions and user-loop resets; ``options`` as in
    env.reset(options=...): None = randomised + accelerated drop, {} = the plain 450 m hover-drop."""
    from PyFlyt.gym_envs.rocket_envs.rocket_landing_env import RocketLandingEnv

    env = RocketLandingEnv(sparse_reward=sparse, angle_representation=angle_representation)
    rng = ril.ScriptedNoise(seed)
    env._np_random = rng
    obs0, _ = env.reset(options=None if options is None else dict(options))
    spawns = [np.concatenate([env.start_pos[0], env.start_orn[0]])]
    arng = np.random.default_rng(action_seed)
    lo, hi = env.action_space.low, env.action_space.high
    obs, rew, term, trunc, info, acts, episode_start, resets_obs = [], [], [], [], [], [], [], []
    noise_splits = [len(rng.normal_log)]
    for i in range(n_steps):
        a = arng.uniform(lo, hi)
        a[3] = 1.0 if arng.random() < ignite_p else 0.0
        o, r, te, tr, inf = env.step(a)
        acts.append(a); obs.append(np.array(o)); rew.append(r); term.append(te); trunc.append(tr)
        info.append(int(inf["out_of_bounds"]) | (int(inf["fatal_collision"]) << 1) | (int(inf["env_complete"]) << 2))
        noise_splits.append(len(rng.normal_log))
        if te or tr:
            o2, _ = env.reset(options=None if options is None else dict(options))
            spawns.append(np.concatenate([env.start_pos[0], env.start_orn[0]]))
            resets_obs.append(np.array(o2))
            episode_start.append(i + 1)
            noise_splits.append(len(rng.normal_log))
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), kind="rocket_landing", sparse=sparse, angle_representation=angle_representation,
        randomize_drop=options is None, accelerate_drop=options is None, spawns=np.array(spawns), reset_obs=np.array(obs0),
        actions=np.array(acts), obs=np.array(obs), reward=np.array(rew), term=np.array(term), trunc=np.array(trunc), info=np.array(info),
        noise=np.array(rng.normal_log), noise_splits=np.array(noise_splits), episode_start=np.array(episode_start, dtype=np.int64),
        after_reset_obs=np.array(resets_obs) if resets_obs else np.zeros((0, len(obs0))),
    )
    print(name, "steps", n_steps, "episodes", len(episode_start) + 1, "infos", sorted(set(info)), "draws", len(rng.normal_log))


def fly_touchdown(name, seed, n_steps, descent_rate, ceiling=4.0, max_displacement=20.0, angle_representation="quaternion", lateral=(0.0, 0.0)):
    """RocketLandingEnv brought down onto the pad by a scripted bang-bang ignition law, with the engine's contact RESPONSE
    switched on (oracle/fakebullet World.contact_response): the rocket touches down at `descent_rate`-ish m/s, rests on its
    legs and the UNMODIFIED env reports env_complete (rocket_landing_env.py:231-263).  Same file format as fly_landing."""
    import pybullet as fb  # oracle/fakebullet
    from PyFlyt.gym_envs.rocket_envs.rocket_landing_env import RocketLandingEnv

    fb.World.contact_response = True
    try:
        env = RocketLandingEnv(ceiling=ceiling, max_displacement=max_displacement, angle_representation=angle_representation)
        rng = ril.ScriptedNoise(seed)
        env._np_random = rng
        obs0, _ = env.reset(options=dict(randomize_drop=False, accelerate_drop=False))
        if lateral != (0.0, 0.0):  # a small sideways push: friction has to stop the slide
            env.env.resetBaseVelocity(env.env.drones[0].Id, [lateral[0], lateral[1], 0.0], [0.0, 0.0, 0.0])
        spawns = [np.concatenate([env.start_pos[0], env.start_orn[0]])]
        att = 4 if angle_representation == "quaternion" else 3
        obs_k, acts, obs, rew, term, trunc, info = np.array(obs0), [], [], [], [], [], []
        noise_splits = [len(rng.normal_log)]
        for i in range(n_steps):
            vz, z = obs_k[3 + att + 2], obs_k[3 + att + 3 + 2]
            h = z - 2.425 - 0.15  # leg soles above the pad
            ign = 1.0 if (vz < -(descent_rate + 1.0 * max(h, 0.0)) and h > 0.02) else 0.0
            a = np.array([0.0, 0.0, 0.0, ign, 0.0, 0.0, 0.0])
            o, r, te, tr, inf = env.step(a)
            obs_k = np.array(o)
            acts.append(a); obs.append(obs_k); rew.append(r); term.append(te); trunc.append(tr)
            info.append(int(inf["out_of_bounds"]) | (int(inf["fatal_collision"]) << 1) | (int(inf["env_complete"]) << 2))
            noise_splits.append(len(rng.normal_log))
            if te or tr:
                break
    finally:
        fb.World.contact_response = False
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"), kind="rocket_landing", sparse=False, angle_representation=angle_representation,
        randomize_drop=False, accelerate_drop=False, spawns=np.array(spawns), reset_obs=np.array(obs0),
        actions=np.array(acts), obs=np.array(obs), reward=np.array(rew), term=np.array(term), trunc=np.array(trunc), info=np.array(info),
        noise=np.array(rng.normal_log), noise_splits=np.array(noise_splits), episode_start=np.zeros(0, dtype=np.int64),
        after_reset_obs=np.zeros((0, len(obs0))), ceiling=ceiling, max_displacement=max_displacement, contact_response=True,
        start_lin_vel=np.array([lateral[0], lateral[1], 0.0]),
    )
    print(name, "steps", len(acts), "last info", info[-1], "pad contact steps", int(sum(o[-1] for o in obs)), "draws", len(rng.normal_log))


def touchdown_fixtures():
    # gentle touchdowns that must end in env_complete, a hard one that must be a fatal collision
    fly_touchdown("landing_touchdown", seed=81, n_steps=300, descent_rate=0.6)
    fly_touchdown("landing_touchdown_soft_euler", seed=82, n_steps=300, descent_rate=0.3, angle_representation="euler")
    fly_touchdown("landing_touchdown_hard", seed=83, n_steps=300, descent_rate=2.5)


def fly_dogfight(name, seed, n_steps, action_seed, team_size=1, sparse=False, action_scale=0.6, lethal_distance=20.0, lethal_angle=0.07,
                 spawn_min_radius=10.0, spawn_max_radius=50.0, damage_per_hit=0.003, pitch_bias=0.0):
    """MAFixedwingDogfightEnv (pz_envs/fixedwing_envs/ma_fixedwing_dogfight_env.py) with scripted actions; a new
    episode is started whenever every agent is done.  The Aviary's own generator (seeded by reset(seed)) is
    wrapped so that its motor-noise draws are recorded."""
    from PyFlyt.pz_envs.fixedwing_envs.ma_fixedwing_dogfight_env import MAFixedwingDogfightEnv

    env = MAFixedwingDogfightEnv(team_size=team_size, sparse_reward=sparse, lethal_distance=lethal_distance, lethal_angle_radians=lethal_angle,
                                 spawn_min_radius=spawn_min_radius, spawn_max_radius=spawn_max_radius, damage_per_hit=damage_per_hit)
    A = 2 * team_size
    real_default_rng = np.random.default_rng
    loggers = []

    def reset(seed_):
        def patched(s=None):
            lg = ril.ScriptedNoise.__new__(ril.ScriptedNoise)
            lg._rng = real_default_rng(s)
            lg.normal_log = []
            loggers.append(lg)
            return lg
        np.random.default_rng = patched
        try:
            obs, _ = env.reset(seed=seed_)
        finally:
            np.random.default_rng = real_default_rng
        return np.stack([obs[f"uav_{i}"] for i in range(A)])

    def drained():
        lg = loggers[-1]
        out = np.array(lg.normal_log)
        lg.normal_log.clear()
        return out

    arng = real_default_rng(action_seed)
    episodes = []
    ep_seed = seed
    obs0 = reset(ep_seed)
    ep = dict(spawn=np.concatenate([env.start_pos, env.start_orn], axis=1), reset_obs=obs0, reset_noise=drained(), actions=[], obs=[], reward=[], term=[], trunc=[], noise=[])
    for i in range(n_steps):
        alive = set(env.agents)
        act = arng.uniform(-1.0, 1.0, (A, 4)) * action_scale
        act[:, 1] = np.clip(act[:, 1] + pitch_bias, -1.0, 1.0)
        o, r, te, tr, _ = env.step({f"uav_{k}": act[k] for k in range(A) if f"uav_{k}" in alive})
        ep["actions"].append(act)
        ep["noise"].append(drained())
        row = lambda d, default: np.array([d.get(f"uav_{k}", default) for k in range(A)])  # noqa: E731
        ep["obs"].append(np.stack([o.get(f"uav_{k}", np.full(obs0.shape[1], np.nan)) for k in range(A)]))
        ep["reward"].append(row(r, np.nan)); ep["term"].append(row(te, True)); ep["trunc"].append(row(tr, False))
        ep["alive"] = ep.get("alive", []) + [np.array([f"uav_{k}" in alive for k in range(A)])]
        if len(env.agents) == 0:
            episodes.append(ep)
            ep_seed += 1
            obs0 = reset(ep_seed)
            ep = dict(spawn=np.concatenate([env.start_pos, env.start_orn], axis=1), reset_obs=obs0, reset_noise=drained(), actions=[], obs=[], reward=[], term=[], trunc=[], noise=[])
    if ep["actions"]:
        episodes.append(ep)
    flat = {}
    for k, e in enumerate(episodes):
        for key, v in e.items():
            flat[f"ep{k}_{key}"] = np.array(v)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kind="dogfight", team_size=team_size, sparse=sparse, n_episodes=len(episodes),
                        lethal_distance=lethal_distance, lethal_angle=lethal_angle, damage_per_hit=damage_per_hit, **flat)
    print(name, "episodes", len(episodes), "steps", [len(e["actions"]) for e in episodes], "reward range", min(np.nanmin(e["reward"]) for e in episodes), max(np.nanmax(e["reward"]) for e in episodes))


def fly_ma_hover(name, seed, n_steps, action_seed, flight_mode=0, angle_representation="quaternion", sparse=False, dome=10.0,
                 max_duration_seconds=30.0, action_scale=0.3, start_pos=None):
    """MAQuadXHoverEnv (pz_envs/quadx_envs/ma_quadx_hover_env.py) with scripted actions; a new episode is started whenever
    every agent is done.  The Aviary's own generator (seeded by reset(seed)) is wrapped so that its draws are recorded."""
    from PyFlyt.pz_envs.quadx_envs.ma_quadx_hover_env import MAQuadXHoverEnv

    kw = dict(sparse_reward=sparse, flight_mode=flight_mode, flight_dome_size=dome, max_duration_seconds=max_duration_seconds,
              angle_representation=angle_representation)
    if start_pos is not None:
        kw.update(start_pos=np.asarray(start_pos, dtype=np.float64), start_orn=np.zeros_like(np.asarray(start_pos, dtype=np.float64)))
    env = MAQuadXHoverEnv(**kw)
    A = len(env.possible_agents)
    real_default_rng = np.random.default_rng
    loggers = []

    def reset(seed_):
        def patched(s=None):
            lg = ril.ScriptedNoise.__new__(ril.ScriptedNoise)
            lg._rng = real_default_rng(s)
            lg.normal_log = []
            loggers.append(lg)
            return lg
        np.random.default_rng = patched
        try:
            obs, _ = env.reset(seed=seed_)
        finally:
            np.random.default_rng = real_default_rng
        return np.stack([obs[f"uav_{i}"] for i in range(A)])

    def drained():
        lg = loggers[-1]
        out = np.array(lg.normal_log)
        lg.normal_log.clear()
        return out

    arng = real_default_rng(action_seed)
    lo, hi = np.array([-np.pi, -np.pi, -np.pi, 0.0]), np.array([np.pi, np.pi, np.pi, 0.8])
    episodes = []
    ep_seed = seed
    obs0 = reset(ep_seed)
    new_ep = lambda o: dict(reset_obs=o, reset_noise=drained(), actions=[], obs=[], reward=[], term=[], trunc=[], noise=[], alive=[])  # noqa: E731
    ep = new_ep(obs0)
    for i in range(n_steps):
        alive = set(env.agents)
        act = arng.uniform(lo, hi, (A, 4)) * np.array([action_scale] * 3 + [1.0])
        o, r, te, tr, _ = env.step({f"uav_{k}": act[k] for k in range(A) if f"uav_{k}" in alive})
        ep["actions"].append(act)
        ep["noise"].append(drained())
        row = lambda d, default: np.array([d.get(f"uav_{k}", default) for k in range(A)])  # noqa: E731
        ep["obs"].append(np.stack([o.get(f"uav_{k}", np.full(obs0.shape[1], np.nan)) for k in range(A)]))
        ep["reward"].append(row(r, np.nan)); ep["term"].append(row(te, True)); ep["trunc"].append(row(tr, False))
        ep["alive"].append(np.array([f"uav_{k}" in alive for k in range(A)]))
        if len(env.agents) == 0:
            episodes.append(ep)
            ep_seed += 1
            ep = new_ep(reset(ep_seed))
    if ep["actions"]:
        episodes.append(ep)
    flat = {}
    for k, e in enumerate(episodes):
        for key, v in e.items():
            flat[f"ep{k}_{key}"] = np.array(v)
    np.savez_compressed(os.path.join(OUT, f"{name}.npz"), kind="ma_quadx_hover", n_agents=A, sparse=sparse, n_episodes=len(episodes), flight_mode=flight_mode,
                        angle_representation=angle_representation, dome=dome, max_duration_seconds=max_duration_seconds,
                        start_pos=np.asarray(env.start_pos, dtype=np.float64), start_orn=np.asarray(env.start_orn, dtype=np.float64), **flat)
    print(name, "episodes", len(episodes), "steps", [len(e["actions"]) for e in episodes], "reward range", min(np.nanmin(e["reward"]) for e in episodes), max(np.nanmax(e["reward"]) for e in episodes))


def fly_hover(name, seed, n_steps, action_seed, angle_representation, flight_mode=0, sparse=False, dome=3.0, action_scale=1.0, wind=None):
    from PyFlyt.gym_envs.quadx_envs.quadx_hover_env import QuadXHoverEnv

    env = QuadXHoverEnv(
        sparse_reward=sparse, flight_mode=flight_mode, flight_dome_size=dome, angle_representation=angle_representation
    )
    rng = ril.ScriptedNoise(seed)
    env._np_random = rng  # the env hands its generator to the Aviary (quadx_base_env.py:192)
    obs0, _ = env.reset()
    if wind is not None:  # the reference envs have no wind argument: a user attaches the field to the env's Aviary after reset()
        env.env.register_wind_field_function(wind)
        env.env.drones[0].update_state()
    arng = np.random.default_rng(action_seed)
    lo, hi = env.action_space.low, env.action_space.high
    obs, rew, term, trunc, info, acts, episode_start = [], [], [], [], [], [], []
    noise_splits = [len(rng.normal_log)]
    resets_obs = []
    for i in range(n_steps):
        a = arng.uniform(lo, hi) * action_scale
        o, r, te, tr, inf = env.step(a)
        acts.append(a)
        obs.append(o)
        rew.append(r)
        term.append(te)
        trunc.append(tr)
        info.append(int(inf["out_of_bounds"]) | (int(inf["collision"]) << 1) | (int(inf["env_complete"]) << 2))
        noise_splits.append(len(rng.normal_log))
        if te or tr:
            assert wind is None, "wind fixtures are single-episode (a reset re-creates the Aviary without the field)"
            # next-episode reset exactly as a user loop would do it
            o2, _ = env.reset()
            resets_obs.append(o2)
            episode_start.append(i + 1)
            noise_splits.append(len(rng.normal_log))
    np.savez_compressed(
        os.path.join(OUT, f"{name}.npz"),
        kind="quadx_hover",
        flight_mode=flight_mode,
        sparse=sparse,
        dome=dome,
        angle_representation=angle_representation,
        reset_obs=obs0,
        actions=np.array(acts),
        obs=np.array(obs),
        reward=np.array(rew),
        term=np.array(term),
        trunc=np.array(trunc),
        info=np.array(info),
        noise=np.array(rng.normal_log),
        noise_splits=np.array(noise_splits),
        episode_start=np.array(episode_start, dtype=np.int64),
        after_reset_obs=np.array(resets_obs) if resets_obs else np.zeros((0, len(obs0))),
        **wind_fields(wind),
    )
    print(name, "steps", n_steps, "episodes", len(episode_start) + 1, "draws", len(rng.normal_log))


def fixedwing_fixtures():
    # Fixedwing (lifting surfaces + one motor), both airframes, mode 0 (RPYT mixing) and -1 (raw surfaces)
    r = np.random.default_rng(21)
    for model in ["fixedwing", "acrowing"]:
        sched0 = {}
        for k in range(0, 600, 40):
            sched0[k] = np.concatenate([r.uniform(-0.6, 0.6, 3), r.uniform(0.3, 1.0, 1)])
        fly_vehicle(f"fixedwing_{model}_mode0", "fixedwing", model, 0, [0, 0, 60.0], [0.05, -0.1, 0.4], sched0, 600, seed=31)
        schedm = {}
        for k in range(0, 400, 50):
            schedm[k] = np.concatenate([r.uniform(-0.8, 0.8, 5), r.uniform(0.0, 1.0, 1)])
        fly_vehicle(f"fixedwing_{model}_mode-1", "fixedwing", model, -1, [0, 0, 80.0], [0.0, 0.2, -1.0], schedm, 400, seed=32)
    # deep stall / tumbling: large attitude offsets and zero airspeed at spawn exercise the post-stall branches
    fly_vehicle(
        "fixedwing_stall", "fixedwing", "fixedwing", 0, [0, 0, 120.0], [1.2, 0.9, -2.0],
        {0: [0.9, -0.9, 0.5, 0.0], 150: [-0.9, 0.9, -0.5, 1.0]}, 500, seed=33, drone_options=dict(starting_velocity=np.array([0.0, 0.0, 0.0])),
    )
    # dive into the floor: contact flag from the link boxes
    fly_vehicle("fixedwing_floor", "fixedwing", "fixedwing", 0, [0, 0, 3.0], [0.0, 0.6, 0.0], {0: [0.0, 0.5, 0.0, 0.2]}, 120, seed=34)
    # Fixedwing-Waypoints env (BASELINE configs[2]); the wide goal radius makes scripted flights reach targets
    fly_waypoints("fwwp_quat_dense", seed=41, n_steps=300, action_seed=5, action_scale=0.3)
    fly_waypoints("fwwp_wide_goal", seed=42, n_steps=400, action_seed=6, goal_reach_distance=70.0, action_scale=0.15, num_targets=3)
    fly_waypoints("fwwp_euler_sparse", seed=43, n_steps=200, action_seed=7, angle_representation="euler", sparse=True, action_scale=0.5)


def rocket_fixtures():
    r = np.random.default_rng(51)

    def sched(n, every, ign_p=0.8):
        out = {}
        for k in range(0, n, every):
            out[k] = np.concatenate([r.uniform(-1, 1, 3), [1.0 if r.random() < ign_p else 0.0], r.uniform(0, 1, 1), r.uniform(-1, 1, 2)])
        return out

    # powered flight from rest: gimbal, fins, throttle changes, full tank (mass/inertia vary slowly)
    fly_vehicle("rocket_powered", "rocket", "rocket", 0, [0, 0, 100.0], [0.05, -0.04, 0.3], sched(600, 40), 600, seed=61)
    # Rocket-Landing style drop: 5 % fuel runs dry (hard cut-off), -100 m/s start hits Bullet's velocity clamp
    def drop(env):
        env.resetBaseVelocity(env.drones[0].Id, [3.0, -2.0, -100.0], [0.2, -0.1, 0.3])
    fly_vehicle("rocket_drop", "rocket", "rocket", 0, [5.0, -8.0, 420.0], [0.2, -0.15, 0.1], sched(500, 25, 0.6), 500, seed=62,
                drone_options=dict(starting_fuel_ratio=0.05), pre_hook=drop)
    # tumbling, engine off: finlets + body drag at large angles of attack
    def spin(env):
        env.resetBaseVelocity(env.drones[0].Id, [20.0, 10.0, -30.0], [1.5, -1.0, 0.5])
    s3 = sched(300, 30, 0.0)
    fly_vehicle("rocket_tumble", "rocket", "rocket", 0, [0, 0, 300.0], [1.0, 0.5, 0.0], s3, 300, seed=63, pre_hook=spin)
    # ground strike (legs / body primitives)
    fly_vehicle("rocket_ground", "rocket", "rocket", 0, [30.0, 0, 6.0], [0.3, 0.0, 0.0], {0: [0, 0, 0, 0, 0, 0, 0]}, 150, seed=64)
    # full tank: the composite mass / COM / inertia change every substep while the booster burns
    fly_vehicle("rocket_full_tank", "rocket", "rocket", 0, [0, 0, 50.0], [0.0, 0.05, 0.0], sched(480, 60, 1.0), 480, seed=65,
                drone_options=dict(starting_fuel_ratio=1.0))
    # Rocket-Landing env (BASELINE configs[3]): randomised accelerated drops (options=None) and the plain drop ({})
    fly_landing("landing_random_drop", seed=71, n_steps=500, action_seed=8, options=None)
    fly_landing("landing_plain_euler", seed=72, n_steps=400, action_seed=9, options={}, angle_representation="euler", ignite_p=0.2)
    fly_landing("landing_sparse", seed=73, n_steps=300, action_seed=10, options=None, sparse=True, ignite_p=0.0)
    # unpowered plain drop straight onto the pad: pad-contact reward, fatal touchdown speed, next episode
    fly_landing("landing_pad_strike", seed=74, n_steps=450, action_seed=11, options={}, ignite_p=0.0)


def quadx_waypoints_fixtures():
    # QuadX-Waypoints: rate mode with random actions (crashes, no targets), position mode chasing the
    # targets (reached targets, env_complete), yaw targets, euler + sparse
    fly_qx_waypoints("qxwp_mode0_random", seed=101, n_steps=300, action_seed=21, action_scale=0.3)
    fly_qx_waypoints("qxwp_mode7_chase", seed=102, n_steps=500, action_seed=22, flight_mode=7, chase=True, goal_reach_distance=0.3, num_targets=3)
    fly_qx_waypoints("qxwp_mode7_yaw_targets", seed=103, n_steps=500, action_seed=23, flight_mode=7, chase=True, use_yaw_targets=True,
                     goal_reach_distance=0.4, goal_reach_angle=0.3, num_targets=2)
    fly_qx_waypoints("qxwp_euler_sparse", seed=104, n_steps=200, action_seed=24, angle_representation="euler", sparse=True, action_scale=0.5)


def ma_hover_fixtures():
    # MAQuadXHover: default 4 agents; crashes end agents one by one (dead agents keep falling)
    fly_ma_hover("mahover_mode0", seed=201, n_steps=160, action_seed=31)
    fly_ma_hover("mahover_euler_sparse", seed=203, n_steps=120, action_seed=32, angle_representation="euler", sparse=True, action_scale=0.5)
    fly_ma_hover("mahover_mode6_trunc", seed=205, n_steps=100, action_seed=33, flight_mode=6, max_duration_seconds=1.0, action_scale=0.1, dome=10.0)
    fly_ma_hover("mahover_two_agents_small_dome", seed=207, n_steps=140, action_seed=34, dome=2.0, start_pos=[[0.0, 0.0, 1.0], [0.5, 0.5, 1.5]])


def dogfight_fixtures():
    # MAFixedwingDogfight (BASELINE configs[4]): 1-vs-1 arenas; a wide lethal cone makes scripted flights score hits
    fly_dogfight("dogfight_1v1", seed=81, n_steps=260, action_seed=12)
    fly_dogfight("dogfight_1v1_wide_cone", seed=83, n_steps=260, action_seed=13, lethal_distance=150.0, lethal_angle=1.2, action_scale=0.3)
    fly_dogfight("dogfight_1v1_sparse", seed=85, n_steps=150, action_seed=14, sparse=True)
    # nose-down bias: ground collisions (-1000), the survivor's team win (+300), several episodes
    fly_dogfight("dogfight_1v1_crash", seed=91, n_steps=250, action_seed=17, action_scale=0.2, pitch_bias=0.8)
    # heavy damage in a 2-vs-2: deaths by health, team wins, dead agents that keep flying
    fly_dogfight("dogfight_2v2_lethal", seed=87, n_steps=200, action_seed=15, team_size=2, lethal_distance=120.0, lethal_angle=0.9, action_scale=0.3, damage_per_hit=0.05)
    fly_dogfight("dogfight_2v2", seed=87, n_steps=200, action_seed=15, team_size=2, lethal_distance=120.0, lethal_angle=0.9, action_scale=0.3)


def main():
    os.makedirs(OUT, exist_ok=True)
    # A: tests/test_core.py:13-31
    fly_quadx("quadx_mode7_hold", 7, "cf2x", [0, 0, 1], [0, 0, 0], {}, 1000, seed=1)
    # B: tests/test_core.py:65-93 / examples/core/03_control.py
    fly_quadx(
        "quadx_mode7_setpoints", 7, "cf2x", [0, 0, 1], [0, 0, 0],
        {0: [1.0, 0.0, 0.0, 1.0], 500: [0.0, 0.0, np.pi / 4, 2.0]}, 1000, seed=2,
    )
    # every flight mode, both quad models, a tilted start and setpoint changes
    sched = {
        -1: {0: [0.3, 0.31, 0.32, 0.3], 120: [0.5, 0.5, 0.45, 0.5]},
        0: {0: [0.3, -0.2, 0.1, 0.45], 150: [-0.5, 0.4, -0.3, 0.3]},
        1: {0: [0.2, -0.1, 0.5, 0.3], 150: [-0.2, 0.2, -0.5, -0.2]},
        2: {0: [0.2, -0.2, 0.3, 6.0], 150: [-0.3, 0.1, 0.0, 4.0]},
        3: {0: [0.15, -0.1, 0.6, 6.0], 150: [0.0, 0.0, -0.6, 4.5]},
        4: {0: [0.8, -0.5, 0.3, 6.0], 150: [-0.6, 0.4, -0.2, 4.5]},
        5: {0: [0.8, -0.5, 0.3, 0.4], 150: [-0.6, 0.4, -0.2, -0.3]},
        6: {0: [0.9, 0.4, 0.5, 0.3], 150: [-0.5, -0.7, -0.4, -0.2]},
        7: {0: [1.0, -1.0, 0.8, 6.0], 150: [-0.5, 0.5, -0.8, 4.0]},
    }
    for model in ["cf2x", "primitive_drone"]:
        for mode in range(-1, 8):
            fly_quadx(f"quadx_{model}_mode{mode}", mode, model, [0.3, -0.2, 5.0], [0.1, -0.15, 0.7], sched[mode], 300, seed=10 + mode)
    # floor strike: contact flag + "no rotational drag while in contact" (quadx.py:509-510)
    fly_quadx("quadx_floor_contact", 0, "cf2x", [0, 0, 0.3], [0.4, 0.2, 0], {0: [1.0, 2.0, 0.5, 0.05]}, 80, seed=3)
    # long open-sky parity scenario: 3000 Aviary steps = 1000 Hover env-steps
    prng = np.random.default_rng(1)
    sched0 = {}
    for k in range(0, 3000, 30):
        a = prng.uniform([-np.pi, -np.pi, -np.pi, 0.0], [np.pi, np.pi, np.pi, 0.8])
        a[:3] *= 0.3
        sched0[k] = a
    fly_quadx("quadx_mode0_long", 0, "cf2x", [0, 0, 50.0], [0, 0, 0], sched0, 3000, seed=4)
    # Hover env: tests/test_gym_envs.py:92-112 shape (seeded env + scripted actions)
    fly_hover("hover_quat_dense", seed=0, n_steps=400, action_seed=1, angle_representation="quaternion")
    fly_hover("hover_euler_sparse", seed=5, n_steps=200, action_seed=2, angle_representation="euler", sparse=True)
    fly_hover("hover_quat_gentle", seed=6, n_steps=450, action_seed=3, angle_representation="quaternion", action_scale=0.05, dome=50.0)
    fly_hover("hover_mode6", seed=7, n_steps=300, action_seed=4, angle_representation="quaternion", flight_mode=6, action_scale=0.3)


def wind_fixtures():
    """the UNMODIFIED reference flown in an analytic wind (register_wind_field_function, aviary.py:324-334;
    tests/test_core.py:262-290 of the reference does the same with exp(z))."""
    from pyflyt_b200.core.wind import AnalyticWind

    # QuadX, position hold in a power-law boundary layer; the drag body feels it, the controller leans into it
    fly_quadx("wind_quadx_power", 7, "cf2x", [0, 0, 2.0], [0, 0, 0.3], {0: [0.5, -0.5, 0.3, 3.0]}, 400, seed=71,
              wind=AnalyticWind("power", base=(3.0, -1.5, 0.2), z_ref=10.0, alpha=1.0 / 7.0))
    # Fixedwing in a logarithmic profile: every lifting surface sees the wind at its own altitude
    r = np.random.default_rng(72)
    sched = {k: np.concatenate([r.uniform(-0.4, 0.4, 3), r.uniform(0.4, 1.0, 1)]) for k in range(0, 500, 50)}
    fly_vehicle("wind_fixedwing_log", "fixedwing", "fixedwing", 0, [0, 0, 50.0], [0.05, -0.05, 0.4], sched, 500, seed=72,
                wind=AnalyticWind("log", base=(-4.0, 2.0, 0.0), z_ref=10.0, z0=0.03))
    # Rocket: the reference test's own field shape, wind_z = exp(z / z_ref), plus a constant cross wind on a second fixture
    def sched_r(n, every, thr):
        rr = np.random.default_rng(73)
        return {k: np.concatenate([rr.uniform(-0.5, 0.5, 3), [1.0, thr], rr.uniform(-0.5, 0.5, 2)]) for k in range(0, n, every)}
    fly_vehicle("wind_rocket_exp", "rocket", "rocket", 0, [0, 0, 100.0], [0.05, -0.04, 0.3], sched_r(400, 40, 0.8), 400, seed=73,
                wind=AnalyticWind("exp", base=(0.0, 0.0, 1.0), z_ref=60.0))
    fly_vehicle("wind_rocket_constant", "rocket", "rocket", 0, [0, 0, 200.0], [0.1, 0.0, 0.0], sched_r(300, 30, 0.5), 300, seed=74,
                wind=AnalyticWind("constant", base=(6.0, -3.0, 0.5)))
    # one env: QuadX-Hover, a single episode with the field attached to the env's Aviary after reset()
    fly_hover("wind_hover_quat", seed=75, n_steps=80, action_seed=6, angle_representation="quaternion", flight_mode=6, action_scale=0.2, dome=50.0,
              wind=AnalyticWind("constant", base=(2.0, 1.0, 0.0)))


if __name__ == "__main__":
    os.makedirs(OUT, exist_ok=True)
    which = sys.argv[1] if len(sys.argv) > 1 else "all"
    if which in ("all", "quadx"):
        main()
    if which in ("all", "fixedwing"):
        fixedwing_fixtures()
    if which in ("all", "rocket"):
        rocket_fixtures()
    if which in ("all", "dogfight"):
        dogfight_fixtures()
    if which in ("all", "mahover"):
        ma_hover_fixtures()
    if which in ("all", "qxwp"):
        quadx_waypoints_fixtures()
    if which in ("all", "wind"):
        wind_fixtures()
    if which in ("all", "touchdown"):
        touchdown_fixtures()
