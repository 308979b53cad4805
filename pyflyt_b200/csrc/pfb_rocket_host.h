// pfb_rocket_host.h — host-side narrowing of PfbModel to RocketParams.  Included after `fail(...)` and
// pfb_fixedwing_host.h (shared surface / contact builders).
#pragma once

#include "pfb_fixedwing_host.h"
#include "pfb_rocket.cuh"

static int rk_build_params_impl(const PfbModel& m, const PfbEnvConfig* env, pfb::RocketParams& p, pfb::LandingParams& l) {
  memset(&p, 0, sizeof(p));
  memset(&l, 0, sizeof(l));
  if (!m.has_booster) return fail("rocket model needs a booster");
  if (m.n_surfaces < 0 || m.n_surfaces > 4) return fail("rocket model supports up to 4 finlets, got %d", m.n_surfaces);
  p.dt = (float)(1.0 / m.physics_hz);
  p.gravity = (float)m.gravity;
  p.vmax = (float)m.max_coord_velocity;
  p.ratio = (int)(m.physics_hz / m.control_hz);
  if (p.ratio < 1 || p.ratio > 4) return fail("physics_hz / control_hz must be in 1..4 (got %d)", p.ratio);
  {
    double hmax = 0.5 * sqrt(3.0) * m.max_coord_velocity / m.physics_hz;
    if (hmax * hmax > 0.25) return fail("max_coord_velocity * dt too large for the attitude series");
  }
  p.dry_mass = (float)m.dry_mass;
  for (int k = 0; k < 3; ++k) {
    p.dry_mc[k] = (float)m.dry_first_moment[k];
    p.fuel_pos[k] = (float)m.fuel_pos[k];
    p.fuel_max_inertia[k] = (float)m.fuel_max_inertia[k];
    p.body_r[k] = (float)m.body_pos[k];
    p.drag_k[k] = (float)m.drag_const[k];
    p.booster_r[k] = (float)m.booster_pos[k];
    p.booster_axis[k] = (float)m.booster_axis[k];
    p.gimbal_u1[k] = (float)m.gimbal_unit1[k];
    p.gimbal_u2[k] = (float)m.gimbal_unit2[k];
  }
  for (int k = 0; k < 9; ++k) p.dry_I[k] = (float)m.dry_inertia[k];
  p.fuel_total_mass = (float)m.fuel_total_mass;
  p.n_surfaces = m.n_surfaces;
  for (int i = 0; i < m.n_surfaces; ++i) pfb_build_surface(m.surfaces[i], p.surf[i]);
  p.booster_lag = (float)m.booster_dt_over_tau;
  p.booster_noise = (float)m.booster_noise_ratio;
  p.booster_min_ratio = (float)(m.booster_min_thrust / m.booster_max_thrust);
  p.booster_max_thrust = (float)m.booster_max_thrust;
  p.fuel_rate = (float)(m.fuel_max_rate / m.fuel_total_mass);
  p.reignitable = m.reignitable;
  p.gimbal_lag = (float)m.gimbal_dt_over_tau;
  p.gimbal_range[0] = (float)m.gimbal_range_rad[0];
  p.gimbal_range[1] = (float)m.gimbal_range_rad[1];
  p.start_fuel = (float)m.starting_fuel_ratio;
  p.noise_loc = 1.0f;  // one booster: normal(*throttle.shape) == normal(loc=1) (boosters.py:241-245)
  if (pfb_build_contact(m, p.contact)) return -1;
  p.contact_response = (env && env->contact_response) ? 1 : 0;
  if (env) {
    l.env_step_ratio = env->env_step_ratio;
    l.max_steps = env->max_steps;
    l.angle_representation = env->angle_representation;
    l.sparse_reward = env->sparse_reward;
    l.warmup_steps = env->warmup_steps;
    l.randomize_drop = env->randomize_drop;
    l.accelerate_drop = env->accelerate_drop;
    l.ceiling = (float)env->ceiling;
    l.max_displacement = (float)env->max_displacement;
  }
  return 0;
}
