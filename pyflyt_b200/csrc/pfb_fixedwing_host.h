// pfb_fixedwing_host.h — host-side narrowing of PfbModel to the fp32 parameter blocks of the fixedwing
// (and, for the shared pieces, the rocket).  Included after `fail(...)` is in scope.
#pragma once

#include <math.h>
#include <string.h>

#include "../../include/pyflyt_b200.h"
#include "pfb_fixedwing.cuh"

// ---------------------------------------------------------------------------------------------------
// host: PfbModel -> FixedwingParams
// ---------------------------------------------------------------------------------------------------
static void invert6(const double A[6][6], double out[6][6]) {
  double M[6][12];
  for (int i = 0; i < 6; ++i)
    for (int j = 0; j < 6; ++j) { M[i][j] = A[i][j]; M[i][6 + j] = i == j ? 1.0 : 0.0; }
  for (int c = 0; c < 6; ++c) {
    int p = c;
    for (int r = c + 1; r < 6; ++r) if (fabs(M[r][c]) > fabs(M[p][c])) p = r;
    for (int k = 0; k < 12; ++k) { double t = M[c][k]; M[c][k] = M[p][k]; M[p][k] = t; }
    double d = M[c][c];
    for (int k = 0; k < 12; ++k) M[c][k] /= d;
    for (int r = 0; r < 6; ++r) {
      if (r == c) continue;
      double f = M[r][c];
      for (int k = 0; k < 12; ++k) M[r][k] -= f * M[c][k];
    }
  }
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) out[i][j] = M[i][6 + j];
}

static int pfb_build_rigid(double M, const double c[3], const double I[9], pfb::RigidParams& rb) {
  double A[6][6];
  memset(A, 0, sizeof(A));
  const double cx[3][3] = {{0.0, -c[2], c[1]}, {c[2], 0.0, -c[0]}, {-c[1], c[0], 0.0}};
  for (int i = 0; i < 3; ++i) {
    A[i][i] = M;
    for (int j = 0; j < 3; ++j) {
      A[i][3 + j] = -M * cx[i][j];
      A[3 + i][j] = M * cx[i][j];
      A[3 + i][3 + j] = I[3 * i + j];
    }
  }
  double Ai[6][6];
  invert6(A, Ai);
  rb.mass = (float)M;
  for (int k = 0; k < 3; ++k) rb.mc[k] = (float)(M * c[k]);
  for (int k = 0; k < 9; ++k) rb.I[k] = (float)I[k];
  for (int i = 0; i < 6; ++i) for (int j = 0; j < 6; ++j) rb.Ainv[6 * i + j] = (float)Ai[i][j];
  return 0;
}

static int pfb_build_surface(const PfbSurface& s, pfb::SurfaceParams& o) {
  const double pi = 3.14159265358979323846;
  for (int k = 0; k < 3; ++k) { o.r[k] = (float)s.pos[k]; o.lift[k] = (float)s.lift_unit[k]; o.fwd[k] = (float)s.drag_unit[k]; o.tq[k] = (float)s.torque_unit[k]; }
  o.lag = (float)s.dt_over_tau;
  o.Cl_alpha_3D = (float)s.Cl_alpha_3D;
  o.inv_pi_aspect = (float)(1.0 / (pi * s.aspect));
  o.defl_rad = (float)(s.deflection_limit_deg * pi / 180.0);
  o.dCl = (float)(s.Cl_alpha_3D * s.aero_tau * s.eta * (s.deflection_limit_deg * pi / 180.0));
  o.flap_to_chord = (float)s.flap_to_chord;
  o.alpha_0_base = (float)s.alpha_0_base;
  o.alpha_stall_P_base = (float)s.alpha_stall_P_base;
  o.alpha_stall_N_base = (float)s.alpha_stall_N_base;
  o.Cd_0 = (float)s.Cd_0;
  o.stall_k = (float)(0.41 * (1.0 - exp(-17.0 / s.aspect)));
  o.q_area = (float)(s.half_rho * s.area);
  o.chord = (float)s.chord;
  return 0;
}

static int pfb_build_contact(const PfbModel& m, pfb::ContactParams& c) {
  if (m.n_shapes > pfb::kMaxShapes) return fail("at most %d collision primitives are supported, got %d", pfb::kMaxShapes, m.n_shapes);
  c.n_shapes = m.n_shapes;
  c.zmax = -1e30f;
  for (int s = 0; s < m.n_shapes; ++s) {
    const PfbShape& sh = m.shapes[s];
    for (int k = 0; k < 9; ++k) c.rot[s][k] = (float)sh.rot[k];
    double disc;
    if (sh.kind == PFB_SHAPE_BOX) disc = sqrt(sh.dims[0] * sh.dims[0] + sh.dims[1] * sh.dims[1] + sh.dims[2] * sh.dims[2]);
    else if (sh.kind == PFB_SHAPE_CYLINDER) disc = sqrt(sh.dims[0] * sh.dims[0] + sh.dims[1] * sh.dims[1]);
    else disc = sh.dims[0];
    c.kind[s] = sh.kind;
    for (int k = 0; k < 3; ++k) { c.dims[s][k] = (float)sh.dims[k]; c.at[s][k] = (float)sh.at[k]; }
    c.thr[s] = (float)(m.contact_factor * disc);
    double reach = sqrt(sh.at[0] * sh.at[0] + sh.at[1] * sh.at[1] + sh.at[2] * sh.at[2]) + disc + m.contact_factor * disc;
    if ((float)(reach * 1.001) > c.zmax) c.zmax = (float)(reach * 1.001);
  }
  return 0;
}

static int fw_build_params_impl(const PfbModel& m, const PfbEnvConfig* env, pfb::FixedwingParams& p, pfb::WaypointParams& w) {
  memset(&p, 0, sizeof(p));
  memset(&w, 0, sizeof(w));
  if (m.n_surfaces < 1 || m.n_surfaces > pfb::kMaxSurfaces) return fail("fixedwing model needs 1..5 lifting surfaces, got %d", m.n_surfaces);
  if (m.n_motors != 1) return fail("fixedwing model must have exactly one motor, got %d", m.n_motors);
  if (m.motor_axis[0][0] != 1.0 || m.motor_axis[0][1] != 0.0 || m.motor_axis[0][2] != 0.0) return fail("fixedwing motor must thrust along +x");
  p.dt = (float)(1.0 / m.physics_hz);
  p.gravity = (float)m.gravity;
  p.vmax = (float)m.max_coord_velocity;
  p.ratio = (int)(m.physics_hz / m.control_hz);
  if (p.ratio < 1 || p.ratio > 4) return fail("physics_hz / control_hz must be in 1..4 (got %d)", p.ratio);
  {
    double hmax = 0.5 * sqrt(3.0) * m.max_coord_velocity / m.physics_hz;
    if (hmax * hmax > 0.25) return fail("max_coord_velocity * dt too large for the attitude series");
  }
  pfb_build_rigid(m.mass, m.com, m.inertia, p.rb);
  p.n_surfaces = m.n_surfaces;
  for (int i = 0; i < m.n_surfaces; ++i) pfb_build_surface(m.surfaces[i], p.surf[i]);
  for (int k = 0; k < 3; ++k) { p.motor_r[k] = (float)m.motor_pos[0][k]; p.start_vel[k] = (float)m.starting_velocity[k]; }
  p.thrust_k = (float)(m.thrust_coef[0] * m.max_rpm[0] * m.max_rpm[0]);
  p.torque_k = (float)(m.torque_coef[0] * m.max_rpm[0] * m.max_rpm[0]);
  p.motor_lag = (float)m.motor_dt_over_tau[0];
  p.noise_ratio = (float)m.motor_noise_ratio[0];
  p.noise_loc = (float)m.n_motors;
  if (pfb_build_contact(m, p.contact)) return -1;
  if (env) {
    w.env_step_ratio = env->env_step_ratio;
    w.max_steps = env->max_steps;
    w.angle_representation = env->angle_representation;
    w.sparse_reward = env->sparse_reward;
    w.warmup_steps = env->warmup_steps;
    w.flight_mode = env->flight_mode;
    w.num_targets = env->num_targets;
    w.dome = (float)env->flight_dome_size;
    w.dome2 = (float)(env->flight_dome_size * env->flight_dome_size);
    w.goal_reach_distance = (float)env->goal_reach_distance;
    w.min_height = 0.5f;  // fixedwing_waypoints_env.py:83
    if (env->env_kind == PFB_ENV_FIXEDWING_WAYPOINTS) {
      if (w.num_targets < 1 || w.num_targets > pfb::kMaxTargets) return fail("num_targets must be in 1..%d, got %d", pfb::kMaxTargets, w.num_targets);
      if (env->flight_mode != 0) return fail("Fixedwing-Waypoints runs flight mode 0 (4-dim action box), got %d", env->flight_mode);
    }
  }
  return 0;
}

