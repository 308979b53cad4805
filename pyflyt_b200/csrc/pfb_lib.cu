// pfb_lib.cu — CUDA kernels (sm_100a) and the C-ABI of libpyflyt_b200.so (include/pyflyt_b200.h).
//
// One thread integrates one env; the whole env step (control ticks, physics substeps, reward,
// termination, observation) happens in registers between one coalesced SoA load and one store.
// Row-major API buffers (actions [N][4], observations [N][O]) are moved with 16-byte vector loads
// and a shared-memory transpose so that global traffic stays fully coalesced.
#include <cuda_runtime.h>

#include <cstdarg>
#include <cstdio>
#include <cstring>
#include <new>

#include "../../include/pyflyt_b200.h"
#include "pfb_context.h"
#include "pfb_noise.cuh"

using namespace pfb;

// ---------------------------------------------------------------------------------------------------
// error plumbing
// ---------------------------------------------------------------------------------------------------
static thread_local char g_err[512] = "";
int pfb_fail(const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return -1;
}
#include "pfb_quadx_host.h"

// ---------------------------------------------------------------------------------------------------
// kernels — Aviary surface
// ---------------------------------------------------------------------------------------------------
// Aviary.reset + QuadX.reset + update_state (aviary.py:218-312, quadx.py:222-231)
__global__ void __launch_bounds__(kBlock) k_quadx_reset(float* __restrict__ st, int32_t* __restrict__ ist,
                                                        float* __restrict__ setpoint, const float* __restrict__ start_pos,
                                                        const float* __restrict__ start_orn,
                                                        const uint8_t* __restrict__ mask, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  if (mask && !mask[i]) return;
  QuadXRegs s;
  quadx_reset(s, start_pos[3 * i + 0], start_pos[3 * i + 1], start_pos[3 * i + 2], start_orn[3 * i + 0],
              start_orn[3 * i + 1], start_orn[3 * i + 2]);
  quadx_store<7>(st, ist, N, i, s);  // mode 7 touches every PID row
  ist[(int64_t)QI_STEP * N + i] = 0;
  if (setpoint) reinterpret_cast<float4*>(setpoint)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
}

// Aviary.set_mode -> QuadX.set_mode (quadx.py:233-373): preset the setpoint, fresh attitude/position PIDs
template <int MODE>
__global__ void __launch_bounds__(kBlock) k_quadx_set_mode(float* __restrict__ st, int32_t* __restrict__ ist,
                                                           float* __restrict__ setpoint, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  QuadXRegs s;
  quadx_load<7>(st, ist, N, i, s);
  float4 sp = reinterpret_cast<const float4*>(setpoint)[i];
  s.sp[0] = sp.x; s.sp[1] = sp.y; s.sp[2] = sp.z; s.sp[3] = sp.w;
  quadx_set_mode<MODE>(s);
  quadx_store<7>(st, ist, N, i, s);
  reinterpret_cast<float4*>(setpoint)[i] = make_float4(s.sp[0], s.sp[1], s.sp[2], s.sp[3]);
}

// n_steps x Aviary.step() (aviary.py:480-531)
template <int MODE, bool INJECT>
__global__ void __launch_bounds__(kBlock, kMinBlocks)
    k_quadx_aviary_step(const __grid_constant__ QuadXParams p, const __grid_constant__ RngParams rng,
                        float* __restrict__ st, int32_t* __restrict__ ist, const float* __restrict__ setpoint,
                        const float* __restrict__ noise, int n_steps, uint32_t seq, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  QuadXRegs s;
  quadx_load<MODE>(st, ist, N, i, s);
  float4 sp = __ldg(reinterpret_cast<const float4*>(setpoint) + i);
  s.sp[0] = sp.x; s.sp[1] = sp.y; s.sp[2] = sp.z; s.sp[3] = sp.w;
  auto nz = make_noise<INJECT>(noise, N, i, rng, seq, TAG_AVIARY, p.noise_loc, p.ratio);
  for (int k = 0; k < n_steps; ++k) quadx_aviary_step<MODE>(p, s, nz);
  quadx_store<MODE>(st, ist, N, i, s);
}

// Aviary.state(i) / aux_state(i) / contact_array  -> row-major API buffers
__global__ void __launch_bounds__(kBlock) k_quadx_observe(const float* __restrict__ st, const int32_t* __restrict__ ist,
                                                          float* __restrict__ drone_state, float* __restrict__ aux,
                                                          uint8_t* __restrict__ contact, int64_t N) {
  int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= N) return;
  QuadXRegs s;
  quadx_load<-1>(st, ist, N, i, s);
  float o[12], a[4];
  quadx_drone_state(s, o, a);
  if (drone_state) {
    float4* d = reinterpret_cast<float4*>(drone_state + 12 * i);
    d[0] = make_float4(o[0], o[1], o[2], o[3]);
    d[1] = make_float4(o[4], o[5], o[6], o[7]);
    d[2] = make_float4(o[8], o[9], o[10], o[11]);
  }
  if (aux) reinterpret_cast<float4*>(aux)[i] = make_float4(a[0], a[1], a[2], a[3]);
  if (contact) contact[i] = (s.flags & FLAG_CONTACT_ARRAY) ? 1 : 0;
}

// ---------------------------------------------------------------------------------------------------
// kernels — QuadX-Hover env
// ---------------------------------------------------------------------------------------------------
constexpr int kObsStride = 24;  // floats of shared memory per env for the observation tile (rows are packed at stride O <= 24)
// 640 threads per SM resident (<= 96 registers): the regular + tail CTAs of a 65 536-env step and the CTAs of the
// concurrent spare rebuild must all be resident at once, or the stragglers form a second wave
constexpr int kHoverBlocks = 640 / kBlock;

// ---- spare post-reset states ("reset pipeline") ---------------------------------------
// env.reset() = start pose + `warmup_steps` Aviary steps (quadx_base_env.py:149-212): 3.3x the work of an env step and a
// strictly serial chain; reset inline, even one finished env stretches the launch to the length of that chain.  So each
// env owns a SPARE, the post-warm-up state of its NEXT episode, with noise keyed by (env id, episode number, Aviary
// step) so that it does not matter when it is computed.  The step kernel's tail CTAs only COPY the spare of a finished
// env; a second launch of the SAME kernel in build mode (all CTAs are tail CTAs, same compiled warm-up loop, hence
// bit-identical results) rebuilds the spares just consumed on a side stream, concurrently with the following step
// launches (the step two launches later waits for it: an env cannot finish again sooner).  If the start pose was
// edited since a spare was built it is ignored and the warm-up runs inline (same episode number, same result).
// Library-owned buffer [N][SP_ROWS], ENV-MAJOR (a 256-byte record per env: a tail thread touches 2-3 lines of DRAM instead of
// one 32-byte sector per field; measured 9 MB -> ~1 MB of DRAM reads per launch with ~2200 resets): words [0, QX_ROWS) the
// spare's state, then:
enum { SP_POSE = QX_ROWS, SP_VALID = QX_ROWS + 6, SP_FLAGS = QX_ROWS + 7, SP_EPISODE = QX_ROWS + 8, SP_ROWS = 64 };
static_assert(QX_ROWS + 9 <= SP_ROWS, "spare record too small");

__device__ __forceinline__ bool spare_usable(const float* __restrict__ spare, const float* __restrict__ start_pos,
                                             const float* __restrict__ start_orn, int64_t N, int64_t i) {
  const float* c = spare + i * SP_ROWS + SP_POSE;
  bool ok = c[6] != 0.0f;
#pragma unroll
  for (int k = 0; k < 3; ++k) ok = ok && (c[k] == start_pos[3 * i + k]) && (c[3 + k] == start_orn[3 * i + k]);
  return ok;
}

// the pose rows are written by the caller at reset time, from the very values the warm-up started from (the user may
// edit start_pos / start_orn while a rebuild is in flight on the side stream)
template <int MODE>
__device__ __forceinline__ void spare_store(float* __restrict__ spare, int32_t* __restrict__ ist, int64_t N, int64_t i,
                                            const QuadXRegs& s, uint32_t episode) {
  quadx_store<MODE>(spare + i * SP_ROWS, ist, N, i, s, false, 1, 0);
  float* c = spare + i * SP_ROWS + SP_POSE;
  c[7] = __uint_as_float(s.flags);
  c[8] = __uint_as_float(episode);
  c[6] = 1.0f;
}

// row of element j in a dense [rows][O] tile; O is one of the four observation widths (constant divisors, no idiv)
__device__ __forceinline__ int obs_row(int j, int O) { return O == 21 ? j / 21 : (O == 20 ? j / 20 : (O == 24 ? j / 24 : j / 23)); }

// env.reset() body for one env: begin_reset + end_reset (quadx_base_env.py:149-212); obs -> `out`
template <int MODE, bool INJECT>
__device__ __forceinline__ void hover_reset_env(const QuadXParams& p, const HoverParams& h, const RngParams& rng,
                                                float* __restrict__ st, int32_t* __restrict__ ist,
                                                const float* __restrict__ start_pos, const float* __restrict__ start_orn,
                                                const float* __restrict__ noise, uint32_t seq, int64_t N, int64_t i,
                                                float* out) {
  QuadXRegs s;
  quadx_reset(s, start_pos[3 * i + 0], start_pos[3 * i + 1], start_pos[3 * i + 2], start_orn[3 * i + 0],
              start_orn[3 * i + 1], start_orn[3 * i + 2]);
  quadx_set_mode<MODE>(s);
  auto nz = make_noise<INJECT>(noise, N, i, rng, seq, TAG_RESET, p.noise_loc, p.ratio);
  for (int k = 0; k < h.warmup_steps; ++k) quadx_aviary_step<MODE>(p, s, nz);
  const float zero[4] = {0.f, 0.f, 0.f, 0.f};  // self.action = zeros (quadx_base_env.py:165)
  if (h.ma) {  // past_actions is NOT cleared by a reset in the reference: it still holds the previous episode's value
    float past[4];
    for (int k = 0; k < 4; ++k) past[k] = st[(int64_t)(QM_PAST + k) * N + i];
    ma_hover_observation(h, s, past, start_pos[3 * i], start_pos[3 * i + 1], start_pos[3 * i + 2], out);
  } else {
    hover_observation(h, s, zero, out);
  }
  quadx_store<7>(st, ist, N, i, s);
  ist[(int64_t)QI_STEP * N + i] = 0;
}

// env.step(action) for every env (quadx_base_env.py:269-301 + quadx_hover_env.py), ONE launch.
//   RANDACT   actions are drawn on device, uniform in the env's action box (quadx_base_env.py:79-102)
//   AUTORESET gymnasium NEXT_STEP autoreset: an env that finished on the previous call is reset on this
//             one (its action is ignored; obs = first observation, reward 0, flags cleared).  Those envs
//             were queued by the previous launch and are handled by dense "tail" CTAs placed at the
//             front of the grid, so the 10 warm-up Aviary steps run in full warps concurrently with the
//             regular CTAs instead of diverging inside them.
// Both roles run the SAME code (role-dependent scalars only): the kernel is instruction-fetch bound, so one
// compact hot loop shared by every warp on the SM matters more than anything else.
template <int MODE, bool INJECT, bool RANDACT, bool AUTORESET>
__global__ void __launch_bounds__(kBlock, kHoverBlocks)
    k_hover_step(const __grid_constant__ QuadXParams p, const __grid_constant__ HoverParams h,
                 const __grid_constant__ RngParams rng, float* __restrict__ st, int32_t* __restrict__ ist,
                 float* __restrict__ actions, const float* __restrict__ noise, float* __restrict__ obs,
                 float* __restrict__ reward, uint8_t* __restrict__ term, uint8_t* __restrict__ trunc,
                 uint8_t* __restrict__ info, const float* __restrict__ start_pos, const float* __restrict__ start_orn,
                 const int32_t* __restrict__ prev_count, const int32_t* __restrict__ prev_list,
                 int32_t* __restrict__ cur_count, int32_t* __restrict__ cur_list, int32_t* __restrict__ next_count,
                 float* __restrict__ spare, int spare_copy, int build, int tail_blocks, uint32_t step_seq, int64_t N) {
  __shared__ __align__(16) float smem[kBlock * kObsStride];
  __shared__ uint8_t row_skip[kBlock];
  const int O = (h.angle_representation == 0 ? 20 : 21) + (h.ma ? 3 : 0);
  const bool tail = AUTORESET && (int)blockIdx.x < tail_blocks;  // CTA-uniform role
  const int64_t block_first = tail ? 0 : (int64_t)((int)blockIdx.x - (AUTORESET ? tail_blocks : 0)) * kBlock;

  // work items: a regular thread owns exactly one env; a tail thread strides over the done list
  int t, t_end, t_stride;
  if (tail) {
    if (blockIdx.x == 0 && threadIdx.x == 0 && !build) *next_count = 0;  // arm the counter the NEXT launch appends to
    t = blockIdx.x * kBlock + threadIdx.x;
    t_end = prev_list ? *prev_count : (int)N;  // build mode after a user reset: every env
    t_stride = tail_blocks * kBlock;
  } else {
    t = 0;
    t_end = (block_first + threadIdx.x < N) ? 1 : 0;
    t_stride = 1;
  }
  bool skip = true;
  float* row = smem + threadIdx.x * O;  // dense [kBlock][O] tile: copied out below as float4
#pragma unroll 1
  for (; t < t_end; t += t_stride) {
    const int64_t i = tail ? (prev_list ? (int64_t)prev_list[t] : (int64_t)t) : block_first + threadIdx.x;
    QuadXRegs s;
    float act[4] = {0.f, 0.f, 0.f, 0.f};
    float past[4] = {0.f, 0.f, 0.f, 0.f};
    int n_aviary, step_count;
    float rew;
    uint32_t nseq = step_seq;
    // ONE load site for every role: a regular thread reads its env's state, a tail thread the env's spare.  The loads are
    // issued before the spare's pose / validity words are examined, so a cold tail thread pays one round trip, not two.
    const bool from_spare = tail && spare;  // env-major record vs field-major state rows: same loads, different strides
    quadx_load<MODE>(from_spare ? spare + i * SP_ROWS : st, ist, N, i, s, from_spare ? 1 : N, from_spare ? 0 : i);
    if (tail) {
      // env.reset(): begin_reset + end_reset (quadx_base_env.py:149-212) — normally a copy of the env's spare
      nseq = step_seq | 0x40000000u;
      bool hit = false;
      if (spare) {
        nseq = __float_as_uint(spare[i * SP_ROWS + SP_EPISODE]) + (build ? 1u : 0u);  // episode number: keys the warm-up noise
        hit = !build && spare_copy && spare_usable(spare, start_pos, start_orn, N, i);
      }
      if (hit) {
        s.flags = __float_as_uint(spare[i * SP_ROWS + SP_FLAGS]);
#pragma unroll
        for (int k = 0; k < 4; ++k) { s.sp[k] = 0.0f; s.pwm[k] = spare[i * SP_ROWS + QX_PWM + k]; }  // quadx_load skips the pwm words
        n_aviary = 0;
      } else {
        const float px = start_pos[3 * i + 0], py = start_pos[3 * i + 1], pz = start_pos[3 * i + 2];
        const float ox = start_orn[3 * i + 0], oy = start_orn[3 * i + 1], oz = start_orn[3 * i + 2];
        if (build) {
          float* c = spare + i * SP_ROWS + SP_POSE;
          c[6] = 0.0f;  // invalid until the warm-up below is stored
          c[0] = px; c[1] = py; c[2] = pz; c[3] = ox; c[4] = oy; c[5] = oz;
        }
        quadx_reset(s, px, py, pz, ox, oy, oz);
        quadx_set_mode<MODE>(s);
        n_aviary = h.warmup_steps;
      }
      step_count = 0;
      rew = 0.0f;
      if (!build) s.flags |= fresh_tag(step_seq);
    } else {
      if (AUTORESET && (s.flags & (FLAG_TERM | FLAG_TRUNC | fresh_tag(step_seq)))) continue;  // a tail CTA owns this env on this call
      s.flags &= ~(uint32_t)FLAG_FRESH_ANY;
      if (RANDACT) {
        uint64_t g = ((uint64_t)rng.env_offset_hi << 32 | rng.env_offset_lo) + (uint64_t)i;
        U4 r = philox4x32_10(U4{(uint32_t)g, (uint32_t)(g >> 32), step_seq, (uint32_t)TAG_ACTION << 24}, rng.k0, rng.k1);
        const float pi = 3.14159265358979323846f;
        if (MODE == -1) {
          act[0] = 0.8f * u32_to_unit_open(r.x); act[1] = 0.8f * u32_to_unit_open(r.y);
          act[2] = 0.8f * u32_to_unit_open(r.z); act[3] = 0.8f * u32_to_unit_open(r.w);
        } else {
          act[0] = pi * (2.0f * u32_to_unit_open(r.x) - 1.0f); act[1] = pi * (2.0f * u32_to_unit_open(r.y) - 1.0f);
          act[2] = pi * (2.0f * u32_to_unit_open(r.z) - 1.0f); act[3] = 0.8f * u32_to_unit_open(r.w);
        }
        reinterpret_cast<float4*>(actions)[i] = make_float4(act[0], act[1], act[2], act[3]);
      } else {
        float4 a4 = __ldg(reinterpret_cast<const float4*>(actions) + i);
        act[0] = a4.x; act[1] = a4.y; act[2] = a4.z; act[3] = a4.w;
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) s.sp[k] = act[k];
      n_aviary = h.env_step_ratio;
      step_count = ist[(int64_t)QI_STEP * N + i];
      rew = -0.1f;
      if (h.ma) {  // MAQuadXHover: flags are re-evaluated every step, rewards add up from 0, the obs shows the PREVIOUS action
        s.flags &= ~(uint32_t)(FLAG_TERM | FLAG_TRUNC | FLAG_OOB | FLAG_COLLISION);
        rew = 0.0f;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          past[k] = st[(int64_t)(QM_CUR + k) * N + i];
          st[(int64_t)(QM_PAST + k) * N + i] = past[k];
          st[(int64_t)(QM_CUR + k) * N + i] = act[k];
        }
      }
    }
    auto nz = make_noise<INJECT>(noise, N, i, rng, nseq, tail ? TAG_RESET : TAG_ENV_STEP, p.noise_loc, p.ratio);
#pragma unroll 1
    for (int k = 0; k < n_aviary; ++k) {
      if (!h.ma && (s.flags & (FLAG_TERM | FLAG_TRUNC))) break;  // quadx_base_env.py:289-290 (never set while resetting)
      quadx_aviary_step<MODE>(p, s, nz);
      if (!tail) {
        if (h.ma) ma_hover_term_trunc_reward(h, s, step_count, start_pos[3 * i], start_pos[3 * i + 1], start_pos[3 * i + 2], rew);
        else hover_term_trunc_reward(h, s, step_count, rew);
      }
    }
    step_count = tail ? 0 : step_count + 1;
    if (tail && build) {  // build mode: the warm-up result is the env's new spare
      spare_store<MODE>(spare, ist, N, i, s, nseq);
      continue;
    }
    if (tail && n_aviary > 0) quadx_requantize(s);  // an inline warm-up must leave exactly what a copied spare holds
    if (h.ma) ma_hover_observation(h, s, past, start_pos[3 * i], start_pos[3 * i + 1], start_pos[3 * i + 2], row);
    else hover_observation(h, s, act, row);
    quadx_store<MODE>(st, ist, N, i, s);
    ist[(int64_t)QI_STEP * N + i] = step_count;
    reward[i] = rew;
    term[i] = (s.flags & FLAG_TERM) ? 1 : 0;
    trunc[i] = (s.flags & FLAG_TRUNC) ? 1 : 0;
    if (info) info[i] = (uint8_t)(((s.flags & FLAG_OOB) ? 1 : 0) | ((s.flags & FLAG_COLLISION) ? 2 : 0));
    if (tail) {
      float* dst = obs + i * O;  // scattered rows: the tail handles ~1-3 % of the envs
      for (int k = 0; k < O; ++k) dst[k] = row[k];
    } else {
      skip = false;
      if (AUTORESET) {  // queue finished episodes for the next launch's tail CTAs (warp-aggregated append)
        bool done = (s.flags & (FLAG_TERM | FLAG_TRUNC)) != 0;
        unsigned m = __ballot_sync(__activemask(), done);
        if (done) {
          int lane = threadIdx.x & 31;
          int leader = __ffs(m) - 1;
          int base = 0;
          if (lane == leader) base = atomicAdd(cur_count, __popc(m));
          base = __shfl_sync(m, base, leader);
          cur_list[base + __popc(m & ((1u << lane) - 1u))] = (int32_t)i;
        }
      }
    }
  }
  if (tail) return;
  // ---- block-cooperative write of this CTA's observation tile obs[block_first .. +rows][O]: the tile is contiguous in
  //      global memory and 16-byte aligned (kBlock * O * 4 bytes per CTA), so it goes out as float4; rows of envs that a
  //      tail CTA owns on this launch are left alone
  row_skip[threadIdx.x] = skip ? 1 : 0;
  const int any_skip = __syncthreads_or(skip ? 1 : 0);
  int64_t rows = N - block_first;
  if (rows > kBlock) rows = kBlock;
  const int total = (int)rows * O;
  const int nvec = total >> 2;
  float* dst = obs + block_first * O;
  const float4* src4 = reinterpret_cast<const float4*>(smem);
  float4* dst4 = reinterpret_cast<float4*>(dst);
  if (!any_skip) {
    for (int v = threadIdx.x; v < nvec; v += kBlock) dst4[v] = src4[v];
  } else {
    for (int v = threadIdx.x; v < nvec; v += kBlock) {
      const int j = v << 2;
      const int r0 = obs_row(j, O), r1 = obs_row(j + 3, O);
      const float4 val = src4[v];
      if (!row_skip[r0] && !row_skip[r1]) {
        dst4[v] = val;
      } else {  // the vector straddles a skipped row: element-wise
        const int split = r1 * O - j;  // elements [0, split) belong to row r0
        const float e[4] = {val.x, val.y, val.z, val.w};
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (!row_skip[q < split ? r0 : r1]) dst[j + q] = e[q];
      }
    }
  }
  for (int j = (nvec << 2) + threadIdx.x; j < total; j += kBlock)  // ragged last CTA only
    if (!row_skip[obs_row(j, O)]) dst[j] = smem[j];
}

// env.reset() for all / masked envs
template <int MODE, bool INJECT>
__global__ void __launch_bounds__(kBlock)
    k_hover_reset(const __grid_constant__ QuadXParams p, const __grid_constant__ HoverParams h,
                  const __grid_constant__ RngParams rng, float* __restrict__ st, int32_t* __restrict__ ist,
                  const float* __restrict__ start_pos, const float* __restrict__ start_orn,
                  const uint8_t* __restrict__ mask, const float* __restrict__ noise, float* __restrict__ obs,
                  uint32_t seq, int64_t N) {
  __shared__ float smem[kBlock * kObsStride];
  const int64_t i = (int64_t)blockIdx.x * kBlock + threadIdx.x;
  if (i >= N) return;
  if (mask && !mask[i]) return;
  const int O = (h.angle_representation == 0 ? 20 : 21) + (h.ma ? 3 : 0);
  float* row = smem + threadIdx.x * kObsStride;
  hover_reset_env<MODE, INJECT>(p, h, rng, st, ist, start_pos, start_orn, noise, seq, N, i, row);
  if (obs) {
    for (int k = 0; k < O; ++k) obs[i * O + k] = row[k];
  }
}

// ---------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------
#define PFB_MODE_SWITCH(mode, BODY)                         \
  switch (mode) {                                           \
    case -1: { constexpr int MODE = -1; BODY; } break;      \
    case 0: { constexpr int MODE = 0; BODY; } break;        \
    case 1: { constexpr int MODE = 1; BODY; } break;        \
    case 2: { constexpr int MODE = 2; BODY; } break;        \
    case 3: { constexpr int MODE = 3; BODY; } break;        \
    case 4: { constexpr int MODE = 4; BODY; } break;        \
    case 5: { constexpr int MODE = 5; BODY; } break;        \
    case 6: { constexpr int MODE = 6; BODY; } break;        \
    case 7: { constexpr int MODE = 7; BODY; } break;        \
    default: return fail("`mode` must be between -1 and 7, got %d", mode); \
  }

extern "C" {

const char* pfb_last_error(void) { return g_err; }
int pfb_abi_version(void) { return PFB_ABI_VERSION; }
int pfb_sizeof_model(void) { return (int)sizeof(PfbModel); }
int pfb_sizeof_env_config(void) { return (int)sizeof(PfbEnvConfig); }
int pfb_sizeof_buffers(void) { return (int)sizeof(PfbBuffers); }

int pfb_create(const PfbModel* model, const PfbEnvConfig* env, int64_t n_envs, int device, uint64_t seed, PfbHandle* out) {
  if (!model || !out) return fail("pfb_create: null argument");
  if (model->abi_version != PFB_ABI_VERSION) return fail("PfbModel ABI %d != library ABI %d", model->abi_version, PFB_ABI_VERSION);
  if (n_envs <= 0) return fail("n_envs must be positive");
  if (model->kind != PFB_KIND_QUADX && model->kind != PFB_KIND_FIXEDWING && model->kind != PFB_KIND_ROCKET)
    return fail("unknown vehicle kind %d", model->kind);
  int count = 0;
  cudaError_t e = cudaGetDeviceCount(&count);
  if (e != cudaSuccess || count == 0)
    return fail("no CUDA device: libpyflyt_b200 has no CPU fallback (%s)", e != cudaSuccess ? cudaGetErrorString(e) : "0 devices");
  if (device < 0 || device >= count) return fail("device %d out of range (have %d)", device, count);
  CUDA_OK(cudaSetDevice(device));
  PfbContext* c = new (std::nothrow) PfbContext();
  if (!c) return fail("out of host memory");
  memset(c, 0, sizeof(*c));
  c->model = *model;
  if (env) c->env = *env;
  c->n = n_envs;
  c->device = device;
  if (model->kind == PFB_KIND_QUADX) {
    if (build_quadx_params(*model, c->qx) != 0) { delete c; return -1; }
  } else if (model->kind == PFB_KIND_FIXEDWING) {
    if (fw_build_params(*model, env, c->fw, c->wp) != 0) { delete c; return -1; }
    if (df_build_params(env, c->df) != 0) { delete c; return -1; }
    if (env && env->env_kind == PFB_ENV_DOGFIGHT && (n_envs % (2 * env->team_size)) != 0) {
      delete c;
      return fail("n_envs (%lld) must be a multiple of the arena size 2*team_size = %d", (long long)n_envs, 2 * env->team_size);
    }
  } else {
    if (rk_build_params(*model, env, c->rk, c->land) != 0) { delete c; return -1; }
  }
  c->hover.env_step_ratio = env ? env->env_step_ratio : 1;
  c->hover.max_steps = env ? env->max_steps : 0;
  c->hover.angle_representation = env ? env->angle_representation : 1;
  c->hover.sparse_reward = env ? env->sparse_reward : 0;
  c->hover.warmup_steps = env ? env->warmup_steps : 0;
  c->hover.flight_mode = env ? env->flight_mode : 0;
  {
    double dome = env ? env->flight_dome_size : INFINITY;
    c->hover.dome2 = (float)(dome * dome);
  }
  c->hover.ma = (env && env->env_kind == PFB_ENV_MA_QUADX_HOVER) ? 1 : 0;
  if (c->hover.ma && env->autoreset) {
    delete c;
    return fail("MAQuadXHover is a per-agent epilogue: arenas are reset by the caller (pfb_env_reset with a mask), autoreset must be 0");
  }
  if (env && env->env_kind != PFB_ENV_NONE) {
    const bool ok = (model->kind == PFB_KIND_QUADX && env->env_kind == PFB_ENV_QUADX_HOVER) ||
                    (model->kind == PFB_KIND_QUADX && env->env_kind == PFB_ENV_QUADX_WAYPOINTS) ||
                    (model->kind == PFB_KIND_QUADX && env->env_kind == PFB_ENV_MA_QUADX_HOVER) ||
                    (model->kind == PFB_KIND_FIXEDWING && env->env_kind == PFB_ENV_FIXEDWING_WAYPOINTS) ||
                    (model->kind == PFB_KIND_FIXEDWING && env->env_kind == PFB_ENV_DOGFIGHT) ||
                    (model->kind == PFB_KIND_ROCKET && env->env_kind == PFB_ENV_ROCKET_LANDING);
    if (!ok) {
      delete c;
      return fail("env kind %d is not available for vehicle kind %d in this library", env->env_kind, model->kind);
    }
  }
  if (env && env->env_kind == PFB_ENV_QUADX_WAYPOINTS) {
    if (env->num_targets < 1 || env->num_targets > kMaxTargets) {
      delete c;
      return fail("num_targets must be in 1..%d, got %d", kMaxTargets, env->num_targets);
    }
    c->qwp.env_step_ratio = env->env_step_ratio;
    c->qwp.max_steps = env->max_steps;
    c->qwp.sparse_reward = env->sparse_reward;
    c->qwp.warmup_steps = env->warmup_steps;
    c->qwp.num_targets = env->num_targets;
    c->qwp.use_yaw_targets = env->use_yaw_targets ? 1 : 0;
    c->qwp.dome = (float)env->flight_dome_size;
    c->qwp.dome2 = (float)(env->flight_dome_size * env->flight_dome_size);
    c->qwp.goal_reach_distance = (float)env->goal_reach_distance;
    c->qwp.goal_reach_angle = (float)env->goal_reach_angle;
    c->qwp.min_height = 0.1f;  // quadx_waypoints_env.py:88
  }
  c->rng.k0 = (uint32_t)seed;
  c->rng.k1 = (uint32_t)(seed >> 32);
  c->mode = 0;
  cudaDeviceProp prop;
  CUDA_OK(cudaGetDeviceProperties(&prop, device));
  c->sm_count = prop.multiProcessorCount;
  CUDA_OK(cudaMalloc(&c->d_counters, 8 * sizeof(int32_t)));  // [0..3] rotating autoreset counters, [4] ticket of the split dogfight
  CUDA_OK(cudaMemset(c->d_counters, 0, 8 * sizeof(int32_t)));
  CUDA_OK(cudaMalloc(&c->d_done_list, 4 * (size_t)n_envs * sizeof(int32_t)));
  if (env && env->autoreset && (env->env_kind == PFB_ENV_QUADX_HOVER || env->env_kind == PFB_ENV_FIXEDWING_WAYPOINTS ||
                                 env->env_kind == PFB_ENV_QUADX_WAYPOINTS || env->env_kind == PFB_ENV_ROCKET_LANDING ||
                                 env->env_kind == PFB_ENV_DOGFIGHT)) {
    // spare post-reset states: one env-major record per env (64 floats; 128 for QuadX-Waypoints, whose record holds 8 x 4 targets)
    const size_t rec = env->env_kind == PFB_ENV_QUADX_WAYPOINTS ? (size_t)qwp_spare_rows()
                       : (env->env_kind == PFB_ENV_DOGFIGHT ? (size_t)df_spare_rows() : (size_t)SP_ROWS);
    CUDA_OK(cudaMalloc(&c->d_spare, rec * (size_t)n_envs * sizeof(float)));
    CUDA_OK(cudaMemset(c->d_spare, 0, rec * (size_t)n_envs * sizeof(float)));
    int prio_lo = 0, prio_hi = 0;  // the rebuild is small and latency-critical: let its CTAs go first when slots free up
    CUDA_OK(cudaDeviceGetStreamPriorityRange(&prio_lo, &prio_hi));
    CUDA_OK(cudaStreamCreateWithPriority(&c->side, cudaStreamNonBlocking, prio_hi));
    CUDA_OK(cudaEventCreateWithFlags(&c->ev_step, cudaEventDisableTiming));
    for (int k = 0; k < 4; ++k) CUDA_OK(cudaEventCreateWithFlags(&c->ev_spare[k], cudaEventDisableTiming));
  }
  *out = c;
  return 0;
}

int pfb_destroy(PfbHandle h) {
  if (!h) return 0;
  cudaSetDevice(h->device);
  if (h->d_spare) {
    cudaStreamSynchronize(h->side);
    cudaStreamDestroy(h->side);
    cudaEventDestroy(h->ev_step);
    for (int k = 0; k < 4; ++k) cudaEventDestroy(h->ev_spare[k]);
    cudaFree(h->d_spare);
  }
  cudaFree(h->d_counters);
  cudaFree(h->d_done_list);
  if (h->prof_ev) {
    for (int i = 0; i < 2 * h->prof_cap; ++i) cudaEventDestroy(h->prof_ev[i]);
    delete[] h->prof_ev;
  }
  delete h;
  return 0;
}

int pfb_set_env_offset(PfbHandle h, uint64_t first_global_env) {
  if (!h) return fail("null handle");
  h->rng.env_offset_lo = (uint32_t)first_global_env;
  h->rng.env_offset_hi = (uint32_t)(first_global_env >> 32);
  return 0;
}

static inline bool is_fw(PfbHandle h) { return h->model.kind == PFB_KIND_FIXEDWING; }
static inline bool is_rk(PfbHandle h) { return h->model.kind == PFB_KIND_ROCKET; }
static inline bool is_qwp(PfbHandle h) { return h->model.kind == PFB_KIND_QUADX && h->env.env_kind == PFB_ENV_QUADX_WAYPOINTS; }
static inline bool is_ma(PfbHandle h) { return h->model.kind == PFB_KIND_QUADX && h->env.env_kind == PFB_ENV_MA_QUADX_HOVER; }
static inline bool is_df(PfbHandle h) { return h->model.kind == PFB_KIND_FIXEDWING && h->env.env_kind == PFB_ENV_DOGFIGHT; }
int pfb_state_rows(PfbHandle h) { return is_rk(h) ? rk_state_rows() : (is_fw(h) ? fw_state_rows() : (is_qwp(h) ? qwp_state_rows() : (is_ma(h) ? QM_ROWS : QX_ROWS))); }
int pfb_istate_rows(PfbHandle h) { return is_rk(h) ? rk_istate_rows() : (is_fw(h) ? fw_istate_rows() : (is_qwp(h) ? qwp_istate_rows() : QI_ROWS)); }
int pfb_setpoint_dim(PfbHandle h) { return is_rk(h) ? 7 : ((is_fw(h) && h->env.env_kind == PFB_ENV_NONE) ? 6 : 4); }
int pfb_obs_dim(PfbHandle h) { return is_df(h) ? df_obs_dim(h) : is_rk(h) ? rk_obs_dim(h) : (is_fw(h) ? fw_obs_dim(h) : (is_qwp(h) ? qwp_obs_dim(h) : (h->hover.angle_representation == 0 ? 20 : 21) + (is_ma(h) ? 3 : 0))); }
int pfb_aux_dim(PfbHandle h) { return is_rk(h) ? 9 : (is_fw(h) ? 6 : 4); }

int pfb_bind(PfbHandle h, const PfbBuffers* b) {
  if (!h || !b) return fail("pfb_bind: null argument");
  if (!b->state || !b->istate || !b->setpoint || !b->start_pos || !b->start_orn)
    return fail("pfb_bind: state, istate, setpoint, start_pos and start_orn are mandatory");
  if (((uintptr_t)b->setpoint & 15) || ((uintptr_t)b->state & 15)) return fail("pfb_bind: buffers must be 16-byte aligned");
  h->buf = *b;
  h->bound = true;
  return 0;
}

#define REQUIRE_BOUND(h)                                   \
  if (!(h)) return fail("null handle");                    \
  if (!(h)->bound) return fail("buffers not bound: call pfb_bind first"); \
  CUDA_OK(cudaSetDevice((h)->device));


int pfb_reset(PfbHandle h, const uint8_t* mask, void* stream) {
  REQUIRE_BOUND(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (is_fw(h)) return fw_reset(h, mask, s);
  if (is_rk(h)) return rk_reset(h, mask, s);
  k_quadx_reset<<<grid_for(h->n), kBlock, 0, s>>>(h->buf.state, h->buf.istate, h->buf.setpoint, h->buf.start_pos,
                                                  h->buf.start_orn, mask, h->n);
  LAUNCH_CHECK(h);
  if (!mask) h->mode = 0;  // QuadX.reset() calls set_mode(0) (quadx.py:224)
  return 0;
}

int pfb_set_mode(PfbHandle h, int mode, void* stream) {
  REQUIRE_BOUND(h);
  cudaStream_t s = (cudaStream_t)stream;
  if (is_fw(h)) return fw_set_mode(h, mode, s);
  if (is_rk(h)) return rk_set_mode(h, mode, s);
  PFB_MODE_SWITCH(mode, (k_quadx_set_mode<MODE><<<grid_for(h->n), kBlock, 0, s>>>(h->buf.state, h->buf.istate,
                                                                                  h->buf.setpoint, h->n)));
  LAUNCH_CHECK(h);
  h->mode = mode;
  return 0;
}

int pfb_aviary_step(PfbHandle h, int n_steps, const float* noise, void* stream) {
  REQUIRE_BOUND(h);
  if (n_steps <= 0) return fail("n_steps must be positive");
  cudaStream_t s = (cudaStream_t)stream;
  if (is_fw(h)) return fw_aviary_step(h, n_steps, noise, s);
  if (is_rk(h)) return rk_aviary_step(h, n_steps, noise, s);
  const int mode = h->mode;
  const uint32_t seq = (uint32_t)h->aviary_seq++;
  if (noise) {
    PFB_MODE_SWITCH(mode, (k_quadx_aviary_step<MODE, true><<<grid_for(h->n), kBlock, 0, s>>>(
                              h->qx, h->rng, h->buf.state, h->buf.istate, h->buf.setpoint, noise, n_steps, seq, h->n)));
  } else {
    PFB_MODE_SWITCH(mode, (k_quadx_aviary_step<MODE, false><<<grid_for(h->n), kBlock, 0, s>>>(
                              h->qx, h->rng, h->buf.state, h->buf.istate, h->buf.setpoint, nullptr, n_steps, seq, h->n)));
  }
  LAUNCH_CHECK(h);
  return 0;
}

int pfb_set_base_velocity(PfbHandle h, const float* lin_vel, const float* ang_vel, void* stream) {
  REQUIRE_BOUND(h);
  if (!lin_vel || !ang_vel) return fail("pfb_set_base_velocity: null argument");
  if (is_rk(h)) return rk_set_velocity(h, lin_vel, ang_vel, (cudaStream_t)stream);
  return fail("pfb_set_base_velocity is only built for the rocket (the one vehicle whose env calls resetBaseVelocity)");
}

int pfb_observe_state(PfbHandle h, void* stream) {
  REQUIRE_BOUND(h);
  if (is_fw(h)) return fw_observe(h, (cudaStream_t)stream);
  if (is_rk(h)) return rk_observe(h, (cudaStream_t)stream);
  k_quadx_observe<<<grid_for(h->n), kBlock, 0, (cudaStream_t)stream>>>(h->buf.state, h->buf.istate, h->buf.drone_state,
                                                                      h->buf.aux_state, h->buf.contact, h->n);
  LAUNCH_CHECK(h);
  return 0;
}

static int require_env(PfbHandle h) {
  if (h->env.env_kind == PFB_ENV_NONE) return fail("handle was created without an env epilogue");
  if (!h->buf.obs || !h->buf.reward || !h->buf.term || !h->buf.trunc) return fail("obs/reward/term/trunc buffers are not bound");
  return 0;
}

int pfb_env_reset(PfbHandle h, const uint8_t* mask, const float* noise, void* stream) {
  REQUIRE_BOUND(h);
  if (require_env(h)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  if (is_df(h)) return df_env_reset(h, mask, noise, s);
  if (is_fw(h)) return fw_env_reset(h, mask, noise, s);
  if (is_rk(h)) return rk_env_reset(h, mask, noise, s);
  if (is_qwp(h)) return qwp_env_reset(h, mask, noise, s);
  const int mode = h->hover.flight_mode;
  // resets draw from their own Philox stream; the high bit keeps them apart from in-step autoresets
  const uint32_t seq = 0x80000000u | (uint32_t)h->reset_seq++;
  if (h->d_spare) {  // the reset kernel rewrites spares: let the side stream's last rebuild finish first
    if (h->step_seq > 0) CUDA_OK(cudaStreamWaitEvent(s, h->ev_spare[(h->step_seq - 1) % 4], 0));
    if (!mask) CUDA_OK(cudaMemsetAsync(h->d_counters, 0, 4 * sizeof(int32_t), s));  // a full reset empties the autoreset queues
  }
  if (noise) {
    PFB_MODE_SWITCH(mode, (k_hover_reset<MODE, true><<<grid_for(h->n), kBlock, 0, s>>>(
                              h->qx, h->hover, h->rng, h->buf.state, h->buf.istate, h->buf.start_pos, h->buf.start_orn, mask,
                              noise, h->buf.obs, seq, h->n)));
  } else {
    PFB_MODE_SWITCH(mode, (k_hover_reset<MODE, false><<<grid_for(h->n), kBlock, 0, s>>>(
                              h->qx, h->hover, h->rng, h->buf.state, h->buf.istate, h->buf.start_pos, h->buf.start_orn, mask,
                              nullptr, h->buf.obs, seq, h->n)));
  }
  LAUNCH_CHECK(h);
  if (h->d_spare) {  // every env gets a fresh spare (build mode of the step kernel over all envs, same stream)
    const int g = grid_for(h->n);
    PFB_MODE_SWITCH(mode, (k_hover_step<MODE, false, false, true><<<g, kBlock, 0, s>>>(
                              h->qx, h->hover, h->rng, h->buf.state, h->buf.istate, h->buf.setpoint, nullptr, h->buf.obs, h->buf.reward,
                              h->buf.term, h->buf.trunc, h->buf.info, h->buf.start_pos, h->buf.start_orn, nullptr, nullptr, nullptr, nullptr,
                              nullptr, h->d_spare, 0, 1, g, 0u, h->n)));
    LAUNCH_CHECK(h);
  }
  h->mode = mode;
  return 0;
}

static int env_step_impl(PfbHandle h, float* actions, const float* noise, bool randact, cudaStream_t s) {
  if (is_df(h)) return df_env_step(h, actions, noise, randact, s);
  if (is_fw(h)) return fw_env_step(h, actions, noise, randact, s);
  if (is_rk(h)) return rk_env_step(h, actions, noise, randact, s);
  if (is_qwp(h)) return qwp_env_step(h, actions, noise, randact, s);
  const int mode = h->hover.flight_mode;
  const bool autoreset = h->env.autoreset != 0;
  const uint64_t k = h->step_seq;
  // four rotating done lists / counters: step k appends to [k % 4], its tail CTAs and the side-stream spare rebuild read
  // [(k - 1) % 4], and it zeroes counter [(k + 1) % 4] (last read by the rebuild of step k - 2, which step k waits for)
  int32_t* cnt_cur = h->d_counters + (k % 4);
  int32_t* cnt_prev = h->d_counters + ((k + 3) % 4);
  int32_t* cnt_next = h->d_counters + ((k + 1) % 4);
  int32_t* list_cur = h->d_done_list + (k % 4) * h->n;
  int32_t* list_prev = h->d_done_list + ((k + 3) % 4) * h->n;
  const uint32_t seq = (uint32_t)k;
  const bool spares = autoreset && h->d_spare != nullptr;
  const int spare_copy = (spares && !h->env.inline_reset) ? 1 : 0;
  if (spares && k >= 2) CUDA_OK(cudaStreamWaitEvent(s, h->ev_spare[(k - 2) % 4], 0));
  // tail CTAs (front of the grid) reset the envs that finished on the previous call; one per SM is
  // plenty for the ~1-3 % of envs that finish per step, and the loop is grid-strided anyway
  int tail = 0;
  if (autoreset) {
    tail = h->sm_count;
    int need = grid_for(h->n);
    if (tail > need) tail = need;
  }
  const int grid = grid_for(h->n) + tail;
  const bool prof = h->prof_ev && h->prof_n < h->prof_cap;
  if (prof) CUDA_OK(cudaEventRecord(h->prof_ev[2 * h->prof_n], s));
#define STEP_ARGS h->qx, h->hover, h->rng, h->buf.state, h->buf.istate, actions, noise, h->buf.obs, h->buf.reward,     \
                  h->buf.term, h->buf.trunc, h->buf.info, h->buf.start_pos, h->buf.start_orn, cnt_prev, list_prev, \
                  cnt_cur, list_cur, cnt_next, h->d_spare, spare_copy, 0, tail, seq, h->n
  if (autoreset) {
    if (noise) return fail("injected noise (parity mode) is only supported with autoreset = 0");
    if (randact) {
      PFB_MODE_SWITCH(mode, (k_hover_step<MODE, false, true, true><<<grid, kBlock, 0, s>>>(STEP_ARGS)));
    } else {
      PFB_MODE_SWITCH(mode, (k_hover_step<MODE, false, false, true><<<grid, kBlock, 0, s>>>(STEP_ARGS)));
    }
  } else {
    if (noise) {
      PFB_MODE_SWITCH(mode, (k_hover_step<MODE, true, false, false><<<grid, kBlock, 0, s>>>(STEP_ARGS)));
    } else if (randact) {
      PFB_MODE_SWITCH(mode, (k_hover_step<MODE, false, true, false><<<grid, kBlock, 0, s>>>(STEP_ARGS)));
    } else {
      PFB_MODE_SWITCH(mode, (k_hover_step<MODE, false, false, false><<<grid, kBlock, 0, s>>>(STEP_ARGS)));
    }
  }
#undef STEP_ARGS
  LAUNCH_CHECK(h);
  if (prof) {
    CUDA_OK(cudaEventRecord(h->prof_ev[2 * h->prof_n + 1], s));
    h->prof_n += 1;
  }
  if (spares) {  // rebuild the spares this launch consumed, on the side stream, while the next launches run
    CUDA_OK(cudaEventRecord(h->ev_step, s));
    CUDA_OK(cudaStreamWaitEvent(h->side, h->ev_step, 0));
    PFB_MODE_SWITCH(mode, (k_hover_step<MODE, false, false, true><<<h->sm_count, kBlock, 0, h->side>>>(
                              h->qx, h->hover, h->rng, h->buf.state, h->buf.istate, actions, nullptr, h->buf.obs, h->buf.reward, h->buf.term,
                              h->buf.trunc, h->buf.info, h->buf.start_pos, h->buf.start_orn, cnt_prev, list_prev, cnt_cur, list_cur, cnt_next,
                              h->d_spare, 0, 1, h->sm_count, seq, h->n)));
    LAUNCH_CHECK(h);
    CUDA_OK(cudaEventRecord(h->ev_spare[k % 4], h->side));
  }
  h->step_seq += 1;
  return 0;
}

int pfb_env_step(PfbHandle h, const float* actions, const float* noise, void* stream) {
  REQUIRE_BOUND(h);
  if (require_env(h)) return -1;
  if (actions && ((uintptr_t)actions & 15)) return fail("pfb_env_step: actions must be 16-byte aligned");
  return env_step_impl(h, actions ? const_cast<float*>(actions) : h->buf.setpoint, noise, false, (cudaStream_t)stream);
}

int pfb_env_rollout(PfbHandle h, int n_steps, void* stream) {
  REQUIRE_BOUND(h);
  if (require_env(h)) return -1;
  for (int k = 0; k < n_steps; ++k)
    if (env_step_impl(h, h->buf.setpoint, nullptr, true, (cudaStream_t)stream)) return -1;
  return 0;
}

int pfb_env_step_host(PfbHandle h, const float* host_actions, float* host_obs, float* host_reward, uint8_t* host_term,
                      uint8_t* host_trunc, void* stream) {
  REQUIRE_BOUND(h);
  if (require_env(h)) return -1;
  cudaStream_t s = (cudaStream_t)stream;
  const int O = pfb_obs_dim(h);
  CUDA_OK(cudaMemcpyAsync(h->buf.setpoint, host_actions, (size_t)h->n * pfb_setpoint_dim(h) * sizeof(float), cudaMemcpyHostToDevice, s));
  if (env_step_impl(h, h->buf.setpoint, nullptr, false, s)) return -1;
  // obs | reward | term | trunc laid out back to back on both sides (what the Python mirror allocates): one copy, one
  // PCIe transaction stream instead of four latency-bound ones
  const size_t ob = (size_t)h->n * O * sizeof(float), rb = (size_t)h->n * sizeof(float), fb = (size_t)h->n;
  const char* d0 = (const char*)h->buf.obs;
  char* h0 = (char*)host_obs;
  const bool packed = (const char*)h->buf.reward == d0 + ob && (const char*)h->buf.term == d0 + ob + rb && (const char*)h->buf.trunc == d0 + ob + rb + fb &&
                      (char*)host_reward == h0 + ob && (char*)host_term == h0 + ob + rb && (char*)host_trunc == h0 + ob + rb + fb;
  if (packed) {
    CUDA_OK(cudaMemcpyAsync(host_obs, h->buf.obs, ob + rb + 2 * fb, cudaMemcpyDeviceToHost, s));
    return 0;
  }
  CUDA_OK(cudaMemcpyAsync(host_obs, h->buf.obs, ob, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaMemcpyAsync(host_reward, h->buf.reward, rb, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaMemcpyAsync(host_term, h->buf.term, fb, cudaMemcpyDeviceToHost, s));
  CUDA_OK(cudaMemcpyAsync(host_trunc, h->buf.trunc, fb, cudaMemcpyDeviceToHost, s));
  return 0;
}

// ---- MAFixedwingDogfight, split variant: an arena's agents on different ranks
int pfb_dogfight_payload_dim(void) { return 20; }

int pfb_dogfight_physics(PfbHandle h, const float* actions, const float* noise, float* payload_out, int first, int do_reset, int aviary_index,
                         void* stream) {
  REQUIRE_BOUND(h);
  if (!is_df(h)) return fail("handle is not a dogfight env");
  if (!payload_out) return fail("pfb_dogfight_physics: null payload buffer");
  return df_split_physics(h, actions ? actions : h->buf.setpoint, noise, payload_out, nullptr, 0, 0, nullptr, 0, 0, first, do_reset, aviary_index,
                          (cudaStream_t)stream);
}

int pfb_dogfight_physics_peer(PfbHandle h, const float* actions, const float* noise, const uint64_t* peer_tables_dev, int world,
                             int64_t slot_offset_floats, const uint64_t* peer_flags_dev, int rank, int epoch, int first, int do_reset,
                             int aviary_index, void* stream) {
  REQUIRE_BOUND(h);
  if (!is_df(h)) return fail("handle is not a dogfight env");
  if (!peer_tables_dev || world < 1) return fail("pfb_dogfight_physics_peer: need the device array of peer table pointers");
  return df_split_physics(h, actions ? actions : h->buf.setpoint, noise, nullptr, peer_tables_dev, world, slot_offset_floats, peer_flags_dev, rank,
                          epoch, first, do_reset, aviary_index, (cudaStream_t)stream);
}

int pfb_dogfight_combat(PfbHandle h, const float* payload_table, int64_t first_global_agent, int64_t num_arenas, int last, void* stream) {
  REQUIRE_BOUND(h);
  if (!is_df(h)) return fail("handle is not a dogfight env");
  if (require_env(h)) return -1;
  return df_split_combat(h, payload_table, first_global_agent, num_arenas, last, nullptr, 0, 0, (cudaStream_t)stream);
}

int pfb_dogfight_combat_wait(PfbHandle h, const float* payload_table, int64_t first_global_agent, int64_t num_arenas, int last,
                             const int32_t* flags, int world, int epoch, void* stream) {
  REQUIRE_BOUND(h);
  if (!is_df(h)) return fail("handle is not a dogfight env");
  if (require_env(h)) return -1;
  if (!flags) return fail("pfb_dogfight_combat_wait: null flag array");
  return df_split_combat(h, payload_table, first_global_agent, num_arenas, last, flags, world, epoch, (cudaStream_t)stream);
}

// One whole env step of the split dogfight with the fused exchange: env_step_ratio x (physics with peer stores + in-kernel
// signal, combat with in-kernel wait).  Nothing between the kernels needs the host, so the step is ONE call.
int pfb_dogfight_split_step(PfbHandle h, const float* actions, const uint64_t* peer_tables_dev, const uint64_t* peer_flags_dev,
                            const float* local_tables, const int32_t* local_flags, int world, int rank, int epoch0,
                            int64_t first_global_agent, int64_t num_arenas, void* stream) {
  REQUIRE_BOUND(h);
  if (!is_df(h)) return fail("handle is not a dogfight env");
  if (require_env(h)) return -1;
  if (!peer_tables_dev || !peer_flags_dev || !local_tables || !local_flags) return fail("pfb_dogfight_split_step: null argument");
  const int64_t na = 2 * num_arenas;
  const int ratio = h->env.env_step_ratio;
  for (int k = 0; k < ratio; ++k) {
    const int epoch = epoch0 + k;          // exchange number, 1-based; its parity selects the half of the double-buffered table
    const int phase = (epoch - 1) & 1;
    if (df_split_physics(h, actions ? actions : h->buf.setpoint, nullptr, nullptr, peer_tables_dev, world, (phase * na + first_global_agent) * 20,
                         peer_flags_dev, rank, epoch, k == 0, 0, k, (cudaStream_t)stream))
      return -1;
    if (df_split_combat(h, local_tables + phase * na * 20, first_global_agent, num_arenas, k == ratio - 1, local_flags, world, epoch,
                        (cudaStream_t)stream))
      return -1;
  }
  return 0;
}

int64_t pfb_launch_count(PfbHandle h) { return h ? h->launches : 0; }

int pfb_profile_begin(PfbHandle h, int capacity) {
  if (!h) return fail("null handle");
  CUDA_OK(cudaSetDevice(h->device));
  if (h->prof_ev) {
    for (int i = 0; i < 2 * h->prof_cap; ++i) cudaEventDestroy(h->prof_ev[i]);
    delete[] h->prof_ev;
    h->prof_ev = nullptr;
  }
  h->prof_cap = 0;
  h->prof_n = 0;
  if (capacity <= 0) return 0;
  h->prof_ev = new (std::nothrow) cudaEvent_t[2 * (size_t)capacity];
  if (!h->prof_ev) return fail("out of host memory");
  for (int i = 0; i < 2 * capacity; ++i) CUDA_OK(cudaEventCreate(&h->prof_ev[i]));
  h->prof_cap = capacity;
  return 0;
}

int pfb_profile_read(PfbHandle h, float* ms_out, int capacity) {
  if (!h) return fail("null handle");
  CUDA_OK(cudaSetDevice(h->device));
  int n = h->prof_n < capacity ? h->prof_n : capacity;
  for (int i = 0; i < n; ++i) CUDA_OK(cudaEventElapsedTime(&ms_out[i], h->prof_ev[2 * i], h->prof_ev[2 * i + 1]));
  return n;
}

}  // extern "C"
