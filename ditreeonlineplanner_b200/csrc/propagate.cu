// propagate.cu -- fused bicycle-model propagation + goal test + two-ball grid collision.
//
// One thread per candidate edge; the 6-float state lives in registers across all S Euler steps
// (car_env.py:356-396), the occupancy grid sits in shared memory (staged by one bulk TMA copy),
// actions are read and the trajectory written through strides so both the reference's
// array-of-structs layouts and coalesced struct-of-arrays layouts are served by one kernel.
//
// Arithmetic: dynamics in fp32 (north-star tolerance 1e-4 relative vs the reference's float64),
// collision and goal flags in float64 on the fp32 state so they are bit-exact vs NumPy given
// the same states.
#include "common.cuh"

#define PROP_THREADS 128

struct PropArgs {
  const float* state0;
  int64_t s_cand, s_comp;
  const float* actions;
  int64_t a_cand, a_step, a_comp;
  int64_t B;
  int S;
  float goal_x, goal_y;
  float* traj;
  int64_t t_cand, t_step, t_comp;
  float* state_out;
  int32_t* first_coll;
  int32_t* done_step;
  int flags;
};

__device__ __forceinline__ void bicycle_euler(float& x, float& y, float& psi, float& v, float& D, float& dl, float u0,
                                              float u1) {
  // clip to the action space (car_env.py:371; bounds car_env.py:594-597)
  u0 = fminf(fmaxf(u0, -10.0f), 10.0f);
  u1 = fminf(fmaxf(u1, -2.0f), 2.0f);
  const float fxd = (0.28f - 0.05f * v) * D - 0.006f * (v * v) - 0.011f * tanhf(5.0f * v);
  float sn, cs;
  sincosf(psi + 0.5f * dl, &sn, &cs);
  const float dx = v * cs, dy = v * sn;
  const float dpsi = v * 15.5f * dl;
  const float dv = (fxd / 0.043f) * cosf(0.5f * dl);
  const float dt = 0.02f;
  x += dt * dx;
  y += dt * dy;
  psi += dt * dpsi;
  v += dt * dv;
  D += dt * u0;
  dl += dt * u1;
}

template <bool kSoAActions>
__global__ void __launch_bounds__(PROP_THREADS)
k_propagate_collide(MapView m, PropArgs a, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  uint16_t* s_nbr = reinterpret_cast<uint16_t*>(s_map + m.bytes);
  dt_stage_map(s_map, &bar, m);
  dt_build_nbr(s_map, s_nbr, m.rows, m.cols);
  __syncthreads();
  const bool stop = (a.flags & DT_PROP_STOP_ON_COLLISION) != 0;
  const double gx = (double)a.goal_x, gy = (double)a.goal_y;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < a.B; b += (int64_t)gridDim.x * blockDim.x) {
    const float* s0 = a.state0 + b * a.s_cand;
    float x = s0[0], y = s0[a.s_comp], psi = s0[2 * a.s_comp], v = s0[3 * a.s_comp], D = s0[4 * a.s_comp],
          dl = s0[5 * a.s_comp];
    const float* act = a.actions + b * a.a_cand;
    float* tr = a.traj ? a.traj + b * a.t_cand : nullptr;
    int first = -1, done = -1;
    bool alive = true;
    for (int i = 0; i < a.S; ++i) {
      if (alive) {
        float u0, u1;
        if (kSoAActions) {
          u0 = act[i * a.a_step];
          u1 = act[i * a.a_step + a.a_comp];
        } else {  // (.., S, 2) rows: one 8-byte load
          const float2 u = *reinterpret_cast<const float2*>(act + i * a.a_step);
          u0 = u.x;
          u1 = u.y;
        }
        bicycle_euler(x, y, psi, v, D, dl, u0, u1);
        // goal test (car_env.py:341-350) and collision (planners/base_planner.py:306) on the new state
        const double ex = xsub((double)x, gx), ey = xsub((double)y, gy);
        // ||p - goal|| < 0.5: compare squares, and take the square root only on the knife edge
        const double d2 = xadd(xmul(ex, ex), xmul(ey, ey));
        const bool in_goal = (fabs(d2 - 0.25) < 1.0e-9) ? (__dsqrt_rn(d2) < 0.5) : (d2 < 0.25);
        const int c = dt_car_test_nbr(s_map, s_nbr, m.rows, m.cols, x, y, psi);
        if (c & 4) atomicMin(status, DT_E_INDEX);
        const bool coll = (c & 1) != 0;
        if (coll && first < 0) first = i;
        if (tr) {
          float* o = tr + i * a.t_step;
          o[0] = x; o[a.t_comp] = y; o[2 * a.t_comp] = psi; o[3 * a.t_comp] = v; o[4 * a.t_comp] = D;
          o[5 * a.t_comp] = dl;
        }
        if (coll && stop) {
          alive = false;          // collision ends the edge, the goal flag is ignored (base_planner.py:306-312)
        } else if (in_goal) {
          done = i;               // goal reached: remaining actions are zeroed, loop breaks (:314-317)
          alive = false;
        }
      } else if (tr) {
        float* o = tr + i * a.t_step;
        o[0] = 0.f; o[a.t_comp] = 0.f; o[2 * a.t_comp] = 0.f; o[3 * a.t_comp] = 0.f; o[4 * a.t_comp] = 0.f;
        o[5 * a.t_comp] = 0.f;
      }
    }
    if (a.state_out) {
      float* so = a.state_out + b * a.s_cand;
      so[0] = x; so[a.s_comp] = y; so[2 * a.s_comp] = psi; so[3 * a.s_comp] = v; so[4 * a.s_comp] = D;
      so[5 * a.s_comp] = dl;
    }
    if (a.first_coll) a.first_coll[b] = first;
    if (a.done_step) a.done_step[b] = done;
  }
}


// ---------------------------------------------------------------------------------------------
// Row-layout specialisation: actions (B, T, 2) rows and trajectory (B, S, 6) rows, the layouts the
// reference hands over.  Per-thread strided accesses (8 B every 512 B, 24 B every 1200 B) waste
// most of every 32-byte sector, so both streams are staged through shared memory in chunks of
// PROP_CH steps: the block loads / stores whole contiguous segments (64 B of actions, 192 B of
// trajectory per candidate and chunk) with consecutive lanes on consecutive addresses, and the
// per-thread accesses hit conflict-free padded shared rows.
// ---------------------------------------------------------------------------------------------
#define PROP_CH 8
#define PROP_APITCH (PROP_CH * 2 + 1)
#define PROP_TPITCH (PROP_CH * 6 + 1)

__global__ void __launch_bounds__(PROP_THREADS)
k_propagate_rows(MapView m, PropArgs a, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ uint64_t bar;
  uint8_t* s_map = s_dyn;
  uint16_t* s_nbr = reinterpret_cast<uint16_t*>(s_dyn + m.bytes);
  float* s_act = reinterpret_cast<float*>(s_dyn + 3 * m.bytes);
  float* s_trj = s_act + PROP_THREADS * PROP_APITCH;
  dt_stage_map(s_map, &bar, m);
  dt_build_nbr(s_map, s_nbr, m.rows, m.cols);
  __syncthreads();
  const bool stop = (a.flags & DT_PROP_STOP_ON_COLLISION) != 0;
  const double gx = (double)a.goal_x, gy = (double)a.goal_y;
  const int tid = threadIdx.x;
  for (int64_t b0 = (int64_t)blockIdx.x * PROP_THREADS; b0 < a.B; b0 += (int64_t)gridDim.x * PROP_THREADS) {
    const int64_t b = b0 + tid;
    const bool live = b < a.B;
    const int nb = (int)((a.B - b0 < PROP_THREADS) ? (a.B - b0) : PROP_THREADS);
    float x = 0.f, y = 0.f, psi = 0.f, v = 0.f, D = 0.f, dl = 0.f;
    if (live) {
      const float* s0 = a.state0 + b * a.s_cand;
      x = s0[0]; y = s0[a.s_comp]; psi = s0[2 * a.s_comp]; v = s0[3 * a.s_comp]; D = s0[4 * a.s_comp];
      dl = s0[5 * a.s_comp];
    }
    int first = -1, done = -1;
    bool alive = live;
    for (int c0 = 0; c0 < a.S; c0 += PROP_CH) {
      const int cs = (a.S - c0 < PROP_CH) ? (a.S - c0) : PROP_CH;  // steps in this chunk
      // coalesced load of the chunk's actions: per candidate 2*cs contiguous floats
      if (cs == PROP_CH) {  // full chunk: compile-time divisors
        constexpr int na = PROP_CH * 2;
        for (int e = tid; e < nb * na; e += PROP_THREADS) {
          const int c = e / na, off = e % na;
          s_act[c * PROP_APITCH + off] = __ldg(a.actions + (b0 + c) * a.a_cand + (int64_t)c0 * 2 + off);
        }
      } else {
        const int na = cs * 2;
        for (int e = tid; e < nb * na; e += PROP_THREADS) {
          const int c = e / na, off = e - c * na;
          s_act[c * PROP_APITCH + off] = __ldg(a.actions + (b0 + c) * a.a_cand + (int64_t)c0 * 2 + off);
        }
      }
      __syncthreads();
      for (int i = 0; i < cs; ++i) {
        float* o = s_trj + tid * PROP_TPITCH + i * 6;
        if (alive) {
          bicycle_euler(x, y, psi, v, D, dl, s_act[tid * PROP_APITCH + 2 * i], s_act[tid * PROP_APITCH + 2 * i + 1]);
          const double ex = xsub((double)x, gx), ey = xsub((double)y, gy);
          const double d2 = xadd(xmul(ex, ex), xmul(ey, ey));
          const bool in_goal = (fabs(d2 - 0.25) < 1.0e-9) ? (__dsqrt_rn(d2) < 0.5) : (d2 < 0.25);
          const int c = dt_car_test_nbr(s_map, s_nbr, m.rows, m.cols, x, y, psi);
          if (c & 4) atomicMin(status, DT_E_INDEX);
          const bool coll = (c & 1) != 0;
          if (coll && first < 0) first = c0 + i;
          o[0] = x; o[1] = y; o[2] = psi; o[3] = v; o[4] = D; o[5] = dl;
          if (coll && stop) {
            alive = false;
          } else if (in_goal) {
            done = c0 + i;
            alive = false;
          }
        } else {
          o[0] = 0.f; o[1] = 0.f; o[2] = 0.f; o[3] = 0.f; o[4] = 0.f; o[5] = 0.f;
        }
      }
      __syncthreads();
      // coalesced store of the chunk's trajectory rows: per candidate 6*cs contiguous floats
      if (a.traj) {
        if (cs == PROP_CH) {
          constexpr int nt = PROP_CH * 6;
          for (int e = tid; e < nb * nt; e += PROP_THREADS) {
            const int c = e / nt, off = e % nt;
            a.traj[(b0 + c) * a.t_cand + (int64_t)c0 * 6 + off] = s_trj[c * PROP_TPITCH + off];
          }
        } else {
          const int nt = cs * 6;
          for (int e = tid; e < nb * nt; e += PROP_THREADS) {
            const int c = e / nt, off = e - c * nt;
            a.traj[(b0 + c) * a.t_cand + (int64_t)c0 * 6 + off] = s_trj[c * PROP_TPITCH + off];
          }
        }
      }
      // (the next chunk's first __syncthreads orders these reads before the staging rows are rewritten)
    }
    if (live) {
      if (a.state_out) {
        float* so = a.state_out + b * a.s_cand;
        so[0] = x; so[a.s_comp] = y; so[2 * a.s_comp] = psi; so[3 * a.s_comp] = v; so[4 * a.s_comp] = D;
        so[5 * a.s_comp] = dl;
      }
      if (a.first_coll) a.first_coll[b] = first;
      if (a.done_step) a.done_step[b] = done;
    }
    __syncthreads();
  }
}

extern "C" int dt_propagate_collide(dt_ctx* ctx, const float* state0, int64_t s_cand, int64_t s_comp,
                                    const float* actions, int64_t a_cand, int64_t a_step, int64_t a_comp, int64_t B,
                                    int S, float goal_x, float goal_y, float* traj_out, int64_t t_cand, int64_t t_step,
                                    int64_t t_comp, float* state_out, int32_t* first_coll, int32_t* done_step,
                                    int flags, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!ctx->d_map) return dt_fail(ctx, DT_E_NOMAP, "dt_set_map has not been called");
  if (B <= 0) return DT_OK;
  if (!state0 || !actions || S < 0) return dt_fail(ctx, DT_E_ARG, "dt_propagate_collide: bad argument");
  PropArgs a;
  a.state0 = state0; a.s_cand = s_cand; a.s_comp = s_comp;
  a.actions = actions; a.a_cand = a_cand; a.a_step = a_step; a.a_comp = a_comp;
  a.B = B; a.S = S; a.goal_x = goal_x; a.goal_y = goal_y;
  a.traj = traj_out; a.t_cand = t_cand; a.t_step = t_step; a.t_comp = t_comp;
  a.state_out = state_out; a.first_coll = first_coll; a.done_step = done_step; a.flags = flags;
  MapView m = dt_map_view(ctx);
  int64_t blocks = (B + PROP_THREADS - 1) / PROP_THREADS;
  const int64_t cap = (int64_t)ctx->sm_count * 16;  // grid-stride over a whole number of waves
  if (blocks > cap) blocks = cap;
  cudaStream_t st = (cudaStream_t)stream;
  // (.., S, 2) action rows that are 8-byte aligned take the vector-load path
  const bool rows2 = (a_comp == 1) && (a_step % 2 == 0) && (a_cand % 2 == 0) && (((uintptr_t)actions & 7) == 0);
  const bool traj_rows = !traj_out || (t_comp == 1 && t_step == 6);
  if (a_comp == 1 && a_step == 2 && traj_rows) {
    const size_t smem = (size_t)3 * m.bytes + (size_t)PROP_THREADS * (PROP_APITCH + PROP_TPITCH) * sizeof(float);
    static bool attr_set = false;
    if (!attr_set) {
      DT_CUDA(cudaFuncSetAttribute(k_propagate_rows, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      attr_set = true;
    }
    const int64_t cap2 = (int64_t)ctx->sm_count * 6;  // 6 resident blocks per SM: a whole number of waves
    int64_t blocks2 = (B + PROP_THREADS - 1) / PROP_THREADS;
    if (blocks2 > cap2) blocks2 = cap2;
    k_propagate_rows<<<(int)blocks2, PROP_THREADS, smem, st>>>(m, a, ctx->d_status);
  } else if (rows2) {
    k_propagate_collide<false><<<(int)blocks, PROP_THREADS, 3 * m.bytes, st>>>(m, a, ctx->d_status);
  } else {
    k_propagate_collide<true><<<(int)blocks, PROP_THREADS, 3 * m.bytes, st>>>(m, a, ctx->d_status);
  }
  DT_LAUNCH_CHECK("k_propagate_collide");
  return DT_OK;
}
