// propagate.cu -- fused bicycle-model propagation + goal test + two-ball grid collision.
//
// One thread per candidate edge; the 6-float state (plus the sine / cosine of its heading) lives in
// registers across all S Euler steps (car_env.py:356-396); the occupancy grid and the quadrant map of the
// collision fast path (carfast.cuh) sit in shared memory, staged by bulk TMA copies.
//
// Arithmetic: dynamics in fp32 (north-star tolerance 1e-4 relative vs the reference's float64; MUFU
// sin / cos / ex2 / rcp, measured error < 1e-5 over 50 steps), collision and goal flags decided exactly as
// the float64 reference decides them on the fp32 state (guard-banded fp32 fast paths that defer to the
// float64 code near a threshold), so flags are bit-exact vs NumPy given the same states.
//
// Two kernels share the step function, so every layout yields the same bits:
//   k_propagate<false>  generic element strides (coalesced for struct-of-arrays buffers)
//   k_propagate<true>   the reference's row layouts -- actions (B, T, 2), trajectory (B, S, 6) -- staged
//                       through warp-private shared memory in chunks of PROP_CH steps, so global memory
//                       sees whole 32-byte sectors (per-thread row accesses would touch 8 / 24 bytes of
//                       each 512 / 1200-byte row per step)
#include "carfast.cuh"

#define PROP_THREADS 128
#define PROP_WARPS (PROP_THREADS / 32)

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

struct Car {
  float x, y, psi, v, D, dl;
  float sn, cs;  // sin / cos of psi (shared by this step's collision test and the next step's dynamics)
};

__device__ __forceinline__ float rcp_approx(float v) {
  float r;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__device__ __forceinline__ float ex2_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}

// CarEnv._update_state (car_env.py:356-396), explicit Euler with dt = 0.02
__device__ __forceinline__ void car_step(Car& c, float u0, float u1) {
  // clip to the action space (car_env.py:371; bounds car_env.py:594-597)
  u0 = fminf(fmaxf(u0, -10.0f), 10.0f);
  u1 = fminf(fmaxf(u1, -2.0f), 2.0f);
  // tanh(5 v) = 1 - 2 / (exp(10 v) + 1): saturates correctly at +-inf, absolute error ~1e-7
  const float th5v = __fmaf_rn(-2.0f, rcp_approx(ex2_approx(14.426950f * c.v) + 1.0f), 1.0f);
  const float fxd = (0.28f - 0.05f * c.v) * c.D - 0.006f * (c.v * c.v) - 0.011f * th5v;
  const float hd = 0.5f * c.dl;
  const float sh = __sinf(hd), ch = __cosf(hd);
  // cos / sin (psi + delta / 2) by angle addition
  const float cb = c.cs * ch - c.sn * sh, sb = c.sn * ch + c.cs * sh;
  const float dt = 0.02f;
  const float dpsi = c.v * 15.5f * c.dl;
  const float dv = (fxd * (1.0f / 0.043f)) * ch;
  c.x += dt * (c.v * cb);
  c.y += dt * (c.v * sb);
  c.psi += dt * dpsi;
  c.v += dt * dv;
  c.D += dt * u0;
  c.dl += dt * u1;
  dt_sincos_fast(c.psi, c.sn, c.cs);
}

// ||p - goal|| < 0.5 (car_env.py:341-350) decided exactly as the float64 reference does: fp32 squared
// distance when it is clear of 0.25 by more than its rounding error, else the float64 expression (and
// on the knife edge the square root itself).
__device__ __forceinline__ bool goal_test(float x, float y, float gxf, float gyf) {
  const float fx = x - gxf, fy = y - gyf;
  const float f2 = fx * fx + fy * fy;
  if (fabsf(f2 - 0.25f) > 1.0e-4f * fmaxf(1.0f, f2)) return f2 < 0.25f;
  const double ex = xsub((double)x, (double)gxf), ey = xsub((double)y, (double)gyf);
  const double d2 = xadd(xmul(ex, ex), xmul(ey, ey));
  return (fabs(d2 - 0.25) < 1.0e-9) ? (__dsqrt_rn(d2) < 0.5) : (d2 < 0.25);
}

struct EdgeState {
  int first, done;
  bool alive;
};

// one step of BasePlanner.propagate_action_sequence_env (planners/base_planner.py:281-317) for a live edge
__device__ __forceinline__ void edge_step(Car& c, EdgeState& e, int i, float u0, float u1, bool stop,
                                          const uint8_t* s_map, uint32_t s_q, const MapView& m,
                                          const QMapView& q, float gx, float gy, int* status) {
  car_step(c, u0, u1);
  // goal test (car_env.py:341-350) and collision (planners/base_planner.py:306) on the new state
  const bool in_goal = goal_test(c.x, c.y, gx, gy);
  const int hit = dt_car_fast(s_map, s_q, q, m.rows, m.cols, c.x, c.y, c.psi, c.sn, c.cs);
  if (hit & 4) atomicMin(status, DT_E_INDEX);
  const bool coll = (hit & 1) != 0;
  if (coll && e.first < 0) e.first = i;
  if (coll && stop) {
    e.alive = false;  // collision ends the edge, the goal flag is ignored (base_planner.py:306-312)
  } else if (in_goal) {
    e.done = i;       // goal reached: remaining actions are zeroed, loop breaks (:314-317)
    e.alive = false;
  }
}

#define PROP_CH 4                        // steps per staged chunk
#define PROP_APITCH (PROP_CH * 2 + 4)    // 12 words: float4 rows, conflict-free for per-lane 16-byte accesses
#define PROP_TPITCH (PROP_CH * 6 + 4)    // 28 words
#define PROP_STAGE_WORDS (32 * (PROP_APITCH + PROP_TPITCH))

template <bool kRows>
__global__ void __launch_bounds__(PROP_THREADS)
k_propagate(MapView m, QMapView q, PropArgs a, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ uint64_t bar;
  uint8_t* s_map = s_dyn;
  uint16_t* s_qp = reinterpret_cast<uint16_t*>(s_dyn + m.bytes);
  dt_stage_maps(s_map, s_qp, &bar, m, q);
  const uint32_t s_q = dt_qmap_addr(s_qp, q);
  const bool stop = (a.flags & DT_PROP_STOP_ON_COLLISION) != 0;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;

  if (!kRows) {
    for (int64_t b = blockIdx.x * (int64_t)PROP_THREADS + threadIdx.x; b < a.B; b += (int64_t)gridDim.x * PROP_THREADS) {
      const float* s0 = a.state0 + b * a.s_cand;
      Car c;
      c.x = s0[0]; c.y = s0[a.s_comp]; c.psi = s0[2 * a.s_comp]; c.v = s0[3 * a.s_comp]; c.D = s0[4 * a.s_comp];
      c.dl = s0[5 * a.s_comp];
      dt_sincos_fast(c.psi, c.sn, c.cs);
      const float* act = a.actions + b * a.a_cand;
      float* tr = a.traj ? a.traj + b * a.t_cand : nullptr;
      EdgeState e = {-1, -1, true};
      for (int i = 0; i < a.S; ++i) {
        if (e.alive) {
          const float u0 = __ldg(act + i * a.a_step), u1 = __ldg(act + i * a.a_step + a.a_comp);
          edge_step(c, e, i, u0, u1, stop, s_map, s_q, m, q, a.goal_x, a.goal_y, status);
          if (tr) {
            float* o = tr + i * a.t_step;
            o[0] = c.x; o[a.t_comp] = c.y; o[2 * a.t_comp] = c.psi; o[3 * a.t_comp] = c.v; o[4 * a.t_comp] = c.D;
            o[5 * a.t_comp] = c.dl;
          }
        } else if (tr) {
          float* o = tr + i * a.t_step;
          o[0] = 0.f; o[a.t_comp] = 0.f; o[2 * a.t_comp] = 0.f; o[3 * a.t_comp] = 0.f; o[4 * a.t_comp] = 0.f;
          o[5 * a.t_comp] = 0.f;
        }
      }
      if (a.state_out) {
        float* so = a.state_out + b * a.s_cand;
        so[0] = c.x; so[a.s_comp] = c.y; so[2 * a.s_comp] = c.psi; so[3 * a.s_comp] = c.v; so[4 * a.s_comp] = c.D;
        so[5 * a.s_comp] = c.dl;
      }
      if (a.first_coll) a.first_coll[b] = e.first;
      if (a.done_step) a.done_step[b] = e.done;
    }
    return;
  }

  // ---- row layouts: each warp owns 32 consecutive candidates and a private staging area ----
  float* s_act = reinterpret_cast<float*>(s_dyn + m.bytes + q.bytes) + (size_t)warp * PROP_STAGE_WORDS;
  float* s_trj = s_act + 32 * PROP_APITCH;
  // copy mapping: lane -> (candidate c8 within a group of 8, 16-byte column q4); a quarter-warp touches 8
  // different candidates, whose rows are 3 (resp. 7) 16-byte bank groups apart modulo 8: conflict-free
  const int c8 = lane & 7, q4 = lane >> 3;
  const int64_t nwarps = (int64_t)gridDim.x * PROP_WARPS;
  for (int64_t b0 = ((int64_t)blockIdx.x * PROP_WARPS + warp) * 32; b0 < a.B; b0 += nwarps * 32) {
    const int64_t b = b0 + lane;
    const bool live = b < a.B;
    const int nb = (int)((a.B - b0 < 32) ? (a.B - b0) : 32);
    Car c = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
    if (live) {
      const float* s0 = a.state0 + b * a.s_cand;
      c.x = s0[0]; c.y = s0[a.s_comp]; c.psi = s0[2 * a.s_comp]; c.v = s0[3 * a.s_comp]; c.D = s0[4 * a.s_comp];
      c.dl = s0[5 * a.s_comp];
      dt_sincos_fast(c.psi, c.sn, c.cs);
    }
    EdgeState e = {-1, -1, live};
    const int full = a.S / PROP_CH;  // whole chunks
    // prefetch of the next chunk's actions: 32 candidates x 2 float4 = 64 float4, two per lane
    // (lane -> candidate lane & 7 (+8, +16, +24 over two loads of two column halves))
    float4 pf[2];
    auto prefetch = [&](int chunk) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int cand = c8 + 8 * (q4 >> 1) + 16 * g, col = q4 & 1;
        pf[g] = make_float4(0.f, 0.f, 0.f, 0.f);
        if (cand < nb)
          pf[g] = __ldg(reinterpret_cast<const float4*>(a.actions + (b0 + cand) * a.a_cand + (int64_t)chunk * (PROP_CH * 2)) + col);
      }
    };
    if (full > 0) prefetch(0);
    for (int ch = 0; ch < full; ++ch) {
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int cand = c8 + 8 * (q4 >> 1) + 16 * g, col = q4 & 1;
        *reinterpret_cast<float4*>(s_act + cand * PROP_APITCH + 4 * col) = pf[g];
      }
      __syncwarp();
      if (ch + 1 < full) prefetch(ch + 1);
      const float4 a01 = *reinterpret_cast<const float4*>(s_act + lane * PROP_APITCH);
      const float4 a23 = *reinterpret_cast<const float4*>(s_act + lane * PROP_APITCH + 4);
      const float us[PROP_CH * 2] = {a01.x, a01.y, a01.z, a01.w, a23.x, a23.y, a23.z, a23.w};
      float o[PROP_CH * 6];
#pragma unroll
      for (int i = 0; i < PROP_CH; ++i) {
        if (e.alive) {
          edge_step(c, e, ch * PROP_CH + i, us[2 * i], us[2 * i + 1], stop, s_map, s_q, m, q, a.goal_x, a.goal_y, status);
          o[6 * i] = c.x; o[6 * i + 1] = c.y; o[6 * i + 2] = c.psi; o[6 * i + 3] = c.v; o[6 * i + 4] = c.D;
          o[6 * i + 5] = c.dl;
        } else {
          o[6 * i] = 0.f; o[6 * i + 1] = 0.f; o[6 * i + 2] = 0.f; o[6 * i + 3] = 0.f; o[6 * i + 4] = 0.f;
          o[6 * i + 5] = 0.f;
        }
      }
      if (a.traj) {
#pragma unroll
        for (int j = 0; j < PROP_CH * 6 / 4; ++j)
          *reinterpret_cast<float4*>(s_trj + lane * PROP_TPITCH + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
        __syncwarp();
        // 32 candidates x 6 float4: per pass 8 candidates x 4 columns; 4 candidate groups x 2 column passes
        // (columns 0-3, then 4-5 on the lower half-warp's column lanes)
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          const int cand = c8 + 8 * g;
          if (cand < nb) {
            float* dst = a.traj + (b0 + cand) * a.t_cand + (int64_t)ch * (PROP_CH * 6);
            const float* src = s_trj + cand * PROP_TPITCH;
            *(reinterpret_cast<float4*>(dst) + q4) = *reinterpret_cast<const float4*>(src + 4 * q4);
            if (q4 < 2) *(reinterpret_cast<float4*>(dst) + 4 + q4) = *reinterpret_cast<const float4*>(src + 16 + 4 * q4);
          }
        }
      }
      __syncwarp();
    }
    // tail steps (S not a multiple of PROP_CH): per-lane row accesses
    for (int i = full * PROP_CH; i < a.S; ++i) {
      if (live) {
        float* o = a.traj ? a.traj + b * a.t_cand + (int64_t)i * 6 : nullptr;
        if (e.alive) {
          const float2 u = __ldg(reinterpret_cast<const float2*>(a.actions + b * a.a_cand + (int64_t)i * 2));
          edge_step(c, e, i, u.x, u.y, stop, s_map, s_q, m, q, a.goal_x, a.goal_y, status);
          if (o) {
            *reinterpret_cast<float2*>(o) = make_float2(c.x, c.y);
            *reinterpret_cast<float2*>(o + 2) = make_float2(c.psi, c.v);
            *reinterpret_cast<float2*>(o + 4) = make_float2(c.D, c.dl);
          }
        } else if (o) {
          *reinterpret_cast<float2*>(o) = make_float2(0.f, 0.f);
          *reinterpret_cast<float2*>(o + 2) = make_float2(0.f, 0.f);
          *reinterpret_cast<float2*>(o + 4) = make_float2(0.f, 0.f);
        }
      }
    }
    if (live) {
      if (a.state_out) {
        float* so = a.state_out + b * a.s_cand;
        so[0] = c.x; so[a.s_comp] = c.y; so[2 * a.s_comp] = c.psi; so[3 * a.s_comp] = c.v; so[4 * a.s_comp] = c.D;
        so[5 * a.s_comp] = c.dl;
      }
      if (a.first_coll) a.first_coll[b] = e.first;
      if (a.done_step) a.done_step[b] = e.done;
    }
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
  const MapView m = dt_map_view(ctx);
  const QMapView q = dt_qmap_view(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  const size_t map_smem = (size_t)m.bytes + (size_t)q.bytes;
  // the reference's row layouts, 16-byte aligned: staged kernel
  const bool rows = (a_comp == 1) && (a_step == 2) && (a_cand % 4 == 0) && (((uintptr_t)actions & 15) == 0) &&
                    (!traj_out || (t_comp == 1 && t_step == 6 && t_cand % 4 == 0 && ((uintptr_t)traj_out & 15) == 0));
  if (!ctx->prop_attr_set) {
    DT_CUDA(cudaFuncSetAttribute(k_propagate<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    DT_CUDA(cudaFuncSetAttribute(k_propagate<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
    ctx->prop_attr_set = true;
  }
  if (rows) {
    const size_t smem = map_smem + (size_t)PROP_WARPS * PROP_STAGE_WORDS * sizeof(float);
    int per_sm = 0;
    DT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_propagate<true>, PROP_THREADS, smem));
    if (per_sm < 1) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_propagate_collide: map too large for shared memory");
    int64_t blocks = (B + PROP_THREADS - 1) / PROP_THREADS;
    const int64_t cap = (int64_t)ctx->sm_count * per_sm;  // persistent: one wave of resident blocks
    if (blocks > cap) blocks = cap;
    k_propagate<true><<<(int)blocks, PROP_THREADS, smem, st>>>(m, q, a, ctx->d_status);
  } else {
    int per_sm = 0;
    DT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, k_propagate<false>, PROP_THREADS, map_smem));
    if (per_sm < 1) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_propagate_collide: map too large for shared memory");
    int64_t blocks = (B + PROP_THREADS - 1) / PROP_THREADS;
    const int64_t cap = (int64_t)ctx->sm_count * per_sm;
    if (blocks > cap) blocks = cap;
    k_propagate<false><<<(int)blocks, PROP_THREADS, map_smem, st>>>(m, q, a, ctx->d_status);
  }
  DT_LAUNCH_CHECK("k_propagate");
  return DT_OK;
}
