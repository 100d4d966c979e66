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
// Two kernels share the step function, so every layout yields the same bits (template flag: the collision
// fast path's table exists for this map):
//   k_propagate_strided  generic element strides (coalesced for struct-of-arrays buffers)
//   k_propagate_rows     the reference's row layouts -- state (B, 6), actions (B, T, 2), trajectory (B, S, 6)
//                        -- staged through warp-private shared memory in chunks of PROP_CH steps, so global
//                        memory sees whole 32-byte sectors (per-thread row accesses would touch 8 / 24
//                        bytes of each 512 / 1200-byte row per step)
#include "carprop.cuh"

#define PROP_THREADS 128
#ifndef PROP_PF
#define PROP_PF 2     // action prefetch distance of the strided kernel, steps (register ring)
#endif
// trajectory stores: plain write-back stores measured 2.5 % faster than streaming (st.global.cs) ones
#ifdef PROP_STREAMING_ST
#define PROP_ST(p, v) __stcs((p), (v))
#else
#define PROP_ST(p, v) (*(p) = (v))
#endif

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

// ---------------------------------------------------------------------------------------------
// Generic element strides (coalesced for struct-of-arrays buffers); next step's action prefetched
// ---------------------------------------------------------------------------------------------
#ifndef PROP_SMINB
#define PROP_SMINB 5
#endif
template <bool kTable, bool kStop>
__global__ void __launch_bounds__(PROP_THREADS, PROP_SMINB)
k_propagate_strided(MapView m, QMapView q, PropArgs a, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ uint64_t bar;
  uint8_t* s_map = s_dyn;
  uint32_t* s_qp = reinterpret_cast<uint32_t*>(s_dyn + m.bytes);
  dt_stage_maps(s_map, s_qp, &bar, m, q);
  const uint32_t s_q = dt_qmap_addr(s_qp, q);
  for (int64_t b = blockIdx.x * (int64_t)PROP_THREADS + threadIdx.x; b < a.B; b += (int64_t)gridDim.x * PROP_THREADS) {
    const float* s0 = a.state0 + b * a.s_cand;
    Car c;
    c.x = s0[0]; c.y = s0[a.s_comp]; c.psi = s0[2 * a.s_comp]; c.v = s0[3 * a.s_comp]; c.D = s0[4 * a.s_comp];
    c.dl = s0[5 * a.s_comp];
    dt_sincos_fast(c.psi, c.sn, c.cs);
    const float* act = a.actions + b * a.a_cand;
    float* tr = a.traj ? a.traj + b * a.t_cand : nullptr;
    EdgeState e = {-1, -1, 1};
    // actions are fetched PROP_PF steps ahead (register ring): under the trajectory's write traffic a
    // global load takes ~1 us, several steps of compute
    float n0[PROP_PF], n1[PROP_PF];
#pragma unroll
    for (int k = 0; k < PROP_PF; ++k) {
      n0[k] = 0.f; n1[k] = 0.f;
      if (k < a.S) { n0[k] = __ldg(act + k * a.a_step); n1[k] = __ldg(act + k * a.a_step + a.a_comp); }
    }
    for (int i0 = 0; i0 < a.S; i0 += PROP_PF) {
#pragma unroll
    for (int k = 0; k < PROP_PF; ++k) {
      const int i = i0 + k;
      if (i >= a.S) break;
      const float u0 = n0[k], u1 = n1[k];
      if (i + PROP_PF < a.S) {
        n0[k] = __ldg(act + (i + PROP_PF) * a.a_step);
        n1[k] = __ldg(act + (i + PROP_PF) * a.a_step + a.a_comp);
      }
      // one store per component for the whole warp (live lanes: the new state, finished lanes: zeros), so
      // a 128-byte line of a struct-of-arrays trajectory is written once, never as two partial writes
      const bool was_alive = e.alive != 0;
      if (was_alive) edge_step<kTable, kStop>(c, e, i, u0, u1, s_map, s_q, m, q, a.goal_x, a.goal_y, status);
      if (tr) {
        float* o = tr + i * a.t_step;
        PROP_ST(o, was_alive ? c.x : 0.f);
        PROP_ST(o + a.t_comp, was_alive ? c.y : 0.f);
        PROP_ST(o + 2 * a.t_comp, was_alive ? c.psi : 0.f);
        PROP_ST(o + 3 * a.t_comp, was_alive ? c.v : 0.f);
        PROP_ST(o + 4 * a.t_comp, was_alive ? c.D : 0.f);
        PROP_ST(o + 5 * a.t_comp, was_alive ? c.dl : 0.f);
      }
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
}

// ---------------------------------------------------------------------------------------------
// Row layouts: state (B, 6), actions (B, T, 2), trajectory (B, S, 6), all rows 16-byte aligned.
// Each warp owns tasks of 32 consecutive candidates and walks them in chunks of PROP_CH steps.  A chunk's
// actions (32 B per candidate) arrive in a double-buffered warp-private staging area by cp.async, issued
// one chunk ahead -- across task boundaries, together with the next task's start states -- so no global
// load latency is exposed; a chunk's trajectory (96 B per candidate) is written to staging rows by the
// owning lanes and copied out by the whole warp as 64-byte / 32-byte row segments.
// Staging rows are padded to 12 / 28 words: per-lane 16-byte accesses (row stride 3 resp. 7 bank groups
// modulo 8) and the copy mapping (a quarter-warp = 8 different candidates) are both conflict-free.
// ---------------------------------------------------------------------------------------------
#define PROP_CH 4                        // steps per chunk
#define PROP_APITCH (PROP_CH * 2 + 4)    // words
#define PROP_TPITCH (PROP_CH * 6 + 4)
#define PROP_SPITCH 6                    // start states: 24-byte rows, as in global memory
#define PROP_WARP_WORDS (32 * (2 * PROP_APITCH + PROP_TPITCH + PROP_SPITCH))
// 4 warps x 4 blocks per SM at <= 128 registers: no spills (a spilled value costs a local load that misses the
// store-churned L1); more, smaller-register warps measured slower (tools/build_variant.sh)
#ifndef PROP_RTHREADS
#define PROP_RTHREADS 128
#endif
#ifndef PROP_RMINB
#define PROP_RMINB 4
#endif
#define PROP_RWARPS (PROP_RTHREADS / 32)

__device__ __forceinline__ void cp_async16(uint32_t dst, const void* src) {
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async8(uint32_t dst, const void* src) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() { asm volatile("cp.async.wait_all;" ::: "memory"); }

template <bool kTable, bool kStop>
__global__ void __launch_bounds__(PROP_RTHREADS, PROP_RMINB)
k_propagate_rows(MapView m, QMapView q, PropArgs a, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ uint64_t bar;
  uint8_t* s_map = s_dyn;
  uint32_t* s_qp = reinterpret_cast<uint32_t*>(s_dyn + m.bytes);
  dt_stage_maps(s_map, s_qp, &bar, m, q);
  const uint32_t s_q = dt_qmap_addr(s_qp, q);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float* s_warp = reinterpret_cast<float*>(s_dyn + m.bytes + q.bytes) + (size_t)warp * PROP_WARP_WORDS;
  float* s_act = s_warp;                          // 2 buffers of 32 x PROP_APITCH
  float* s_trj = s_act + 2 * 32 * PROP_APITCH;    // 32 x PROP_TPITCH
  float* s_st = s_trj + 32 * PROP_TPITCH;         // 32 x 6
  const uint32_t s_act_u = dt_smem_u32(s_act), s_st_u = dt_smem_u32(s_st);
  // copy mapping: lane -> candidate c8 (+8, +16, +24) and 16-byte column q4
  const int c8 = lane & 7, q4 = lane >> 3;
  // (loop bounds are re-derived from kernel parameters -- constant bank operands -- instead of being kept
  // in registers: a spilled loop invariant costs a local-memory load that misses the store-churned L1)
#define PROP_TSTRIDE (gridDim.x * (PROP_RWARPS * 32u))
  const uint32_t Bn = (uint32_t)a.B;  // the launcher sends B >= 2^31 to the strided kernel

  // asynchronous copy of chunk `ch` of the task starting at candidate t0 into action buffer `buf`
  // (plus that task's start states when ch == 0)
  auto issue = [&](uint32_t t0, int ch, int buf) {
    const int nb = (int)((Bn - t0 < 32u) ? (Bn - t0) : 32u);
    const int cs = (a.S - ch * PROP_CH < PROP_CH) ? (a.S - ch * PROP_CH) : PROP_CH;
    const uint32_t dst = s_act_u + (uint32_t)buf * (32 * PROP_APITCH * 4);
    if (cs == PROP_CH) {
      // 32 candidates x 2 float4: lane -> candidates c8 + 8 * (q4 >> 1) (+16), column q4 & 1
#pragma unroll
      for (int g = 0; g < 2; ++g) {
        const int cand = c8 + 8 * (q4 >> 1) + 16 * g, col = q4 & 1;
        if (cand < nb)
          cp_async16(dst + (uint32_t)(cand * PROP_APITCH + 4 * col) * 4,
                     a.actions + (int64_t)(t0 + cand) * a.a_cand + ch * (PROP_CH * 2) + 4 * col);
      }
    } else if (lane < nb) {  // ragged last chunk: each lane fetches its own cs steps (never past step S)
      for (int i = 0; i < cs; ++i)
        cp_async8(dst + (uint32_t)(lane * PROP_APITCH + 2 * i) * 4,
                  a.actions + (int64_t)(t0 + lane) * a.a_cand + (ch * PROP_CH + i) * 2);
    }
    if (ch == 0 && lane < nb) {
      const float* src = a.state0 + (int64_t)(t0 + lane) * 6;
#pragma unroll
      for (int k = 0; k < 3; ++k) cp_async8(s_st_u + (uint32_t)(lane * PROP_SPITCH + 2 * k) * 4, src + 2 * k);
    }
  };

  uint32_t t0 = (blockIdx.x * PROP_RWARPS + warp) * 32u;
  if (t0 < Bn && a.S > 0) issue(t0, 0, 0);
  int buf = 0;
  Car c = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 1.f};
  EdgeState e = {-1, -1, 0};
  while (t0 < Bn) {
    const int nb = (int)((Bn - t0 < 32u) ? (Bn - t0) : 32u);
    const bool live = lane < nb;
    for (int ch = 0; ch * PROP_CH < a.S; ++ch) {
      cp_async_wait_all();
      __syncwarp();
      if (ch == 0) {
        if (live) {
          const float2 p0 = *reinterpret_cast<const float2*>(s_st + lane * PROP_SPITCH);
          const float2 p1 = *reinterpret_cast<const float2*>(s_st + lane * PROP_SPITCH + 2);
          const float2 p2 = *reinterpret_cast<const float2*>(s_st + lane * PROP_SPITCH + 4);
          c.x = p0.x; c.y = p0.y; c.psi = p1.x; c.v = p1.y; c.D = p2.x; c.dl = p2.y;
          dt_sincos_fast(c.psi, c.sn, c.cs);
        }
        e.first = -1; e.done = -1; e.alive = live ? 1 : 0;
        __syncwarp();  // start states consumed before the next task's may land
      }
      // next chunk (possibly the next task's first) into the other buffer
      if ((ch + 1) * PROP_CH < a.S) issue(t0, ch + 1, buf ^ 1);
      else if (t0 + PROP_TSTRIDE < Bn && t0 + PROP_TSTRIDE > t0) issue(t0 + PROP_TSTRIDE, 0, buf ^ 1);
      const int cs = (a.S - ch * PROP_CH < PROP_CH) ? (a.S - ch * PROP_CH) : PROP_CH;
      const float* ua = s_act + buf * (32 * PROP_APITCH) + lane * PROP_APITCH;
      float* my = s_trj + lane * PROP_TPITCH;
      float o[PROP_CH * 6];
#pragma unroll
      for (int i = 0; i < PROP_CH; ++i) {
        if (e.alive && i < cs) {
          const float2 u = *reinterpret_cast<const float2*>(ua + 2 * i);
          edge_step<kTable, kStop>(c, e, ch * PROP_CH + i, u.x, u.y, s_map, s_q, m, q, a.goal_x, a.goal_y, status);
          o[6 * i] = c.x; o[6 * i + 1] = c.y; o[6 * i + 2] = c.psi; o[6 * i + 3] = c.v; o[6 * i + 4] = c.D;
          o[6 * i + 5] = c.dl;
        } else {
          o[6 * i] = 0.f; o[6 * i + 1] = 0.f; o[6 * i + 2] = 0.f; o[6 * i + 3] = 0.f; o[6 * i + 4] = 0.f;
          o[6 * i + 5] = 0.f;
        }
        // a float4 of the staging row is stored as soon as it is complete
#pragma unroll
        for (int j = (6 * i) / 4; j < (6 * i + 6) / 4; ++j)
          *reinterpret_cast<float4*>(my + 4 * j) = make_float4(o[4 * j], o[4 * j + 1], o[4 * j + 2], o[4 * j + 3]);
      }
      if (a.traj) {
        __syncwarp();
        float* dst0 = a.traj + (int64_t)t0 * a.t_cand + ch * (PROP_CH * 6);
        if (cs == PROP_CH) {
          // 32 candidates x 6 float4: per pass 8 candidates x columns 0-3, plus columns 4-5 on half the lanes
#pragma unroll
          for (int g = 0; g < 4; ++g) {
            const int cand = c8 + 8 * g;
            if (cand < nb) {
              float* dst = dst0 + (int64_t)cand * a.t_cand;
              const float* src = s_trj + cand * PROP_TPITCH;
              PROP_ST(reinterpret_cast<float4*>(dst) + q4, *reinterpret_cast<const float4*>(src + 4 * q4));
              if (q4 < 2) PROP_ST(reinterpret_cast<float4*>(dst) + 4 + q4, *reinterpret_cast<const float4*>(src + 16 + 4 * q4));
            }
          }
        } else {  // ragged last chunk: cs * 3 float2 per candidate
          const int per = cs * 3;
          for (int idx = lane; idx < nb * per; idx += 32) {
            const int cand = idx / per, k = idx - cand * per;
            *(reinterpret_cast<float2*>(dst0 + (int64_t)cand * a.t_cand) + k) = *reinterpret_cast<const float2*>(s_trj + cand * PROP_TPITCH + 2 * k);
          }
        }
      }
      __syncwarp();
      buf ^= 1;
    }
    if (live) {
      const int64_t b = (int64_t)t0 + lane;
      if (a.state_out) {
        float* so = a.state_out + b * 6;
        *reinterpret_cast<float2*>(so) = make_float2(c.x, c.y);
        *reinterpret_cast<float2*>(so + 2) = make_float2(c.psi, c.v);
        *reinterpret_cast<float2*>(so + 4) = make_float2(c.D, c.dl);
      }
      if (a.first_coll) a.first_coll[b] = e.first;
      if (a.done_step) a.done_step[b] = e.done;
    }
    if (t0 + PROP_TSTRIDE < t0) break;  // 32-bit wrap
    t0 += PROP_TSTRIDE;
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
  // the reference's row layouts, 16-byte aligned rows: staged kernel
  const bool rows = (B < (int64_t)0x7fffffff) && (s_cand == 6) && (s_comp == 1) && (((uintptr_t)state0 & 7) == 0) &&
                    (!state_out || ((uintptr_t)state_out & 7) == 0) && (a_comp == 1) && (a_step == 2) &&
                    (a_cand % 4 == 0) && (((uintptr_t)actions & 15) == 0) &&
                    (!traj_out || (t_comp == 1 && t_step == 6 && t_cand % 4 == 0 && ((uintptr_t)traj_out & 15) == 0));
  typedef void (*PropKernel)(MapView, QMapView, PropArgs, int*);
  static const PropKernel kRowsK[2][2] = {{k_propagate_rows<false, false>, k_propagate_rows<false, true>},
                                          {k_propagate_rows<true, false>, k_propagate_rows<true, true>}};
  static const PropKernel kStridedK[2][2] = {{k_propagate_strided<false, false>, k_propagate_strided<false, true>},
                                             {k_propagate_strided<true, false>, k_propagate_strided<true, true>}};
  if (!ctx->prop_attr_set) {
    for (int i = 0; i < 4; ++i) {
      DT_CUDA(cudaFuncSetAttribute(kRowsK[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
      DT_CUDA(cudaFuncSetAttribute(kStridedK[i >> 1][i & 1], cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    }
    ctx->prop_attr_set = true;
  }
  const int table = q.g != nullptr ? 1 : 0, stopf = (flags & DT_PROP_STOP_ON_COLLISION) ? 1 : 0;
  const int threads = rows ? PROP_RTHREADS : PROP_THREADS;
  const size_t smem = map_smem + (rows ? (size_t)PROP_RWARPS * PROP_WARP_WORDS * sizeof(float) : 0);
  PropKernel kern = rows ? kRowsK[table][stopf] : kStridedK[table][stopf];
  int per_sm = 0;
  DT_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kern, threads, smem));
  if (per_sm < 1) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_propagate_collide: map too large for shared memory");
  int64_t blocks = (B + threads - 1) / threads;
  const int64_t cap = (int64_t)ctx->sm_count * per_sm;  // persistent: one wave of resident blocks
  if (blocks > cap) blocks = cap;
  kern<<<(int)blocks, threads, smem, st>>>(m, q, a, ctx->d_status);
  DT_LAUNCH_CHECK("k_propagate");
  return DT_OK;
}

// ---------------------------------------------------------------------------------------------
// MPPI rollout cost (planners' MPPI baseline, call sites run_scenarios_with_lidar_MPPI.py:339-341,422; PARITY
// UNPINNED: the reference's MPPI module is not in its repository).  One thread per rollout: T bicycle steps from the
// common start state with controls u[t] + noise[k, t] (state in registers, collision test per step, the rollout
// freezes at a collision or in the goal disc exactly like dt_propagate_collide), then
//     cost[k] = ||p_T - target||^2 + collision_cost * collided + effort_cost * sum_t |u_t + noise_kt|^2
// with target = ref[min(nearest(ref, p_0) + lookahead, n_ref - 1)], found by every block for itself (a few hundred
// path points).  Replaces a broadcast copy, an add, the propagate kernel and five elementwise / reduction launches.
// ---------------------------------------------------------------------------------------------
struct MppiArgs {
  const float* state;   // (6) device
  const float* u;       // (T, 2)
  const float* noise;   // (K, T, 2)
  int64_t K;
  int T;
  const float* ref;     // (n_ref, 2)
  int n_ref, lookahead;
  float goal_x, goal_y, collision_cost, effort_cost;
  float* cost;          // (K)
  float* target_out;    // (2) or null
};

template <bool kTable>
__global__ void __launch_bounds__(PROP_THREADS)
k_mppi_rollout_cost(MapView m, QMapView q, MppiArgs a, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ uint64_t bar;
  __shared__ float s_d[PROP_THREADS / 32];
  __shared__ int s_i[PROP_THREADS / 32];
  __shared__ float s_target[2];
  uint8_t* s_map = s_dyn;
  uint32_t* s_qp = reinterpret_cast<uint32_t*>(s_dyn + m.bytes);
  dt_stage_maps(s_map, s_qp, &bar, m, q);
  const uint32_t s_q = dt_qmap_addr(s_qp, q);
  const float x0 = a.state[0], y0 = a.state[1];
  // nearest reference point to the current position: first index of the minimum squared distance (fp32)
  float best = 3.4e38f;
  int bi = 0x7fffffff;
  for (int j = threadIdx.x; j < a.n_ref; j += blockDim.x) {
    const float dx = a.ref[2 * j] - x0, dy = a.ref[2 * j + 1] - y0;
    const float d = dx * dx + dy * dy;
    if (d < best) {
      best = d;
      bi = j;
    }
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const float ov = __shfl_xor_sync(0xffffffffu, best, off);
    const int oi = __shfl_xor_sync(0xffffffffu, bi, off);
    if (ov < best || (ov == best && oi < bi)) {
      best = ov;
      bi = oi;
    }
  }
  if ((threadIdx.x & 31) == 0) {
    s_d[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
      if (s_d[w] < s_d[0] || (s_d[w] == s_d[0] && s_i[w] < s_i[0])) {
        s_d[0] = s_d[w];
        s_i[0] = s_i[w];
      }
    int t = (s_i[0] == 0x7fffffff ? 0 : s_i[0]) + a.lookahead;
    t = t > a.n_ref - 1 ? a.n_ref - 1 : t;
    s_target[0] = a.ref[2 * t];
    s_target[1] = a.ref[2 * t + 1];
    if (blockIdx.x == 0 && a.target_out) {
      a.target_out[0] = s_target[0];
      a.target_out[1] = s_target[1];
    }
  }
  __syncthreads();
  const float tx = s_target[0], ty = s_target[1];
  for (int64_t k = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; k < a.K; k += (int64_t)gridDim.x * blockDim.x) {
    Car c;
    c.x = x0; c.y = y0; c.psi = a.state[2]; c.v = a.state[3]; c.D = a.state[4]; c.dl = a.state[5];
    dt_sincos_fast(c.psi, c.sn, c.cs);
    EdgeState e = {-1, -1, 1};
    const float2* nz = reinterpret_cast<const float2*>(a.noise + k * (int64_t)a.T * 2);
    float effort = 0.f;
    for (int i = 0; i < a.T; ++i) {
      const float2 n2 = __ldg(nz + i);
      const float u0 = a.u[2 * i] + n2.x, u1 = a.u[2 * i + 1] + n2.y;
      effort += u0 * u0 + u1 * u1;
      if (e.alive) edge_step<kTable, true>(c, e, i, u0, u1, s_map, s_q, m, q, a.goal_x, a.goal_y, status);
    }
    const float dx = c.x - tx, dy = c.y - ty;
    float cost = dx * dx + dy * dy;
    cost += a.collision_cost * (e.first >= 0 ? 1.0f : 0.0f);
    cost += a.effort_cost * effort;
    a.cost[k] = cost;
  }
}

extern "C" int dt_mppi_rollout_cost(dt_ctx* ctx, const float* state, const float* u, const float* noise, int64_t K, int T,
                                    const float* ref_xy, int n_ref, int lookahead, float goal_x, float goal_y,
                                    float collision_cost, float effort_cost, float* cost_out, float* target_out,
                                    void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!ctx->d_map) return dt_fail(ctx, DT_E_NOMAP, "dt_set_map has not been called");
  if (K <= 0) return DT_OK;
  if (!state || !u || !noise || !ref_xy || !cost_out || T < 1 || n_ref < 1 || lookahead < 0 ||
      (reinterpret_cast<uintptr_t>(noise) & 7))
    return dt_fail(ctx, DT_E_ARG, "dt_mppi_rollout_cost: bad argument");
  MppiArgs a;
  a.state = state; a.u = u; a.noise = noise; a.K = K; a.T = T; a.ref = ref_xy; a.n_ref = n_ref; a.lookahead = lookahead;
  a.goal_x = goal_x; a.goal_y = goal_y; a.collision_cost = collision_cost; a.effort_cost = effort_cost;
  a.cost = cost_out; a.target_out = target_out;
  const MapView m = dt_map_view(ctx);
  const QMapView q = dt_qmap_view(ctx);
  const size_t smem = (size_t)m.bytes + (size_t)q.bytes;
  static unsigned long long attr_set = 0;
  const unsigned long long dev_bit = 1ull << (ctx->device & 63);
  if (!(attr_set & dev_bit)) {
    DT_CUDA(cudaFuncSetAttribute(k_mppi_rollout_cost<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    DT_CUDA(cudaFuncSetAttribute(k_mppi_rollout_cost<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024));
    attr_set |= dev_bit;
  }
  int64_t blocks = (K + PROP_THREADS - 1) / PROP_THREADS;
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  if (q.g) k_mppi_rollout_cost<true><<<(int)blocks, PROP_THREADS, smem, (cudaStream_t)stream>>>(m, q, a, ctx->d_status);
  else k_mppi_rollout_cost<false><<<(int)blocks, PROP_THREADS, smem, (cudaStream_t)stream>>>(m, q, a, ctx->d_status);
  DT_LAUNCH_CHECK("k_mppi_rollout_cost");
  return DT_OK;
}

// first action out, control sequence shifted left by one step (the last step repeats): the end of an MPPI tick
__global__ void k_mppi_shift(float* __restrict__ u, int T, int A, float* __restrict__ action_out) {
  __shared__ float s_u[1024];
  const int n = T * A;
  for (int i = threadIdx.x; i < n; i += blockDim.x) s_u[i] = u[i];
  __syncthreads();
  for (int i = threadIdx.x; i < n; i += blockDim.x) {
    if (i < A && action_out) action_out[i] = s_u[i];
    const int src = i + A < n ? i + A : i;
    u[i] = s_u[src];
  }
}

extern "C" int dt_mppi_shift(dt_ctx* ctx, float* u_inout, int T, int A, float* action_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!u_inout || T < 1 || A < 1 || T * A > 1024) return dt_fail(ctx, DT_E_ARG, "dt_mppi_shift: bad argument");
  k_mppi_shift<<<1, 256, 0, (cudaStream_t)stream>>>(u_inout, T, A, action_out);
  DT_LAUNCH_CHECK("k_mppi_shift");
  return DT_OK;
}
