// planner.cu -- the device-resident, multi-scenario RRT expansion loop.
//
// Reference: RRT_Planner.plan (planners/RRT.py:113-257) around BasePlanner.random_node_sample
// (planners/base_planner.py:162-207), nearest_node (RRT.py:49-55), propagate_action_sequence_env
// (base_planner.py:257-320), node insertion (RRT.py:195-207), the final node selection (:220-257) and
// generate_final_path_env (base_planner.py:342-363); the driver loop over scenarios is run_scenarios.py:202-395.
//
// U "unit slots" run U independent trees -- (scenario, run) units of the suite, possibly on different mazes -- in ONE
// device pass; every unit slot owns S edge slots, so a pass advances P = U * S candidate edges by one chunk of h
// actions.  Everything the reference does per iteration on the host happens here on the device:
//
//   k_plan_refill   one warp per edge slot: the sampler's initial noise for this pass (Philox4x32-10 + Box-Muller,
//                   keyed by (unit seed, slot, pass): a unit's result does not depend on which units run beside it);
//                   for a free slot: state sample with goal bias, conditioning-goal coin, nearest node of the unit's
//                   tree (coalesced SoA x[] / y[], float64 distances, warp-shuffle arg-min, lowest index on ties),
//                   visit count, edge length from the propagation schedule, start state / previous action.
//   (library)       local maps with a per-unit map slot, conditioning vectors, K ODE steps of the denoiser, batched
//                   over all P slots (geom.cu, cond.cu, denoiser.cu).
//   k_plan_advance  one block per unit slot, one thread per edge slot: h bicycle steps with collision / goal tests
//                   (the unit's grid and quadrant table staged by TMA), chunk booked into the slot's edge record;
//                   finished edges become tree nodes in SLOT ORDER (block-wide scan: node indices, hence all later
//                   nearest-node ties, are deterministic); goal / iteration-cap detection; on a finished unit the
//                   final node (goal node, or arg-min goal distance), the back-trace and the path copy-out, then the
//                   next unit descriptor is popped from the device queue and the unit slot restarts.
//
// The host only enqueues passes and reads a few counters with a lag; no device->host copy sits between the passes.
// Deviations from the reference loop, all inherited from the batched planner (planners/RRT.py docstring): the S
// samples of a pass see the tree as of the start of the pass; random numbers come from per-unit Philox streams, not
// from the three host generators.  run_type 0 (uniform state sampler) and run_types 1-3 without a previous main path
// (probability-map cell sampler for run_type >= 2 -- np.random.choice over the unit's CDF, dt_plan_set_cdf --, the
// sample as conditioning goal, the obstacle-ahead probe per new node and its penalty in the final selection); replans
// along a previous main path (the online driver) use the host-driven planners.
#include <cooperative_groups.h>

#include "carprop.cuh"

extern "C" int dt_local_map_slots(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride,
                                  int64_t B, int N, double scale, const int32_t* slot_of_group, int group_size,
                                  void* out_bf16, void* stream);

#define PLAN_MAX_SCHED 8
#define PLAN_MAX_DEPTH 1024      // longest root-to-leaf chain the path copy-out handles
#define PLAN_MAX_CDF 64          // probability maps (their CDFs) staged at once
#define PLAN_THREADS_REFILL 256

enum { CNT_HEAD = 0, CNT_TAIL = 1, CNT_DONE = 2, CNT_PASS = 3, CNT_ERR = 4, CNT_ACTIVE = 5, CNT_N = 8 };
enum { PLAN_ERR_NODE_CAP = 1, PLAN_ERR_DEPTH = 2, PLAN_ERR_PATH_CAP = 4 };

struct PlanUnit {   // one unit slot (device)
  int unit_id;      // global id of the unit running here, -1 = idle
  int map_slot;
  uint32_t seed;
  int n_nodes, passes, first_pass, fresh;
  int cdf_slot;     // probability-map sampler (run_type >= 2): slot of the unit's CDF, -1 = uniform sampler
  float goal[2], half_w, half_h;
  int chunks, collisions;   // statistics: chunk expansions booked / chunks that ended in a collision
};

struct PlanDev {   // everything the kernels need, by value
  int U, S, ncap, h, nmax, T, A, erec, max_path, max_units, iter_cap, n_sched, run_type;
  const double* cdf;              // [PLAN_MAX_CDF][DT_MAX_MAP_CELLS] normalised cumulative sums
  const int* cdf_n;               // cells per slot
  uint8_t* node_ahead;            // check_obstacle_ahead of every node (run_type >= 1)
  int sched[PLAN_MAX_SCHED];      // chunks per edge by visit count of the parent (prop_duration // action_horizon)
  float goal_sample_rate, goal_cond_bias;
  float act_mean[2];
  PlanUnit* units;
  int32_t* unit_slot_map;         // [U] map slot of each unit slot (what dt_local_map_slots reads)
  const dt_plan_unit* queue;
  int* cnt;
  // tree, per unit slot
  float *node_x, *node_y, *node_state, *node_lastact, *node_edge;
  int *node_parent, *node_visit, *node_len;   // node_len: (states, actions) per node
  // edge slots
  float *slot_state, *slot_prev, *slot_goal, *slot_edge;
  int *slot_parent, *slot_chunk, *slot_nchunks, *slot_free;
  float *noise, *actions;
  // results, indexed by unit id
  dt_plan_result* res;
  float *res_path, *res_act;
  const MapEntry* table;
};

struct dt_plan {
  dt_ctx* ctx;
  dt_plan_cfg cfg;
  PlanDev d;
  std::vector<void*> allocs;
  float* cond = nullptr;
  void* lm = nullptr;
  int* h_cnt = nullptr;          // pinned: 4 snapshots of the counters
  cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};
  long long passes_enqueued = 0;
  int pushed = 0;
  double norm[16];               // obs mean / std (6 + 6), action mean / std (2 + 2): dt_build_cond_car's layout
  double act_norm[4];
  size_t smem_advance = 0;
};

// ---------------------------------------------------------------------------------------------------------
// Philox4x32-10 (Salmon et al., SC'11): counter-based, so every (unit, slot, pass) owns its own stream
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ uint4 philox4x32(uint4 c, uint2 k) {
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const uint32_t hi0 = __umulhi(0xD2511F53u, c.x), lo0 = 0xD2511F53u * c.x;
    const uint32_t hi1 = __umulhi(0xCD9E8D57u, c.z), lo1 = 0xCD9E8D57u * c.z;
    c = make_uint4(hi1 ^ c.y ^ k.x, lo1, hi0 ^ c.w ^ k.y, lo0);
    k.x += 0x9E3779B9u;
    k.y += 0xBB67AE85u;
  }
  return c;
}
__device__ __forceinline__ float u01(uint32_t x) { return ((float)(x >> 8) + 0.5f) * 5.9604645e-8f; }  // (0, 1)

__device__ __forceinline__ void plan_warp_argmin(double& v, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double ov = __shfl_xor_sync(0xffffffffu, v, off);
    const int oi = __shfl_xor_sync(0xffffffffu, i, off);
    if (ov < v || (ov == v && oi < i)) {
      v = ov;
      i = oi;
    }
  }
}

// ---------------------------------------------------------------------------------------------------------
// noise + sampling + nearest node + slot refill: one warp per edge slot
// ---------------------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(PLAN_THREADS_REFILL)
k_plan_refill(PlanDev p) {
  const int lane = threadIdx.x & 31;
  const int slot = blockIdx.x * (PLAN_THREADS_REFILL / 32) + (threadIdx.x >> 5);
  if (slot >= p.U * p.S) return;
  const int u = slot / p.S, sl = slot - u * p.S;
  const PlanUnit un = p.units[u];
  if (un.unit_id < 0) return;
  const uint2 key = make_uint2(un.seed, 0x44695472u);
  // initial sample of the flow-matching ODE: torch.randn(B, T, A) in the reference (fm_policy.py:158)
  const int na = p.T * p.A;
  float* nz = p.noise + (size_t)slot * na;
  for (int e = 4 * lane; e < na; e += 128) {
    const uint4 r = philox4x32(make_uint4((uint32_t)sl, (uint32_t)un.passes, (uint32_t)(e >> 2), 1u), key);
    const float r0 = sqrtf(-2.0f * __logf(u01(r.x))), r1 = sqrtf(-2.0f * __logf(u01(r.z)));
    float s0, c0, s1, c1;
    __sincosf(6.2831853f * u01(r.y), &s0, &c0);
    __sincosf(6.2831853f * u01(r.w), &s1, &c1);
    const float v[4] = {r0 * c0, r0 * s0, r1 * c1, r1 * s1};
#pragma unroll
    for (int j = 0; j < 4; ++j)
      if (e + j < na) nz[e + j] = v[j];
  }
  if (!(un.fresh || p.slot_free[slot])) return;
  // random_node_sample (base_planner.py:162-207, run_type 0): goal with probability goal_sample_rate, else a
  // uniform position over the map (the other four components of the sample are never used by the planner);
  // conditioning goal (RRT.py:154-157): the sample with probability 1 - goal_conditioning_bias, else the goal
  const uint4 r = philox4x32(make_uint4((uint32_t)sl, (uint32_t)un.passes, 0u, 2u), key);
  const bool explore = u01(r.x) > p.goal_sample_rate;
  float sx = un.goal[0], sy = un.goal[1];
  if (explore) {
    if (un.cdf_slot >= 0) {
      // sample_row_col_from_probability_map (base_planner.py:157-160): np.random.choice(size, p = prob_map) =
      // searchsorted(cdf, u, side = 'right') with a 53-bit uniform variate, then the cell's centre
      // (cell_rowcol_to_xy, car_env.py:189-194)
      const double uu = ((double)(r.y >> 5) * 67108864.0 + (double)(r.z >> 6)) * (1.0 / 9007199254740992.0);
      const double* cdf = p.cdf + (size_t)un.cdf_slot * DT_MAX_MAP_CELLS;
      int lo = 0, hi = p.cdf_n[un.cdf_slot];
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (cdf[mid] <= uu) lo = mid + 1; else hi = mid;
      }
      const int cols = (int)(2.0f * un.half_w + 0.5f);
      const int cell = lo < p.cdf_n[un.cdf_slot] ? lo : p.cdf_n[un.cdf_slot] - 1;
      const int row = cell / cols, col = cell - row * cols;
      sx = ((float)col + 0.5f) - un.half_w;
      sy = un.half_h - ((float)row + 0.5f);
    } else {
      sx = fmaf(2.0f * un.half_w, u01(r.y), -un.half_w);
      sy = fmaf(2.0f * un.half_h, u01(r.z), -un.half_h);
    }
  }
  // run_type 0: the sample with probability 1 - goal_conditioning_bias, else the goal; other run types: the sample
  const bool cond_sample = p.run_type != 0 || u01(r.w) > p.goal_cond_bias;
  // nearest node (RRT.py:49-55): KDTree.query on (x, y); squared distances in float64, lowest index on ties
  const float* nx = p.node_x + (size_t)u * p.ncap;
  const float* ny = p.node_y + (size_t)u * p.ncap;
  double best = __longlong_as_double(0x7ff0000000000000LL);
  int bi = 0x7fffffff;
  for (int j = lane; j < un.n_nodes; j += 32) {
    const double dx = xsub((double)sx, (double)nx[j]), dy = xsub((double)sy, (double)ny[j]);
    const double d = xadd(xmul(dx, dx), xmul(dy, dy));
    if (d < best) {
      best = d;
      bi = j;
    }
  }
  plan_warp_argmin(best, bi);
  if (bi == 0x7fffffff) bi = 0;
  const size_t node = (size_t)u * p.ncap + bi;
  if (lane < 6) p.slot_state[(size_t)slot * 6 + lane] = p.node_state[node * 6 + lane];
  if (lane < 2) {
    p.slot_prev[(size_t)slot * 2 + lane] = p.node_lastact[node * 2 + lane];
    p.slot_goal[(size_t)slot * 2 + lane] = cond_sample ? (lane ? sy : sx) : un.goal[lane];
  }
  if (lane == 0) {
    // edge length from the schedule by the parent's visit count (RRT.py:148-151), then the visit is counted
    const int visit = atomicAdd(&p.node_visit[node], 1);
    const int k = visit < 0 ? 0 : (visit >= p.n_sched ? p.n_sched - 1 : visit);
    p.slot_parent[slot] = bi;
    p.slot_chunk[slot] = 0;
    p.slot_nchunks[slot] = p.sched[k] < 1 ? 1 : p.sched[k];
    p.slot_free[slot] = 0;
  }
}

// ---------------------------------------------------------------------------------------------------------
// propagate one chunk, book it, insert nodes, finish / restart units: one block per unit slot
// ---------------------------------------------------------------------------------------------------------
__device__ __forceinline__ void plan_start_unit(const PlanDev& p, int u, const dt_plan_unit& d, int pass) {
  // called by ONE thread: the root node and the unit-slot record of a fresh unit
  PlanUnit& un = p.units[u];
  const size_t root = (size_t)u * p.ncap;
  for (int k = 0; k < 6; ++k) p.node_state[root * 6 + k] = d.start[k];
  p.node_x[root] = d.start[0];
  p.node_y[root] = d.start[1];
  p.node_parent[root] = -1;
  p.node_visit[root] = 0;
  p.node_len[root * 2] = 0;
  p.node_len[root * 2 + 1] = 0;
  // a root has no incoming edge: the reference feeds prev_actions = None, i.e. zeros AFTER normalisation, which is
  // the action mean before it (fm_policy.py:114-121)
  p.node_lastact[root * 2] = p.act_mean[0];
  p.node_lastact[root * 2 + 1] = p.act_mean[1];
  un.map_slot = d.map_slot;
  un.seed = d.seed;
  un.n_nodes = 1;
  un.passes = 0;
  un.first_pass = pass;
  un.fresh = 1;
  un.goal[0] = d.goal[0];
  un.goal[1] = d.goal[1];
  un.half_w = d.half_w;
  un.half_h = d.half_h;
  un.cdf_slot = (p.run_type >= 2 && d.cdf_slot >= 0 && d.cdf_slot < PLAN_MAX_CDF) ? d.cdf_slot : -1;
  p.node_ahead[root] = 0;
  un.chunks = 0;
  un.collisions = 0;
  p.unit_slot_map[u] = d.map_slot;
  __threadfence();
  un.unit_id = d.unit_id;
}

// pop the next queued unit into unit slot u (one thread); returns false when the queue is empty
__device__ __forceinline__ bool plan_pop(const PlanDev& p, int u, int pass) {
  const int i = atomicAdd(&p.cnt[CNT_HEAD], 1);
  if (i >= p.cnt[CNT_TAIL]) {
    atomicSub(&p.cnt[CNT_HEAD], 1);
    p.units[u].unit_id = -1;
    return false;
  }
  plan_start_unit(p, u, p.queue[i], pass);
  return true;
}

__global__ void __launch_bounds__(256)
k_plan_advance(PlanDev p, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_dyn[];
  __shared__ uint64_t bar;
  __shared__ PlanUnit s_un;
  __shared__ int s_warp_cnt[8], s_goal_slot, s_goal_node, s_total, s_final;
  __shared__ double s_bd[8];
  __shared__ int s_bi[8];
  __shared__ int s_chain[PLAN_MAX_DEPTH], s_off_s[PLAN_MAX_DEPTH], s_off_a[PLAN_MAX_DEPTH];
  __shared__ int s_depth, s_rows_s, s_rows_a;
  const int u = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int pass = p.cnt[CNT_PASS];
  if (tid == 0) {
    s_un = p.units[u];
    s_goal_slot = 0x7fffffff;
    s_goal_node = -1;
  }
  __syncthreads();
  if (s_un.unit_id < 0) {   // idle unit slot: look for work the host may have queued since
    if (tid == 0) plan_pop(p, u, pass + 1);
    __syncthreads();
    if (p.units[u].unit_id >= 0) p.slot_free[u * p.S + tid] = 1;
    return;
  }
  const MapEntry me = p.table[s_un.map_slot];
  const MapView m = me.m;
  const QMapView q = me.q;
  uint8_t* s_map = s_dyn;
  uint32_t* s_qp = reinterpret_cast<uint32_t*>(s_dyn + m.bytes);
  dt_stage_maps(s_map, s_qp, &bar, m, q);
  const uint32_t s_q = dt_qmap_addr(s_qp, q);

  // ---- one chunk of h steps from the slot's state (base_planner.py:257-320) ----
  const int slot = u * p.S + tid;
  const int h = p.h;
  const float* s0 = p.slot_state + (size_t)slot * 6;
  Car c;
  c.x = s0[0]; c.y = s0[1]; c.psi = s0[2]; c.v = s0[3]; c.D = s0[4]; c.dl = s0[5];
  const float start[6] = {c.x, c.y, c.psi, c.v, c.D, c.dl};
  dt_sincos_fast(c.psi, c.sn, c.cs);
  EdgeState e = {-1, -1, 1};
  const float* act = p.actions + (size_t)slot * p.T * p.A;
  const int chunk = p.slot_chunk[slot];
  float* rec = p.slot_edge + (size_t)slot * p.erec;
  float* rec_s = rec + (size_t)chunk * (h + 1) * 6;                       // this chunk's state rows: start, then h steps
  float* rec_a = rec + (size_t)p.nmax * (h + 1) * 6 + (size_t)chunk * h * 2;
#pragma unroll
  for (int k = 0; k < 6; ++k) rec_s[k] = start[k];
  float last_u0 = 0.f, last_u1 = 0.f;
  for (int i = 0; i < h; ++i) {
    const float u0 = act[2 * i], u1 = act[2 * i + 1];
    float* row = rec_s + (i + 1) * 6;
    if (e.alive) {
      if (q.g) edge_step<true, true>(c, e, i, u0, u1, s_map, s_q, m, q, s_un.goal[0], s_un.goal[1], status);
      else edge_step<false, true>(c, e, i, u0, u1, s_map, s_q, m, q, s_un.goal[0], s_un.goal[1], status);
      row[0] = c.x; row[1] = c.y; row[2] = c.psi; row[3] = c.v; row[4] = c.D; row[5] = c.dl;
      rec_a[2 * i] = u0;
      rec_a[2 * i + 1] = u1;
      last_u0 = u0;
      last_u1 = u1;
    }
  }
  // ---- book the chunk (RRT.py:179-211) ----
  const bool coll = e.first >= 0;                 // collision: the whole edge is dropped
  const bool done = !coll && e.done >= 0;
  const int n = done ? e.done + 1 : h;            // steps of this chunk that count
  const int chunks_now = chunk + 1;
  const bool ends = !coll && (done || chunks_now >= p.slot_nchunks[slot]);
  // node indices in slot order: warp ballots + a scan over the 8 warp totals
  const unsigned bal = __ballot_sync(0xffffffffu, ends);
  const int rank_w = __popc(bal & ((1u << lane) - 1u));
  if (lane == 0) s_warp_cnt[warp] = __popc(bal);
  const unsigned bal_c = __ballot_sync(0xffffffffu, coll);
  __syncthreads();
  int base = s_un.n_nodes;
  for (int w = 0; w < warp; ++w) base += s_warp_cnt[w];
  if (tid == 0) {
    int t = 0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += s_warp_cnt[w];
    s_total = t;
  }
  int my_node = ends ? base + rank_w : -1;
  if (my_node >= p.ncap) {   // tree storage exhausted: the edge is dropped and the unit flagged
    atomicOr(&p.cnt[CNT_ERR], PLAN_ERR_NODE_CAP);
    my_node = -1;
  }
  const int len_s = (chunks_now - 1) * (h + 1) + 1 + n, len_a = (chunks_now - 1) * h + n;
  if (my_node >= 0) {
    const size_t node = (size_t)u * p.ncap + my_node;
    p.node_x[node] = c.x;
    p.node_y[node] = c.y;
    float* ns = p.node_state + node * 6;
    ns[0] = c.x; ns[1] = c.y; ns[2] = c.psi; ns[3] = c.v; ns[4] = c.D; ns[5] = c.dl;
    p.node_parent[node] = p.slot_parent[slot];
    p.node_visit[node] = 0;
    p.node_lastact[node * 2] = last_u0;       // parent_action_seq[-1] of the new node
    p.node_lastact[node * 2 + 1] = last_u1;
    p.node_len[node * 2] = len_s;
    p.node_len[node * 2 + 1] = len_a;
    // has_obstacle_ahead of the new node (RRT.py:201-205): always False for run_type 0
    p.node_ahead[node] = p.run_type == 0 ? 0 : (uint8_t)dt_ray_probe_one(s_map, m.rows, m.cols, c.x, c.y, c.psi);
    if (done) atomicMin(&s_goal_slot, tid);
  }
  // the edge records of the new nodes: warp-cooperative copies (coalesced), one ending slot at a time
  {
    unsigned todo = __ballot_sync(0xffffffffu, my_node >= 0);
    while (todo) {
      const int src_lane = __ffs(todo) - 1;
      todo &= todo - 1;
      const int nd = __shfl_sync(0xffffffffu, my_node, src_lane);
      const int ls = __shfl_sync(0xffffffffu, len_s, src_lane), la = __shfl_sync(0xffffffffu, len_a, src_lane);
      const float* src = p.slot_edge + (size_t)(u * p.S + warp * 32 + src_lane) * p.erec;
      float* dst = p.node_edge + ((size_t)u * p.ncap + nd) * p.erec;
      __syncwarp();   // the owning lane's record writes above are visible to the warp
      for (int k = lane; k < ls * 6; k += 32) dst[k] = src[k];
      const int ao = p.nmax * (h + 1) * 6;
      for (int k = lane; k < la * 2; k += 32) dst[ao + k] = src[ao + k];
    }
  }
  // the slot's next pass
  if (coll || ends) {
    p.slot_free[slot] = 1;
  } else {
    float* ss = p.slot_state + (size_t)slot * 6;
    ss[0] = c.x; ss[1] = c.y; ss[2] = c.psi; ss[3] = c.v; ss[4] = c.D; ss[5] = c.dl;
    p.slot_prev[(size_t)slot * 2] = act[2 * (h - 1)];
    p.slot_prev[(size_t)slot * 2 + 1] = act[2 * (h - 1) + 1];
    p.slot_chunk[slot] = chunks_now;
  }
  __syncthreads();
  if (tid == s_goal_slot) s_goal_node = my_node;
  // collisions of this pass (statistics), counted once per warp
  if (lane == 0 && bal_c) atomicAdd(&p.units[u].collisions, __popc(bal_c));
  __syncthreads();
  const int n_nodes = min(s_un.n_nodes + s_total, p.ncap);
  const int passes = s_un.passes + 1;
  const bool finished = s_goal_node >= 0 || passes * p.S >= p.iter_cap;
  if (tid == 0) {
    PlanUnit& un = p.units[u];
    un.n_nodes = n_nodes;
    un.passes = passes;
    un.fresh = 0;
    un.chunks = s_un.chunks + p.S;
  }
  if (!finished) return;

  // ---- the unit is over: final node, back-trace, path copy-out (RRT.py:220-257, base_planner.py:342-363) ----
  const float* nx = p.node_x + (size_t)u * p.ncap;
  const float* ny = p.node_y + (size_t)u * p.ncap;
  int final_node = s_goal_node;
  if (final_node < 0) {
    // arg-min over the nodes but the root of ||p - goal|| + 1e4 * obstacle_ahead (float64, first index on ties,
    // RRT.py:233-237).  No node but the root, or an obstacle ahead of every node: np.all(has_obstacle_ahead) is True
    // and the reference returns (None, None) (:221-226).
    const uint8_t* ahead = p.node_ahead + (size_t)u * p.ncap;
    double best = __longlong_as_double(0x7ff0000000000000LL);
    int bi = 0x7fffffff;
    int any_clear = 0;
    for (int j = 1 + tid; j < n_nodes; j += blockDim.x) {
      const double dx = xsub((double)nx[j], (double)s_un.goal[0]), dy = xsub((double)ny[j], (double)s_un.goal[1]);
      double d = __dsqrt_rn(xadd(xmul(dx, dx), xmul(dy, dy)));
      if (ahead[j]) d = xadd(d, xmul(10e3, 1.0));
      else any_clear = 1;
      if (d < best) {
        best = d;
        bi = j;
      }
    }
    if (!__syncthreads_or(any_clear)) bi = 0x7fffffff;
    plan_warp_argmin(best, bi);
    if (lane == 0) {
      s_bd[warp] = best;
      s_bi[warp] = bi;
    }
    __syncthreads();
    if (tid == 0) {
      for (int w = 1; w < (int)(blockDim.x >> 5); ++w)
        if (s_bd[w] < s_bd[0] || (s_bd[w] == s_bd[0] && s_bi[w] < s_bi[0])) {
          s_bd[0] = s_bd[w];
          s_bi[0] = s_bi[w];
        }
      s_final = (s_bi[0] == 0x7fffffff) ? -1 : s_bi[0];
    }
    __syncthreads();
    final_node = s_final;
  }
  const int uid = s_un.unit_id;
  const int* par = p.node_parent + (size_t)u * p.ncap;
  const int* nlen = p.node_len + (size_t)u * p.ncap * 2;
  if (tid == 0) {
    int depth = 0, rows_s = 0, rows_a = 0, overflow = 0;
    if (final_node >= 0) {
      for (int nd = final_node; nd >= 0; nd = par[nd]) {
        if (depth == PLAN_MAX_DEPTH) {
          overflow = PLAN_ERR_DEPTH;
          break;
        }
        s_chain[depth++] = nd;
      }
      // root first
      for (int a = 0, b = depth - 1; a < b; ++a, --b) {
        const int t = s_chain[a];
        s_chain[a] = s_chain[b];
        s_chain[b] = t;
      }
      for (int k = 0; k < depth; ++k) {
        s_off_s[k] = rows_s;
        s_off_a[k] = rows_a;
        rows_s += nlen[2 * s_chain[k]] + 1;      // the edge's states, then the node's own state
        rows_a += nlen[2 * s_chain[k] + 1];
      }
      if (rows_s > p.max_path || rows_a > p.max_path) overflow |= PLAN_ERR_PATH_CAP;
    }
    if (overflow) atomicOr(&p.cnt[CNT_ERR], overflow);
    s_depth = overflow ? 0 : depth;
    s_rows_s = overflow ? 0 : rows_s;
    s_rows_a = overflow ? 0 : rows_a;
    dt_plan_result r;
    r.unit_id = uid;
    r.goal_reached = s_goal_node >= 0 ? 1 : 0;
    r.has_path = (final_node >= 0 && !overflow) ? 1 : 0;
    r.n_states = s_rows_s;
    r.n_actions = s_rows_a;
    r.n_nodes = n_nodes;
    r.iterations = passes * p.S;
    r.first_pass = s_un.first_pass;
    r.last_pass = pass;
    r.collisions = p.units[u].collisions;
    r.chunks = passes * p.S;
    r.error = overflow;
    if (uid < p.max_units) p.res[uid] = r;
  }
  __syncthreads();
  if (uid < p.max_units) {
    float* out_s = p.res_path + (size_t)uid * p.max_path * 6;
    float* out_a = p.res_act + (size_t)uid * p.max_path * 2;
    const int ao = p.nmax * (h + 1) * 6;
    for (int k = 0; k < s_depth; ++k) {
      const int nd = s_chain[k];
      const float* er = p.node_edge + ((size_t)u * p.ncap + nd) * p.erec;
      const int ls = nlen[2 * nd], la = nlen[2 * nd + 1];
      float* ds = out_s + (size_t)s_off_s[k] * 6;
      for (int j = tid; j < ls * 6; j += blockDim.x) ds[j] = er[j];
      if (tid < 6) ds[ls * 6 + tid] = p.node_state[((size_t)u * p.ncap + nd) * 6 + tid];
      float* da = out_a + (size_t)s_off_a[k] * 2;
      for (int j = tid; j < la * 2; j += blockDim.x) da[j] = er[ao + j];
    }
  }
  __syncthreads();
  // the next unit takes over this unit slot
  if (tid == 0) {
    __threadfence();
    atomicAdd(&p.cnt[CNT_DONE], 1);
    plan_pop(p, u, pass + 1);
  }
  __syncthreads();
  p.slot_free[slot] = 1;
}

__global__ void k_plan_tick(int* cnt) { cnt[CNT_PASS] += 1; }

// pops initial units into idle unit slots (also after a push)
__global__ void __launch_bounds__(256) k_plan_fill_idle(PlanDev p) {
  const int u = blockIdx.x;
  __shared__ int s_started;
  if (threadIdx.x == 0) {
    s_started = 0;
    if (p.units[u].unit_id < 0) s_started = plan_pop(p, u, p.cnt[CNT_PASS]) ? 1 : 0;
  }
  __syncthreads();
  if (s_started && (int)threadIdx.x < p.S) p.slot_free[u * p.S + threadIdx.x] = 1;
}

// ---------------------------------------------------------------------------------------------------------
// host side
// ---------------------------------------------------------------------------------------------------------
template <typename Tp>
static int plan_alloc(dt_plan* pl, Tp** out, size_t count, bool zero = true) {
  dt_ctx* ctx = pl->ctx;
  void* ptr = nullptr;
  DT_CUDA(cudaMalloc(&ptr, count * sizeof(Tp)));
  pl->allocs.push_back(ptr);
  if (zero) DT_CUDA(cudaMemset(ptr, 0, count * sizeof(Tp)));
  *out = (Tp*)ptr;
  return DT_OK;
}

extern "C" void dt_plan_destroy(dt_plan* pl) {
  if (!pl) return;
  cudaSetDevice(pl->ctx->device);
  cudaDeviceSynchronize();
  for (void* a : pl->allocs) cudaFree(a);
  if (pl->h_cnt) cudaFreeHost(pl->h_cnt);
  for (cudaEvent_t e : pl->ev)
    if (e) cudaEventDestroy(e);
  delete pl;
}

extern "C" int dt_plan_create(dt_ctx* ctx, const dt_plan_cfg* cfg, dt_plan** out) {
  if (!ctx || !cfg || !out) return DT_E_ARG;
  *out = nullptr;
  if (!ctx->den) return dt_fail(ctx, DT_E_NOMODEL, "dt_plan_create: dt_load_denoiser has not been called");
  const dt_model_cfg mc = dt_denoiser_cfg(ctx);
  if (cfg->unit_slots < 1 || cfg->unit_slots > 64 || cfg->edge_slots != 256)
    return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_plan_create: 1..64 unit slots of exactly 256 edge slots");
  if (cfg->action_horizon < 1 || cfg->action_horizon > mc.horizon || cfg->n_sched < 1 || cfg->n_sched > PLAN_MAX_SCHED ||
      cfg->node_cap < 2 || cfg->max_units < 1 || cfg->max_path < 16 || cfg->ode_steps < 1 || mc.action_dim != 2 ||
      cfg->run_type < 0 || cfg->run_type > 3)
    return dt_fail(ctx, DT_E_ARG, "dt_plan_create: bad configuration");
  DT_CUDA(cudaSetDevice(ctx->device));
  dt_plan* pl = new dt_plan();
  pl->ctx = ctx;
  pl->cfg = *cfg;
  PlanDev& d = pl->d;
  memset(&d, 0, sizeof d);
  d.U = cfg->unit_slots; d.S = cfg->edge_slots; d.ncap = cfg->node_cap; d.h = cfg->action_horizon;
  d.T = mc.horizon; d.A = mc.action_dim; d.max_path = cfg->max_path; d.max_units = cfg->max_units;
  d.iter_cap = cfg->iteration_cap; d.n_sched = cfg->n_sched;
  d.run_type = cfg->run_type;
  d.nmax = 1;
  for (int i = 0; i < cfg->n_sched; ++i) {
    d.sched[i] = cfg->sched_chunks[i];
    if (d.sched[i] > d.nmax) d.nmax = d.sched[i];
  }
  if (d.nmax > 64) { delete pl; return dt_fail(ctx, DT_E_ARG, "dt_plan_create: an edge is at most 64 chunks"); }
  d.erec = d.nmax * (d.h + 1) * 6 + d.nmax * d.h * 2;
  d.goal_sample_rate = cfg->goal_sample_rate;
  d.goal_cond_bias = cfg->goal_conditioning_bias;
  memcpy(pl->norm, cfg->norm, sizeof pl->norm);
  pl->act_norm[0] = cfg->norm[12]; pl->act_norm[1] = cfg->norm[13];
  pl->act_norm[2] = cfg->norm[14]; pl->act_norm[3] = cfg->norm[15];
  d.act_mean[0] = (float)cfg->norm[12];
  d.act_mean[1] = (float)cfg->norm[13];
  const size_t P = (size_t)d.U * d.S, NN = (size_t)d.U * d.ncap;
  int rc = DT_OK;
#define PA(field, count) if (!rc) rc = plan_alloc(pl, &field, (count))
  PA(d.units, (size_t)d.U);
  PA(d.unit_slot_map, (size_t)d.U);
  dt_plan_unit* queue = nullptr;
  PA(queue, (size_t)d.max_units);
  d.queue = queue;
  PA(d.cnt, (size_t)CNT_N);
  PA(d.node_x, NN); PA(d.node_y, NN); PA(d.node_state, NN * 6); PA(d.node_lastact, NN * 2);
  PA(d.node_edge, NN * d.erec);
  PA(d.node_parent, NN); PA(d.node_visit, NN); PA(d.node_len, NN * 2);
  PA(d.node_ahead, NN);
  if (cfg->run_type >= 2) {
    double* cdf = nullptr;
    int* cdf_n = nullptr;
    PA(cdf, (size_t)PLAN_MAX_CDF * DT_MAX_MAP_CELLS);
    PA(cdf_n, (size_t)PLAN_MAX_CDF);
    d.cdf = cdf;
    d.cdf_n = cdf_n;
  }
  PA(d.slot_state, P * 6); PA(d.slot_prev, P * 2); PA(d.slot_goal, P * 2); PA(d.slot_edge, P * d.erec);
  PA(d.slot_parent, P); PA(d.slot_chunk, P); PA(d.slot_nchunks, P); PA(d.slot_free, P);
  PA(d.noise, P * d.T * d.A); PA(d.actions, P * d.T * d.A);
  PA(d.res, (size_t)d.max_units); PA(d.res_path, (size_t)d.max_units * d.max_path * 6);
  PA(d.res_act, (size_t)d.max_units * d.max_path * 2);
  PA(pl->cond, P * mc.cond_dim);
  __nv_bfloat16* lm = nullptr;
  PA(lm, P * mc.map_size * mc.map_size);
  pl->lm = lm;
#undef PA
  if (!rc && cudaMallocHost(&pl->h_cnt, 4 * CNT_N * sizeof(int)) != cudaSuccess) rc = dt_fail(ctx, DT_E_CUDA, "cudaMallocHost");
  for (int i = 0; i < 4 && !rc; ++i)
    if (cudaEventCreateWithFlags(&pl->ev[i], cudaEventDisableTiming) != cudaSuccess) rc = dt_fail(ctx, DT_E_CUDA, "cudaEventCreate");
  if (rc) {
    dt_plan_destroy(pl);
    return rc;
  }
  memset(pl->h_cnt, 0, 4 * CNT_N * sizeof(int));
  // every unit slot starts idle
  std::vector<PlanUnit> idle((size_t)d.U);
  memset(idle.data(), 0, idle.size() * sizeof(PlanUnit));
  for (auto& un : idle) un.unit_id = -1;
  DT_CUDA(cudaMemcpy(d.units, idle.data(), idle.size() * sizeof(PlanUnit), cudaMemcpyHostToDevice));
  DT_CUDA(cudaFuncSetAttribute(k_plan_advance, cudaFuncAttributeMaxDynamicSharedMemorySize, 160 * 1024));
  *out = pl;
  return DT_OK;
}

extern "C" int dt_plan_push(dt_plan* pl, const dt_plan_unit* units_host, int n, void* stream) {
  if (!pl) return DT_E_ARG;
  dt_ctx* ctx = pl->ctx;
  if (n <= 0) return DT_OK;
  if (!units_host || pl->pushed + n > pl->d.max_units) return dt_fail(ctx, DT_E_ARG, "dt_plan_push: more units than max_units");
  for (int i = 0; i < n; ++i) {
    const dt_plan_unit& un = units_host[i];
    if (un.unit_id < 0 || un.unit_id >= pl->d.max_units || un.map_slot < 0 || un.map_slot >= DT_MAX_MAP_SLOTS ||
        ctx->slots[un.map_slot].d_map == nullptr)
      return dt_fail(ctx, DT_E_ARG, "dt_plan_push: bad unit id or map slot not staged (dt_set_map_slot)");
    if (ctx->slots[un.map_slot].s_global != 1.0) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_plan_push: car maps use 1 m cells");
    if (pl->d.run_type >= 2 && (un.cdf_slot < 0 || un.cdf_slot >= PLAN_MAX_CDF))
      return dt_fail(ctx, DT_E_ARG, "dt_plan_push: run_type >= 2 needs the unit's probability map (dt_plan_set_cdf slot)");
  }
  cudaStream_t st = (cudaStream_t)stream;
  pl->d.table = (const MapEntry*)ctx->d_map_table;
  pl->smem_advance = (size_t)((ctx->slots_max_bytes + 15) / 16) * 16;
  // synchronous (the caller's array may die at return); pushes happen a few times per suite, not per pass
  DT_CUDA(cudaMemcpyAsync((void*)(pl->d.queue + pl->pushed), units_host, (size_t)n * sizeof(dt_plan_unit),
                          cudaMemcpyHostToDevice, st));
  pl->pushed += n;
  DT_CUDA(cudaMemcpyAsync(pl->d.cnt + CNT_TAIL, &pl->pushed, sizeof(int), cudaMemcpyHostToDevice, st));
  DT_CUDA(cudaStreamSynchronize(st));
  k_plan_fill_idle<<<pl->d.U, 256, 0, st>>>(pl->d);
  DT_LAUNCH_CHECK("k_plan_fill_idle");
  return DT_OK;
}

extern "C" int dt_plan_set_cdf(dt_plan* pl, int slot, const double* prob_host, int n, void* stream) {
  if (!pl) return DT_E_ARG;
  dt_ctx* ctx = pl->ctx;
  if (!pl->d.cdf) return dt_fail(ctx, DT_E_ARG, "dt_plan_set_cdf: the plan was created with run_type < 2");
  if (slot < 0 || slot >= PLAN_MAX_CDF || !prob_host || n < 1 || n > DT_MAX_MAP_CELLS)
    return dt_fail(ctx, DT_E_ARG, "dt_plan_set_cdf: bad slot or size");
  // RandomState.choice: cdf = p.cumsum(); cdf /= cdf[-1]  (sequential float64 sums, as NumPy does them)
  std::vector<double> cdf((size_t)n);
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    if (!(prob_host[i] >= 0.0)) return dt_fail(ctx, DT_E_ARG, "dt_plan_set_cdf: probabilities are not non-negative");
    acc += prob_host[i];
    cdf[i] = acc;
  }
  if (!(acc > 0.0)) return dt_fail(ctx, DT_E_ARG, "dt_plan_set_cdf: probabilities sum to zero");
  for (int i = 0; i < n; ++i) cdf[i] /= acc;
  cudaStream_t st = (cudaStream_t)stream;
  DT_CUDA(cudaMemcpyAsync((void*)(pl->d.cdf + (size_t)slot * DT_MAX_MAP_CELLS), cdf.data(), (size_t)n * sizeof(double),
                          cudaMemcpyHostToDevice, st));
  DT_CUDA(cudaMemcpyAsync((void*)(pl->d.cdf_n + slot), &n, sizeof(int), cudaMemcpyHostToDevice, st));
  DT_CUDA(cudaStreamSynchronize(st));
  return DT_OK;
}

// test hook: the conditioning goals (P, 2) and parents (P) of all edge slots as of the last enqueued pass
extern "C" int dt_plan_peek_slots(dt_plan* pl, float* goals_out, int32_t* parents_out, void* stream) {
  if (!pl) return DT_E_ARG;
  dt_ctx* ctx = pl->ctx;
  const size_t P = (size_t)pl->d.U * pl->d.S;
  cudaStream_t st = (cudaStream_t)stream;
  if (goals_out) DT_CUDA(cudaMemcpyAsync(goals_out, pl->d.slot_goal, P * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
  if (parents_out) DT_CUDA(cudaMemcpyAsync(parents_out, pl->d.slot_parent, P * sizeof(int), cudaMemcpyDeviceToHost, st));
  DT_CUDA(cudaStreamSynchronize(st));
  return DT_OK;
}

extern "C" int dt_plan_pass(dt_plan* pl, void* stream) {
  if (!pl) return DT_E_ARG;
  dt_ctx* ctx = pl->ctx;
  if (!pl->d.table) return dt_fail(ctx, DT_E_ARG, "dt_plan_pass: no unit has been pushed");
  cudaStream_t st = (cudaStream_t)stream;
  const PlanDev& d = pl->d;
  const int64_t P = (int64_t)d.U * d.S;
  const dt_model_cfg mc = dt_denoiser_cfg(ctx);
  k_plan_refill<<<(unsigned)((P * 32 + PLAN_THREADS_REFILL - 1) / PLAN_THREADS_REFILL), PLAN_THREADS_REFILL, 0, st>>>(d);
  DT_LAUNCH_CHECK("k_plan_refill");
  int rc = dt_local_map_slots(ctx, d.slot_state, d.slot_state + 1, d.slot_state + 2, 6, P, mc.map_size,
                              pl->cfg.local_map_scale, d.unit_slot_map, d.S, pl->lm, stream);
  if (rc) return rc;
  rc = dt_build_cond_car(ctx, d.slot_state, 6, 1, d.slot_prev, d.slot_goal, 2, P, pl->norm, (double)mc.map_size, pl->cond,
                         stream);
  if (rc) return rc;
  rc = dt_fm_sample(ctx, d.noise, pl->cond, pl->lm, P, pl->cfg.ode_steps, 4.0, pl->act_norm, d.actions, stream);
  if (rc) return rc;
  k_plan_advance<<<d.U, d.S, pl->smem_advance, st>>>(d, ctx->d_status);
  DT_LAUNCH_CHECK("k_plan_advance");
  k_plan_tick<<<1, 1, 0, st>>>(d.cnt);
  DT_LAUNCH_CHECK("k_plan_tick");
  const int ring = (int)(pl->passes_enqueued & 3);
  DT_CUDA(cudaMemcpyAsync(pl->h_cnt + ring * CNT_N, d.cnt, CNT_N * sizeof(int), cudaMemcpyDeviceToHost, st));
  DT_CUDA(cudaEventRecord(pl->ev[ring], st));
  pl->passes_enqueued++;
  return DT_OK;
}

extern "C" int dt_plan_counters(dt_plan* pl, int64_t pass_index, int wait, int32_t* out5) {
  if (!pl || !out5) return DT_E_ARG;
  dt_ctx* ctx = pl->ctx;
  if (pass_index < 0 || pass_index >= pl->passes_enqueued) return dt_fail(ctx, DT_E_ARG, "dt_plan_counters: pass not enqueued");
  if (pl->passes_enqueued - pass_index > 4) return dt_fail(ctx, DT_E_ARG, "dt_plan_counters: snapshot already overwritten (4 deep)");
  const int ring = (int)(pass_index & 3);
  if (wait) {
    DT_CUDA(cudaEventSynchronize(pl->ev[ring]));
  } else {
    const cudaError_t e = cudaEventQuery(pl->ev[ring]);
    if (e == cudaErrorNotReady) return 1;   // not there yet
    if (e != cudaSuccess) return dt_fail_cuda(ctx, e, "cudaEventQuery");
  }
  const int* c = pl->h_cnt + ring * CNT_N;
  out5[0] = c[CNT_HEAD]; out5[1] = c[CNT_TAIL]; out5[2] = c[CNT_DONE]; out5[3] = c[CNT_PASS]; out5[4] = c[CNT_ERR];
  return DT_OK;
}

extern "C" int dt_plan_fetch(dt_plan* pl, int unit_id, dt_plan_result* hdr_out, float* path_out, float* actions_out,
                             int cap_rows, void* stream) {
  if (!pl || !hdr_out) return DT_E_ARG;
  dt_ctx* ctx = pl->ctx;
  if (unit_id < 0 || unit_id >= pl->d.max_units) return dt_fail(ctx, DT_E_ARG, "dt_plan_fetch: bad unit id");
  cudaStream_t st = (cudaStream_t)stream;
  DT_CUDA(cudaMemcpyAsync(hdr_out, pl->d.res + unit_id, sizeof(dt_plan_result), cudaMemcpyDeviceToHost, st));
  DT_CUDA(cudaStreamSynchronize(st));
  if (hdr_out->has_path && path_out && actions_out) {
    if (hdr_out->n_states > cap_rows || hdr_out->n_actions > cap_rows) return dt_fail(ctx, DT_E_ARG, "dt_plan_fetch: output too small");
    DT_CUDA(cudaMemcpyAsync(path_out, pl->d.res_path + (size_t)unit_id * pl->d.max_path * 6,
                            (size_t)hdr_out->n_states * 6 * sizeof(float), cudaMemcpyDeviceToHost, st));
    DT_CUDA(cudaMemcpyAsync(actions_out, pl->d.res_act + (size_t)unit_id * pl->d.max_path * 2,
                            (size_t)hdr_out->n_actions * 2 * sizeof(float), cudaMemcpyDeviceToHost, st));
    DT_CUDA(cudaStreamSynchronize(st));
  }
  return DT_OK;
}

// test hook: the tree of the unit currently in unit slot `u` (node arrays, device -> host)
extern "C" int dt_plan_peek_tree(dt_plan* pl, int u, int32_t* n_nodes_out, int32_t* unit_id_out, float* xy_out,
                                 int32_t* parent_out, int cap, void* stream) {
  if (!pl || !n_nodes_out) return DT_E_ARG;
  dt_ctx* ctx = pl->ctx;
  if (u < 0 || u >= pl->d.U) return dt_fail(ctx, DT_E_ARG, "dt_plan_peek_tree: bad unit slot");
  cudaStream_t st = (cudaStream_t)stream;
  PlanUnit un;
  DT_CUDA(cudaMemcpyAsync(&un, pl->d.units + u, sizeof un, cudaMemcpyDeviceToHost, st));
  DT_CUDA(cudaStreamSynchronize(st));
  *n_nodes_out = un.n_nodes;
  if (unit_id_out) *unit_id_out = un.unit_id;
  const int n = un.n_nodes < cap ? un.n_nodes : cap;
  if (xy_out && n > 0) {
    DT_CUDA(cudaMemcpyAsync(xy_out, pl->d.node_x + (size_t)u * pl->d.ncap, n * sizeof(float), cudaMemcpyDeviceToHost, st));
    DT_CUDA(cudaMemcpyAsync(xy_out + cap, pl->d.node_y + (size_t)u * pl->d.ncap, n * sizeof(float), cudaMemcpyDeviceToHost, st));
  }
  if (parent_out && n > 0)
    DT_CUDA(cudaMemcpyAsync(parent_out, pl->d.node_parent + (size_t)u * pl->d.ncap, n * sizeof(int), cudaMemcpyDeviceToHost, st));
  DT_CUDA(cudaStreamSynchronize(st));
  return DT_OK;
}
