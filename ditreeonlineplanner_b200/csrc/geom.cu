// geom.cu -- grid geometry kernels: collision tests, robot-centric local maps, obstacle probes,
// 2-D lidar ray-marching.  All coordinate arithmetic is float64 with explicit round-to-nearest
// intrinsics (no FMA contraction) so the flags and cell indices are bit-identical to the
// reference's NumPy float64 code when both are fed the same float32 values.
#include "carfast.cuh"

#define GEOM_THREADS 256

// -------------------------------------------------------------------------------------------
// is_colliding_car  (common/map_utils.py:103-115)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEOM_THREADS)
k_collide_car(MapView m, QMapView q, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ th,
              int64_t stride, int64_t B, uint8_t* __restrict__ out, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  uint32_t* s_q = reinterpret_cast<uint32_t*>(s_map + m.bytes);
  dt_stage_maps(s_map, s_q, &bar, m, q);
  const uint32_t q_addr = dt_qmap_addr(s_q, q);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const int r = dt_car_any(s_map, q_addr, q, m.rows, m.cols, x[i * stride], y[i * stride], th[i * stride]);
    if (r & 4) atomicMin(status, DT_E_INDEX);
    out[i] = (uint8_t)(r & 1);
  }
}

// Four consecutive states per thread: the four guard-banded fast tests are straight-line code the scheduler
// interleaves (the single-state kernel is latency-bound at half occupancy), the rare exact decisions are
// taken afterwards behind one branch, and the four flags leave as one 32-bit store.  Needs the quadrant
// table and a 4-byte aligned flag array; states B - B % 4 .. B - 1 take the single-state path.
// (Measured: 0.100 -> 0.079 ms for 2^24 states; a lane-strided assignment -- 32 states apart, byte stores --
// is slower, 0.092 ms: the kernel is issue-bound, ~160 instructions per state, not load-coalescing bound.)
#ifndef COLL4_MINB
#define COLL4_MINB 3  // resident blocks per SM the register allocation is held to (80 registers, no spills)
#endif
// LAYOUT 0: x, y, theta anywhere (element stride `stride`); 1: rows of exactly (x, y, theta) -- y = x + 1, theta = x + 2,
// stride 3, 16-byte aligned: a thread's four states are three 128-bit loads; 2: x, y, theta adjacent in rows of any even
// stride (the planner's (B, 6) states), 8-byte aligned: one 64-bit + one 32-bit load per state off one base address.
template <int LAYOUT>
__global__ void __launch_bounds__(GEOM_THREADS, COLL4_MINB)
k_collide_car4(MapView m, QMapView q, const float* __restrict__ x, const float* __restrict__ y,
               const float* __restrict__ th, int64_t stride, int64_t B, uint8_t* __restrict__ out,
               int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  uint32_t* s_q = reinterpret_cast<uint32_t*>(s_map + m.bytes);
  dt_stage_maps(s_map, s_q, &bar, m, q);
  const uint32_t q_addr = dt_qmap_addr(s_q, q);
  const int64_t quads = B >> 2;
  const int64_t gtid = blockIdx.x * (int64_t)blockDim.x + threadIdx.x, total = (int64_t)gridDim.x * blockDim.x;
  for (int64_t t = gtid; t < quads; t += total) {
    float xs[4], ys[4], ts[4];
    if (LAYOUT == 1) {
      // (loading the next iteration's states one iteration ahead measured no gain: 3.19 TB/s either way)
      const float4* p = reinterpret_cast<const float4*>(x) + 3 * t;
      const float4 a = __ldg(p), b = __ldg(p + 1), c = __ldg(p + 2);
      xs[0] = a.x; ys[0] = a.y; ts[0] = a.z;
      xs[1] = a.w; ys[1] = b.x; ts[1] = b.y;
      xs[2] = b.z; ys[2] = b.w; ts[2] = c.x;
      xs[3] = c.y; ys[3] = c.z; ts[3] = c.w;
    } else if (LAYOUT == 2) {
      const float* p = x + 4 * t * stride;
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 xy = __ldg(reinterpret_cast<const float2*>(p + j * stride));
        xs[j] = xy.x;
        ys[j] = xy.y;
        ts[j] = __ldg(p + j * stride + 2);
      }
    } else {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const int64_t o = (4 * t + j) * stride;
        xs[j] = x[o];
        ys[j] = y[o];
        ts[j] = th[o];
      }
    }
    uint32_t flags = 0, rare = 0;
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const bool th_ok = fabsf(ts[j]) <= DT_SC_MAX;  // false for NaN / huge headings: the exact code decides,
      float sn, cs;                                  // the fast test runs on heading 0 (its table address must stay valid)
      dt_sincos_mufu(th_ok ? ts[j] : 0.f, sn, cs);
      bool amb;
      const bool hit = dt_car_fast(q_addr, q, xs[j], ys[j], sn, cs, amb);
      flags |= (hit ? 1u : 0u) << (8 * j);
      rare |= ((amb || !th_ok) ? 1u : 0u) << j;
    }
    if (rare) {
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        if ((rare >> j) & 1u) {
          const int r = dt_car_test_exact(s_map, m.rows, m.cols, xs[j], ys[j], ts[j]);
          if (r & 4) atomicMin(status, DT_E_INDEX);
          flags = (flags & ~(0xFFu << (8 * j))) | ((uint32_t)(r & 1) << (8 * j));
        }
      }
    }
    *reinterpret_cast<uint32_t*>(out + 4 * t) = flags;
  }
  const int64_t i = 4 * quads + gtid;
  if (i < B) {
    const int r = dt_car_any(s_map, q_addr, q, m.rows, m.cols, x[i * stride], y[i * stride], th[i * stride]);
    if (r & 4) atomicMin(status, DT_E_INDEX);
    out[i] = (uint8_t)(r & 1);
  }
}

// -------------------------------------------------------------------------------------------
// is_colliding_parallel  (common/map_utils.py:221-329), two passes for the batch early return
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEOM_THREADS)
k_collide_points(MapView m, const float* __restrict__ x, const float* __restrict__ y, int64_t stride, int64_t N,
                 double scale, double r, uint8_t* __restrict__ out, int* __restrict__ any_oob,
                 int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  dt_stage_map(s_map, &bar, m);
  int block_oob = 0;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const int t = dt_ball_test(s_map, m.rows, m.cols, scale, r, (double)x[i * stride], (double)y[i * stride]);
    if (t & 4) atomicMin(status, DT_E_INDEX);
    out[i] = (uint8_t)(t & 3);
    block_oob |= (t & 2);
  }
  if (__syncthreads_or(block_oob) && threadIdx.x == 0) atomicOr(any_oob, 1);
}

__global__ void k_select_points(uint8_t* __restrict__ out, int64_t N, const int* __restrict__ any_oob) {
  const int oob = *any_oob;  // map_utils.py:255-259: if any point is outside, only that mask is returned
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < N; i += (int64_t)gridDim.x * blockDim.x) {
    const uint8_t v = out[i];
    out[i] = oob ? ((v >> 1) & 1) : (v & 1);
  }
}

// -------------------------------------------------------------------------------------------
// is_colliding_ant -> is_colliding_maze  (common/map_utils.py:126-218)
// -------------------------------------------------------------------------------------------
__device__ __forceinline__ int maze_test(const uint8_t* __restrict__ g, int R, int C, double s, double r, double x,
                                         double y) {
  const double cx = xmul(xdiv((double)C, 2.0), s), cy = xmul(xdiv((double)R, 2.0), s);
  const int row = dt_floor_i(xdiv(xsub(cy, y), s));
  const int col = dt_floor_i(xdiv(xadd(x, cx), s));
  if (row < 0 || row >= R || col < 0 || col >= C) return 1;
  const double mid_x = xsub(xmul(xadd((double)col, 0.5), s), cx);
  const double mid_y = xsub(cy, xmul(xadd((double)row, 0.5), s));
  const double h = xdiv(s, 2.0);
  const double x_lo = xsub(mid_x, h), x_hi = xadd(mid_x, h), y_lo = xsub(mid_y, h), y_hi = xadd(mid_y, h);
  if (xadd(x, r) > x_hi && (col + 1 >= C || g[row * C + col + 1] == 1)) return 1;
  if (xsub(x, r) < x_lo && (col - 1 < 0 || g[row * C + col - 1] == 1)) return 1;
  if (xadd(y, r) > y_hi && (row - 1 < 0 || g[(row - 1) * C + col] == 1)) return 1;
  if (xsub(y, r) < y_lo && (row + 1 >= R || g[(row + 1) * C + col] == 1)) return 1;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double kx = (k & 1) ? x_lo : x_hi, ky = (k & 2) ? y_lo : y_hi;
    const int ci = row + ((k & 2) ? 1 : -1), cj = col + ((k & 1) ? -1 : 1);
    const double dx = xsub(kx, x), dy = xsub(ky, y);
    const double d = __dsqrt_rn(xadd(xmul(dx, dx), xmul(dy, dy)));
    if (d < r && ci >= 0 && ci < R && cj >= 0 && cj < C && g[ci * C + cj] == 1) return 1;
  }
  return 0;
}

__global__ void __launch_bounds__(GEOM_THREADS)
k_collide_ant(MapView m, const float* __restrict__ st, int64_t row_stride, int64_t B, double radius,
              uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  dt_stage_map(s_map, &bar, m);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const float* p = st + i * row_stride;
    // slots 3..6 are unpacked as (w, x, y, z) by the matrix builder (common/se3_utils.py:155-164)
    const double qx = (double)p[4], qy = (double)p[5];
    const double up = xsub(1.0, xmul(2.0, xadd(xmul(qx, qx), xmul(qy, qy))));
    int hit = 1;
    if (!(up < 0.0)) hit = maze_test(s_map, m.rows, m.cols, m.s, radius, (double)p[0], (double)p[1]);
    out[i] = (uint8_t)hit;
  }
}

// -------------------------------------------------------------------------------------------
// create_local_map  (common/map_utils.py:391-459): one warp per robot pose
// -------------------------------------------------------------------------------------------
struct Axis {
  double v[32];  // linspace(-L/2 + scale/2, L/2 - scale/2, N), N <= 32
  float startf, stepf;  // fp32 start / step of the same linspace (fast path: v[k] ~ startf + k stepf)
  float amax;    // max |v|
};

// exact cell of one local-map point, as the float64 reference computes it
__device__ __forceinline__ int local_map_cell_exact(const MapView& m, double cs, double sn, double px, double py,
                                                    double cx, double cy, double xl, double yl) {
  const double xg = xadd(xsub(xmul(cs, xl), xmul(sn, yl)), px);
  const double yg = xadd(xadd(xmul(sn, xl), xmul(cs, yl)), py);
  int yi = dt_floor_i(xdiv(xsub(cy, yg), m.s));
  int xi = dt_floor_i(xdiv(xadd(xg, cx), m.s));
  xi = dt_clampi(xi, 0, m.cols - 1);
  yi = dt_clampi(yi, 0, m.rows - 1);
  return yi * m.cols + xi;
}

// Fast path: the rotated / shifted grid coordinates in fp32, trusted when they are farther than `eps` from
// every cell border that can change the clipped index (integers 1 .. dim-1); otherwise the float64 code
// decides that point.  eps bounds the fp32 error per pose: five roundings of magnitude <= |x| + |y| + L +
// (cols + rows) s / 2, each <= 2^-24 of it, plus the fp32 rounding of sin / cos / axis values (<= 2^-24
// relative each, times |axis| <= L / 2), divided by the cell size -- with a factor 2 of slack.
// kMulti: the poses come in groups of `group_size` consecutive candidates (one scenario of the device-resident
// planner each, a multiple of the 8 poses a block handles) and group g uses the staged map slot slot_of_group[g]
// (dt_set_map_slot): the block resolves its map through the device table before staging it.  One block per 8
// poses, no grid stride, so a block never crosses a group.
struct MultiMap {
  const MapEntry* table;
  const int32_t* slot_of_group;
  int group_size;
};

template <typename OutT, bool kMulti>
__global__ void __launch_bounds__(GEOM_THREADS)
k_local_map(MapView m, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ th,
            int64_t stride, int64_t B, int N, Axis ax, int pair, OutT* __restrict__ out, MultiMap mm) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  if (kMulti) {
    const int grp = (int)((blockIdx.x * (int64_t)(blockDim.x >> 5)) / mm.group_size);
    int slot = mm.slot_of_group[grp];
    slot = (slot < 0 || slot >= DT_MAX_MAP_SLOTS) ? 0 : slot;
    m = mm.table[slot].m;
    if (m.g == nullptr) return;  // an empty slot: nothing to crop from (idle group)
  }
  dt_stage_map(s_map, &bar, m);
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const double cx = xmul(xdiv((double)m.cols, 2.0), m.s), cy = xmul(xdiv((double)m.rows, 2.0), m.s);
  const float inv_s = (float)(1.0 / m.s);
  const uint32_t s_map_u = dt_smem_u32(s_map);
  const float cy2 = (float)(cy / m.s + 2.0), cx2 = (float)(cx / m.s + 2.0);
  const float span = (float)(cx + cy) + 2.0f * ax.amax;
  const float rmax = (float)m.rows, cmax = (float)m.cols;
  for (int64_t b = blockIdx.x * (int64_t)warps_per_block + (threadIdx.x >> 5); b < B;
       b += (int64_t)gridDim.x * warps_per_block) {
    const float pxf = x[b * stride], pyf = y[b * stride], thf = th[b * stride];
    const double px = (double)pxf, py = (double)pyf;
    // fast path: libm's fp32 sincosf (<= 2 ulp); the float64 pair is computed only if some point needs it
    float snf, csf;
    sincosf(thf, &snf, &csf);
    double sn = 0.0, cs = 0.0;
    bool have_exact_trig = false;
    const float eps = 2.0f * 5.9604645e-8f * 12.0f * (fabsf(pxf) + fabsf(pyf) + span) * inv_s + 1.0e-7f;
    // non-finite poses or an eps that swallows a cell: everything through the exact code
    const bool fast_ok = eps < 0.25f;
    OutT* o = out + b * (int64_t)(N * N);
    // fast-path cell of point (i, j); `exact` is set when the float64 code must decide it
    auto fast_cell = [&](int i, int j, bool& exact) -> int {
      // axis values arithmetically (an indexed constant-bank load per point stalls the FMA chain): within
      // 1 ulp of float(linspace), covered by eps
      const float xl = fmaf((float)j, ax.stepf, ax.startf), yl = fmaf((float)i, ax.stepf, ax.startf);
      const float xg = fmaf(csf, xl, fmaf(-snf, yl, pxf));
      const float yg = fmaf(snf, xl, fmaf(csf, yl, pyf));
      // grid coordinates shifted by +2 and clamped half a cell outside the map: floor by a round-down add
      // of 2^23 (the integer lands in the mantissa; no conversion instructions), in-cell offset exact
      const float M23 = 8388608.0f;
      const float u = fminf(fmaxf(fmaf(-yg, inv_s, cy2), 1.5f), rmax + 2.5f);
      const float w = fminf(fmaxf(fmaf(xg, inv_s, cx2), 1.5f), cmax + 2.5f);
      const float tu = __fadd_rd(u, M23), tw = __fadd_rd(w, M23);
      const float du = u - (tu - M23), dw = w - (tw - M23);
      // any cell border closer than eps sends the point to the float64 code
      exact = fminf(fminf(du, 1.0f - du), fminf(dw, 1.0f - dw)) < eps;
      const int yi = min(max(__float_as_int(tu) - (0x4B000000 + 2), 0), m.rows - 1);
      const int xi = min(max(__float_as_int(tw) - (0x4B000000 + 2), 0), m.cols - 1);
      return yi * m.cols + xi;
    };
    auto exact_cell = [&](int i, int j) -> int {
      if (!have_exact_trig) {
        sincos((double)thf, &sn, &cs);
        have_exact_trig = true;
      }
      return local_map_cell_exact(m, cs, sn, px, py, cx, cy, ax.v[j], ax.v[i]);
    };
    auto occupancy = [&](int cell) -> float {
      uint32_t occ_u;
      asm("ld.shared.u8 %0, [%1];" : "=r"(occ_u) : "r"(s_map_u + (uint32_t)cell));
      return (float)occ_u;
    };
    const int NN = N * N;
    if (pair) {
      // two neighbouring points per lane and step: two independent chains in flight, one packed store
      int i = (2 * lane) / N, j = 2 * lane - i * N;  // advanced incrementally (no division)
      for (int p = 2 * lane; p < NN; p += 64) {
        int i1 = i, j1 = j + 1;
        if (j1 == N) {
          j1 = 0;
          ++i1;
        }
        int c0 = 0, c1 = 0;
        bool e0 = !fast_ok, e1 = !fast_ok;
        if (fast_ok) {
          c0 = fast_cell(i, j, e0);
          c1 = fast_cell(i1, j1, e1);
        }
        if (e0 | e1) {
          if (e0) c0 = exact_cell(i, j);
          if (e1) c1 = exact_cell(i1, j1);
        }
        const float v0 = occupancy(c0), v1 = occupancy(c1);
        if (sizeof(OutT) == 4) {
          *reinterpret_cast<float2*>(o + p) = make_float2(v0, v1);
        } else {  // sampler's rescale to [-1, 1] (fm_policy.py:152)
          *reinterpret_cast<__nv_bfloat162*>(o + p) = __floats2bfloat162_rn(v0 * 2.0f - 1.0f, v1 * 2.0f - 1.0f);
        }
        j += 64;
        while (j >= N) {
          j -= N;
          ++i;
        }
      }
    } else {
      int i = lane / N, j = lane - i * N;  // out[i, j] uses ys[i], xs[j]
      for (int p = lane; p < NN; p += 32) {
        bool exact = !fast_ok;
        int cell = 0;
        if (fast_ok) cell = fast_cell(i, j, exact);
        if (exact) cell = exact_cell(i, j);
        const float occ = occupancy(cell);
        if (sizeof(OutT) == 4) {
          o[p] = (OutT)occ;
        } else {
          o[p] = (OutT)(occ * 2.0f - 1.0f);
        }
        j += 32;
        while (j >= N) {
          j -= N;
          ++i;
        }
      }
    }
  }
}

// -------------------------------------------------------------------------------------------
// check_obstacle_ahead  (planners/RRT.py:61-81)
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEOM_THREADS)
k_ray_probe(MapView m, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ th,
            int64_t stride, int64_t B, uint8_t* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  dt_stage_map(s_map, &bar, m);
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < B; i += (int64_t)gridDim.x * blockDim.x) {
    const int hit = dt_ray_probe_one(s_map, m.rows, m.cols, x[i * stride], y[i * stride], th[i * stride]);
    out[i] = (uint8_t)hit;
  }
}

// -------------------------------------------------------------------------------------------
// check_no_obstacles_in_path  (run_scenarios_with_lidar_DiTree.py:158-181): single block
// -------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(GEOM_THREADS)
k_path_first_obstacle(MapView m, const float* __restrict__ x, const float* __restrict__ y, int64_t stride, int64_t n,
                      int32_t* __restrict__ idx_out, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  __shared__ int best;
  dt_stage_map(s_map, &bar, m);
  if (threadIdx.x == 0) best = 0x7fffffff;
  __syncthreads();
  const double cx = xdiv((double)m.cols, 2.0), cy = xdiv((double)m.rows, 2.0);
  int mine = 0x7fffffff;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    int r = dt_floor_i(xsub(cy, (double)y[i * stride]));
    int c = dt_floor_i(xadd((double)x[i * stride], cx));
    if (r < 0) r += m.rows;  // NumPy negative indices wrap
    if (c < 0) c += m.cols;
    if (r < 0 || r >= m.rows || c < 0 || c >= m.cols) {
      atomicMin(status, DT_E_INDEX);
      continue;
    }
    if (s_map[r * m.cols + c] == 1 && (int)i < mine) mine = (int)i;
  }
  atomicMin(&best, mine);
  __syncthreads();
  if (threadIdx.x == 0) idx_out[0] = (best == 0x7fffffff) ? -1 : best;
}

// the same against a caller-supplied grid in global memory (a few hundred bytes, read through L1 / L2)
__global__ void __launch_bounds__(GEOM_THREADS)
k_path_first_obstacle_grid(const uint8_t* __restrict__ grid, int rows, int cols, const float* __restrict__ x,
                           const float* __restrict__ y, int64_t stride, int64_t n, int32_t* __restrict__ idx_out,
                           int* __restrict__ status) {
  __shared__ int best;
  if (threadIdx.x == 0) best = 0x7fffffff;
  __syncthreads();
  const double cx = xdiv((double)cols, 2.0), cy = xdiv((double)rows, 2.0);
  int mine = 0x7fffffff;
  for (int64_t i = threadIdx.x; i < n; i += blockDim.x) {
    int r = dt_floor_i(xsub(cy, (double)y[i * stride]));
    int c = dt_floor_i(xadd((double)x[i * stride], cx));
    if (r < 0) r += rows;  // NumPy negative indices wrap
    if (c < 0) c += cols;
    if (r < 0 || r >= rows || c < 0 || c >= cols) {
      atomicMin(status, DT_E_INDEX);
      continue;
    }
    if (grid[r * cols + c] == 1 && (int)i < mine) mine = (int)i;
  }
  atomicMin(&best, mine);
  __syncthreads();
  if (threadIdx.x == 0) idx_out[0] = (best == 0x7fffffff) ? -1 : best;
}

// -------------------------------------------------------------------------------------------
// Lidar2DSim.scan  (lidar_sim/lidar_2d_sim.py:18-98): one warp per bundle of 32 rays
// -------------------------------------------------------------------------------------------
#define LIDAR_RAYS 181
#define LIDAR_BUNDLES 6  // ceil(181 / 32)

// 2x2 solve with partial pivoting in the operation order of LAPACK's dgesv (which is what
// numpy.linalg.solve calls): pivot on the larger |a_i0|, scale by the reciprocal, eliminate, back-solve.
__device__ __forceinline__ bool solve2(double a00, double a01, double a10, double a11, double b0, double b1,
                                       double* t, double* s) {
  if (fabs(a10) > fabs(a00)) {
    double tmp;
    tmp = a00; a00 = a10; a10 = tmp;
    tmp = a01; a01 = a11; a11 = tmp;
    tmp = b0; b0 = b1; b1 = tmp;
  }
  if (a00 == 0.0) return false;
  const double l = xmul(a10, xdiv(1.0, a00));
  const double u11 = xsub(a11, xmul(l, a01));
  if (u11 == 0.0) return false;
  const double y1 = xsub(b1, xmul(l, b0));
  const double x1 = xdiv(y1, u11);
  const double x0 = xdiv(xsub(b0, xmul(a01, x1)), a00);
  *t = x0;
  *s = x1;
  return true;
}

__global__ void __launch_bounds__(LIDAR_BUNDLES * 32)
k_lidar_scan(MapView m, const float* __restrict__ pose, int64_t B, double* __restrict__ dist_out,
             double* __restrict__ end_out, uint8_t* __restrict__ visited, int* __restrict__ status) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  dt_stage_map(s_map, &bar, m);
  const int ray = threadIdx.x;  // warp w handles rays 32w .. 32w+31
  // the reference names maze.shape[0] "width" and uses it as the x extent (lidar_2d_sim.py:51)
  const double w = (double)m.rows, h = (double)m.cols;
  const int wi = m.rows, hi = m.cols;
  for (int64_t b = blockIdx.x; b < B; b += gridDim.x) {
    if (ray >= LIDAR_RAYS) continue;
    const double x0 = (double)pose[b * 3 + 0], y0 = (double)pose[b * 3 + 1], yaw = (double)pose[b * 3 + 2];
    const double angle_deg = xadd(-180.0, xmul((double)ray, 2.0));  // arange(-180, 182, 2)
    const double ang = xmul(xadd(yaw, angle_deg), 0.017453292519943295);  // deg2rad: x * (pi / 180)
    double rs, rc;
    sincos(ang, &rs, &rc);
    // borders in the reference's order: left, right, bottom, top
    const double bax[4] = {0.0, w, 0.0, 0.0}, bay[4] = {0.0, 0.0, 0.0, h};
    const double bdx[4] = {0.0, 0.0, w, w}, bdy[4] = {h, h, 0.0, 0.0};
    bool found = false;
    double lx = 0.0, ly = 0.0;
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      if (found) continue;
      double t, s;
      if (!solve2(rc, -bdx[k], rs, -bdy[k], xsub(bax[k], x0), xsub(bay[k], y0), &t, &s)) continue;
      if (t >= 0.0 && 1.0 >= s && s >= 0.0) {
        lx = xadd(xmul(t, rc), x0);
        ly = xadd(xmul(t, rs), y0);
        found = true;
      }
    }
    double hx = lx, hy = ly;
    if (!found) {
      // the reference would raise (last_point is None); flag it and report a zero-length ray
      atomicMin(status, DT_E_INDEX);
      hx = x0;
      hy = y0;
    } else {
      const double ddx = xsub(lx, x0), ddy = xsub(ly, y0);
      const double len = __dsqrt_rn(xadd(xmul(ddx, ddx), xmul(ddy, ddy)));
      const double step = xdiv(0.1, len);
      // numpy.arange(0, 1, step): ceil((1 - 0) / step) samples t_k = k * step
      const double cnt_d = ceil(xdiv(1.0, step));
      const int cnt = cnt_d > 1.0e6 ? 1000000 : (int)cnt_d;
      uint8_t* vis = visited ? visited + b * (int64_t)(m.rows * m.cols) : nullptr;
      for (int k = 0; k < cnt; ++k) {
        const double t = xmul((double)k, step);
        const double sx = xadd(x0, xmul(t, ddx)), sy = xadd(y0, xmul(t, ddy));
        int qx = dt_clampi(dt_floor_i(sx), 0, wi - 1);
        int qy = dt_clampi(dt_floor_i(sy), 0, hi - 1);
        if (qy >= m.rows || qx >= m.cols) {  // non-square map: the reference indexes out of range
          atomicMin(status, DT_E_INDEX);
          break;
        }
        if (s_map[qy * m.cols + qx] == 1) {
          hx = sx;
          hy = sy;
          break;
        }
        if (vis) vis[qy * m.cols + qx] = 1;
      }
    }
    const double ex = xsub(hx, x0), ey = xsub(hy, y0);
    double d = __dsqrt_rn(xadd(xmul(ex, ex), xmul(ey, ey)));
    d = fmin(fmax(d, 0.0), 300.0);  // noise_std = 0, clip to max_range
    dist_out[b * LIDAR_RAYS + ray] = d;
    end_out[(b * LIDAR_RAYS + ray) * 2 + 0] = xadd(x0, xmul(d, rc));
    end_out[(b * LIDAR_RAYS + ray) * 2 + 1] = xadd(y0, xmul(d, rs));
  }
}

// -------------------------------------------------------------------------------------------
// host entry points
// -------------------------------------------------------------------------------------------
static inline int grid_for(int64_t n, int threads, const dt_ctx* ctx, int per_sm = 8) {
  int64_t blocks = (n + threads - 1) / threads;
  int64_t cap = (int64_t)ctx->sm_count * per_sm;  // grid-stride: a whole number of waves
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  return (int)blocks;
}

#define NEED_MAP()                                                              \
  if (!ctx) return DT_E_ARG;                                                    \
  if (!ctx->d_map) return dt_fail(ctx, DT_E_NOMAP, "dt_set_map has not been called")

extern "C" int dt_collide_car(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride,
                              int64_t B, uint8_t* flags_out, void* stream) {
  NEED_MAP();
  if (B <= 0) return DT_OK;
  if (!x || !y || !theta || !flags_out) return dt_fail(ctx, DT_E_ARG, "dt_collide_car: null pointer");
  MapView m = dt_map_view(ctx);
  QMapView q = dt_qmap_view(ctx);
  if (m.bytes + q.bytes > 48 * 1024) {  // beyond the default dynamic shared memory limit: exact code only
    q.g = nullptr;
    q.bytes = 0;
  }
  if (q.g && B >= 4 && (reinterpret_cast<uintptr_t>(flags_out) & 3u) == 0) {
    const int grid = grid_for((B + 3) / 4, GEOM_THREADS, ctx, COLL4_MINB);
    const size_t smem = m.bytes + q.bytes;
    const bool rows = (y == x + 1) && (theta == x + 2);
    const uintptr_t xa = reinterpret_cast<uintptr_t>(x);
    if (rows && stride == 3 && (xa & 15u) == 0)
      k_collide_car4<1><<<grid, GEOM_THREADS, smem, (cudaStream_t)stream>>>(m, q, x, y, theta, stride, B, flags_out, ctx->d_status);
    else if (rows && stride % 2 == 0 && (xa & 7u) == 0)
      k_collide_car4<2><<<grid, GEOM_THREADS, smem, (cudaStream_t)stream>>>(m, q, x, y, theta, stride, B, flags_out, ctx->d_status);
    else
      k_collide_car4<0><<<grid, GEOM_THREADS, smem, (cudaStream_t)stream>>>(m, q, x, y, theta, stride, B, flags_out, ctx->d_status);
  } else {
    k_collide_car<<<grid_for(B, GEOM_THREADS, ctx), GEOM_THREADS, m.bytes + q.bytes, (cudaStream_t)stream>>>(
        m, q, x, y, theta, stride, B, flags_out, ctx->d_status);
  }
  DT_LAUNCH_CHECK("k_collide_car");
  return DT_OK;
}

extern "C" int dt_collide_points(dt_ctx* ctx, const float* x, const float* y, int64_t stride, int64_t N, double scale,
                                 double r, uint8_t* flags_out, void* stream) {
  NEED_MAP();
  if (N <= 0) return DT_OK;
  if (!x || !y || !flags_out) return dt_fail(ctx, DT_E_ARG, "dt_collide_points: null pointer");
  MapView m = dt_map_view(ctx);
  int rc = dt_ensure_scratch(ctx, 256);
  if (rc) return rc;
  int* any_oob = (int*)ctx->d_scratch;
  cudaStream_t st = (cudaStream_t)stream;
  DT_CUDA(cudaMemsetAsync(any_oob, 0, sizeof(int), st));
  k_collide_points<<<grid_for(N, GEOM_THREADS, ctx), GEOM_THREADS, m.bytes, st>>>(m, x, y, stride, N, scale, r,
                                                                                  flags_out, any_oob, ctx->d_status);
  DT_LAUNCH_CHECK("k_collide_points");
  k_select_points<<<grid_for(N, GEOM_THREADS, ctx), GEOM_THREADS, 0, st>>>(flags_out, N, any_oob);
  DT_LAUNCH_CHECK("k_select_points");
  return DT_OK;
}

extern "C" int dt_collide_ant(dt_ctx* ctx, const float* states, int64_t row_stride, int64_t B, double radius,
                              uint8_t* flags_out, void* stream) {
  NEED_MAP();
  if (B <= 0) return DT_OK;
  if (!states || !flags_out || row_stride < 7) return dt_fail(ctx, DT_E_ARG, "dt_collide_ant: bad argument");
  MapView m = dt_map_view(ctx);
  k_collide_ant<<<grid_for(B, GEOM_THREADS, ctx), GEOM_THREADS, m.bytes, (cudaStream_t)stream>>>(
      m, states, row_stride, B, radius, flags_out);
  DT_LAUNCH_CHECK("k_collide_ant");
  return DT_OK;
}

// ---- fast variant ------------------------------------------------------------------------------------------
// The N x N points of a local map are an affine lattice in grid coordinates:
//     u(i, j) = U0 + i Ui + j Uj   (row coordinate - 0.5),      w(i, j) = W0 + i Wi + j Wj   (column coordinate - 0.5)
// with six per-pose coefficients, so a point costs two FMAs per coordinate; the nearest integer of a shifted
// coordinate IS its cell (one magic add, no clamps: the lookup table is the map padded by LM_PAD rings of
// edge-replicated cells -- numpy's clip -- and already holds the OUTPUT value as a float), and the distance to the
// cell centre tells whether a cell border is closer than the fp32 error bound `eps`, in which case the float64 code of
// the reference decides that point.  (i, j) of the pair of points a lane handles in iteration k are the same for
// every pose: kept in registers.  ~16 instructions per point instead of ~40.
// Poses whose map can leave the padded table (centre more than a cell outside the map, or a local map wider than the
// padding) and non-finite poses take the exact code for every point.
#define LM_PAD 5
// the reference's float64 decision for one point, out of line (it is needed for ~1 point in 2500: inlined 14 times it
// quadrupled the kernel and the hot loop missed the instruction cache)
static __device__ __noinline__ int lm_exact_cell(int rows, int cols, double s, float thf, float pxf, float pyf, double xl,
                                                 double yl) {
  MapView m;
  m.g = nullptr; m.rows = rows; m.cols = cols; m.bytes = 0; m.s = s;
  double sn, cs;
  sincos((double)thf, &sn, &cs);
  const double cx = xmul(xdiv((double)cols, 2.0), s), cy = xmul(xdiv((double)rows, 2.0), s);
  return local_map_cell_exact(m, cs, sn, (double)pxf, (double)pyf, cx, cy, xl, yl);
}
#ifndef LM_MINB
#define LM_MINB 3
#endif
template <typename OutT, bool kMulti, int IT>
__global__ void __launch_bounds__(GEOM_THREADS, LM_MINB)
k_local_map_fast(MapView m, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ th,
                 int64_t stride, int64_t B, int N, Axis ax, OutT* __restrict__ out, MultiMap mm) {
  extern __shared__ __align__(16) uint8_t s_raw[];
  __shared__ uint64_t bar;
  if (kMulti) {
    const int grp = (int)((blockIdx.x * (int64_t)(blockDim.x >> 5)) / mm.group_size);
    int slot = mm.slot_of_group[grp];
    slot = (slot < 0 || slot >= DT_MAX_MAP_SLOTS) ? 0 : slot;
    m = mm.table[slot].m;
    if (m.g == nullptr) return;
  }
  uint8_t* s_map = s_raw;
  float* s_tab = reinterpret_cast<float*>(s_raw + ((m.bytes + 15) & ~15));
  dt_stage_map(s_map, &bar, m);
  const int pitch = m.cols + 2 * LM_PAD, trows = m.rows + 2 * LM_PAD;
  for (int e = threadIdx.x; e < trows * pitch; e += blockDim.x) {
    const int r = dt_clampi(e / pitch - LM_PAD, 0, m.rows - 1), c = dt_clampi(e % pitch - LM_PAD, 0, m.cols - 1);
    const float v = (float)s_map[r * m.cols + c];
    s_tab[e] = sizeof(OutT) == 4 ? v : v * 2.0f - 1.0f;   // bf16 output: the sampler's rescale (fm_policy.py:152)
  }
  __syncthreads();
  const int lane = threadIdx.x & 31;
  const int warps_per_block = blockDim.x >> 5;
  const int NN = N * N;
  // the pair of points (p, p + 1), p = 2 lane + 64 k, of this lane in iteration k
  float fi[IT], fj[IT];
  uint32_t wrap = 0;   // bit k: the second point of the pair starts the next row
#pragma unroll
  for (int k = 0; k < IT; ++k) {
    const int p0 = 2 * lane + 64 * k;
    const int i = p0 / N, j = p0 - i * N;
    fi[k] = (float)i;
    fj[k] = (float)j;
    wrap |= (j + 1 == N ? 1u : 0u) << k;
  }
  const double cx = xmul(xdiv((double)m.cols, 2.0), m.s), cy = xmul(xdiv((double)m.rows, 2.0), m.s);
  const float inv_s = (float)(1.0 / m.s);
  const float cxs = (float)(cx / m.s - 0.5), cys = (float)(cy / m.s - 0.5);
  const float st_s = ax.startf * inv_s, sp_s = ax.stepf * inv_s;
  const float span = (float)((cx + cy) / m.s) + 2.0f * ax.amax * inv_s;
  const float MAGIC = 12582912.0f;                       // 1.5 * 2^23: (v + MAGIC) - MAGIC = rint(v)
  const uint32_t tab_u = dt_smem_u32(s_tab) + 4u * (((uint32_t)LM_PAD - 0x4B400000u) * (uint32_t)(pitch + 1));  // wraps
  // the lattice stays inside the padded table when the centre is at most one cell outside the map and the local
  // map's half diagonal (+ the half cell of the shift) fits the rest of the padding
  const bool reach_ok = 1.4143f * ax.amax * inv_s + 1.6f < (float)LM_PAD;
  for (int64_t b = blockIdx.x * (int64_t)warps_per_block + (threadIdx.x >> 5); b < B;
       b += kMulti ? B : (int64_t)gridDim.x * warps_per_block) {
    const float pxf = x[b * stride], pyf = y[b * stride], thf = th[b * stride];
    float snf, csf;
    const bool th_ok = fabsf(thf) <= DT_SC_MAX;
    dt_sincos_mufu(th_ok ? thf : 0.f, snf, csf);          // |error| <= DT_SC_ERR (carfast.cuh)
    const float pxs = pxf * inv_s, pys = pyf * inv_s;
    const float U0 = cys - pys - (snf + csf) * st_s, Uj = -snf * sp_s, Ui = -csf * sp_s;
    const float W0 = cxs + pxs + (csf - snf) * st_s, Wj = csf * sp_s, Wi = -snf * sp_s;
    // fp32 error of a lattice coordinate: <= 16 roundings of magnitude <= |px| + |py| + span (cells), the MUFU
    // sin / cos error times the lattice radius, doubled for slack
    const float mag = fabsf(pxs) + fabsf(pys) + span;
    const float eps = 2.0f * (16.0f * 5.9604645e-8f * mag + 2.0f * DT_SC_ERR * ax.amax * inv_s) + 1.0e-7f;
    const bool fast_ok = th_ok && reach_ok && eps < 0.25f && fabsf(pxs) <= 0.5f * (float)m.cols + 1.0f &&
                         fabsf(pys) <= 0.5f * (float)m.rows + 1.0f;
    const float lim = 0.5f - eps;
    auto exact_value = [&](int i, int j) -> float {
      const float v = (float)s_map[lm_exact_cell(m.rows, m.cols, m.s, thf, pxf, pyf, ax.v[j], ax.v[i])];
      return sizeof(OutT) == 4 ? v : v * 2.0f - 1.0f;
    };
    OutT* o = out + b * (int64_t)NN;
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const int p0 = 2 * lane + 64 * k;
      if (p0 < NN) {
        const bool wr = (wrap >> k) & 1u;
        const float uA = fmaf(fi[k], Ui, fmaf(fj[k], Uj, U0)), wA = fmaf(fi[k], Wi, fmaf(fj[k], Wj, W0));
        const float uB = wr ? fmaf(fi[k] + 1.0f, Ui, U0) : uA + Uj, wB = wr ? fmaf(fi[k] + 1.0f, Wi, W0) : wA + Wj;
        const float tuA = uA + MAGIC, twA = wA + MAGIC, tuB = uB + MAGIC, twB = wB + MAGIC;
        // distance to the cell centre: beyond 0.5 - eps a border is within the error bound
        const bool amb = !(fmaxf(fmaxf(fabsf(uA - (tuA - MAGIC)), fabsf(wA - (twA - MAGIC))),
                                       fmaxf(fabsf(uB - (tuB - MAGIC)), fabsf(wB - (twB - MAGIC)))) < lim);
        float vA, vB;
        if (fast_ok && !amb) {
          const uint32_t aA = tab_u + 4u * (__float_as_uint(tuA) * (uint32_t)pitch + __float_as_uint(twA));
          const uint32_t aB = tab_u + 4u * (__float_as_uint(tuB) * (uint32_t)pitch + __float_as_uint(twB));
          asm("ld.shared.f32 %0, [%1];" : "=f"(vA) : "r"(aA));
          asm("ld.shared.f32 %0, [%1];" : "=f"(vB) : "r"(aB));
        } else {
          const int i = p0 / N, j = p0 - i * N;
          vA = exact_value(i, j);
          vB = (j + 1 == N) ? exact_value(i + 1, 0) : exact_value(i, j + 1);
        }
        if (sizeof(OutT) == 4) *reinterpret_cast<float2*>(o + p0) = make_float2(vA, vB);
        else *reinterpret_cast<__nv_bfloat162*>(o + p0) = __floats2bfloat162_rn(vA, vB);
      }
    }
  }
}

// ---- quad variant (N % 4 == 0: the reference's 20 x 20 and 16 x 16 maps) -------------------------------------------
// Same lattice arithmetic, laid out for fewer issue slots per point (the kernel is issue-bound, not HBM-bound):
//   * a HALF-warp per pose (two poses per warp): 100 quads of a 20 x 20 map fill 7 x 16 lanes to 89 % (a whole warp
//     would run 4 x 32 at 78 %), and the per-pose set-up (MUFU sin / cos, six lattice coefficients, error bound) is
//     paid once per two poses;
//   * a lane owns FOUR consecutive points of one row per iteration: one base coordinate pair + three increments, one
//     128-bit (fp32) or 64-bit (bf16) store, no row-wrap cases;
//   * a point closer to a cell border than the fp32 error bound is NOT re-decided in line (one such lane stalled its
//     whole warp for the ~400-instruction float64 code in 1 of 20 warp iterations): the quad gets its provisional
//     table value and goes onto a per-block list; after the block's poses are done the list is worked off by all
//     threads in parallel, one float64 decision each, overwriting the provisional value (ordered by the barrier).
#define LMQ_CAP 1024
template <typename OutT, bool kMulti, int NT>
__global__ void __launch_bounds__(GEOM_THREADS, LM_MINB)
k_local_map_quad(MapView m, const float* __restrict__ x, const float* __restrict__ y, const float* __restrict__ th,
                 int64_t stride, int64_t B, Axis ax, OutT* __restrict__ out, MultiMap mm) {
  constexpr int NN = NT * NT, NQ = NT / 4, Q = NT * NQ, IT = (Q + 15) / 16;
  extern __shared__ __align__(16) uint8_t s_raw[];
  __shared__ uint64_t bar;
  __shared__ unsigned int s_cnt;
  __shared__ unsigned long long s_list[LMQ_CAP];   // pose * Q + quad
  __shared__ float2 s_ij[16 * IT];
  __shared__ uint32_t s_base;
  const int hw_per_block = blockDim.x >> 4;        // half-warps = poses in flight per block
  if (kMulti) {
    const int grp = (int)((blockIdx.x * (int64_t)hw_per_block) / mm.group_size);
    int slot = mm.slot_of_group[grp];
    slot = (slot < 0 || slot >= DT_MAX_MAP_SLOTS) ? 0 : slot;
    m = mm.table[slot].m;
    if (m.g == nullptr) return;
  }
  uint8_t* s_map = s_raw;
  float* s_tab = reinterpret_cast<float*>(s_raw + ((m.bytes + 15) & ~15));
  if (threadIdx.x == 0) s_cnt = 0;
  dt_stage_map(s_map, &bar, m);
  const int pitch = m.cols + 2 * LM_PAD, trows = m.rows + 2 * LM_PAD;
  for (int e = threadIdx.x; e < trows * pitch; e += blockDim.x) {
    const int r = dt_clampi(e / pitch - LM_PAD, 0, m.rows - 1), c = dt_clampi(e % pitch - LM_PAD, 0, m.cols - 1);
    const float v = (float)s_map[r * m.cols + c];
    s_tab[e] = sizeof(OutT) == 4 ? v : v * 2.0f - 1.0f;   // bf16 output: the sampler's rescale (fm_policy.py:152)
  }
  // (row, first column) of the quad l16 + 16 k of this lane, as floats: the same for every pose -- one 64-bit shared
  // load per quad (fourteen registers would cost the third block per SM, re-deriving them five instructions per quad)
  if (threadIdx.x < 16 * IT) {
    const int qd = threadIdx.x;
    const int i = qd / NQ, j = 4 * (qd - i * NQ);
    s_ij[qd] = make_float2((float)i, (float)j);
  }
  if (threadIdx.x == 0)
    s_base = dt_smem_u32(s_tab) + 4u * (((uint32_t)LM_PAD - 0x4B400000u) * (uint32_t)(pitch + 1));  // wraps
  __syncthreads();
  const int l16 = threadIdx.x & 15;
  const double cx = xmul(xdiv((double)m.cols, 2.0), m.s), cy = xmul(xdiv((double)m.rows, 2.0), m.s);
  const float inv_s = (float)(1.0 / m.s);
  const float cxs = (float)(cx / m.s - 0.5), cys = (float)(cy / m.s - 0.5);
  const float st_s = ax.startf * inv_s, sp_s = ax.stepf * inv_s;
  const float span = (float)((cx + cy) / m.s) + 2.0f * ax.amax * inv_s;
  const float MAGIC = 12582912.0f;                       // 1.5 * 2^23: (v + MAGIC) - MAGIC = rint(v)
  // one base register: address = (bits(tu) * pitch + bits(tw)) * 4 + base (wrapping).  Read back from shared memory so
  // that ptxas cannot take the constant apart again (it re-split it into an extra add per point)
  const uint32_t tab_u = *reinterpret_cast<volatile uint32_t*>(&s_base);
  const uint32_t upitch = (uint32_t)pitch;
  const bool reach_ok = 1.4143f * ax.amax * inv_s + 1.6f < (float)LM_PAD;
  auto exact_value = [&](float thf, float pxf, float pyf, int i, int j) -> float {
    const float v = (float)s_map[lm_exact_cell(m.rows, m.cols, m.s, thf, pxf, pyf, ax.v[j], ax.v[i])];
    return sizeof(OutT) == 4 ? v : v * 2.0f - 1.0f;
  };
  auto store4 = [&](OutT* o, int qd, const float (&v)[4]) {
    if (sizeof(OutT) == 4) {
      *reinterpret_cast<float4*>(o + 4 * qd) = make_float4(v[0], v[1], v[2], v[3]);
    } else {
      const __nv_bfloat162 lo = __floats2bfloat162_rn(v[0], v[1]), hi = __floats2bfloat162_rn(v[2], v[3]);
      uint2 pk;
      pk.x = *reinterpret_cast<const uint32_t*>(&lo);
      pk.y = *reinterpret_cast<const uint32_t*>(&hi);
      *reinterpret_cast<uint2*>(o + 4 * qd) = pk;
    }
  };
  // the next pose is loaded while this one is computed (ncu: a fifth of the stall samples sat on the first use of a
  // pose -- the only global loads of the loop)
  const int64_t b_step = kMulti ? B : (int64_t)gridDim.x * hw_per_block;
  int64_t b = blockIdx.x * (int64_t)hw_per_block + (threadIdx.x >> 4);
  float nxf = 0.f, nyf = 0.f, ntf = 0.f;
  if (b < B) {
    nxf = x[b * stride];
    nyf = y[b * stride];
    ntf = th[b * stride];
  }
  for (; b < B; b += b_step) {
    const float pxf = nxf, pyf = nyf, thf = ntf;
    if (b + b_step < B) {
      nxf = x[(b + b_step) * stride];
      nyf = y[(b + b_step) * stride];
      ntf = th[(b + b_step) * stride];
    }
    float snf, csf;
    const bool th_ok = fabsf(thf) <= DT_SC_MAX;
    dt_sincos_mufu(th_ok ? thf : 0.f, snf, csf);          // |error| <= DT_SC_ERR (carfast.cuh)
    const float pxs = pxf * inv_s, pys = pyf * inv_s;
    const float U0 = cys - pys - (snf + csf) * st_s, Uj = -snf * sp_s, Ui = -csf * sp_s;
    const float W0 = cxs + pxs + (csf - snf) * st_s, Wj = csf * sp_s, Wi = -snf * sp_s;
    // fp32 error bound of a lattice coordinate: as in k_local_map_fast (a coordinate is built from five roundings of
    // the sixteen budgeted there)
    const float mag = fabsf(pxs) + fabsf(pys) + span;
    const float eps = 2.0f * (16.0f * 5.9604645e-8f * mag + 2.0f * DT_SC_ERR * ax.amax * inv_s) + 1.0e-7f;
    const bool fast_ok = th_ok && reach_ok && eps < 0.25f && fabsf(pxs) <= 0.5f * (float)m.cols + 1.0f &&
                         fabsf(pys) <= 0.5f * (float)m.rows + 1.0f;
    const float lim = 0.5f - eps;
    OutT* o = out + b * (int64_t)NN;
    if (!fast_ok) {
      // the lattice may leave the padded table (or the heading the MUFU range): the float64 code decides every point
#pragma unroll 1
      for (int qd = l16; qd < Q; qd += 16) {
        const int i = qd / NQ, j = 4 * (qd - i * NQ);
        float v[4];
#pragma unroll
        for (int e = 0; e < 4; ++e) v[e] = exact_value(thf, pxf, pyf, i, j + e);
        store4(o, qd, v);
      }
      continue;
    }
#pragma unroll
    for (int k = 0; k < IT; ++k) {
      const int qd = l16 + 16 * k;
      if (16 * k + 15 < Q || qd < Q) {   // only the last iteration can be partial
        float v[4], u[4], w[4], tu[4], tw[4], du[4], dw[4];
        const float2 ij = s_ij[qd];
        u[0] = fmaf(ij.x, Ui, fmaf(ij.y, Uj, U0));
        w[0] = fmaf(ij.x, Wi, fmaf(ij.y, Wj, W0));
#pragma unroll
        for (int e = 1; e < 4; ++e) {
          u[e] = fmaf((float)e, Uj, u[0]);
          w[e] = fmaf((float)e, Wj, w[0]);
        }
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          tu[e] = u[e] + MAGIC;
          tw[e] = w[e] + MAGIC;
          du[e] = u[e] - (tu[e] - MAGIC);   // distance to the cell centre: beyond 0.5 - eps a border is within the bound
          dw[e] = w[e] - (tw[e] - MAGIC);
        }
        const float m0 = fmaxf(fmaxf(fabsf(du[0]), fabsf(dw[0])), fabsf(du[1]));
        const float m1 = fmaxf(fmaxf(fabsf(dw[1]), fabsf(du[2])), fabsf(dw[2]));
        const float m2 = fmaxf(fmaxf(fabsf(du[3]), fabsf(dw[3])), m0);
        const bool amb = !(fmaxf(m1, m2) < lim);
#pragma unroll
        for (int e = 0; e < 4; ++e) {
          // address = (bits(tu) * pitch + bits(tw)) * 4 + base, the magic constants' bits folded into the (wrapping) base
          asm("{\n\t.reg .u32 t;\n\t"
              "mad.lo.u32 t, %1, %2, %3;\n\t"
              "shl.b32 t, t, 2;\n\t"
              "add.u32 t, t, %4;\n\t"
              "ld.shared.f32 %0, [t];\n\t}"
              : "=f"(v[e])
              : "r"(__float_as_uint(tu[e])), "r"(upitch), "r"(__float_as_uint(tw[e])), "r"(tab_u));
        }
        if (amb) {
          const unsigned int slot = atomicAdd(&s_cnt, 1u);
          if (slot < LMQ_CAP) {
            s_list[slot] = (unsigned long long)b * (unsigned long long)Q + (unsigned long long)qd;
          } else {   // list full: decide in line
            const int i = qd / NQ, j = 4 * (qd - i * NQ);
#pragma unroll
            for (int e = 0; e < 4; ++e) v[e] = exact_value(thf, pxf, pyf, i, j + e);
          }
        }
        store4(o, qd, v);
      }
    }
  }
  // ---- the listed quads: one float64 decision per thread, over the provisional values ----
  __syncthreads();
  const unsigned int n_list = s_cnt < (unsigned)LMQ_CAP ? s_cnt : (unsigned)LMQ_CAP;
  for (unsigned int e = threadIdx.x; e < 4u * n_list; e += blockDim.x) {
    const unsigned long long ent = s_list[e >> 2];
    const int64_t b = (int64_t)(ent / (unsigned long long)Q);
    const int qd = (int)(ent - (unsigned long long)b * (unsigned long long)Q);
    const int i = qd / NQ, j = 4 * (qd - i * NQ) + (int)(e & 3u);
    const float val = exact_value(th[b * stride], x[b * stride], y[b * stride], i, j);
    OutT* o = out + b * (int64_t)NN + (i * NT + j);
    if (sizeof(OutT) == 4) *reinterpret_cast<float*>(o) = val;
    else *reinterpret_cast<__nv_bfloat16*>(o) = __float2bfloat16_rn(val);
  }
}

// numpy.linspace(start, stop, N): arange(N) * step + start, last element forced to stop
static Axis make_axis(int N, double scale) {
  Axis ax;
  volatile double L = (double)N * scale;
  volatile double start = -L / 2 + scale / 2, stop = L / 2 - scale / 2;
  volatile double step = (stop - start) / (double)(N - 1);
  for (int i = 0; i < N; ++i) {
    volatile double prod = (double)i * step;
    ax.v[i] = prod + start;
  }
  ax.v[N - 1] = stop;
  for (int i = N; i < 32; ++i) ax.v[i] = 0.0;
  ax.amax = 0.f;
  for (int i = 0; i < 32; ++i)
    if (fabs(ax.v[i]) > ax.amax) ax.amax = (float)fabs(ax.v[i]) * 1.000001f;
  ax.startf = (float)start;
  ax.stepf = (float)step;
  return ax;
}

// -> DT_OK launched, 1 = not applicable (N * N / 64 beyond the instantiated iteration counts), < 0 error
template <typename OutT, bool kMulti>
static int local_map_fast_launch_t(dt_ctx* ctx, const MapView& m, const float* x, const float* y, const float* theta,
                                   int64_t stride, int64_t B, int N, const Axis& ax, OutT* out, const MultiMap& mm,
                                   int blocks, size_t smem, cudaStream_t st) {
  // quads: the reference's map sizes (N = 20 car, 16 ant), 16- / 8-byte aligned output; multi-map launches are cut in
  // blocks of 8 poses (one per warp) by the caller: the quad kernel takes 16 per block, so a group must hold a
  // multiple of 16
  const bool quad_ok = (N == 20 || N == 16) && (reinterpret_cast<uintptr_t>(out) & (sizeof(OutT) == 4 ? 15u : 7u)) == 0 &&
                       (!kMulti || (mm.group_size % 16 == 0 && B % 16 == 0));
  if (quad_ok) {
    static unsigned long long qattr_done = 0;   // one bit per device: the attribute is per device (and per instantiation)
    const unsigned long long qdev_bit = 1ull << (ctx->device & 63);
    if (!(qattr_done & qdev_bit)) {
      DT_CUDA(cudaFuncSetAttribute(k_local_map_quad<OutT, kMulti, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      DT_CUDA(cudaFuncSetAttribute(k_local_map_quad<OutT, kMulti, 20>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
      qattr_done |= qdev_bit;
    }
    // two poses per warp; a single wave of resident blocks (LM_MINB per SM) strides over the poses: a second,
    // partial wave would leave most SMs with one block
    int qblocks = kMulti ? (int)(B / 16) : (int)((B + 15) / 16);
    if (!kMulti && qblocks > ctx->sm_count * LM_MINB) qblocks = ctx->sm_count * LM_MINB;
    if (N == 16) k_local_map_quad<OutT, kMulti, 16><<<qblocks, GEOM_THREADS, smem, st>>>(m, x, y, theta, stride, B, ax, out, mm);
    else k_local_map_quad<OutT, kMulti, 20><<<qblocks, GEOM_THREADS, smem, st>>>(m, x, y, theta, stride, B, ax, out, mm);
    DT_LAUNCH_CHECK("k_local_map_quad");
    return DT_OK;
  }
  const int it = (N * N + 63) / 64;
  static unsigned long long attr_done = 0;
  const unsigned long long dev_bit = 1ull << (ctx->device & 63);
  if (!(attr_done & dev_bit)) {
    DT_CUDA(cudaFuncSetAttribute(k_local_map_fast<OutT, kMulti, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DT_CUDA(cudaFuncSetAttribute(k_local_map_fast<OutT, kMulti, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    DT_CUDA(cudaFuncSetAttribute(k_local_map_fast<OutT, kMulti, 16>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_done |= dev_bit;
  }
  if (it <= 4) k_local_map_fast<OutT, kMulti, 4><<<blocks, GEOM_THREADS, smem, st>>>(m, x, y, theta, stride, B, N, ax, out, mm);
  else if (it <= 7) k_local_map_fast<OutT, kMulti, 7><<<blocks, GEOM_THREADS, smem, st>>>(m, x, y, theta, stride, B, N, ax, out, mm);
  else if (it <= 16) k_local_map_fast<OutT, kMulti, 16><<<blocks, GEOM_THREADS, smem, st>>>(m, x, y, theta, stride, B, N, ax, out, mm);
  else return 1;
  DT_LAUNCH_CHECK("k_local_map_fast");
  return DT_OK;
}

static int local_map_fast_launch(dt_ctx* ctx, const MapView& m, const float* x, const float* y, const float* theta,
                                 int64_t stride, int64_t B, int N, const Axis& ax, int out_dtype, void* out,
                                 const MultiMap& mm, int blocks, size_t smem, cudaStream_t st) {
  if (mm.table) {
    return local_map_fast_launch_t<__nv_bfloat16, true>(ctx, m, x, y, theta, stride, B, N, ax, (__nv_bfloat16*)out, mm, blocks, smem, st);
  }
  if (out_dtype == DT_F32)
    return local_map_fast_launch_t<float, false>(ctx, m, x, y, theta, stride, B, N, ax, (float*)out, mm, blocks, smem, st);
  return local_map_fast_launch_t<__nv_bfloat16, false>(ctx, m, x, y, theta, stride, B, N, ax, (__nv_bfloat16*)out, mm, blocks, smem, st);
}

extern "C" int dt_local_map(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride, int64_t B,
                            int N, double scale, int out_dtype, void* out, void* stream) {
  NEED_MAP();
  if (B <= 0) return DT_OK;
  if (!x || !y || !theta || !out) return dt_fail(ctx, DT_E_ARG, "dt_local_map: null pointer");
  if (N < 2 || N > 32) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_local_map: N must be in [2, 32]");
  const Axis ax = make_axis(N, scale);
  MapView m = dt_map_view(ctx);
  const int warps = GEOM_THREADS / 32;
  int64_t blocks = (B + warps - 1) / warps;
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (out_dtype != DT_F32 && out_dtype != DT_BF16) return dt_fail(ctx, DT_E_ARG, "dt_local_map: unknown out_dtype");
  // paired points (packed 8- / 4-byte stores) need an even point count per pose and an aligned output
  const int even = ((N * N) & 1) == 0;
  const int pair = even && (reinterpret_cast<uintptr_t>(out) & (out_dtype == DT_F32 ? 7u : 3u)) == 0;
  const size_t tab_bytes = (size_t)((m.bytes + 15) & ~15) + (size_t)(m.rows + 2 * LM_PAD) * (m.cols + 2 * LM_PAD) * 4;
  if (pair && tab_bytes <= 96 * 1024) {
    const MultiMap none{nullptr, nullptr, 0};
    int rc = local_map_fast_launch(ctx, m, x, y, theta, stride, B, N, ax, out_dtype, out, none, (int)blocks, tab_bytes, st);
    if (rc != 1) return rc;
  }
  if (out_dtype == DT_F32) {
    k_local_map<float, false><<<(int)blocks, GEOM_THREADS, m.bytes, st>>>(m, x, y, theta, stride, B, N, ax, pair,
                                                                            (float*)out, MultiMap{nullptr, nullptr, 0});
  } else {
    k_local_map<__nv_bfloat16, false><<<(int)blocks, GEOM_THREADS, m.bytes, st>>>(
        m, x, y, theta, stride, B, N, ax, pair, (__nv_bfloat16*)out, MultiMap{nullptr, nullptr, 0});
  }
  DT_LAUNCH_CHECK("k_local_map");
  return DT_OK;
}

extern "C" int dt_local_map_slots(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride,
                                  int64_t B, int N, double scale, const int32_t* slot_of_group, int group_size,
                                  void* out_bf16, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!ctx->d_map_table) return dt_fail(ctx, DT_E_NOMAP, "dt_set_map_slot has not been called");
  if (B <= 0) return DT_OK;
  if (!x || !y || !theta || !out_bf16 || !slot_of_group) return dt_fail(ctx, DT_E_ARG, "dt_local_map_slots: null pointer");
  if (N < 2 || N > 32) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_local_map_slots: N must be in [2, 32]");
  const int warps = GEOM_THREADS / 32;
  if (group_size <= 0 || group_size % warps != 0 || B % group_size != 0)
    return dt_fail(ctx, DT_E_ARG, "dt_local_map_slots: group_size must be a multiple of 8 that divides B");
  const Axis ax = make_axis(N, scale);
  MapView none;
  memset(&none, 0, sizeof none);
  const int pair = (((N * N) & 1) == 0) && (reinterpret_cast<uintptr_t>(out_bf16) & 3u) == 0;
  const int64_t blocks = B / warps;
  if (blocks > 0x7fffffffLL) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_local_map_slots: batch too large");
  const MultiMap mm{(const MapEntry*)ctx->d_map_table, slot_of_group, group_size};
  // the largest staged slot decides the shared memory of every block
  int max_cells = 0, max_tab = 0;
  for (const dt_map_slot& sl : ctx->slots) {
    if (!sl.d_map) continue;
    if (sl.map_bytes > max_cells) max_cells = sl.map_bytes;
    const int t = (sl.rows + 2 * LM_PAD) * (sl.cols + 2 * LM_PAD) * 4;
    if (t > max_tab) max_tab = t;
  }
  const size_t tab_bytes = (size_t)((max_cells + 15) & ~15) + (size_t)max_tab;
  if (pair && tab_bytes <= 96 * 1024) {
    int rc = local_map_fast_launch(ctx, none, x, y, theta, stride, B, N, ax, DT_BF16, out_bf16, mm, (int)blocks, tab_bytes,
                                   (cudaStream_t)stream);
    if (rc != 1) return rc;
  }
  k_local_map<__nv_bfloat16, true><<<(int)blocks, GEOM_THREADS, ((ctx->slots_max_bytes + 15) / 16) * 16, (cudaStream_t)stream>>>(
      none, x, y, theta, stride, B, N, ax, pair, (__nv_bfloat16*)out_bf16, mm);
  DT_LAUNCH_CHECK("k_local_map(slots)");
  return DT_OK;
}

extern "C" int dt_ray_probe(dt_ctx* ctx, const float* x, const float* y, const float* theta, int64_t stride, int64_t B,
                            uint8_t* flags_out, void* stream) {
  NEED_MAP();
  if (B <= 0) return DT_OK;
  if (!x || !y || !theta || !flags_out) return dt_fail(ctx, DT_E_ARG, "dt_ray_probe: null pointer");
  MapView m = dt_map_view(ctx);
  k_ray_probe<<<grid_for(B, GEOM_THREADS, ctx), GEOM_THREADS, m.bytes, (cudaStream_t)stream>>>(m, x, y, theta, stride,
                                                                                               B, flags_out);
  DT_LAUNCH_CHECK("k_ray_probe");
  return DT_OK;
}

extern "C" int dt_path_first_obstacle(dt_ctx* ctx, const float* x, const float* y, int64_t stride, int64_t n,
                                      int32_t* idx_out, void* stream) {
  NEED_MAP();
  if (!idx_out) return dt_fail(ctx, DT_E_ARG, "dt_path_first_obstacle: null pointer");
  if (n > 0 && (!x || !y)) return dt_fail(ctx, DT_E_ARG, "dt_path_first_obstacle: null pointer");
  MapView m = dt_map_view(ctx);
  k_path_first_obstacle<<<1, GEOM_THREADS, m.bytes, (cudaStream_t)stream>>>(m, x, y, stride, n, idx_out,
                                                                           ctx->d_status);
  DT_LAUNCH_CHECK("k_path_first_obstacle");
  return DT_OK;
}

extern "C" int dt_path_first_obstacle_grid(dt_ctx* ctx, const uint8_t* grid_u8, int rows, int cols, const float* x,
                                           const float* y, int64_t stride, int64_t n, int32_t* idx_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!idx_out || !grid_u8 || rows <= 0 || cols <= 0) return dt_fail(ctx, DT_E_ARG, "dt_path_first_obstacle_grid: bad argument");
  if (n > 0 && (!x || !y)) return dt_fail(ctx, DT_E_ARG, "dt_path_first_obstacle_grid: null pointer");
  k_path_first_obstacle_grid<<<1, GEOM_THREADS, 0, (cudaStream_t)stream>>>(grid_u8, rows, cols, x, y, stride, n, idx_out,
                                                                          ctx->d_status);
  DT_LAUNCH_CHECK("k_path_first_obstacle_grid");
  return DT_OK;
}

extern "C" int dt_lidar_scan(dt_ctx* ctx, const float* pose, int64_t B, double* dist_out, double* end_out,
                             uint8_t* visited_out, void* stream) {
  NEED_MAP();
  if (B <= 0) return DT_OK;
  if (!pose || !dist_out || !end_out) return dt_fail(ctx, DT_E_ARG, "dt_lidar_scan: null pointer");
  MapView m = dt_map_view(ctx);
  cudaStream_t st = (cudaStream_t)stream;
  if (visited_out) DT_CUDA(cudaMemsetAsync(visited_out, 0, (size_t)B * m.rows * m.cols, st));
  int64_t blocks = B;
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  k_lidar_scan<<<(int)blocks, LIDAR_BUNDLES * 32, m.bytes, st>>>(m, pose, B, dist_out, end_out, visited_out,
                                                                 ctx->d_status);
  DT_LAUNCH_CHECK("k_lidar_scan");
  return DT_OK;
}
