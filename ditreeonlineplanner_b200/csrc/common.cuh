// common.cuh -- context, error plumbing and the shared-memory map staging used by every kernel.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/ditree.h"

#define DT_MAX_MAP_CELLS 16384  // 128 x 128; the reference's grids are <= 31 x 31

struct dt_denoiser;  // denoiser.cu

struct dt_ctx {
  int device = 0;
  std::string err;
  // occupancy grid staged as bytes (1 = wall), padded to a multiple of 16 bytes for bulk copies
  uint8_t* d_map = nullptr;
  int rows = 0, cols = 0;
  double s_global = 1.0;
  int map_bytes = 0;  // padded size
  // device-side status word (DT_E_*), plus small scratch for reductions
  int* d_status = nullptr;
  int* h_status = nullptr;  // pinned
  void* d_scratch = nullptr;
  size_t scratch_bytes = 0;
  int64_t launches = 0;
  int sm_count = 148;
  dt_denoiser* den = nullptr;
  // optional per-launch event timing of the GEMM kernels (dt_profile_begin / dt_profile_end)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_events;  // pairs (start, stop)
  size_t prof_used = 0;
  struct ProfRec { int bn, epi, gw; long long M; int N; long long K; float ms; };
  std::vector<ProfRec> prof_recs;
};

static inline int dt_fail(dt_ctx* ctx, int code, const char* what) {
  if (ctx) ctx->err = what;
  return code;
}
static inline int dt_fail_cuda(dt_ctx* ctx, cudaError_t e, const char* where) {
  if (ctx) {
    ctx->err = std::string(where) + ": " + cudaGetErrorString(e);
  }
  return DT_E_CUDA;
}

#define DT_CUDA(call)                                                      \
  do {                                                                     \
    cudaError_t e__ = (call);                                              \
    if (e__ != cudaSuccess) return dt_fail_cuda(ctx, e__, #call);          \
  } while (0)

#define DT_LAUNCH_CHECK(name)                                              \
  do {                                                                     \
    cudaError_t e__ = cudaGetLastError();                                  \
    if (e__ != cudaSuccess) return dt_fail_cuda(ctx, e__, name);           \
    ctx->launches++;                                                       \
  } while (0)

int dt_ensure_scratch(dt_ctx* ctx, size_t bytes);

// ---------------------------------------------------------------------------------------------
// device helpers
// ---------------------------------------------------------------------------------------------
struct MapView {
  const uint8_t* g;  // global pointer (padded to map_bytes)
  int rows, cols, bytes;
  double s;  // metres per cell
};

static inline MapView dt_map_view(const dt_ctx* ctx) {
  MapView m;
  m.g = ctx->d_map;
  m.rows = ctx->rows;
  m.cols = ctx->cols;
  m.bytes = ctx->map_bytes;
  m.s = ctx->s_global;
  return m;
}

#ifdef __CUDACC__
__device__ __forceinline__ uint32_t dt_smem_u32(const void* p) {
  return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

// Stage the occupancy grid into shared memory with one 1-D bulk TMA copy (cp.async.bulk,
// SASS UBLKCP) completing on an mbarrier.  Call from every thread of the block; `bar` is a
// __shared__ uint64_t, `dst` a 16-byte aligned shared buffer of >= m.bytes.
__device__ __forceinline__ void dt_stage_map(uint8_t* dst, uint64_t* bar, const MapView& m) {
  const uint32_t bar_a = dt_smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"((uint32_t)m.bytes)
                 : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            dt_smem_u32(dst)),
        "l"(m.g), "r"((uint32_t)m.bytes), "r"(bar_a)
        : "memory");
  }
  // every thread waits for phase 0 of the barrier
  uint32_t done = 0;
  while (!done) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], 0;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar_a)
        : "memory");
  }
}

// ---- bit-exact float64 geometry (no FMA contraction: every op is an explicit _rn intrinsic) ----
__device__ __forceinline__ double xadd(double a, double b) { return __dadd_rn(a, b); }
__device__ __forceinline__ double xsub(double a, double b) { return __dsub_rn(a, b); }
__device__ __forceinline__ double xmul(double a, double b) { return __dmul_rn(a, b); }
__device__ __forceinline__ double xdiv(double a, double b) { return __ddiv_rn(a, b); }

__device__ __forceinline__ int dt_clampi(int v, int lo, int hi) { return v < lo ? lo : (v > hi ? hi : v); }

// floor() of a double to int with saturation (NumPy's astype(int) is int64; values here are tiny
// unless the state has diverged, in which case they are out of bounds either way).
__device__ __forceinline__ int dt_floor_i(double v) {
  double f = floor(v);
  if (!(f > -1.0e9)) return -1000000000;  // also catches NaN
  if (f > 1.0e9) return 1000000000;
  return (int)f;
}

// One ball of is_colliding_parallel (common/map_utils.py:221-329) with cell size s, radius r.
// Returns bit0 = full test result (inside wall | sides | corners), bit1 = out of bounds,
// bit2 = the diagonal stage would raise IndexError in NumPy (column clipped with the row count),
// bit3 = already colliding before the diagonal stage (inside wall | sides).
__device__ __forceinline__ int dt_ball_test(const uint8_t* __restrict__ grid, int R, int C, double s, double r,
                                            double ax, double ay) {
  const double cx = xmul(xdiv((double)C, 2.0), s);
  const double cy = xmul(xdiv((double)R, 2.0), s);
  const int row = dt_floor_i(xdiv(xsub(cy, ay), s));
  const int col = dt_floor_i(xdiv(xadd(ax, cx), s));
  if (row < 0 || row >= R || col < 0 || col >= C) return 2;
  int hit = grid[row * C + col] == 1;
  const double mid_x = xsub(xmul(xadd((double)col, 0.5), s), cx);
  const double mid_y = xsub(cy, xmul(xadd((double)row, 0.5), s));
  const double h = xdiv(s, 2.0);
  const double x_lo = xsub(mid_x, h), x_hi = xadd(mid_x, h);
  const double y_lo = xsub(mid_y, h), y_hi = xadd(mid_y, h);
  const int cR = dt_clampi(col + 1, 0, C - 1), cL = dt_clampi(col - 1, 0, C - 1);
  const int rU = dt_clampi(row - 1, 0, R - 1), rD = dt_clampi(row + 1, 0, R - 1);
  hit |= (xadd(ax, r) > x_hi) & (grid[row * C + cR] == 1);
  hit |= (xsub(ax, r) < x_lo) & (grid[row * C + cL] == 1);
  hit |= (xadd(ay, r) > y_hi) & (grid[rU * C + col] == 1);
  hit |= (xsub(ay, r) < y_lo) & (grid[rD * C + col] == 1);
  int err = hit ? 8 : 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double kx = (k & 1) ? x_lo : x_hi;   // TR, TL, BR, BL
    const double ky = (k & 2) ? y_lo : y_hi;
    const int ci = row + ((k & 2) ? 1 : -1);
    const int cj = col + ((k & 1) ? -1 : 1);
    const int outside = (ci < 0) | (ci >= R) | (cj < 0) | (cj >= C);
    const int ci_c = dt_clampi(ci, 0, R - 1);
    const int cj_c = dt_clampi(cj, 0, R - 1);  // sic: clipped with the ROW count (map_utils.py:326)
    if (cj_c > C - 1) {                        // NumPy would raise IndexError here
      err |= 4;
      continue;
    }
    const double d = hypot(xsub(kx, ax), xsub(ky, ay));
    hit |= outside | ((d < r) & (grid[ci_c * C + cj_c] == 1));
  }
  return hit | err;
}

// is_colliding_car (common/map_utils.py:103-115) on a float32 state up-cast to float64.
// Returns 0/1, or 1|4 when the reference would raise.
__device__ __forceinline__ int dt_car_test(const uint8_t* __restrict__ grid, int R, int C, float xf, float yf,
                                           float thf) {
  const double th = (double)thf;
  double sn, cs;
  sincos(th, &sn, &cs);
  const double half = xmul(0.15, 0.5);
  const double ox = xmul(half, cs), oy = xmul(half, sn);
  const double x = (double)xf, y = (double)yf;
  const int a = dt_ball_test(grid, R, C, 1.0, 0.1, xadd(x, ox), xadd(y, oy));
  const int b = dt_ball_test(grid, R, C, 1.0, 0.1, xsub(x, ox), xsub(y, oy));
  if ((a | b) & 2) return 1;  // a ball out of bounds decides the pair (map_utils.py:255-259 + .any())
  // the reference returns before the diagonal stage when both balls already collide (:266,:311)
  const int raises = ((a | b) & 4) && !((a & 8) && (b & 8));
  return ((a | b) & 1) | (raises ? 4 : 0);
}

// ---------------------------------------------------------------------------------------------
// Exactness-preserving fast path.  The exact test above spends most of its time in the float64
// sincos and in eight hypot calls.  Here the heading's sine / cosine come from fp32 sincosf
// (|error| <= 2 ulp ~ 1.2e-7, i.e. <= 1e-8 m on the 0.075 m ball offset) and every decision the
// reference takes (cell index floors, the four side comparisons, the four corner distances) is
// evaluated with a guard band DT_EPS = 1e-7 m >> that error.  If every decision is clear of its
// threshold by more than the guard band, the outcome is provably the one the exact float64 code
// produces; otherwise (probability ~1e-5 per state) the exact code runs.  Corner distances use
// d^2 vs r^2 instead of hypot, and only when the diagonal cell can matter.
// ---------------------------------------------------------------------------------------------
#define DT_EPS 1.0e-7
#define DT_AMBIG 16

__device__ __forceinline__ int dt_ball_test_guarded(const uint8_t* __restrict__ grid, int R, int C, double ax,
                                                    double ay) {
  const double r = 0.1;
  const double cx = 0.5 * (double)C, cy = 0.5 * (double)R;  // cell size 1
  const double u = cy - ay, w = ax + cx;
  const double fu = floor(u), fw = floor(w);
  const double du = u - fu, dw = w - fw;  // position inside the cell, in [0, 1)
  int amb = (du < DT_EPS) | (du > 1.0 - DT_EPS) | (dw < DT_EPS) | (dw > 1.0 - DT_EPS);
  if (!(fu > -1.0e9 && fu < 1.0e9 && fw > -1.0e9 && fw < 1.0e9)) return amb ? DT_AMBIG : 2;
  const int row = (int)fu, col = (int)fw;
  if (row < 0 || row >= R || col < 0 || col >= C) return amb ? DT_AMBIG : 2;
  int hit = grid[row * C + col] == 1;
  // distances to the four cell edges: right = 1 - dw, left = dw, top = du (y grows upwards, rows down), bottom = 1 - du
  const double e_r = 1.0 - dw, e_l = dw, e_t = du, e_b = 1.0 - du;
  const int cR = dt_clampi(col + 1, 0, C - 1), cL = dt_clampi(col - 1, 0, C - 1);
  const int rU = dt_clampi(row - 1, 0, R - 1), rD = dt_clampi(row + 1, 0, R - 1);
  const int wR = grid[row * C + cR] == 1, wL = grid[row * C + cL] == 1;
  const int wU = grid[rU * C + col] == 1, wD = grid[rD * C + col] == 1;
  // side tests: ball reaches past the edge  <=>  edge distance < r
  hit |= (e_r < r) & wR;
  hit |= (e_l < r) & wL;
  hit |= (e_t < r) & wU;
  hit |= (e_b < r) & wD;
  amb |= (fabs(e_r - r) < DT_EPS) & wR;
  amb |= (fabs(e_l - r) < DT_EPS) & wL;
  amb |= (fabs(e_t - r) < DT_EPS) & wU;
  amb |= (fabs(e_b - r) < DT_EPS) & wD;
  int err = hit ? 8 : 0;
#pragma unroll
  for (int k = 0; k < 4; ++k) {
    const double ex = (k & 1) ? e_l : e_r;   // TR, TL, BR, BL
    const double ey = (k & 2) ? e_b : e_t;
    const int ci = row + ((k & 2) ? 1 : -1);
    const int cj = col + ((k & 1) ? -1 : 1);
    const int outside = (ci < 0) | (ci >= R) | (cj < 0) | (cj >= C);
    const int ci_c = dt_clampi(ci, 0, R - 1);
    const int cj_c = dt_clampi(cj, 0, R - 1);  // sic: clipped with the ROW count (map_utils.py:326)
    if (cj_c > C - 1) {
      err |= 4;
      continue;
    }
    hit |= outside;
    if (grid[ci_c * C + cj_c] == 1 && ex < r + DT_EPS && ey < r + DT_EPS) {
      const double d2 = ex * ex + ey * ey;
      hit |= d2 < r * r;
      amb |= fabs(d2 - r * r) < 4.0 * r * DT_EPS;
    }
  }
  return amb ? DT_AMBIG : (hit | err);
}

__device__ __forceinline__ int dt_car_test_fast(const uint8_t* __restrict__ grid, int R, int C, float xf, float yf,
                                                float thf) {
  float snf, csf;
  sincosf(thf, &snf, &csf);
  const double ox = 0.075 * (double)csf, oy = 0.075 * (double)snf;
  const double x = (double)xf, y = (double)yf;
  const int a = dt_ball_test_guarded(grid, R, C, x + ox, y + oy);
  const int b = dt_ball_test_guarded(grid, R, C, x - ox, y - oy);
  if ((a | b) & DT_AMBIG) return dt_car_test(grid, R, C, xf, yf, thf);  // rare: decide with the exact code
  if ((a | b) & 2) return 1;
  const int raises = ((a | b) & 4) && !((a & 8) && (b & 8));
  return ((a | b) & 1) | (raises ? 4 : 0);
}

// ---------------------------------------------------------------------------------------------
// Second-generation fast path: per-cell neighbourhood masks + fp32 in-cell geometry.
//
// dt_build_nbr precomputes, once per block, a 16-bit mask per grid cell with everything the
// reference looks up around a ball whose centre lies in that cell (own cell, the 4 side cells with
// the reference's index clipping, the 4 diagonal cells with its row-count column clip, the
// "diagonal outside the grid" flags and the would-raise-IndexError flag).  The per-ball test is
// then: float64 only to locate the cell and the in-cell offsets (du, dw in [0,1)), one 16-bit
// shared load, and ~40 fp32 compare / FMA operations with the guard band below.  A decision
// closer than the guard band to its threshold defers to the exact float64 code (dt_car_test).
//   error budget on an in-cell offset: 1e-8 (fp32 sincos on the 0.075 m offset) + 6e-8 (fp32
//   rounding of a value in [0,1)) < DT_EPSF = 4e-7.
// ---------------------------------------------------------------------------------------------
#define NB_SELF 1u
#define NB_R 2u
#define NB_L 4u
#define NB_U 8u
#define NB_D 16u
#define NB_CW(k) (32u << (k))    // diagonal k (0 TR, 1 TL, 2 BR, 3 BL) is a wall
#define NB_CO(k) (512u << (k))   // diagonal k lies outside the grid (collides regardless of distance)
#define NB_ERR 8192u             // the reference would index out of range for a diagonal of this cell
#define DT_EPSF 4.0e-7f

__device__ __forceinline__ void dt_build_nbr(const uint8_t* __restrict__ grid, uint16_t* __restrict__ nbr, int R,
                                             int C) {
  for (int cell = threadIdx.x; cell < R * C; cell += blockDim.x) {
    const int row = cell / C, col = cell - row * C;
    unsigned m = grid[cell] == 1 ? NB_SELF : 0u;
    m |= grid[row * C + dt_clampi(col + 1, 0, C - 1)] == 1 ? NB_R : 0u;
    m |= grid[row * C + dt_clampi(col - 1, 0, C - 1)] == 1 ? NB_L : 0u;
    m |= grid[dt_clampi(row - 1, 0, R - 1) * C + col] == 1 ? NB_U : 0u;
    m |= grid[dt_clampi(row + 1, 0, R - 1) * C + col] == 1 ? NB_D : 0u;
    for (int k = 0; k < 4; ++k) {
      const int ci = row + ((k & 2) ? 1 : -1), cj = col + ((k & 1) ? -1 : 1);
      const int cj_c = dt_clampi(cj, 0, R - 1);  // sic (map_utils.py:326)
      if (cj_c > C - 1) {
        m |= NB_ERR;
        continue;
      }
      if ((ci < 0) | (ci >= R) | (cj < 0) | (cj >= C)) m |= NB_CO(k);
      if (grid[dt_clampi(ci, 0, R - 1) * C + cj_c] == 1) m |= NB_CW(k);
    }
    nbr[cell] = (uint16_t)m;
  }
}

__device__ __forceinline__ unsigned dt_sign(float v) { return __float_as_uint(v) >> 31; }  // 1 if v < 0

// returns bit0 hit, bit1 out of bounds, bit2 would raise, bit3 hit before the diagonal stage, DT_AMBIG
__device__ __forceinline__ int dt_ball_test_nbr(const uint16_t* __restrict__ nbr, int R, int C, double ax, double ay) {
  const double u = 0.5 * (double)R - ay, w = ax + 0.5 * (double)C;
  if (!(u >= 0.0 && u < (double)R && w >= 0.0 && w < (double)C)) {
    const bool near = (u > -1.0e-6) && (u < (double)R + 1.0e-6) && (w > -1.0e-6) && (w < (double)C + 1.0e-6);
    return near ? DT_AMBIG : 2;  // NaN lands here too (every comparison false) -> out of bounds
  }
  const double fu = floor(u), fw = floor(w);
  const float du = (float)(u - fu), dw = (float)(w - fw);
  const unsigned nb = nbr[(int)fu * C + (int)fw];
  const float r = 0.1f, r2 = r * r;
  const float e_r = 1.0f - dw, e_l = dw, e_t = du, e_b = 1.0f - du;
  // geometry mask in the bit layout of `nb`: bit set <=> the ball reaches past that edge / into that corner disc
  const float xr = e_r * e_r, xl = e_l * e_l, yt = e_t * e_t, yb = e_b * e_b;
  const float d0 = xr + yt, d1 = xl + yt, d2 = xr + yb, d3 = xl + yb;
  const unsigned gm = (dt_sign(e_r - r) << 1) | (dt_sign(e_l - r) << 2) | (dt_sign(e_t - r) << 3) |
                      (dt_sign(e_b - r) << 4) | (dt_sign(d0 - r2) << 5) | (dt_sign(d1 - r2) << 6) |
                      (dt_sign(d2 - r2) << 7) | (dt_sign(d3 - r2) << 8);
  const unsigned hits = (nb & NB_SELF) | (nb & gm & 0x1FEu) | (nb & (NB_CO(0) | NB_CO(1) | NB_CO(2) | NB_CO(3)));
  // guard band, evaluated on the smallest margin of ALL decisions (also those whose neighbour cell is
  // free -- that only defers a few more states to the exact code): cell boundaries, edges, corner discs
  const float m_cell = fminf(fminf(e_r, e_l), fminf(e_t, e_b));
  const float m_side = fminf(fminf(fabsf(e_r - r), fabsf(e_l - r)), fminf(fabsf(e_t - r), fabsf(e_b - r)));
  const float m_corner = fminf(fminf(fabsf(d0 - r2), fabsf(d1 - r2)), fminf(fabsf(d2 - r2), fabsf(d3 - r2)));
  if (fminf(m_cell, m_side) < DT_EPSF || m_corner < 2.0f * r * DT_EPSF) return DT_AMBIG;
  const unsigned pre = (nb & NB_SELF) | (nb & gm & 0x1Eu);
  return (hits ? 1 : 0) | (pre ? 8 : 0) | ((nb & NB_ERR) ? 4 : 0);
}

__device__ __forceinline__ int dt_car_test_nbr(const uint8_t* __restrict__ grid, const uint16_t* __restrict__ nbr,
                                               int R, int C, float xf, float yf, float thf) {
  float snf, csf;
  sincosf(thf, &snf, &csf);
  const double ox = 0.075 * (double)csf, oy = 0.075 * (double)snf;
  const double x = (double)xf, y = (double)yf;
  const int a = dt_ball_test_nbr(nbr, R, C, x + ox, y + oy);
  const int b = dt_ball_test_nbr(nbr, R, C, x - ox, y - oy);
  if ((a | b) & DT_AMBIG) return dt_car_test(grid, R, C, xf, yf, thf);  // rare: decide with the exact code
  if ((a | b) & 2) return 1;
  const int raises = ((a | b) & 4) && !((a & 8) && (b & 8));
  return ((a | b) & 1) | (raises ? 4 : 0);
}
#endif  // __CUDACC__
