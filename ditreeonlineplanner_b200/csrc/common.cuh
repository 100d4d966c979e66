// common.cuh -- context, error plumbing and the shared-memory map staging used by every kernel.
#pragma once
#include <cuda_runtime.h>
#include <stdlib.h>
#include <cuda_bf16.h>
#include <stdint.h>
#include <stdio.h>
#include <string.h>
#include <string>
#include <vector>

#include "../../include/ditree.h"

#define DT_MAX_MAP_CELLS 16384  // 128 x 128; the reference's grids are <= 31 x 31
#define DT_QPAD 2               // padding rings of the collision fast path's quadrant table (carfast.cuh)
#define DT_QMAP_MAX_BYTES (96u * 1024u)  // larger tables (maps beyond ~72 x 72) use the exact code only

struct dt_denoiser;  // denoiser.cu

// Additional occupancy grids staged next to the main one (dt_set_map_slot): the device-resident planner runs
// several scenarios -- on different mazes -- in ONE device pass, a block of candidates per scenario, and every
// block stages the grid of its scenario (a per-group map index resolved through a device table of views).
#define DT_MAX_MAP_SLOTS 32
struct dt_map_slot {
  uint8_t* d_map = nullptr;
  uint32_t* d_qmap = nullptr;
  int rows = 0, cols = 0, map_bytes = 0, qmap_bytes = 0;
  size_t map_cap = 0, qmap_cap = 0;
  double s_global = 1.0;
};

struct dt_ctx {
  int device = 0;
  std::string err;
  // occupancy grid staged as bytes (1 = wall), padded to a multiple of 16 bytes for bulk copies
  uint8_t* d_map = nullptr;
  int rows = 0, cols = 0;
  double s_global = 1.0;
  int map_bytes = 0;  // padded size
  // quadrant map of the car collision fast path (carfast.cuh), (rows+5) x (cols+5) x 4 uint32;
  // absent (qmap_bytes == 0) on tall maps where the reference's diagonal lookup can raise IndexError
  // and on maps too large for shared memory
  uint32_t* d_qmap = nullptr;
  int qmap_bytes = 0;
  size_t map_capacity = 0, qmap_capacity = 0;
  dt_map_slot slots[DT_MAX_MAP_SLOTS];
  void* d_map_table = nullptr;        // MapEntry[DT_MAX_MAP_SLOTS] (carfast.cuh), the views of the staged slots
  int slots_max_bytes = 0;            // max over slots of map_bytes + qmap_bytes (dynamic shared memory of a block)
  bool prop_attr_set = false;  // dynamic shared memory opt-in of the propagate kernels done on this device
  // device-side status word (DT_E_*), plus small scratch for reductions
  int* d_status = nullptr;
  int* h_status = nullptr;  // pinned
  void* d_scratch = nullptr;
  size_t scratch_bytes = 0;
  void* d_wide = nullptr;  // fp32 pre-normalisation scratch of the wide-GroupNorm GEMM fallback (gemm.cu)
  size_t wide_bytes = 0;
  void* d_splitk = nullptr;   // fixed-size fp32 scratch of the split-K path (gemm.cu)
  void* d_splitk2 = nullptr;  // the same for layers running concurrently on the denoiser's side stream
  bool pdl_now = true;        // pdl_on && this launch's ConvGemm::pdl (set by dt_conv_gemm)
  bool fork_on = true;        // dt_set_option(ctx, "fork", 0): no side stream for the residual 1 x 1 convs
  bool splitk_on = true;      // dt_set_option(ctx, "splitk", 0) restores batch-size independent bits
  // programmatic dependent launch for the denoiser's kernel chain (dt_set_option(ctx, "pdl", 0) or DITREE_PDL=0
  // turn it off): a kernel's CTAs are scheduled, and run their prologue (barrier init, TMEM allocation, tensor-map
  // prefetch), while the previous kernel's last CTAs drain; griddepcontrol.wait orders all memory traffic
  bool pdl_on = true;
  int64_t launches = 0;
  int sm_count = 148;
  dt_denoiser* den = nullptr;
  // optional per-launch event timing of the GEMM kernels (dt_profile_begin / dt_profile_end)
  bool prof_on = false;
  std::vector<cudaEvent_t> prof_events;  // pairs (start, stop)
  size_t prof_used = 0;
  struct ProfRec { int bn, epi, gw; long long M; int N; long long K; float ms; int ksplit; };  // epi 2 = k_splitk_epi
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

// Launch with (or without) the programmatic-stream-serialization attribute.  ONLY for kernels that execute
// dt_pdl_wait() before their first global-memory access: a kernel launched with the attribute that never waits
// would race with its predecessor.
template <typename... KArgs, typename... Args>
static inline cudaError_t dt_launch(bool pdl, void (*kern)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st,
                                    Args&&... args) {
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kern, static_cast<KArgs>(args)...);
}

int dt_ensure_scratch(dt_ctx* ctx, size_t bytes);
dt_model_cfg dt_denoiser_cfg(dt_ctx* ctx);  // denoiser.cu: configuration of the loaded denoiser (ctx->den != nullptr)

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

// Programmatic dependent launch (sm_90+).  dt_pdl_launch: the next kernel in the stream may be scheduled from now on
// (its CTAs take whatever SM resources are free and block in dt_pdl_wait).  dt_pdl_wait: returns when every kernel
// this one depends on has COMPLETED and its writes are visible; a no-op for a kernel launched without the attribute.
__device__ __forceinline__ void dt_pdl_launch() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void dt_pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

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

// check_obstacle_ahead (planners/RRT.py:61-81) for one state: 30 samples on linspace(0, 1.5) along
// (cos(-theta), sin(-theta)) from the robot's un-floored (col, row); truncation toward zero, clip, any wall.
__device__ __forceinline__ int dt_ray_probe_one(const uint8_t* __restrict__ grid, int R, int C, float xf, float yf,
                                                float thf) {
  const double cx = xdiv((double)C, 2.0), cy = xdiv((double)R, 2.0);  // the env's cell size is 1 for the car
  const double step = xdiv(1.5, 29.0);
  const double row = xsub(cy, (double)yf);
  const double col = xadd((double)xf, cx);
  double sn, cs;
  sincos(-(double)thf, &sn, &cs);
  int hit = 0;
  for (int k = 0; k < 30; ++k) {
    const double t = (k == 29) ? 1.5 : xmul((double)k, step);
    const double sx = xadd(xmul(t, cs), col), sy = xadd(xmul(t, sn), row);
    int qx = (int)sx, qy = (int)sy;  // astype('int'): truncation toward zero
    qx = dt_clampi(qx, 0, C - 1);
    qy = dt_clampi(qy, 0, R - 1);
    hit |= (grid[qy * C + qx] != 0);
  }
  return hit;
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

#endif  // __CUDACC__
