// carfast.cuh -- fp32 fast path of is_colliding_car (common/map_utils.py:103-115 -> is_colliding_parallel
// :221-329) that provably returns the flag of the exact float64 code (dt_car_test, common.cuh).
//
// The reference decides, per ball centre: which cell holds it (two floors), whether it is within r = 0.1 of
// each cell edge whose neighbour is a wall (four comparisons) and within r of each cell corner whose diagonal
// neighbour is a wall (four hypot comparisons).  A ball is within r < 0.5 of at most ONE edge per axis, so
// only the cell quadrant the centre lies in matters.  In grid coordinates (w = x + C/2, u = R/2 - y, cell
// size 1) let (k, j) be the nearest grid vertex and (su, sw) in [-0.5, 0.5] the signed offsets from it:
// |sw|, |su| are the distances to the nearest vertical / horizontal cell edge, su^2 + sw^2 the squared
// distance to the nearest corner, and the signs select one of the four cells around the vertex.  A table
// built by dt_set_map (ctx.cu) holds, per vertex, a 4-bit neighbourhood code for each of those quadrants.
// Rounding to the nearest integer is one fp32 add of 1.5 * 2^23 (the integer lands in the mantissa).
//
// Everything is fp32; a decision whose margin to its threshold is below the guard band `eps` is
// "ambiguous" and the state is re-decided by the exact code.  Error budget of a ball centre in grid
// coordinates, for a car centre inside the padded map (|coordinate| <= max(R,C)/2 + 2, beyond that the
// centre is clamped: both balls are then >= 1.4 cells outside the grid, nowhere near a threshold):
//   sin/cos of the heading      <= DT_SC_ERR  (Cody-Waite reduction + MUFU; measured exhaustively over every
//                                              fp32 in [-8192, 8192] by tools/mufu_bound.cu)  -> x 0.075 m
//   shift of the centre to grid coordinates (one add)         <= 0.5 ulp(max(R,C) + 4)
//   centre +- 0.075 * sin|cos (one fma)                       <= 0.5 ulp(max(R,C) + 4)
//   nearest integer, signed offset : exact;  |offset| - r     <= 0.5 ulp(0.5) + |0.1f - 0.1| = 3.2e-8
// so  eps = 2^-23 * (max(R,C) + 4) + 0.075 * DT_SC_ERR + 1e-7   (3.1e-6 on a 20-cell map), computed on the
// host (dt_qmap_view).  The corner test compares su^2 + sw^2 with r^2; its sensitivity to a position error e
// is <= 2 (|su| + |sw|) e <= 0.29 e at the threshold, covered by a band of 0.4 eps.
#pragma once
#include "common.cuh"

#define DT_SC_ERR 1.5e-6f      // bound on |sin|,|cos| error of dt_sincos_fast for |theta| <= DT_SC_MAX
#define DT_SC_MAX 8192.0f

struct QMapView {
  const uint16_t* g;  // (rows + 2 DT_QPAD + 1) x (cols + 2 DT_QPAD + 1) vertex codes, padded to `bytes`;
                      // nullptr -> exact code only
  int bytes, pitch;
  int bias;           // -(bits of 1.5 * 2^23) * (pitch + 1): turns the two magic-add bit patterns into an index
  float cxp, cyp;     // grid coordinates of the padded map: w = x + cxp, u = cyp - y
  float wmax, umax;   // clamp of the car centre: [0.5, wmax] x [0.5, umax]
  float eps;
};

static inline QMapView dt_qmap_view(const dt_ctx* ctx) {
  QMapView q;
  q.g = ctx->qmap_bytes ? ctx->d_qmap : nullptr;
  q.bytes = ctx->qmap_bytes;
  q.pitch = ctx->cols + 2 * DT_QPAD + 1;
  q.bias = (int)(0u - 0x4B400000u * (unsigned)(q.pitch + 1));
  q.cxp = 0.5f * (float)ctx->cols + (float)DT_QPAD;
  q.cyp = 0.5f * (float)ctx->rows + (float)DT_QPAD;
  q.wmax = (float)(ctx->cols + 2 * DT_QPAD) - 0.5f;
  q.umax = (float)(ctx->rows + 2 * DT_QPAD) - 0.5f;
  const int mx = ctx->rows > ctx->cols ? ctx->rows : ctx->cols;
  q.eps = 1.1920929e-7f * ((float)mx + 4.0f) + 0.075f * DT_SC_ERR + 1.0e-7f;
  return q;
}

#ifdef __CUDACC__
// Stage the occupancy grid and the quadrant map with two bulk TMA copies on one mbarrier.
__device__ __forceinline__ void dt_stage_maps(uint8_t* dst_map, uint16_t* dst_q, uint64_t* bar, const MapView& m,
                                              const QMapView& q) {
  const uint32_t bar_a = dt_smem_u32(bar);
  if (threadIdx.x == 0) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar_a));
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t total = (uint32_t)m.bytes + (q.g ? (uint32_t)q.bytes : 0u);
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar_a), "r"(total) : "memory");
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            dt_smem_u32(dst_map)),
        "l"(m.g), "r"((uint32_t)m.bytes), "r"(bar_a)
        : "memory");
    if (q.g)
      asm volatile(
          "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
              dt_smem_u32(dst_q)),
          "l"(q.g), "r"((uint32_t)q.bytes), "r"(bar_a)
          : "memory");
  }
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

static __device__ __noinline__ float2 dt_sincos_slow(float th) {
  float2 r;
  sincosf(th, &r.x, &r.y);
  return r;
}

// sin / cos of a heading: two-term Cody-Waite reduction to [-pi, pi] and the MUFU approximations.
// |error| <= DT_SC_ERR for |th| <= DT_SC_MAX; larger (or non-finite) headings take libm's sincosf.
__device__ __forceinline__ void dt_sincos_fast(float th, float& sn, float& cs) {
  if (fabsf(th) <= DT_SC_MAX) {
    const float kf = __fadd_rn(__fmaf_rn(th, 0.15915494f, 12582912.0f), -12582912.0f);  // rint(th / 2pi)
    float r = __fmaf_rn(kf, -6.2831855f, th);
    r = __fmaf_rn(kf, 1.7484555e-7f, r);
    sn = __sinf(r);
    cs = __cosf(r);
  } else {
    const float2 r = dt_sincos_slow(th);
    sn = r.x;
    cs = r.y;
  }
}

// One ball at grid coordinates (w, u), both inside [0.4, pitch - 1.4].  Sets bit 0 of `res` when it collides
// and bit 1 when a decision is inside the guard band.  Branch-free.
__device__ __forceinline__ unsigned dt_ball_fast(uint32_t q_addr, const QMapView& q, float w, float u) {
  const float MR = 12582912.0f;  // 1.5 * 2^23: adding it (round to nearest) leaves rint(v) in the mantissa
  const float tu = __fadd_rn(u, MR), tw = __fadd_rn(w, MR);
  const float su = __fsub_rn(u, __fsub_rn(tu, MR)), sw = __fsub_rn(w, __fsub_rn(tw, MR));  // exact, in [-0.5, 0.5]
  const uint32_t addr = q_addr + 2u * (uint32_t)(__float_as_int(tu) * q.pitch + __float_as_int(tw));
  uint32_t code;
  asm("ld.shared.u16 %0, [%1];" : "=r"(code) : "r"(addr));
  // quadrant: sign of su -> bit 3, sign of sw -> bit 2 of the shift (bits 30 of both are 0: |s| < 2)
  const unsigned nib = code >> (((__float_as_uint(su) >> 28) | (__float_as_uint(sw) >> 29)) & 12u);
  const float r = 0.1f;
  const float fx = fabsf(sw), fy = fabsf(su);
  const float tx = fx - r, ty = fy - r;
  const float td = __fmaf_rn(sw, sw, su * su) - r * r;
  const bool b0 = (nib & 1u) != 0, b1 = (nib & 2u) != 0, b2 = (nib & 4u) != 0, b3 = (nib & 8u) != 0;
  const bool hit = b3 | (b0 & (tx < 0.0f)) | (b1 & (ty < 0.0f)) | (b2 & (td < 0.0f));
  // guard band: cell borders (fx, fy ~ 0), side thresholds (only where that side is a wall), corner disc
  const bool amb = (fminf(fx, fy) < q.eps) | (b0 & (fabsf(tx) < q.eps)) | (b1 & (fabsf(ty) < q.eps)) |
                   (b2 & (fabsf(td) < 0.4f * q.eps));
  return (hit ? 1u : 0u) | (amb ? 2u : 0u);
}

// exact decision for the rare ambiguous state (kept out of line: float64 sincos + 8 hypot)
static __device__ __noinline__ int dt_car_test_exact(const uint8_t* __restrict__ grid, int R, int C, float x, float y,
                                              float th) {
  return dt_car_test(grid, R, C, x, y, th);
}

// is_colliding_car given the heading's sine / cosine from dt_sincos_fast.  `q_addr` = dt_qmap_addr(...).
// Returns 0/1, or 1|4 when the reference would raise IndexError (only on maps without a quadrant table).
__device__ __forceinline__ uint32_t dt_qmap_addr(const uint16_t* s_q, const QMapView& q) {
  return dt_smem_u32(s_q) + 2u * (uint32_t)q.bias;
}

__device__ __forceinline__ int dt_car_fast(const uint8_t* __restrict__ grid, uint32_t q_addr, const QMapView& q, int R,
                                           int C, float x, float y, float th, float sn, float cs) {
  if (!q.g) return dt_car_test_exact(grid, R, C, x, y, th);
  // car centre in padded grid coordinates, clamped into the padding (NaN -> 0.5: a padding cell, collides)
  const float wc = fminf(fmaxf(__fadd_rn(x, q.cxp), 0.5f), q.wmax);
  const float uc = fminf(fmaxf(__fsub_rn(q.cyp, y), 0.5f), q.umax);
  const unsigned a = dt_ball_fast(q_addr, q, __fmaf_rn(cs, 0.075f, wc), __fmaf_rn(sn, -0.075f, uc));
  const unsigned b = dt_ball_fast(q_addr, q, __fmaf_rn(cs, -0.075f, wc), __fmaf_rn(sn, 0.075f, uc));
  if ((a | b) & 2u) return dt_car_test_exact(grid, R, C, x, y, th);
  return (int)((a | b) & 1u);
}
#endif  // __CUDACC__
