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
// built by dt_set_map (ctx.cu) holds, per vertex and quadrant, one 32-bit word with the neighbourhood flags at
// the sign-bit positions of its four bytes; the signed margins of the three geometric tests are gathered
// into one register by two byte permutes (their top bytes carry the signs) and ANDed with that word, so the
// whole decision is branch-free and costs ~20 instructions per ball.
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
  const uint32_t* g;  // (rows + 2 DT_QPAD + 1) x (cols + 2 DT_QPAD + 1) vertices x 4 quadrant words (`bytes`);
                      // nullptr -> exact code only
  int bytes, pitch;
  int bias;           // -(bits of 1.5 * 2^23) * (pitch + 1): turns the two magic-add bit patterns into an index
  float cxp, cyp;     // grid coordinates of the padded map: w = x + cxp, u = cyp - y
  float wmax, umax;   // clamp of the car centre: [0.5, wmax] x [0.5, umax]
  float eps;
  uint32_t amb_t3;    // T * 0x010101 with 2^(2T - 127) >= eps: a margin whose top byte (sign | exponent >> 1), sign
                      // dropped, is below T has magnitude < 2^(2T - 127) -- the guard band as an exponent test
};

static inline QMapView dt_qmap_view_of(const uint32_t* d_qmap, int qmap_bytes, int rows, int cols) {
  QMapView q;
  q.g = qmap_bytes ? d_qmap : nullptr;
  q.bytes = qmap_bytes;
  q.pitch = cols + 2 * DT_QPAD + 1;
  q.bias = (int)(0u - 0x4B400000u * (unsigned)(q.pitch + 1));
  q.cxp = 0.5f * (float)cols + (float)DT_QPAD;
  q.cyp = 0.5f * (float)rows + (float)DT_QPAD;
  q.wmax = (float)(cols + 2 * DT_QPAD) - 0.5f;
  q.umax = (float)(rows + 2 * DT_QPAD) - 0.5f;
  const int mx = rows > cols ? rows : cols;
  q.eps = 1.1920929e-7f * ((float)mx + 4.0f) + 0.075f * DT_SC_ERR + 1.0e-7f;
  {
    int ex = 0;
    frexpf(q.eps, &ex);                    // eps = m * 2^ex, m in [0.5, 1): eps < 2^ex
    int T = (127 + ex + 1) / 2;            // smallest T with 2 T - 127 >= ex
    if (T < 1) T = 1;
    if (T > 127) T = 127;
    q.amb_t3 = (uint32_t)T * 0x010101u;
  }
  return q;
}

static inline QMapView dt_qmap_view(const dt_ctx* ctx) {
  return dt_qmap_view_of(ctx->d_qmap, ctx->qmap_bytes, ctx->rows, ctx->cols);
}

// one staged map slot as the kernels see it (dt_ctx::d_map_table)
struct MapEntry {
  MapView m;
  QMapView q;
};

#ifdef __CUDACC__
// Stage the occupancy grid and the quadrant table with two bulk TMA copies on one mbarrier.
__device__ __forceinline__ void dt_stage_maps(uint8_t* dst_map, uint32_t* dst_q, uint64_t* bar, const MapView& m,
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

// sin / cos of a heading with |th| <= DT_SC_MAX: two-term Cody-Waite reduction to [-pi, pi] and the MUFU
// approximations; |error| <= DT_SC_ERR.  (Garbage, but finite or NaN, outside that range: callers check.)
__device__ __forceinline__ void dt_sincos_mufu(float th, float& sn, float& cs) {
  const float kf = __fadd_rn(__fmaf_rn(th, 0.15915494f, 12582912.0f), -12582912.0f);  // rint(th / 2pi)
  float r = __fmaf_rn(kf, -6.2831855f, th);
  r = __fmaf_rn(kf, 1.7484555e-7f, r);
  sn = __sinf(r);
  cs = __cosf(r);
}

// any heading: larger (or non-finite) ones take libm's sincosf
__device__ __forceinline__ void dt_sincos_fast(float th, float& sn, float& cs) {
  if (fabsf(th) <= DT_SC_MAX) {
    dt_sincos_mufu(th, sn, cs);
  } else {
    const float2 r = dt_sincos_slow(th);
    sn = r.x;
    cs = r.y;
  }
}

__device__ __forceinline__ uint32_t dt_prmt(uint32_t a, uint32_t b, uint32_t sel) {
  uint32_t d;
  asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
  return d;
}

// One ball at grid coordinates (w, u), both inside [0.4, pitch - 1.4].
//   hit : some flag bit set <=> the ball collides          (bits 7, 15, 23, 31)
//   amb : some flag bit set <=> a wall-dependent decision is inside the guard band   (bits 7, 15, 23)
//   cell: distance to the nearest cell border (a floor decision; ambiguous when < eps)
__device__ __forceinline__ void dt_ball_fast(uint32_t q_addr, const QMapView& q, float w, float u, uint32_t& hit,
                                             uint32_t& amb, float& cell) {
  const float MR = 12582912.0f;  // 1.5 * 2^23: adding it (round to nearest) leaves rint(v) in the mantissa
  const float tu = __fadd_rn(u, MR), tw = __fadd_rn(w, MR);
  const float su = __fsub_rn(u, __fsub_rn(tu, MR)), sw = __fsub_rn(w, __fsub_rn(tw, MR));  // exact, in [-0.5, 0.5]
  // vertex (tu, tw) -> 16-byte entry; quadrant word: sign of su -> +8, sign of sw -> +4 (bits 30 are 0: |s| < 2)
  const uint32_t quad = ((__float_as_uint(su) >> 28) | (__float_as_uint(sw) >> 29)) & 12u;
  const uint32_t addr = q_addr + 16u * (uint32_t)(__float_as_int(tu) * q.pitch + __float_as_int(tw)) + quad;
  uint32_t word;
  asm("ld.shared.u32 %0, [%1];" : "=r"(word) : "r"(addr));
  const float r = 0.1f;
  const float fx = fabsf(sw), fy = fabsf(su);
  const float tx = fx - r, ty = fy - r;                         // < 0: reaches past that cell edge
  const float td = __fmaf_rn(sw, sw, su * su) - r * r;          // < 0: inside the corner disc
  // top bytes (sign + exponent) of the margins, one per byte lane: the table word keeps only the signs
  const uint32_t sg = dt_prmt(dt_prmt(__float_as_uint(tx), __float_as_uint(ty), 0x7373u), __float_as_uint(td), 0x7710u);
  hit = (sg | 0x80000000u) & word;
  // guard band: |margin| < eps (0.4 eps for the corner test) makes the decision ambiguous.  Tested on the exponents
  // already gathered in `sg`: byte b of ((sg | 0x00808080) & 0x00ffffff) - T3 keeps bit 7 iff (exponent >> 1) >= T,
  // i.e. |margin| >= 2^(2T - 127) >= eps (no borrow crosses a byte: each byte is 128 + e - T >= 1); everything below
  // that power of two counts as ambiguous -- a wider band than eps (at most 4x), three instructions instead of six
  const uint32_t clear = (((sg | 0x00808080u) & 0x00ffffffu) - q.amb_t3);
  amb = ~clear & word & 0x00808080u;
  cell = fminf(fx, fy);
}

// exact decision for the rare ambiguous state (kept out of line: float64 sincos + 8 hypot)
static __device__ __noinline__ int dt_car_test_exact(const uint8_t* __restrict__ grid, int R, int C, float x, float y,
                                                     float th) {
  return dt_car_test(grid, R, C, x, y, th);
}

__device__ __forceinline__ uint32_t dt_qmap_addr(const uint32_t* s_q, const QMapView& q) {
  return dt_smem_u32(s_q) + 16u * (uint32_t)q.bias;
}

// Fast part of is_colliding_car given the heading's sine / cosine: returns the collision flag and sets
// `ambiguous` when the exact code (dt_car_test_exact) must decide instead.  Needs the quadrant table.
__device__ __forceinline__ bool dt_car_fast(uint32_t q_addr, const QMapView& q, float x, float y, float sn, float cs,
                                            bool& ambiguous) {
  // car centre in padded grid coordinates, clamped into the padding (NaN -> 0.5: a padding cell, collides)
  const float wc = fminf(fmaxf(__fadd_rn(x, q.cxp), 0.5f), q.wmax);
  const float uc = fminf(fmaxf(__fsub_rn(q.cyp, y), 0.5f), q.umax);
  uint32_t ha, hb, aa, ab;
  float ca, cb;
  dt_ball_fast(q_addr, q, __fmaf_rn(cs, 0.075f, wc), __fmaf_rn(sn, -0.075f, uc), ha, aa, ca);
  dt_ball_fast(q_addr, q, __fmaf_rn(cs, -0.075f, wc), __fmaf_rn(sn, 0.075f, uc), hb, ab, cb);
  ambiguous = ((aa | ab) != 0u) | (fminf(ca, cb) < q.eps);
  return (ha | hb) != 0u;
}

// is_colliding_car for one state (any heading); returns 0/1, or 1|4 when the reference would raise IndexError
__device__ __forceinline__ int dt_car_any(const uint8_t* __restrict__ grid, uint32_t q_addr, const QMapView& q, int R,
                                          int C, float x, float y, float th) {
  if (!q.g || !(fabsf(th) <= DT_SC_MAX)) return dt_car_test_exact(grid, R, C, x, y, th);
  float sn, cs;
  dt_sincos_mufu(th, sn, cs);
  bool amb;
  const bool hit = dt_car_fast(q_addr, q, x, y, sn, cs, amb);
  if (amb) return dt_car_test_exact(grid, R, C, x, y, th);
  return hit ? 1 : 0;
}
#endif  // __CUDACC__
