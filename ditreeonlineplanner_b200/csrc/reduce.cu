// reduce.cu -- nearest-neighbour search over the tree's SoA node buffers, final-node cost arg-min
// and the MPPI cost reduction: warp-shuffle (value, index) reductions, lowest index on ties
// (NumPy argmin / KDTree-on-random-data semantics).
#include "common.cuh"

#define RED_THREADS 256

__device__ __forceinline__ void argmin_combine(double& d, int& i, double od, int oi) {
  if (od < d || (od == d && oi < i)) {
    d = od;
    i = oi;
  }
}

__device__ __forceinline__ void warp_argmin(double& d, int& i) {
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) {
    const double od = __shfl_xor_sync(0xffffffffu, d, off);
    const int oi = __shfl_xor_sync(0xffffffffu, i, off);
    argmin_combine(d, i, od, oi);
  }
}

// One warp per query, lanes stride over the nodes (coalesced SoA reads).  planners/RRT.py:49-55.
__global__ void __launch_bounds__(RED_THREADS)
k_nearest(const float* __restrict__ nx, const float* __restrict__ ny, int64_t n, const float* __restrict__ qx,
          const float* __restrict__ qy, int64_t q_stride, int64_t Q, int32_t* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t q = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); q < Q; q += (int64_t)gridDim.x * wpb) {
    const double x = (double)qx[q * q_stride], y = (double)qy[q * q_stride];
    double best = __longlong_as_double(0x7ff0000000000000LL);  // +inf
    int bi = 0x7fffffff;
    for (int64_t j = lane; j < n; j += 32) {
      const double dx = xsub(x, (double)nx[j]), dy = xsub(y, (double)ny[j]);
      const double d = xadd(xmul(dx, dx), xmul(dy, dy));
      if (d < best) {  // strict: the lowest index of a lane's stripe wins ties
        best = d;
        bi = (int)j;
      }
    }
    warp_argmin(best, bi);
    if (lane == 0) out[q] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

// k nearest nodes per query (KDTree.query(x, k) of planners/RRT.py:50 for k > 1): one warp per query; every
// lane keeps the DT_KNN_MAX best (d^2, index) pairs of its stripe sorted in registers, then the warp extracts
// the global k smallest with k rounds of shuffle arg-min (the winning lane pops its head).  Ascending
// distance, lowest index first on ties; slots beyond n get index n (SciPy's "missing neighbour" marker).
#define DT_KNN_MAX 16

template <int KMAX>
__global__ void __launch_bounds__(RED_THREADS)
k_nearest_k(const float* __restrict__ nx, const float* __restrict__ ny, int64_t n, const float* __restrict__ qx,
            const float* __restrict__ qy, int64_t q_stride, int64_t Q, int k, int32_t* __restrict__ out) {
  const double INF = __longlong_as_double(0x7ff0000000000000LL);
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t q = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); q < Q; q += (int64_t)gridDim.x * wpb) {
    const double x = (double)qx[q * q_stride], y = (double)qy[q * q_stride];
    double bd[KMAX];
    int bi[KMAX];
#pragma unroll
    for (int s = 0; s < KMAX; ++s) { bd[s] = INF; bi[s] = 0x7fffffff; }
    for (int64_t j = lane; j < n; j += 32) {
      const double dx = xsub(x, (double)nx[j]), dy = xsub(y, (double)ny[j]);
      double d = xadd(xmul(dx, dx), xmul(dy, dy));
      int id = (int)j;
      if (d < bd[KMAX - 1]) {   // insertion into the sorted list (indices ascend within a stripe: strict <)
#pragma unroll
        for (int s = 0; s < KMAX; ++s) {
          if (d < bd[s]) {
            const double td = bd[s]; const int ti = bi[s];
            bd[s] = d; bi[s] = id;
            d = td; id = ti;
          }
        }
      }
    }
    for (int r = 0; r < k; ++r) {
      double d = bd[0];
      int id = bi[0];
      warp_argmin(d, id);
      if (bi[0] == id && id != 0x7fffffff) {  // the owner pops its head
#pragma unroll
        for (int s = 0; s + 1 < KMAX; ++s) { bd[s] = bd[s + 1]; bi[s] = bi[s + 1]; }
        bd[KMAX - 1] = INF; bi[KMAX - 1] = 0x7fffffff;
      }
      if (lane == 0) out[q * k + r] = (id == 0x7fffffff) ? (int32_t)n : id;
    }
  }
}

// Block-wide arg-min of a per-node cost.  planners/RRT.py:233-237.
__global__ void __launch_bounds__(1024)
k_goal_cost_argmin(const float* __restrict__ nx, const float* __restrict__ ny, int64_t n, double gx, double gy,
                   const uint8_t* __restrict__ ahead, int32_t* __restrict__ out) {
  __shared__ double s_d[32];
  __shared__ int s_i[32];
  double best = __longlong_as_double(0x7ff0000000000000LL);
  int bi = 0x7fffffff;
  for (int64_t j = threadIdx.x; j < n; j += blockDim.x) {
    const double dx = xsub((double)nx[j], gx), dy = xsub((double)ny[j], gy);
    double c = __dsqrt_rn(xadd(xmul(dx, dx), xmul(dy, dy)));
    if (ahead) c = xadd(c, xmul(10e3, (double)(ahead[j] != 0)));
    if (c < best) {
      best = c;
      bi = (int)j;
    }
  }
  warp_argmin(best, bi);
  if ((threadIdx.x & 31) == 0) {
    s_d[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    const int nw = blockDim.x >> 5;
    best = threadIdx.x < nw ? s_d[threadIdx.x] : __longlong_as_double(0x7ff0000000000000LL);
    bi = threadIdx.x < nw ? s_i[threadIdx.x] : 0x7fffffff;
    warp_argmin(best, bi);
    if (threadIdx.x == 0) out[0] = (bi == 0x7fffffff) ? 0 : bi;
  }
}

// ---- MPPI ------------------------------------------------------------------------------------
// pass 1 (one block): min / argmin of the K costs.  scratch[0] = min (as float bits), argmin_out.
__global__ void __launch_bounds__(1024)
k_mppi_min(const float* __restrict__ cost, int64_t K, float* __restrict__ min_out, int32_t* __restrict__ argmin_out) {
  __shared__ double s_d[32];
  __shared__ int s_i[32];
  double best = __longlong_as_double(0x7ff0000000000000LL);
  int bi = 0x7fffffff;
  for (int64_t j = threadIdx.x; j < K; j += blockDim.x) {
    const double c = (double)cost[j];
    if (c < best) {
      best = c;
      bi = (int)j;
    }
  }
  warp_argmin(best, bi);
  if ((threadIdx.x & 31) == 0) {
    s_d[threadIdx.x >> 5] = best;
    s_i[threadIdx.x >> 5] = bi;
  }
  __syncthreads();
  if (threadIdx.x < 32) {
    best = s_d[threadIdx.x];
    bi = s_i[threadIdx.x];
    warp_argmin(best, bi);
    if (threadIdx.x == 0) {
      min_out[0] = (float)best;
      if (argmin_out) argmin_out[0] = bi;
    }
  }
}

// pass 2: each block reduces a slice of rollouts into partial[blk][0..TA) = sum w*noise and
// partial[blk][TA] = sum w (un-normalised weights w = exp(-(c - min)/lambda)).
#define MPPI_MAX_TA 128
__global__ void __launch_bounds__(RED_THREADS)
k_mppi_partial(const float* __restrict__ cost, const float* __restrict__ noise, int64_t K, int TA, float inv_lambda,
               const float* __restrict__ min_in, float* __restrict__ partial, float* __restrict__ weights_out) {
  __shared__ float s_w[RED_THREADS];
  __shared__ float s_acc[RED_THREADS / 32][MPPI_MAX_TA + 1];
  const float cmin = min_in[0];
  const int64_t per_block = (K + gridDim.x - 1) / gridDim.x;
  const int64_t k0 = blockIdx.x * per_block;
  const int64_t k1 = (k0 + per_block < K) ? k0 + per_block : K;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
  float wsum = 0.f;
  float acc[(MPPI_MAX_TA + 31) / 32];
#pragma unroll
  for (int c = 0; c < (MPPI_MAX_TA + 31) / 32; ++c) acc[c] = 0.f;
  for (int64_t base = k0; base < k1; base += blockDim.x) {
    const int64_t k = base + threadIdx.x;
    float w = 0.f;
    if (k < k1) {
      w = __expf(-(cost[k] - cmin) * inv_lambda);
      if (weights_out) weights_out[k] = w;
    }
    s_w[threadIdx.x] = w;
    wsum += w;
    __syncthreads();
    // warp `warp` accumulates rollouts base+warp, base+warp+nwarp, ...; lanes stride over TA (coalesced rows)
    const int cnt = (int)((k1 - base < blockDim.x) ? (k1 - base) : blockDim.x);
    for (int r = warp; r < cnt; r += nwarp) {
      const float wr = s_w[r];
      const float* row = noise + (base + r) * (int64_t)TA;
#pragma unroll
      for (int c = 0; c < (MPPI_MAX_TA + 31) / 32; ++c) {
        const int col = c * 32 + lane;
        if (col < TA) acc[c] += wr * row[col];
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int c = 0; c < (MPPI_MAX_TA + 31) / 32; ++c) {
    const int col = c * 32 + lane;
    if (col < TA) s_acc[warp][col] = acc[c];
  }
#pragma unroll
  for (int off = 16; off > 0; off >>= 1) wsum += __shfl_xor_sync(0xffffffffu, wsum, off);
  if (lane == 0) s_acc[warp][TA] = wsum;
  __syncthreads();
  for (int col = threadIdx.x; col <= TA; col += blockDim.x) {
    float s = 0.f;
    for (int w = 0; w < nwarp; ++w) s += s_acc[w][col];
    partial[blockIdx.x * (int64_t)(TA + 1) + col] = s;
  }
}

// pass 3 (one block): fixed-order sum of the partials, normalise, update u, normalise weights.
__global__ void __launch_bounds__(RED_THREADS)
k_mppi_final(const float* __restrict__ partial, int nblocks, int TA, float* __restrict__ u, float* __restrict__ wsum_out) {
  __shared__ float s_tot;
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int b = 0; b < nblocks; ++b) t += partial[b * (int64_t)(TA + 1) + TA];
    s_tot = t;
    wsum_out[0] = t;
  }
  __syncthreads();
  const float inv = 1.0f / s_tot;
  for (int col = threadIdx.x; col < TA; col += blockDim.x) {
    float s = 0.f;
    for (int b = 0; b < nblocks; ++b) s += partial[b * (int64_t)(TA + 1) + col];
    u[col] += s * inv;
  }
}

__global__ void k_mppi_norm_weights(float* __restrict__ w, int64_t K, const float* __restrict__ wsum) {
  const float inv = 1.0f / wsum[0];
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < K; i += (int64_t)gridDim.x * blockDim.x)
    w[i] *= inv;
}

// -------------------------------------------------------------------------------------------
extern "C" int dt_nearest(dt_ctx* ctx, const float* node_x, const float* node_y, int64_t n, const float* qx,
                          const float* qy, int64_t q_stride, int64_t Q, int32_t* idx_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (Q <= 0) return DT_OK;
  if (n <= 0 || !node_x || !node_y || !qx || !qy || !idx_out)
    return dt_fail(ctx, DT_E_ARG, "dt_nearest: bad argument");
  const int wpb = RED_THREADS / 32;
  int64_t blocks = (Q + wpb - 1) / wpb;
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  k_nearest<<<(int)blocks, RED_THREADS, 0, (cudaStream_t)stream>>>(node_x, node_y, n, qx, qy, q_stride, Q, idx_out);
  DT_LAUNCH_CHECK("k_nearest");
  return DT_OK;
}

extern "C" int dt_nearest_k(dt_ctx* ctx, const float* node_x, const float* node_y, int64_t n, const float* qx,
                            const float* qy, int64_t q_stride, int64_t Q, int k, int32_t* idx_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (Q <= 0) return DT_OK;
  if (n <= 0 || !node_x || !node_y || !qx || !qy || !idx_out || k < 1)
    return dt_fail(ctx, DT_E_ARG, "dt_nearest_k: bad argument");
  if (k > DT_KNN_MAX) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_nearest_k: k <= 16");
  const int wpb = RED_THREADS / 32;
  int64_t blocks = (Q + wpb - 1) / wpb;
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  cudaStream_t st = (cudaStream_t)stream;
  if (k <= 4) k_nearest_k<4><<<(int)blocks, RED_THREADS, 0, st>>>(node_x, node_y, n, qx, qy, q_stride, Q, k, idx_out);
  else k_nearest_k<DT_KNN_MAX><<<(int)blocks, RED_THREADS, 0, st>>>(node_x, node_y, n, qx, qy, q_stride, Q, k, idx_out);
  DT_LAUNCH_CHECK("k_nearest_k");
  return DT_OK;
}

extern "C" int dt_goal_cost_argmin(dt_ctx* ctx, const float* node_x, const float* node_y, int64_t n, float goal_x,
                                   float goal_y, const uint8_t* ahead, int32_t* idx_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (n <= 0 || !node_x || !node_y || !idx_out) return dt_fail(ctx, DT_E_ARG, "dt_goal_cost_argmin: bad argument");
  k_goal_cost_argmin<<<1, 1024, 0, (cudaStream_t)stream>>>(node_x, node_y, n, (double)goal_x, (double)goal_y, ahead,
                                                          idx_out);
  DT_LAUNCH_CHECK("k_goal_cost_argmin");
  return DT_OK;
}

extern "C" int dt_mppi_reduce(dt_ctx* ctx, const float* cost, const float* noise, int64_t K, int TA, float lambda,
                              float* u_inout, int32_t* argmin_out, float* weights_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (K <= 0 || TA <= 0 || TA > MPPI_MAX_TA || !cost || !noise || !u_inout || !(lambda > 0.f))
    return dt_fail(ctx, DT_E_ARG, "dt_mppi_reduce: bad argument");
  int nblocks = (int)((K + 4 * RED_THREADS - 1) / (4 * RED_THREADS));
  if (nblocks > ctx->sm_count * 2) nblocks = ctx->sm_count * 2;
  if (nblocks < 1) nblocks = 1;
  const size_t need = 256 + (size_t)nblocks * (TA + 1) * sizeof(float);
  int rc = dt_ensure_scratch(ctx, need);
  if (rc) return rc;
  float* minv = (float*)ctx->d_scratch;          // [0] min cost, [1] weight sum
  float* partial = (float*)((char*)ctx->d_scratch + 256);
  cudaStream_t st = (cudaStream_t)stream;
  k_mppi_min<<<1, 1024, 0, st>>>(cost, K, minv, argmin_out);
  DT_LAUNCH_CHECK("k_mppi_min");
  k_mppi_partial<<<nblocks, RED_THREADS, 0, st>>>(cost, noise, K, TA, 1.0f / lambda, minv, partial, weights_out);
  DT_LAUNCH_CHECK("k_mppi_partial");
  k_mppi_final<<<1, RED_THREADS, 0, st>>>(partial, nblocks, TA, u_inout, minv + 1);
  DT_LAUNCH_CHECK("k_mppi_final");
  if (weights_out) {
    int64_t blocks = (K + RED_THREADS - 1) / RED_THREADS;
    if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
    k_mppi_norm_weights<<<(int)blocks, RED_THREADS, 0, st>>>(weights_out, K, minv + 1);
    DT_LAUNCH_CHECK("k_mppi_norm_weights");
  }
  return DT_OK;
}
