// probmap.cu -- the probability-map state sampler of run_type >= 2:
//   dt_edt_prior     CarEnv.prior = distance_transform_edt(1 - maze) / sum       (car_env.py:100-101,120-121)
//   dt_prob_map      gaussian_map + combine_log_blend                             (prob_sampling_utils.py:48-93,150-172)
//   dt_sample_cells  np.random.choice(size, p = prob_map.ravel()) for given uniform draws (planners/base_planner.py:157-160)
// Grids are <= 128 x 128 cells (31 x 31 in the reference's data), so each map is one thread block's work; the
// sampler is one thread per draw.  All arithmetic float64, in the reference's order where the order is
// defined (cumulative sum), so cell indices are bit-exact vs NumPy given the same uniform draws.
#include "common.cuh"

#define PM_THREADS 256

__device__ __forceinline__ double pm_block_sum(double v, double* s_red) {
  // fixed-order tree reduction (deterministic)
  const int t = threadIdx.x;
  s_red[t] = v;
  __syncthreads();
  for (int o = PM_THREADS / 2; o > 0; o >>= 1) {
    if (t < o) s_red[t] = s_red[t] + s_red[t + o];
    __syncthreads();
  }
  const double r = s_red[0];
  __syncthreads();
  return r;
}

// Exact Euclidean distance transform by exhaustive search over the wall cells: the squared distance is an
// integer, its square root the correctly rounded float64 SciPy returns.  Foreground = cells whose value
// differs from 1 (SciPy treats every non-zero of `1 - maze` as foreground); without any wall cell SciPy 1.x's
// transform returns sqrt((row + 1)^2 + col^2) (an artefact of its feature-transform initialisation; no
// reference map is wall-free) -- reproduced below.
__global__ void __launch_bounds__(PM_THREADS)
k_edt_prior(MapView m, double* __restrict__ out) {
  extern __shared__ __align__(16) uint8_t s_map[];
  __shared__ uint64_t bar;
  __shared__ double s_red[PM_THREADS];
  __shared__ int s_any;
  dt_stage_map(s_map, &bar, m);
  const int n = m.rows * m.cols;
  if (threadIdx.x == 0) s_any = 0;
  __syncthreads();
  for (int c = threadIdx.x; c < n; c += PM_THREADS)
    if (s_map[c] == 1) s_any = 1;
  __syncthreads();
  const bool any_wall = s_any != 0;
  double part = 0.0;
  for (int c = threadIdx.x; c < n; c += PM_THREADS) {
    const int r0 = c / m.cols, c0 = c - r0 * m.cols;
    double d = 0.0;
    if (s_map[c] != 1) {
      long long best = -1;
      if (any_wall) {
        for (int w = 0; w < n; ++w) {
          if (s_map[w] != 1) continue;
          const int r1 = w / m.cols, c1 = w - r1 * m.cols;
          const long long dr = r1 - r0, dc = c1 - c0, q = dr * dr + dc * dc;
          if (best < 0 || q < best) best = q;
        }
      } else {
        const long long dr = r0 + 1, dc = c0;
        best = dr * dr + dc * dc;
      }
      d = sqrt((double)best);
    }
    out[c] = d;
    part += d;
  }
  const double tot = pm_block_sum(part, s_red);
  for (int c = threadIdx.x; c < n; c += PM_THREADS) out[c] = out[c] / tot;
}

// gaussian_map(robot, goal, size) then combine_log_blend(prior, pdf, beta): robot / goal are the (x, y) the
// reference passes (it indexes the pdf as [int(y), int(x)]).  2 x 2 covariance inverted in closed form.
__global__ void __launch_bounds__(PM_THREADS)
k_prob_map(int H, int W, const double* __restrict__ prior, double rx, double ry, double gx, double gy, double beta,
           double eps, double* __restrict__ out, double* __restrict__ gauss_out) {
  __shared__ double s_red[PM_THREADS];
  const int n = H * W;
  const double dx = gx - rx, dy = gy - ry;
  const double d = sqrt(dx * dx + dy * dy) + 1e-6;
  const bool far = d > 1e-6;                        // false only for robot == goal
  const double ux = far ? dx / d : 1.0, uy = far ? dy / d : 0.0;
  const double vx = -uy, vy = ux;
  const double mx = (rx + gx) / 2, my = (ry + gy) / 2;
  const double w = -exp(-d / 15) + 1;
  const double mean_x = (1 - w) * gx + w * mx, mean_y = (1 - w) * gy + w * my;
  const double sl = 1.0 + 0.7 * log1p(d), ss = 0.7 * sl;
  const double a2 = sl * sl, b2 = ss * ss;
  // Sigma = R diag(a2, b2) R^T with R = [u v]
  const double s00 = ux * ux * a2 + vx * vx * b2, s01 = ux * uy * a2 + vx * vy * b2, s11 = uy * uy * a2 + vy * vy * b2;
  const double det = s00 * s11 - s01 * s01;
  const double i00 = s11 / det, i01 = -s01 / det, i11 = s00 / det;
  const int hole = (int)ry * W + (int)rx;         // pdf[int(ry), int(rx)] = 0 (the reference raises if outside)
  double part = 0.0;
  for (int c = threadIdx.x; c < n; c += PM_THREADS) {
    const int yy = c / W, xx = c - yy * W;
    const double ex = (double)xx - mean_x, ey = (double)yy - mean_y;
    const double tx = ex * i00 + ey * i01, ty = ex * i01 + ey * i11;
    double p = exp(-0.5 * (tx * ex + ty * ey));
    if (c == hole) p = 0.0;
    gauss_out[c] = p;
    part += p;
  }
  const double gs = pm_block_sum(part, s_red);
  part = 0.0;
  for (int c = threadIdx.x; c < n; c += PM_THREADS) {
    const double g = gauss_out[c] / gs;
    gauss_out[c] = g;
    const double pr = prior[c];
    const double lp = beta * log(pr + eps) + (1.0 - beta) * log(g + eps);
    const double post = (pr > 0.0) ? exp(lp) : 0.0;
    out[c] = post;
    part += post;
  }
  const double s = pm_block_sum(part, s_red);
  if (s <= eps) {  // fallback chain of combine_log_blend: the prior itself, then uniform
    part = 0.0;
    for (int c = threadIdx.x; c < n; c += PM_THREADS) part += prior[c];
    const double ps = pm_block_sum(part, s_red);
    for (int c = threadIdx.x; c < n; c += PM_THREADS) out[c] = (ps <= eps) ? 1.0 / (double)n : prior[c] / ps;
  } else {
    for (int c = threadIdx.x; c < n; c += PM_THREADS) out[c] = out[c] / s;
  }
}

// combine_log_blend(prior, gauss, beta, obstacle_mask, eps) alone (prob_sampling_utils.py:150-172), for callers
// that bring their own Gaussian: post = exp(beta log(prior + eps) + (1 - beta) log(gauss + eps)) * (prior > 0),
// masked, normalised; fallbacks: the (masked) prior, then uniform over the mask.
__global__ void __launch_bounds__(PM_THREADS)
k_log_blend(int n, const double* __restrict__ prior, const double* __restrict__ gauss, const uint8_t* __restrict__ mask,
            double beta, double eps, double* __restrict__ out) {
  __shared__ double s_red[PM_THREADS];
  double part = 0.0;
  for (int c = threadIdx.x; c < n; c += PM_THREADS) {
    const double pr = prior[c];
    double post = (pr > 0.0) ? exp(beta * log(pr + eps) + (1.0 - beta) * log(gauss[c] + eps)) : 0.0;
    if (mask && !mask[c]) post = 0.0;
    out[c] = post;
    part += post;
  }
  double s = pm_block_sum(part, s_red);
  if (s <= eps) {
    part = 0.0;
    for (int c = threadIdx.x; c < n; c += PM_THREADS) {
      const double v = (mask && !mask[c]) ? 0.0 : prior[c];
      out[c] = v;
      part += v;
    }
    s = pm_block_sum(part, s_red);
    if (s <= eps) {
      part = 0.0;
      for (int c = threadIdx.x; c < n; c += PM_THREADS) {
        const double v = (mask && !mask[c]) ? 0.0 : 1.0;
        out[c] = v;
        part += v;
      }
      s = pm_block_sum(part, s_red);
    }
  }
  for (int c = threadIdx.x; c < n; c += PM_THREADS) out[c] = out[c] / s;
}

// cdf = cumsum(p) (sequential, NumPy's order), cdf /= cdf[-1]  -- one thread; n <= 16384
__global__ void k_cdf(const double* __restrict__ p, int n, double* __restrict__ cdf) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  double acc = 0.0;
  for (int i = 0; i < n; ++i) {
    acc = __dadd_rn(acc, p[i]);
    cdf[i] = acc;
  }
  const double last = cdf[n - 1];
  for (int i = 0; i < n; ++i) cdf[i] = __ddiv_rn(cdf[i], last);
}

// idx = searchsorted(cdf, u, side='right'): number of entries <= u
__global__ void __launch_bounds__(PM_THREADS)
k_sample_cells(const double* __restrict__ cdf, int n, const double* __restrict__ u, int64_t B, int32_t* __restrict__ idx) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const double v = u[b];
    int lo = 0, hi = n;
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (cdf[mid] <= v) lo = mid + 1; else hi = mid;
    }
    idx[b] = lo;
  }
}

extern "C" int dt_edt_prior(dt_ctx* ctx, double* prior_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!ctx->d_map) return dt_fail(ctx, DT_E_NOMAP, "dt_set_map has not been called");
  if (!prior_out) return dt_fail(ctx, DT_E_ARG, "dt_edt_prior: null pointer");
  MapView m = dt_map_view(ctx);
  k_edt_prior<<<1, PM_THREADS, m.bytes, (cudaStream_t)stream>>>(m, prior_out);
  DT_LAUNCH_CHECK("k_edt_prior");
  return DT_OK;
}

extern "C" int dt_prob_map(dt_ctx* ctx, int rows, int cols, const double* prior, double robot_x, double robot_y,
                           double goal_x, double goal_y, double beta, double* prob_out, double* gauss_out,
                           void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!prior || !prob_out || !gauss_out || rows < 1 || cols < 1 || (int64_t)rows * cols > DT_MAX_MAP_CELLS)
    return dt_fail(ctx, DT_E_ARG, "dt_prob_map: bad argument");
  // the reference zeroes pdf[int(ry), int(rx)]: NumPy raises IndexError outside the map
  if (!(robot_x > -(double)cols - 1 && robot_x < (double)cols && robot_y > -(double)rows - 1 && robot_y < (double)rows))
    return dt_fail(ctx, DT_E_INDEX, "dt_prob_map: robot cell outside the map (the reference raises IndexError)");
  if ((int)robot_x < 0 || (int)robot_y < 0)
    return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_prob_map: negative robot cell (NumPy would wrap the index)");
  k_prob_map<<<1, PM_THREADS, 0, (cudaStream_t)stream>>>(rows, cols, prior, robot_x, robot_y, goal_x, goal_y, beta,
                                                        1e-12, prob_out, gauss_out);
  DT_LAUNCH_CHECK("k_prob_map");
  return DT_OK;
}

extern "C" int dt_log_blend(dt_ctx* ctx, const double* prior, const double* gauss, const uint8_t* obstacle_mask, int n,
                            double beta, double eps, double* prob_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!prior || !gauss || !prob_out || n < 1 || n > DT_MAX_MAP_CELLS) return dt_fail(ctx, DT_E_ARG, "dt_log_blend: bad argument");
  k_log_blend<<<1, PM_THREADS, 0, (cudaStream_t)stream>>>(n, prior, gauss, obstacle_mask, beta, eps, prob_out);
  DT_LAUNCH_CHECK("k_log_blend");
  return DT_OK;
}

extern "C" int dt_sample_cells(dt_ctx* ctx, const double* prob, int n, const double* u, int64_t B, int32_t* idx_out,
                               void* stream) {
  if (!ctx) return DT_E_ARG;
  if (B <= 0) return DT_OK;
  if (!prob || !u || !idx_out || n < 1 || n > DT_MAX_MAP_CELLS) return dt_fail(ctx, DT_E_ARG, "dt_sample_cells: bad argument");
  int rc = dt_ensure_scratch(ctx, (size_t)n * sizeof(double));
  if (rc) return rc;
  cudaStream_t st = (cudaStream_t)stream;
  double* cdf = reinterpret_cast<double*>(ctx->d_scratch);
  k_cdf<<<1, 32, 0, st>>>(prob, n, cdf);
  DT_LAUNCH_CHECK("k_cdf");
  int64_t blocks = (B + PM_THREADS - 1) / PM_THREADS;
  if (blocks > (int64_t)ctx->sm_count * 8) blocks = (int64_t)ctx->sm_count * 8;
  k_sample_cells<<<(int)blocks, PM_THREADS, 0, st>>>(cdf, n, u, B, idx_out);
  DT_LAUNCH_CHECK("k_sample_cells");
  return DT_OK;
}
