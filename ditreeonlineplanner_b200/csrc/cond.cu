// cond.cu -- sampler conditioning vectors (policies/fm_policy.py:53-143), one thread per candidate.
#include "common.cuh"

struct CarNorm {
  double obs_mean[6], obs_std[6], act_mean[2], act_std[2];
};

__global__ void __launch_bounds__(256)
k_build_cond_car(const float* __restrict__ st, int64_t s_cand, int64_t s_comp, const float* __restrict__ prev,
                 const float* __restrict__ goal, int goal_stride, int64_t B, CarNorm nm, float map_size,
                 float* __restrict__ out) {
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    const float* s = st + b * s_cand;
    float* o = out + b * 7;
    // (obs - mean) / std in float64, then the cast to float32 the reference does (fm_policy.py:74,112)
#pragma unroll
    for (int d = 3; d < 6; ++d) o[d - 3] = (float)(((double)s[d * s_comp] - nm.obs_mean[d]) / nm.obs_std[d]);
    if (prev) {
      o[3] = (float)(((double)prev[b * 2 + 0] - nm.act_mean[0]) / nm.act_std[0]);
      o[4] = (float)(((double)prev[b * 2 + 1] - nm.act_mean[1]) / nm.act_std[1]);
    } else {  // zeros, left un-normalised (fm_policy.py:114-121)
      o[3] = 0.f;
      o[4] = 0.f;
    }
    // robot-frame goal: float64 difference -> float32, rotation by -yaw and tanh in float32 (:126-143)
    const float gx = (float)((double)goal[b * goal_stride + 0] - (double)s[0]);
    const float gy = (float)((double)goal[b * goal_stride + 1] - (double)s[s_comp]);
    float sn, cs;
    sincosf(s[2 * s_comp], &sn, &cs);
    const float rx = __fadd_rn(__fmul_rn(cs, gx), __fmul_rn(sn, gy));
    const float ry = __fadd_rn(__fmul_rn(-sn, gx), __fmul_rn(cs, gy));
    o[5] = tanhf(rx / map_size);
    o[6] = tanhf(ry / map_size);
  }
}

struct AntNorm {
  double obs_mean[27], obs_std[27], act_mean[8], act_std[8];
};

// obs_seq (B,h,29): x, y, then 27 normalised dims of which slots 1..4 (state[3:7]) are the
// quaternion; per history slot the feature is [z_n, rot6d(6), 22 dims] = 29.
__global__ void __launch_bounds__(128)
k_build_cond_ant(const float* __restrict__ obs, int h, int H, const float* __restrict__ prev,
                 const float* __restrict__ goal, int goal_stride, int64_t B, AntNorm nm, float map_size,
                 float* __restrict__ out) {
  const int G = H * 29 + 10;
  for (int64_t b = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; b < B; b += (int64_t)gridDim.x * blockDim.x) {
    float* o = out + b * G;
    const int pad = H - h;
    for (int slot = 0; slot < H; ++slot) {
      float* f = o + slot * 29;
      const int src = (pad > 0) ? slot - pad : (h - H) + slot;
      if (src < 0) {
        for (int d = 0; d < 29; ++d) f[d] = 0.f;
        continue;
      }
      const float* s = obs + (b * h + src) * 29;
      double n[27];
#pragma unroll
      for (int d = 0; d < 27; ++d) n[d] = ((double)s[2 + d] - nm.obs_mean[d]) / nm.obs_std[d];
      f[0] = (float)n[0];
      // rot6d of the normalised slots read as (x, y, z, w) (common/se3_utils.py:177-189)
      const double qx = n[1], qy = n[2], qz = n[3], qw = n[4];
      f[1] = (float)(1 - 2 * (qy * qy + qz * qz));
      f[2] = (float)(2 * (qx * qy + qw * qz));
      f[3] = (float)(2 * (qx * qz - qw * qy));
      f[4] = (float)(2 * (qx * qy - qw * qz));
      f[5] = (float)(1 - 2 * (qx * qx + qz * qz));
      f[6] = (float)(2 * (qy * qz + qw * qx));
#pragma unroll
      for (int d = 5; d < 27; ++d) f[7 + d - 5] = (float)n[d];
    }
    float* a = o + H * 29;
    for (int d = 0; d < 8; ++d)
      a[d] = prev ? (float)(((double)prev[b * 8 + d] - nm.act_mean[d]) / nm.act_std[d]) : 0.f;
    const float* last = obs + (b * h + (h - 1)) * 29;
    const float gx = (float)((double)goal[b * goal_stride + 0] - (double)last[0]);
    const float gy = (float)((double)goal[b * goal_stride + 1] - (double)last[1]);
    a[8] = tanhf(gx / map_size);   // yaw = 0 for the ant: identity rotation (fm_policy.py:81)
    a[9] = tanhf(gy / map_size);
  }
}

extern "C" int dt_build_cond_car(dt_ctx* ctx, const float* state, int64_t s_cand, int64_t s_comp,
                                 const float* prev_action, const float* goal, int goal_stride, int64_t B,
                                 const double* norm_host, double map_size, float* cond_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (B <= 0) return DT_OK;
  if (!state || !goal || !norm_host || !cond_out || (goal_stride != 0 && goal_stride != 2))
    return dt_fail(ctx, DT_E_ARG, "dt_build_cond_car: bad argument");
  CarNorm nm;
  memcpy(nm.obs_mean, norm_host, 6 * sizeof(double));
  memcpy(nm.obs_std, norm_host + 6, 6 * sizeof(double));
  memcpy(nm.act_mean, norm_host + 12, 2 * sizeof(double));
  memcpy(nm.act_std, norm_host + 14, 2 * sizeof(double));
  int64_t blocks = (B + 255) / 256;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  k_build_cond_car<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(state, s_cand, s_comp, prev_action, goal,
                                                                  goal_stride, B, nm, (float)map_size, cond_out);
  DT_LAUNCH_CHECK("k_build_cond_car");
  return DT_OK;
}

extern "C" int dt_build_cond_ant(dt_ctx* ctx, const float* obs_seq, int h, int obs_history, const float* prev_action,
                                 const float* goal, int goal_stride, int64_t B, const double* norm_host,
                                 double map_size, float* cond_out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (B <= 0) return DT_OK;
  if (!obs_seq || !goal || !norm_host || !cond_out || h < 1 || obs_history < 1 ||
      (goal_stride != 0 && goal_stride != 2))
    return dt_fail(ctx, DT_E_ARG, "dt_build_cond_ant: bad argument");
  AntNorm nm;
  memcpy(nm.obs_mean, norm_host, 27 * sizeof(double));
  memcpy(nm.obs_std, norm_host + 27, 27 * sizeof(double));
  memcpy(nm.act_mean, norm_host + 54, 8 * sizeof(double));
  memcpy(nm.act_std, norm_host + 62, 8 * sizeof(double));
  int64_t blocks = (B + 127) / 128;
  if (blocks > ctx->sm_count * 8) blocks = ctx->sm_count * 8;
  k_build_cond_ant<<<(int)blocks, 128, 0, (cudaStream_t)stream>>>(obs_seq, h, obs_history, prev_action, goal,
                                                                  goal_stride, B, nm, (float)map_size, cond_out);
  DT_LAUNCH_CHECK("k_build_cond_ant");
  return DT_OK;
}
