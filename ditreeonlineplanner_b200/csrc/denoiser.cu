// denoiser.cu -- the flow-matching denoiser on tcgen05 tensor cores.
//
// Reference modules restated here (file:line of the upstream repository):
//   ConditionalUnet1DWithLocalMap.forward        local_map_encoder.py:101-109
//   ResNet18Encoder (+ BatchNorm -> GroupNorm)    local_map_encoder.py:63-76,112-122
//   ConditionalUnet1D.forward                     model/diffusion/conditional_unet1d.py:268-347
//   ConditionalResidualBlock1D.forward            model/diffusion/conditional_unet1d.py:103-142
//   Conv1dBlock / Downsample1d / Upsample1d       model/diffusion/conv1d_components.py:7-40
//   SinusoidalPosEmb                              model/diffusion/positional_embedding.py:5-17
//   DiffusionSampler flow-matching loop           policies/fm_policy.py:183-203
//   get_timesteps('exp')                          common/fm_utils.py:4-17
//
// Layout: activations are bf16, channel-last (B, T, C); every conv is an implicit GEMM launched
// through dt_conv_gemm (gemm.cu).  Exact algebraic hoists: the encoder runs
// once per sample call, the time MLP once per step for the whole batch, the FiLM linear is split
// into a per-candidate part (one GEMM per call) and a per-step part (batch-shared), the encoder's
// first conv is folded over its three identical input channels, and each ConvTranspose1d(4,2,1)
// becomes two 2-tap convs writing the even / odd output rows.
#include <algorithm>
#include <map>
#include <math.h>

#include "gemm.cuh"

#define DEN_KMAX 64  // max ODE steps per call
#define DEN_FORK_MAXB 128   // side-stream forks up to this batch (measured: -7 % at 1, -5 % at 64, +1 % at 256)
#define DEN_GRAPH_MAXB 256  // sampler calls up to this batch are replayed as CUDA graphs (host cost ~0.3 -> 0.05 ms)

// ------------------------------------------------------------------------------------------
// packed parameters
// ------------------------------------------------------------------------------------------
struct ConvW {
  __nv_bfloat16* w = nullptr;  // [N][Ktot] bf16
  float* bias = nullptr;       // [N]
  float* gamma = nullptr;      // GroupNorm affine
  float* beta = nullptr;
  int N = 0, Ktot = 0;
};

struct ResBlockW {
  ConvW c1, c2, res;
  bool has_res = false;
  int cin = 0, cout = 0, film_off = 0;
};

struct EncBlockW {
  ConvW c1, c2, ds;
  bool has_ds = false;
  int cin = 0, cout = 0, stride = 1;
};

struct dt_denoiser {
  dt_model_cfg cfg;
  int A, T, G, E, NM, C[3], MB;
  int Tl[3];
  // U-Net
  ResBlockW down[3][2], mid[2], up[2][2];
  ConvW downs[2], ups[2][2], final_blk;
  float* final_w = nullptr;  // [A][C0] fp32
  float* final_b = nullptr;
  // FiLM
  int F = 0, kc_pad = 0;
  ConvW film_c;              // per-candidate part: N = F, K = kc_pad, bias = cond_encoder biases
  float* film_wt = nullptr;  // per-step part, fp32 [F][256]
  float *t_w1 = nullptr, *t_b1 = nullptr, *t_w2 = nullptr, *t_b2 = nullptr;  // time MLP fp32
  // encoder
  ConvW enc_conv1;
  EncBlockW enc[4][2];
  ConvW enc_fc;
  int emb_pad = 0;
  // scratch
  std::vector<void*> allocs;
  __nv_bfloat16 *X, *H0, *R0, *A0, *B0, *V0, *F0;
  __nv_bfloat16 *D0, *H1u, *R1u, *U1a, *U1b, *H1, *R1, *A1, *B1, *V1;
  __nv_bfloat16 *D1, *H2u, *R2u, *U2a, *U2b, *H2, *R2, *A2, *B2, *M2, *M1;
  float* film_cand = nullptr;  // [MB][F]
  float* film_time = nullptr;  // [DEN_KMAX][F]
  int film_time_K = 0;         // schedule the table currently holds (film_times)
  void* film_time_stream = nullptr;
  float film_time_ts[64];
  // CUDA graphs of small-batch sampler calls (dt_fm_sample): static input / output buffers + one executable
  // graph per (B, K, schedule, un-normalise) key; a key runs eagerly once (first-use initialisation is not
  // capturable), is captured on its second call and replayed afterwards
  float* g_noise = nullptr;    // [GB][T][A]
  float* g_cond = nullptr;     // [GB][G]
  __nv_bfloat16* g_lm = nullptr;  // [GB][NM][NM]
  float* g_out = nullptr;      // [GB][T][A]
  struct GraphEntry { int B, K; unsigned ebits; int norm; int state; /* 0 seen once, 1 graph, -1 not capturable */
                      cudaGraphExec_t exec; long long launches; float* film_table; /* [K][F], filled at capture time */ };
  std::vector<GraphEntry> graphs;
  // side stream of small-batch passes: a residual block's 1 x 1 residual conv runs beside its first 3-tap conv
  // (fork / join by events; inside a captured graph these become parallel branches)
  cudaStream_t side = nullptr;
  cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
  cudaStream_t cap_stream = nullptr;
  float* norm_dev = nullptr;   // [2 A] action mean / std on the device, cached (dt_fm_sample)
  float norm_host[16];
  bool norm_valid = false;
  float* mish_t = nullptr;     // [DEN_KMAX][256]
  float* time_h = nullptr;     // [DEN_KMAX][1024] hidden layer of the time MLP
  __nv_bfloat16* cond_in = nullptr;  // [MB][kc_pad]
  __nv_bfloat16* col = nullptr;      // im2col buffer
  __nv_bfloat16* e[4] = {nullptr, nullptr, nullptr, nullptr};
  float* emb = nullptr;              // [MB][emb_pad]
  float* lin = nullptr;              // [MB][emb_pad] raw fc output / generic f32 scratch
};

// ------------------------------------------------------------------------------------------
// small CUDA-core kernels around the GEMMs
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ float mishf(float x) {
  const float e = __expf(fminf(x, 20.0f));
  const float n = e * (e + 2.0f);
  return x * __fdividef(n, n + 2.0f);
}

// fp32 sample (B,T,A) -> bf16 (B,T,64), channels >= A zero
__global__ void k_prep_sample(const float* __restrict__ a, int64_t rows, int A, __nv_bfloat16* __restrict__ x) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int64_t n = rows * 64;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i & 63);
    const int64_t r = i >> 6;
    x[i] = __float2bfloat16(c < A ? a[r * A + c] : 0.f);
  }
}

// final 1x1 conv (C0 -> A) fused with the Euler update  a += v * dt  (fm_policy.py:194) or, for
// the test hook, a plain store of v; last step un-normalises (fm_policy.py:202-203).  One warp per row.
__global__ void __launch_bounds__(256)
k_final_euler(const __nv_bfloat16* __restrict__ f, int64_t rows, int C0, int A, const float* __restrict__ w,
              const float* __restrict__ bias, float dt, float* __restrict__ a, float* __restrict__ vel_out,
              const float* __restrict__ norm /* mean[A], std[A] or null */) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  for (int64_t r = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    const __nv_bfloat16* row = f + r * C0;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    for (int c = lane * 2; c < C0; c += 64) {
      const float2 v = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + c));
#pragma unroll
      for (int j = 0; j < 8; ++j)
        if (j < A) acc[j] += v.x * w[j * C0 + c] + v.y * w[j * C0 + c + 1];
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      if (j < A) {
#pragma unroll
        for (int off = 16; off > 0; off >>= 1) acc[j] += __shfl_xor_sync(0xffffffffu, acc[j], off);
      }
    }
    if (lane < A) {
      float v = 0.f;
#pragma unroll
      for (int j = 0; j < 8; ++j) v = (lane == j) ? acc[j] : v;
      v += bias[lane];
      if (vel_out) {
        vel_out[r * A + lane] = v;
      } else {
        float s = a[r * A + lane] + v * dt;
        if (norm) s = s * norm[A + lane] + norm[lane];
        a[r * A + lane] = s;
      }
    }
  }
}

// Specialisation for A = 2 (car): each lane keeps the weights of its channels in registers, so a
// row costs NI coalesced 4-byte loads and 4*NI FMAs (the generic kernel re-reads the weights per row).
template <int NI>  // C0 = 64 * NI
__global__ void __launch_bounds__(256)
k_final_euler_a2(const __nv_bfloat16* __restrict__ f, int64_t rows, const float* __restrict__ w,
                 const float* __restrict__ bias, float dt, float* __restrict__ a, float* __restrict__ vel_out,
                 const float* __restrict__ norm) {
  dt_pdl_launch();
  dt_pdl_wait();
  constexpr int C0 = 64 * NI;
  const int lane = threadIdx.x & 31;
  const int wpb = blockDim.x >> 5;
  float w0[NI][2], w1[NI][2];
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int c = lane * 2 + 64 * i;
    w0[i][0] = w[c]; w0[i][1] = w[c + 1];
    w1[i][0] = w[C0 + c]; w1[i][1] = w[C0 + c + 1];
  }
  const float b0 = bias[0], b1 = bias[1];
  for (int64_t r = blockIdx.x * (int64_t)wpb + (threadIdx.x >> 5); r < rows; r += (int64_t)gridDim.x * wpb) {
    const __nv_bfloat16* row = f + r * C0;
    float2 v[NI];
#pragma unroll
    for (int i = 0; i < NI; ++i) v[i] = __bfloat1622float2(*reinterpret_cast<const __nv_bfloat162*>(row + lane * 2 + 64 * i));
    float s0 = 0.f, s1 = 0.f;
#pragma unroll
    for (int i = 0; i < NI; ++i) {
      s0 += v[i].x * w0[i][0] + v[i].y * w0[i][1];
      s1 += v[i].x * w1[i][0] + v[i].y * w1[i][1];
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s0 += __shfl_xor_sync(0xffffffffu, s0, off);
      s1 += __shfl_xor_sync(0xffffffffu, s1, off);
    }
    if (lane < 2) {
      float vv = (lane == 0 ? s0 + b0 : s1 + b1);
      if (vel_out) {
        vel_out[r * 2 + lane] = vv;
      } else {
        float st = a[r * 2 + lane] + vv * dt;
        if (norm) st = st * norm[2 + lane] + norm[lane];
        a[r * 2 + lane] = st;
      }
    }
  }
}

static int launch_final_euler(dt_ctx* ctx, const __nv_bfloat16* f, int64_t rows, int C0, int A, const float* w,
                              const float* bias, float dt, float* a, float* vel_out, const float* norm,
                              cudaStream_t st);

// Mish([emb | cond]) as the bf16 A operand of the per-candidate FiLM GEMM, zero padded to kc_pad
__global__ void k_film_input(const float* __restrict__ emb, int emb_ld, int E, const float* __restrict__ cond, int G,
                             int64_t B, int kc_pad, __nv_bfloat16* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int64_t n = B * kc_pad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % kc_pad);
    const int64_t b = i / kc_pad;
    float v = 0.f;
    if (c < E) v = mishf(emb[b * emb_ld + c]);
    else if (c < E + G) v = mishf(cond[b * G + (c - E)]);
    out[i] = __float2bfloat16(v);
  }
}

// three word-wise copies in one launch (the inputs of a graph-replayed sampler call into its static buffers)
__global__ void k_stage3(const uint32_t* __restrict__ s0, uint32_t* __restrict__ d0, size_t n0,
                         const uint32_t* __restrict__ s1, uint32_t* __restrict__ d1, size_t n1,
                         const uint32_t* __restrict__ s2, uint32_t* __restrict__ d2, size_t n2) {
  const size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x;
  if (i < n0) d0[i] = s0[i];
  else if (i < n0 + n1) d1[i - n0] = s1[i - n0];
  else if (i < n0 + n1 + n2) d2[i - n0 - n1] = s2[i - n0 - n1];
}

// time MLP for all K steps: Mish(Linear(Mish(Linear(sinusoid(t_k))))) -> mish_t[k][256]
// (the outer Mish is the one cond_encoder applies to the global feature, conditional_unet1d.py:69)
struct TimeSteps {
  float t[DEN_KMAX];  // passed by value: no host staging buffer whose lifetime a launch would have to outlive
};

// Two launches, one warp per output neuron (coalesced reads of its weight row, shuffle reduction): the 2 MB of fp32
// weights are pulled by 128 / 32 blocks per step instead of one (one block per step took 275 us -- a fifth of a
// graph-replayed small-batch sampler call, which recomputes the table every time).
__global__ void __launch_bounds__(256)
k_time_mlp1(TimeSteps ts, const float* __restrict__ w1, const float* __restrict__ b1, float* __restrict__ hid) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x >> 7;                                     // 128 blocks of 8 neurons per step
  const int o = ((blockIdx.x & 127) << 3) + (threadIdx.x >> 5);      // hidden neuron 0..1023
  const float t = ts.t[k];
  const float* wr = w1 + o * 256;
  float acc = 0.f;
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int c = j * 32 + lane;                                     // sinusoid: sin on 0..127, cos on 128..255
    const float arg = t * expf((float)(c & 127) * -(logf(10000.0f) / 127.0f));
    acc += wr[c] * (c < 128 ? sinf(arg) : cosf(arg));
  }
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) hid[k * 1024 + o] = mishf(acc + b1[o]);
}

__global__ void __launch_bounds__(256)
k_time_mlp2(const float* __restrict__ hid, const float* __restrict__ w2, const float* __restrict__ b2,
            float* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int k = blockIdx.x >> 5;                                     // 32 blocks of 8 outputs per step
  const int o = ((blockIdx.x & 31) << 3) + (threadIdx.x >> 5);       // output 0..255
  const float* wr = w2 + o * 1024;
  const float* h = hid + k * 1024;
  float acc = 0.f;
#pragma unroll 8
  for (int j = 0; j < 32; ++j) acc += wr[j * 32 + lane] * h[j * 32 + lane];
#pragma unroll
  for (int d = 16; d > 0; d >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, d);
  if (lane == 0) out[k * 256 + o] = mishf(acc + b2[o]);
}

// per-step FiLM part: film_time[k][n] = sum_j Wt[n][j] * mish_t[k][j]; one warp per n
__global__ void __launch_bounds__(256)
k_film_time(const float* __restrict__ wt, const float* __restrict__ mish_t, int F, int K, float* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int lane = threadIdx.x & 31;
  const int n = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (n >= F) return;
  float w[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) w[j] = wt[n * 256 + j * 32 + lane];
  for (int k = 0; k < K; ++k) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < 8; ++j) acc += w[j] * mish_t[k * 256 + j * 32 + lane];
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) acc += __shfl_xor_sync(0xffffffffu, acc, off);
    if (lane == 0) out[(int64_t)k * F + n] = acc;
  }
}

// ---- encoder helpers (channel-last bf16 images (B, H, W, C)) ---------------------------------
// im2col: rows (b, oy, ox), K order (ky, kx, c), zero padded to kpad
__global__ void k_im2col(const __nv_bfloat16* __restrict__ in, int64_t B, int H, int W, int C, int k, int stride,
                         int pad, int OH, int OW, int kpad, __nv_bfloat16* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int64_t n = B * OH * OW * (int64_t)kpad;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int kk = (int)(i % kpad);
    int64_t r = i / kpad;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int64_t b = r / OH;
    __nv_bfloat16 v = __float2bfloat16(0.f);
    if (kk < k * k * C) {
      const int c = kk % C;
      const int tap = kk / C;
      const int ky = tap / k, kx = tap % k;
      const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
      if (iy >= 0 && iy < H && ix >= 0 && ix < W) v = in[((b * H + iy) * W + ix) * C + c];
    }
    out[i] = v;
  }
}

// vectorised im2col for C % 8 == 0 (every encoder conv except conv1): one 16-byte copy per thread
__global__ void k_im2col_v8(const __nv_bfloat16* __restrict__ in, int64_t B, int H, int W, int C, int k, int stride,
                            int pad, int OH, int OW, __nv_bfloat16* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int c8n = C / 8;
  const int64_t n = B * OH * OW * (int64_t)(k * k) * c8n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c8 = (int)(i % c8n);
    int64_t r = i / c8n;
    const int tap = (int)(r % (k * k));
    r /= (k * k);
    const int ox = (int)(r % OW);
    int64_t r2 = r / OW;
    const int oy = (int)(r2 % OH);
    const int64_t b = r2 / OH;
    const int ky = tap / k, kx = tap - ky * k;
    const int iy = oy * stride - pad + ky, ix = ox * stride - pad + kx;
    uint4 v = make_uint4(0u, 0u, 0u, 0u);
    if (iy >= 0 && iy < H && ix >= 0 && ix < W)
      v = *reinterpret_cast<const uint4*>(in + ((b * H + iy) * W + ix) * C + c8 * 8);
    *reinterpret_cast<uint4*>(out + (r * (int64_t)(k * k) + tap) * C + c8 * 8) = v;
  }
}

// GroupNorm(C/16 groups) over (HW x 16 channels) per sample, optional residual add, optional ReLU.
// x: (B, HW, C) bf16 conv output (bias-free convs in resnet).  One warp per (sample, group).
__global__ void __launch_bounds__(256)
k_gn2d(const __nv_bfloat16* __restrict__ x, int64_t B, int HW, int C, const float* __restrict__ gamma,
       const float* __restrict__ beta, const __nv_bfloat16* __restrict__ resid, int relu,
       __nv_bfloat16* __restrict__ out) {
  const int lane = threadIdx.x & 31;
  const int groups = C / 16;
  const int64_t total = B * groups;
  for (int64_t wg = blockIdx.x * (int64_t)(blockDim.x >> 5) + (threadIdx.x >> 5); wg < total;
       wg += (int64_t)gridDim.x * (blockDim.x >> 5)) {
    const int64_t b = wg / groups;
    const int gi = (int)(wg % groups);
    const int n = HW * 16;
    float s = 0.f, q = 0.f;
    for (int i = lane; i < n; i += 32) {
      const float v = __bfloat162float(x[(b * HW + i / 16) * C + gi * 16 + (i & 15)]);
      s += v;
      q += v * v;
    }
#pragma unroll
    for (int off = 16; off > 0; off >>= 1) {
      s += __shfl_xor_sync(0xffffffffu, s, off);
      q += __shfl_xor_sync(0xffffffffu, q, off);
    }
    const float mean = s / n;
    const float rstd = rsqrtf(fmaxf(q / n - mean * mean, 0.f) + 1e-5f);
    for (int i = lane; i < n; i += 32) {
      const int c = gi * 16 + (i & 15);
      const int64_t idx = (b * HW + i / 16) * C + c;
      float v = (__bfloat162float(x[idx]) - mean) * rstd * gamma[c] + beta[c];
      if (resid) v += __bfloat162float(resid[idx]);
      if (relu) v = fmaxf(v, 0.f);
      out[idx] = __float2bfloat16(v);
    }
  }
}

__global__ void k_maxpool3s2(const __nv_bfloat16* __restrict__ in, int64_t B, int H, int W, int C, int OH, int OW,
                             __nv_bfloat16* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int64_t n = B * OH * OW * (int64_t)C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t r = i / C;
    const int ox = (int)(r % OW);
    r /= OW;
    const int oy = (int)(r % OH);
    const int64_t b = r / OH;
    float m = -INFINITY;
    for (int ky = 0; ky < 3; ++ky)
      for (int kx = 0; kx < 3; ++kx) {
        const int iy = oy * 2 - 1 + ky, ix = ox * 2 - 1 + kx;
        if (iy >= 0 && iy < H && ix >= 0 && ix < W) m = fmaxf(m, __bfloat162float(in[((b * H + iy) * W + ix) * C + c]));
      }
    out[i] = __float2bfloat16(m);
  }
}

__global__ void k_avgpool(const __nv_bfloat16* __restrict__ in, int64_t B, int HW, int C,
                          __nv_bfloat16* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int64_t n = B * C;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    const int64_t b = i / C;
    float s = 0.f;
    for (int p = 0; p < HW; ++p) s += __bfloat162float(in[(b * HW + p) * C + c]);
    out[i] = __float2bfloat16(s / HW);
  }
}

// local map (B,N,N) bf16 -> itself viewed as (B, N, N, 1); conv1's im2col reads it directly (C = 1)
__global__ void k_copy_cols(const float* __restrict__ in, int ld_in, int64_t B, int n, float* __restrict__ out) {
  dt_pdl_launch();
  dt_pdl_wait();
  const int64_t tot = B * n;
  for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < tot; i += (int64_t)gridDim.x * blockDim.x)
    out[i] = in[(i / n) * ld_in + (i % n)];
}

// ------------------------------------------------------------------------------------------
// loading: reference state_dict -> packed bf16 operands
// ------------------------------------------------------------------------------------------
typedef std::map<std::string, const dt_tensor_desc*> TMap;

struct Loader {
  dt_ctx* ctx;
  dt_denoiser* d;
  const TMap& tm;
  cudaStream_t st;
  std::string err;

  const dt_tensor_desc* get(const std::string& name, int ndim, std::initializer_list<int64_t> shape) {
    auto it = tm.find(name);
    if (it == tm.end()) {
      if (err.empty()) err = "missing tensor " + name;
      return nullptr;
    }
    const dt_tensor_desc* t = it->second;
    bool ok = t->ndim == ndim;
    int i = 0;
    for (int64_t s : shape) ok = ok && (t->shape[i++] == s);
    if (!ok && err.empty()) err = "unexpected shape for " + name;
    return ok ? t : nullptr;
  }
  template <typename TT>
  TT* upload(const std::vector<TT>& h) {
    void* p = nullptr;
    if (cudaMalloc(&p, h.size() * sizeof(TT)) != cudaSuccess) {
      if (err.empty()) err = "cudaMalloc failed while packing weights";
      return nullptr;
    }
    d->allocs.push_back(p);
    cudaMemcpy(p, h.data(), h.size() * sizeof(TT), cudaMemcpyHostToDevice);
    return (TT*)p;
  }
  float* upload_f32(const std::string& name, int64_t n) {
    const dt_tensor_desc* t = get(name, 1, {n});
    if (!t) return nullptr;
    std::vector<float> h(t->data, t->data + n);
    return upload(h);
  }
  static int pad64(int c) { return (c + 63) / 64 * 64; }

  // Conv1d weight (Cout, Cin, k) -> [Cout][k * cin_pad], K order (tap, channel)
  bool conv1d(const std::string& prefix, int cout, int cin, int k, ConvW* out, bool with_bias = true) {
    const dt_tensor_desc* w = get(prefix + ".weight", 3, {cout, cin, k});
    if (!w) return false;
    const int cp = pad64(cin);
    std::vector<__nv_bfloat16> h((size_t)cout * k * cp, __float2bfloat16(0.f));
    for (int o = 0; o < cout; ++o)
      for (int c = 0; c < cin; ++c)
        for (int j = 0; j < k; ++j)
          h[((size_t)o * k + j) * cp + c] = __float2bfloat16(w->data[((size_t)o * cin + c) * k + j]);
    out->w = upload(h);
    out->N = cout;
    out->Ktot = k * cp;
    if (with_bias) out->bias = upload_f32(prefix + ".bias", cout);
    return out->w != nullptr;
  }
  // ConvTranspose1d weight (Cin, Cout, 4), stride 2, padding 1 -> two 2-tap convs:
  //   out[2i]   = W[:,:,1]^T x[i] + W[:,:,3]^T x[i-1]
  //   out[2i+1] = W[:,:,2]^T x[i] + W[:,:,0]^T x[i+1]
  bool convT(const std::string& prefix, int c, ConvW out[2]) {
    const dt_tensor_desc* w = get(prefix + ".weight", 3, {c, c, 4});
    if (!w) return false;
    const int widx[2][2] = {{1, 3}, {2, 0}};
    float* bias = upload_f32(prefix + ".bias", c);
    for (int e = 0; e < 2; ++e) {
      std::vector<__nv_bfloat16> h((size_t)c * 2 * c);
      for (int o = 0; o < c; ++o)
        for (int j = 0; j < 2; ++j)
          for (int i = 0; i < c; ++i)
            h[((size_t)o * 2 + j) * c + i] = __float2bfloat16(w->data[((size_t)i * c + o) * 4 + widx[e][j]]);
      out[e].w = upload(h);
      out[e].N = c;
      out[e].Ktot = 2 * c;
      out[e].bias = bias;
      if (!out[e].w) return false;
    }
    return true;
  }
  bool conv_block(const std::string& prefix, int cout, int cin, ConvW* out) {
    if (!conv1d(prefix + ".block.0", cout, cin, 3, out)) return false;
    out->gamma = upload_f32(prefix + ".block.1.weight", cout);
    out->beta = upload_f32(prefix + ".block.1.bias", cout);
    return out->gamma && out->beta;
  }
  bool res_block(const std::string& prefix, int cin, int cout, ResBlockW* rb, int* film_cursor) {
    rb->cin = cin;
    rb->cout = cout;
    if (!conv_block(prefix + ".blocks.0", cout, cin, &rb->c1)) return false;
    if (!conv_block(prefix + ".blocks.1", cout, cout, &rb->c2)) return false;
    rb->has_res = cin != cout;
    if (rb->has_res && !conv1d(prefix + ".residual_conv", cout, cin, 1, &rb->res)) return false;
    rb->film_off = *film_cursor;
    *film_cursor += 2 * cout;
    film_names.push_back(prefix + ".cond_encoder.1");
    film_widths.push_back(2 * cout);
    return true;
  }
  std::vector<std::string> film_names;
  std::vector<int> film_widths;

  // Conv2d weight (Cout, Cin, k, k) -> [Cout][kpad], K order (ky, kx, c); fold = sum over Cin
  bool conv2d(const std::string& name, int cout, int cin, int k, bool fold, ConvW* out) {
    const dt_tensor_desc* w = get(name, 4, {cout, cin, k, k});
    if (!w) return false;
    const int ce = fold ? 1 : cin;
    const int kp = pad64(k * k * ce);
    std::vector<__nv_bfloat16> h((size_t)cout * kp, __float2bfloat16(0.f));
    for (int o = 0; o < cout; ++o)
      for (int ky = 0; ky < k; ++ky)
        for (int kx = 0; kx < k; ++kx) {
          if (fold) {
            float s = 0.f;
            for (int c = 0; c < cin; ++c) s += w->data[(((size_t)o * cin + c) * k + ky) * k + kx];
            h[(size_t)o * kp + (ky * k + kx)] = __float2bfloat16(s);
          } else {
            for (int c = 0; c < cin; ++c)
              h[(size_t)o * kp + (ky * k + kx) * cin + c] =
                  __float2bfloat16(w->data[(((size_t)o * cin + c) * k + ky) * k + kx]);
          }
        }
    out->w = upload(h);
    out->N = cout;
    out->Ktot = kp;
    return out->w != nullptr;
  }
};

template <typename TT>
static TT* arena(dt_denoiser* d, size_t n, bool* ok) {
  void* p = nullptr;
  if (cudaMalloc(&p, n * sizeof(TT)) != cudaSuccess) {
    *ok = false;
    return nullptr;
  }
  d->allocs.push_back(p);
  return (TT*)p;
}

// captured sampler graphs bake in the kernel selection: dropped whenever an option that changes it is set
void dt_denoiser_drop_graphs(dt_ctx* ctx) {
  if (!ctx || !ctx->den) return;
  for (auto& g : ctx->den->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.film_table) cudaFree(g.film_table);
  }
  ctx->den->graphs.clear();
}

void dt_denoiser_free(dt_ctx* ctx) {
  if (!ctx || !ctx->den) return;
  for (auto& g : ctx->den->graphs) {
    if (g.exec) cudaGraphExecDestroy(g.exec);
    if (g.film_table) cudaFree(g.film_table);
  }
  if (ctx->den->cap_stream) cudaStreamDestroy(ctx->den->cap_stream);
  if (ctx->den->side) cudaStreamDestroy(ctx->den->side);
  if (ctx->den->ev_fork) cudaEventDestroy(ctx->den->ev_fork);
  if (ctx->den->ev_join) cudaEventDestroy(ctx->den->ev_join);
  for (void* p : ctx->den->allocs) cudaFree(p);
  delete ctx->den;
  ctx->den = nullptr;
}

extern "C" int dt_load_denoiser(dt_ctx* ctx, const dt_tensor_desc* tensors, int n_tensors, const dt_model_cfg* cfg,
                                void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!tensors || !cfg || n_tensors <= 0) return dt_fail(ctx, DT_E_ARG, "dt_load_denoiser: null argument");
  DT_CUDA(cudaSetDevice(ctx->device));
  DT_CUDA(cudaDeviceSynchronize());
  dt_denoiser_free(ctx);
  const int A = cfg->action_dim, T = cfg->horizon, G = cfg->cond_dim, E = cfg->emb_dim, NM = cfg->map_size;
  const int* C = cfg->down_dims;
  if (A < 1 || A > 8) return dt_fail(ctx, DT_E_UNSUPPORTED, "action_dim must be 1..8");
  if (T % 4 != 0 || T > 128 || 128 % T != 0 || 128 % (T / 4) != 0)
    return dt_fail(ctx, DT_E_UNSUPPORTED, "pred_horizon must be a power of two in 4..128");
  for (int l = 0; l < 3; ++l) {
    const int gw = C[l] / 8;
    // group widths up to 256 are normalised inside the GEMM epilogue; wider ones (xlarge's 512) take the
    // plain-epilogue GEMM + k_gn_mish_wide fallback (gemm.cu)
    if (C[l] % 64 != 0 || (gw != 8 && gw != 16 && gw != 32 && gw != 64 && gw != 128 && gw != 256 && gw != 512 && gw != 1024))
      return dt_fail(ctx, DT_E_UNSUPPORTED,
                     "down_dims must be multiples of 64 with GroupNorm group width (C/8) in {8,...,1024}");
  }
  if (cfg->max_batch < 1) return dt_fail(ctx, DT_E_ARG, "max_batch must be positive");
  TMap tm;
  for (int i = 0; i < n_tensors; ++i) tm[tensors[i].name] = &tensors[i];

  dt_denoiser* d = new dt_denoiser();
  ctx->den = d;
  d->cfg = *cfg;
  d->A = A; d->T = T; d->G = G; d->E = E; d->NM = NM; d->MB = cfg->max_batch;
  for (int l = 0; l < 3; ++l) {
    d->C[l] = C[l];
    d->Tl[l] = T >> l;
  }
  Loader L{ctx, d, tm, (cudaStream_t)stream, ""};
  bool ok = true;
  int fc = 0;  // FiLM cursor, in forward order
  const int dims[4] = {A, C[0], C[1], C[2]};
  for (int l = 0; l < 3 && ok; ++l) {
    const std::string p = "unet.down_modules." + std::to_string(l);
    ok = ok && L.res_block(p + ".0", dims[l], dims[l + 1], &d->down[l][0], &fc);
    ok = ok && L.res_block(p + ".1", dims[l + 1], dims[l + 1], &d->down[l][1], &fc);
    if (l < 2) ok = ok && L.conv1d(p + ".2.conv", dims[l + 1], dims[l + 1], 3, &d->downs[l]);
  }
  ok = ok && L.res_block("unet.mid_modules.0", C[2], C[2], &d->mid[0], &fc);
  ok = ok && L.res_block("unet.mid_modules.1", C[2], C[2], &d->mid[1], &fc);
  for (int u = 0; u < 2 && ok; ++u) {
    const int cin = dims[3 - u] * 2, cout = dims[2 - u];
    const std::string p = "unet.up_modules." + std::to_string(u);
    ok = ok && L.res_block(p + ".0", cin, cout, &d->up[u][0], &fc);
    ok = ok && L.res_block(p + ".1", cout, cout, &d->up[u][1], &fc);
    ok = ok && L.convT(p + ".2.conv", cout, d->ups[u]);
  }
  ok = ok && L.conv_block("unet.final_conv.0", C[0], C[0], &d->final_blk);
  if (ok) {
    const dt_tensor_desc* fw = L.get("unet.final_conv.1.weight", 3, {A, C[0], 1});
    if (fw) {
      std::vector<float> h(fw->data, fw->data + (size_t)A * C[0]);
      d->final_w = L.upload(h);
    }
    d->final_b = L.upload_f32("unet.final_conv.1.bias", A);
    ok = d->final_w && d->final_b;
  }
  // FiLM: split every cond_encoder Linear(256 + E + G -> 2 Cout) into time / per-candidate parts
  d->F = fc;
  const int gdim = 256 + E + G;
  d->kc_pad = Loader::pad64(E + G);
  if (ok) {
    if (d->F % 64 != 0) { ok = false; L.err = "FiLM width is not a multiple of 64"; }
  }
  if (ok) {
    std::vector<__nv_bfloat16> wc((size_t)d->F * d->kc_pad, __float2bfloat16(0.f));
    std::vector<float> wt((size_t)d->F * 256), bias(d->F);
    int row0 = 0;
    for (size_t i = 0; i < L.film_names.size() && ok; ++i) {
      const int wdt = L.film_widths[i];
      const dt_tensor_desc* w = L.get(L.film_names[i] + ".weight", 2, {wdt, gdim});
      const dt_tensor_desc* b = L.get(L.film_names[i] + ".bias", 1, {wdt});
      if (!w || !b) { ok = false; break; }
      for (int r = 0; r < wdt; ++r) {
        const float* src = w->data + (size_t)r * gdim;
        for (int j = 0; j < 256; ++j) wt[(size_t)(row0 + r) * 256 + j] = src[j];
        for (int j = 0; j < E + G; ++j) wc[(size_t)(row0 + r) * d->kc_pad + j] = __float2bfloat16(src[256 + j]);
        bias[row0 + r] = b->data[r];
      }
      row0 += wdt;
    }
    if (ok) {
      d->film_c.w = L.upload(wc);
      d->film_c.bias = L.upload(bias);
      d->film_c.N = d->F;
      d->film_c.Ktot = d->kc_pad;
      d->film_wt = L.upload(wt);
      ok = d->film_c.w && d->film_c.bias && d->film_wt;
    }
  }
  if (ok) {
    const dt_tensor_desc* w1 = L.get("unet.diffusion_step_encoder.1.weight", 2, {1024, 256});
    const dt_tensor_desc* w2 = L.get("unet.diffusion_step_encoder.3.weight", 2, {256, 1024});
    if (w1 && w2) {
      d->t_w1 = L.upload(std::vector<float>(w1->data, w1->data + 1024 * 256));
      d->t_w2 = L.upload(std::vector<float>(w2->data, w2->data + 1024 * 256));
    }
    d->t_b1 = L.upload_f32("unet.diffusion_step_encoder.1.bias", 1024);
    d->t_b2 = L.upload_f32("unet.diffusion_step_encoder.3.bias", 256);
    ok = d->t_w1 && d->t_w2 && d->t_b1 && d->t_b2;
  }
  // encoder
  if (ok) {
    const std::string p = "encoder.resnet18.";
    ok = L.conv2d(p + "conv1.weight", 64, 3, 7, true, &d->enc_conv1);
    d->enc_conv1.gamma = L.upload_f32(p + "bn1.weight", 64);
    d->enc_conv1.beta = L.upload_f32(p + "bn1.bias", 64);
    int cin = 64;
    const int chans[4] = {64, 128, 256, 512};
    for (int li = 0; li < 4 && ok; ++li) {
      for (int b = 0; b < 2 && ok; ++b) {
        EncBlockW& eb = d->enc[li][b];
        const std::string q = p + "layer" + std::to_string(li + 1) + "." + std::to_string(b) + ".";
        eb.cin = (b == 0) ? cin : chans[li];
        eb.cout = chans[li];
        eb.stride = (b == 0 && li > 0) ? 2 : 1;
        ok = ok && L.conv2d(q + "conv1.weight", eb.cout, eb.cin, 3, false, &eb.c1);
        eb.c1.gamma = L.upload_f32(q + "bn1.weight", eb.cout);
        eb.c1.beta = L.upload_f32(q + "bn1.bias", eb.cout);
        ok = ok && L.conv2d(q + "conv2.weight", eb.cout, eb.cout, 3, false, &eb.c2);
        eb.c2.gamma = L.upload_f32(q + "bn2.weight", eb.cout);
        eb.c2.beta = L.upload_f32(q + "bn2.bias", eb.cout);
        eb.has_ds = (b == 0 && li > 0);
        if (eb.has_ds) {
          ok = ok && L.conv2d(q + "downsample.0.weight", eb.cout, eb.cin, 1, false, &eb.ds);
          eb.ds.gamma = L.upload_f32(q + "downsample.1.weight", eb.cout);
          eb.ds.beta = L.upload_f32(q + "downsample.1.bias", eb.cout);
        }
      }
      cin = chans[li];
    }
    if (ok) {
      // fc: Linear(512 -> E), N padded to a multiple of 64 with zero rows
      d->emb_pad = Loader::pad64(E);
      const dt_tensor_desc* w = L.get(p + "fc.weight", 2, {E, 512});
      const dt_tensor_desc* b = L.get(p + "fc.bias", 1, {E});
      if (w && b) {
        std::vector<__nv_bfloat16> h((size_t)d->emb_pad * 512, __float2bfloat16(0.f));
        std::vector<float> hb(d->emb_pad, 0.f);
        for (int o = 0; o < E; ++o) {
          for (int c = 0; c < 512; ++c) h[(size_t)o * 512 + c] = __float2bfloat16(w->data[(size_t)o * 512 + c]);
          hb[o] = b->data[o];
        }
        d->enc_fc.w = L.upload(h);
        d->enc_fc.bias = L.upload(hb);
        d->enc_fc.N = d->emb_pad;
        d->enc_fc.Ktot = 512;
      }
      ok = d->enc_fc.w && d->enc_fc.bias;
    }
  }
  if (!ok || !L.err.empty()) {
    std::string msg = "dt_load_denoiser: " + (L.err.empty() ? std::string("packing failed") : L.err);
    dt_denoiser_free(ctx);
    return dt_fail(ctx, DT_E_ARG, msg.c_str());
  }
  // scratch arena
  const size_t MB = d->MB;
  const size_t n0 = MB * d->Tl[0] * C[0], n10 = MB * d->Tl[1] * C[0], n11 = MB * d->Tl[1] * C[1];
  const size_t n21 = MB * d->Tl[2] * C[1], n22 = MB * d->Tl[2] * C[2];
  bool aok = true;
  typedef __nv_bfloat16 bf;
  d->X = arena<bf>(d, MB * T * 64, &aok);
  d->H0 = arena<bf>(d, n0, &aok); d->R0 = arena<bf>(d, n0, &aok); d->A0 = arena<bf>(d, n0, &aok);
  d->B0 = arena<bf>(d, n0, &aok); d->V0 = arena<bf>(d, n0, &aok); d->F0 = arena<bf>(d, n0, &aok);
  d->D0 = arena<bf>(d, n10, &aok); d->H1u = arena<bf>(d, n10, &aok); d->R1u = arena<bf>(d, n10, &aok);
  d->U1a = arena<bf>(d, n10, &aok); d->U1b = arena<bf>(d, n10, &aok);
  d->H1 = arena<bf>(d, n11, &aok); d->R1 = arena<bf>(d, n11, &aok); d->A1 = arena<bf>(d, n11, &aok);
  d->B1 = arena<bf>(d, n11, &aok); d->V1 = arena<bf>(d, n11, &aok);
  d->D1 = arena<bf>(d, n21, &aok); d->H2u = arena<bf>(d, n21, &aok); d->R2u = arena<bf>(d, n21, &aok);
  d->U2a = arena<bf>(d, n21, &aok); d->U2b = arena<bf>(d, n21, &aok);
  d->H2 = arena<bf>(d, n22, &aok); d->R2 = arena<bf>(d, n22, &aok); d->A2 = arena<bf>(d, n22, &aok);
  d->B2 = arena<bf>(d, n22, &aok); d->M2 = arena<bf>(d, n22, &aok); d->M1 = arena<bf>(d, n22, &aok);
  d->film_cand = arena<float>(d, MB * d->F, &aok);
  d->film_time = arena<float>(d, (size_t)DEN_KMAX * d->F, &aok);
  d->mish_t = arena<float>(d, (size_t)DEN_KMAX * 256, &aok);
  d->time_h = arena<float>(d, (size_t)DEN_KMAX * 1024, &aok);
  d->norm_dev = arena<float>(d, 16, &aok);
  d->g_noise = arena<float>(d, (size_t)DEN_GRAPH_MAXB * T * A, &aok);
  d->g_out = arena<float>(d, (size_t)DEN_GRAPH_MAXB * T * A, &aok);
  d->g_cond = arena<float>(d, (size_t)DEN_GRAPH_MAXB * G, &aok);
  d->g_lm = arena<bf>(d, (size_t)DEN_GRAPH_MAXB * NM * NM, &aok);
  d->cond_in = arena<bf>(d, MB * d->kc_pad, &aok);
  const int oh1 = (NM + 6 - 7) / 2 + 1;
  const int ph = (oh1 + 2 - 3) / 2 + 1;
  const size_t col_elems = MB * std::max<size_t>((size_t)oh1 * oh1 * 64, (size_t)ph * ph * 1152);
  d->col = arena<bf>(d, std::max<size_t>(col_elems, MB * 4608), &aok);
  for (int i = 0; i < 4; ++i) d->e[i] = arena<bf>(d, MB * (size_t)oh1 * oh1 * 64, &aok);
  d->emb = arena<float>(d, MB * d->emb_pad, &aok);
  d->lin = arena<float>(d, MB * d->emb_pad, &aok);
  // GroupNorm groups wider than one N tile (`xlarge`: 4096 / 8 = 512 channels) take the unfused path through an
  // fp32 scratch of the largest such layer (gemm.cu, conv_gemm_wide_gn): sized here, once, so that no launch path
  // ever allocates or synchronises (and stays graph-capturable)
  {
    size_t wide = 0;
    for (int l = 0; l < 3; ++l)
      if (C[l] / 8 > 256) wide = std::max(wide, MB * (size_t)d->Tl[l] * C[l] * sizeof(float));
    if (wide > ctx->wide_bytes) {
      if (ctx->d_wide) cudaFree(ctx->d_wide);
      ctx->d_wide = nullptr;
      ctx->wide_bytes = 0;
      if (cudaMalloc(&ctx->d_wide, wide) == cudaSuccess) ctx->wide_bytes = wide;
      else aok = false;
    }
  }
  if (!aok) {
    dt_denoiser_free(ctx);
    return dt_fail(ctx, DT_E_CUDA, "dt_load_denoiser: out of device memory for scratch (lower max_batch)");
  }
  return DT_OK;
}

// ------------------------------------------------------------------------------------------
// forward
// ------------------------------------------------------------------------------------------
struct Act {
  const __nv_bfloat16* p;
  int T, C;
};

static inline int ew_grid(int64_t n, const dt_ctx* ctx) {
  int64_t b = (n + 255) / 256;
  const int64_t cap = (int64_t)ctx->sm_count * 16;
  return (int)(b > cap ? cap : (b < 1 ? 1 : b));
}

static int launch_final_euler(dt_ctx* ctx, const __nv_bfloat16* f, int64_t rows, int C0, int A, const float* w,
                              const float* bias, float dt, float* a, float* vel_out, const float* norm,
                              cudaStream_t st) {
  const int grid = ew_grid(rows * 32, ctx);
  const bool pdl = ctx->pdl_on;
  if (A == 2 && C0 == 512) dt_launch(pdl, k_final_euler_a2<8>, grid, 256, 0, st, f, rows, w, bias, dt, a, vel_out, norm);
  else if (A == 2 && C0 == 256) dt_launch(pdl, k_final_euler_a2<4>, grid, 256, 0, st, f, rows, w, bias, dt, a, vel_out, norm);
  else if (A == 2 && C0 == 64) dt_launch(pdl, k_final_euler_a2<1>, grid, 256, 0, st, f, rows, w, bias, dt, a, vel_out, norm);
  else dt_launch(pdl, k_final_euler, grid, 256, 0, st, f, rows, C0, A, w, bias, dt, a, vel_out, norm);
  DT_LAUNCH_CHECK("k_final_euler");
  return DT_OK;
}

// conv1d k=3 (or k=1) stride 1 over one or two (channel-concatenated) sources
static int conv_s1(dt_ctx* ctx, const ConvW& w, int k, Act in0, const Act* in1, int64_t B, ConvGemm g,
                   cudaStream_t st) {
  g.a[0] = ActSrc{in0.p, in0.C, in0.T, 1};
  g.n_src = 1;
  if (in1) {
    g.a[1] = ActSrc{in1->p, in1->C, in1->T, 1};
    g.n_src = 2;
  }
  g.w = w.w;
  g.N = w.N;
  g.nseg = 0;
  for (int j = 0; j < k; ++j) {
    g.seg[g.nseg++] = GemmSeg{0, 0, j - k / 2, in0.C / 64};
    if (in1) g.seg[g.nseg++] = GemmSeg{1, 0, j - k / 2, in1->C / 64};
  }
  g.B = B;
  g.T = in0.T;
  g.bias = w.bias;
  g.ldc = w.N;
  g.out_b_stride = in0.T;
  g.out_t_stride = 1;
  g.out_off = 0;
  return dt_conv_gemm(ctx, g, st);
}

// the side stream and its fork / join events (created on first use); false when forking is off or unavailable
static bool side_stream_ready(dt_ctx* ctx, dt_denoiser* d) {
  if (!ctx->fork_on || ctx->prof_on) return false;
  if (!d->side) {
    if (cudaStreamCreateWithFlags(&d->side, cudaStreamNonBlocking) != cudaSuccess ||
        cudaEventCreateWithFlags(&d->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&d->ev_join, cudaEventDisableTiming) != cudaSuccess) {
      cudaGetLastError();
      ctx->fork_on = false;
      return false;
    }
  }
  return true;
}

static int res_block(dt_ctx* ctx, dt_denoiser* d, const ResBlockW& rb, Act in0, const Act* in1, int64_t B,
                     const float* film_t_k, __nv_bfloat16* H, __nv_bfloat16* R, __nv_bfloat16* out, cudaStream_t st) {
  const int T = in0.T;
  int rc;
  // Small batches (every launch is latency-bound, the SMs are mostly idle): the 1 x 1 residual conv does not depend
  // on the block's first conv, so it runs on a side stream beside it and joins before the second conv's epilogue
  // reads it.  Its launches carry no programmatic-launch attribute (their predecessor is an event, not a kernel).
  const bool fork = rb.has_res && B <= DEN_FORK_MAXB && side_stream_ready(ctx, d);
  if (fork) {
    DT_CUDA(cudaEventRecord(d->ev_fork, st));
    DT_CUDA(cudaStreamWaitEvent(d->side, d->ev_fork, 0));
    ConvGemm gr;
    gr.epi = EPI_PLAIN;
    gr.out_bf16 = R;
    gr.pdl = false;
    gr.scratch = 1;
    if ((rc = conv_s1(ctx, rb.res, 1, in0, in1, B, gr, d->side))) return rc;
    DT_CUDA(cudaEventRecord(d->ev_join, d->side));
  }
  // h = FiLM(Mish(GN(conv3(x))))                         conditional_unet1d.py:111-120
  ConvGemm g1;
  g1.epi = EPI_GN_MISH;
  g1.gamma = rb.c1.gamma;
  g1.beta = rb.c1.beta;
  g1.group_width = rb.cout / 8;
  g1.film = d->film_cand + rb.film_off;
  g1.film_ld = d->F;
  g1.film_t = film_t_k + rb.film_off;
  g1.out_bf16 = H;
  if ((rc = conv_s1(ctx, rb.c1, 3, in0, in1, B, g1, st))) return rc;
  // residual path                                          conditional_unet1d.py:100-101,141
  const __nv_bfloat16* resid;
  if (fork) {
    DT_CUDA(cudaStreamWaitEvent(st, d->ev_join, 0));
    resid = R;
  } else if (rb.has_res) {
    ConvGemm gr;
    gr.epi = EPI_PLAIN;
    gr.out_bf16 = R;
    if ((rc = conv_s1(ctx, rb.res, 1, in0, in1, B, gr, st))) return rc;
    resid = R;
  } else {
    resid = in0.p;
  }
  // out = Mish(GN(conv3(h))) + residual                   conditional_unet1d.py:140-142
  ConvGemm g2;
  g2.epi = EPI_GN_MISH;
  g2.gamma = rb.c2.gamma;
  g2.beta = rb.c2.beta;
  g2.group_width = rb.cout / 8;
  g2.resid = resid;
  g2.ld_res = rb.cout;
  g2.out_bf16 = out;
  g2.pdl = !fork;   // after a join the predecessors are a kernel AND an event: plain full dependencies
  Act h{H, T, rb.cout};
  return conv_s1(ctx, rb.c2, 3, h, nullptr, B, g2, st);
}

// Downsample1d: Conv1d(C, C, 3, stride 2, padding 1)       conv1d_components.py:7-13
static int downsample(dt_ctx* ctx, const ConvW& w, Act in, int64_t B, __nv_bfloat16* out, cudaStream_t st) {
  ConvGemm g;
  g.a[0] = ActSrc{in.p, in.C, in.T, 2};
  g.n_src = 1;
  g.w = w.w;
  g.N = w.N;
  g.nseg = 3;
  g.seg[0] = GemmSeg{0, 1, -1, in.C / 64};  // tap 0: t = 2 t_o - 1
  g.seg[1] = GemmSeg{0, 0, 0, in.C / 64};   // tap 1: t = 2 t_o
  g.seg[2] = GemmSeg{0, 1, 0, in.C / 64};   // tap 2: t = 2 t_o + 1
  g.B = B;
  g.T = in.T / 2;
  g.epi = EPI_PLAIN;
  g.bias = w.bias;
  g.out_bf16 = out;
  g.ldc = w.N;
  g.out_b_stride = in.T / 2;
  return dt_conv_gemm(ctx, g, st);
}

// Upsample1d: ConvTranspose1d(C, C, 4, 2, 1) as two 2-tap convs   conv1d_components.py:15-21
static int upsample(dt_ctx* ctx, const ConvW w[2], Act in, int64_t B, __nv_bfloat16* out, cudaStream_t st) {
  for (int e = 0; e < 2; ++e) {
    ConvGemm g;
    g.a[0] = ActSrc{in.p, in.C, in.T, 1};
    g.n_src = 1;
    g.w = w[e].w;
    g.N = w[e].N;
    g.nseg = 2;
    g.seg[0] = GemmSeg{0, 0, 0, in.C / 64};
    g.seg[1] = GemmSeg{0, 0, e == 0 ? -1 : 1, in.C / 64};
    g.B = B;
    g.T = in.T;
    g.epi = EPI_PLAIN;
    g.bias = w[e].bias;
    g.out_bf16 = out;
    g.ldc = w[e].N;
    g.out_b_stride = 2 * in.T;
    g.out_t_stride = 2;
    g.out_off = e;
    int rc = dt_conv_gemm(ctx, g, st);
    if (rc) return rc;
  }
  return DT_OK;
}

// one U-Net evaluation; X holds the bf16 sample, the result is left in F0 (B, T, C0)
static int unet_body(dt_ctx* ctx, dt_denoiser* d, int64_t B, const float* film_t_k, cudaStream_t st) {
  const int* C = d->C;
  const int* T = d->Tl;
  int rc;
  Act x{d->X, T[0], 64};
  if ((rc = res_block(ctx, d, d->down[0][0], x, nullptr, B, film_t_k, d->H0, d->R0, d->A0, st))) return rc;
  Act a0{d->A0, T[0], C[0]};
  if ((rc = res_block(ctx, d, d->down[0][1], a0, nullptr, B, film_t_k, d->H0, d->R0, d->B0, st))) return rc;
  Act b0{d->B0, T[0], C[0]};
  if ((rc = downsample(ctx, d->downs[0], b0, B, d->D0, st))) return rc;
  Act d0{d->D0, T[1], C[0]};
  if ((rc = res_block(ctx, d, d->down[1][0], d0, nullptr, B, film_t_k, d->H1, d->R1, d->A1, st))) return rc;
  Act a1{d->A1, T[1], C[1]};
  if ((rc = res_block(ctx, d, d->down[1][1], a1, nullptr, B, film_t_k, d->H1, d->R1, d->B1, st))) return rc;
  Act b1{d->B1, T[1], C[1]};
  if ((rc = downsample(ctx, d->downs[1], b1, B, d->D1, st))) return rc;
  Act d1{d->D1, T[2], C[1]};
  if ((rc = res_block(ctx, d, d->down[2][0], d1, nullptr, B, film_t_k, d->H2, d->R2, d->A2, st))) return rc;
  Act a2{d->A2, T[2], C[2]};
  if ((rc = res_block(ctx, d, d->down[2][1], a2, nullptr, B, film_t_k, d->H2, d->R2, d->B2, st))) return rc;
  Act b2{d->B2, T[2], C[2]};
  if ((rc = res_block(ctx, d, d->mid[0], b2, nullptr, B, film_t_k, d->H2, d->R2, d->M1, st))) return rc;
  Act m1{d->M1, T[2], C[2]};
  if ((rc = res_block(ctx, d, d->mid[1], m1, nullptr, B, film_t_k, d->H2, d->R2, d->M2, st))) return rc;
  // up path: cat(x, skip) is two TMA sources, never materialised      conditional_unet1d.py:326-340
  Act m2{d->M2, T[2], C[2]};
  if ((rc = res_block(ctx, d, d->up[0][0], m2, &b2, B, film_t_k, d->H2u, d->R2u, d->U2a, st))) return rc;
  Act u2a{d->U2a, T[2], C[1]};
  if ((rc = res_block(ctx, d, d->up[0][1], u2a, nullptr, B, film_t_k, d->H2u, d->R2u, d->U2b, st))) return rc;
  Act u2b{d->U2b, T[2], C[1]};
  if ((rc = upsample(ctx, d->ups[0], u2b, B, d->V1, st))) return rc;
  Act v1{d->V1, T[1], C[1]};
  if ((rc = res_block(ctx, d, d->up[1][0], v1, &b1, B, film_t_k, d->H1u, d->R1u, d->U1a, st))) return rc;
  Act u1a{d->U1a, T[1], C[0]};
  if ((rc = res_block(ctx, d, d->up[1][1], u1a, nullptr, B, film_t_k, d->H1u, d->R1u, d->U1b, st))) return rc;
  Act u1b{d->U1b, T[1], C[0]};
  if ((rc = upsample(ctx, d->ups[1], u1b, B, d->V0, st))) return rc;
  // final Conv1dBlock (no FiLM); the 1x1 projection is fused with the Euler update by the caller
  Act v0{d->V0, T[0], C[0]};
  ConvGemm gf;
  gf.epi = EPI_GN_MISH;
  gf.gamma = d->final_blk.gamma;
  gf.beta = d->final_blk.beta;
  gf.group_width = C[0] / 8;
  gf.out_bf16 = d->F0;
  return conv_s1(ctx, d->final_blk, 3, v0, nullptr, B, gf, st);
}

// ---- encoder ---------------------------------------------------------------------------------
// One ResNet conv + its GroupNorm(C / 16) (+ residual) (+ ReLU) as ONE tcgen05 GEMM launch (gemm.cu, EPI_GN_RELU):
//   * Cin a multiple of 64 (every conv but the stem): the input (B, H, W, Cin) is the GEMM's A operand through a 4-D
//     TMA tensor map (channel, x, y, sample) -- a conv tap is an (x, y) offset of the box, the conv stride is the
//     TMA traversal stride, the zero padding is TMA's out-of-bounds fill; nothing is im2col'ed;
//   * one output pixel per sample (the last stage on small local maps): the taps that only meet padding are left
//     out (K = valid taps x Cin) and the input is read as H W "time steps";
//   * the stem (Cin = 1, 7 x 7): im2col (49 -> 64 columns), then the same fused epilogue.
// Maps with more than 128 output pixels per sample do not fit the one-tile-holds-whole-samples epilogue and take the
// unfused path (im2col -> GEMM -> k_gn2d).
static int gn2d(dt_ctx* ctx, const ConvW& w, __nv_bfloat16* x, int64_t B, int HW, int C, const __nv_bfloat16* resid,
                int relu, __nv_bfloat16* out, cudaStream_t st) {
  const int64_t warps = B * (C / 16);
  int64_t blocks = (warps + 7) / 8;
  if (blocks > (int64_t)ctx->sm_count * 16) blocks = (int64_t)ctx->sm_count * 16;
  k_gn2d<<<(int)blocks, 256, 0, st>>>(x, B, HW, C, w.gamma, w.beta, resid, relu, out);
  DT_LAUNCH_CHECK("k_gn2d");
  return DT_OK;
}

// `after_event`: this launch does not follow a kernel of the chain in its stream (it runs on the side stream, or right
// after a join): no programmatic-launch attribute; side-stream launches also take the second split-K scratch
static int enc_conv_gn(dt_ctx* ctx, dt_denoiser* d, const ConvW& w, const __nv_bfloat16* in, int64_t B, int H, int W,
                       int Cin, int k, int stride, int pad, int OH, int OW, const __nv_bfloat16* resid, int relu,
                       __nv_bfloat16* out, cudaStream_t st, bool after_event = false, bool side = false) {
  const int64_t rows = B * OH * OW;
  const int T = OH * OW;
  ConvGemm g;
  g.pdl = !after_event;
  g.scratch = side ? 1 : 0;
  g.n_src = 1;
  g.w = w.w;
  g.N = w.N;
  g.B = B;
  g.T = T;
  g.bias = w.bias;
  g.out_bf16 = out;
  g.ldc = w.N;
  g.out_b_stride = T;
  g.out_t_stride = 1;
  const bool fused = T <= 128 && w.N % 64 == 0;
  if (fused) {
    g.epi = EPI_GN_RELU;
    g.gamma = w.gamma;
    g.beta = w.beta;
    g.group_width = 16;
    g.resid = resid;
    g.ld_res = w.N;
    g.relu = relu;
  } else {
    g.epi = EPI_PLAIN;
  }
  bool direct = false;
  if (fused && OH == 1 && OW == 1 && H * W <= GEMM_MAX_SEG && Cin % 64 == 0 && w.Ktot == k * k * Cin) {
    // one output pixel per sample: valid taps only, the input read as H W time steps
    g.a[0] = ActSrc{in, Cin, H * W, 1};
    g.w_ktot = w.Ktot;
    g.nseg = 0;
    int next_tap = 0;  // first weight tap not yet consumed or skipped
    for (int ky = 0; ky < k; ++ky) {
      for (int kx = 0; kx < k; ++kx) {
        const int iy = ky - pad, ix = kx - pad;  // output pixel (0, 0): input pixel (iy, ix), any stride
        if (iy < 0 || iy >= H || ix < 0 || ix >= W) continue;
        const int tap = ky * k + kx;
        g.seg[g.nseg++] = GemmSeg{0, 0, iy * W + ix, Cin / 64, (tap - next_tap) * (Cin / 64)};
        next_tap = tap + 1;
      }
    }
    direct = true;
  } else if (fused && Cin % 64 == 0 && w.Ktot == k * k * Cin && k * k <= GEMM_MAX_SEG) {
    ActSrc a{in, Cin, 0, 1};
    a.W2 = W; a.H2 = H; a.stride2 = stride; a.ow2 = OW; a.oh2 = OH;
    g.a[0] = a;
    g.nseg = 0;
    for (int ky = 0; ky < k; ++ky)
      for (int kx = 0; kx < k; ++kx) g.seg[g.nseg++] = GemmSeg{0, kx - pad, ky - pad, Cin / 64, 0};
    direct = true;
  }
  if (!direct) {
    if (Cin % 8 == 0 && w.Ktot == k * k * Cin) {
      dt_launch(ctx->pdl_on, k_im2col_v8, ew_grid(rows * (w.Ktot / 8), ctx), 256, 0, st, in, B, H, W, Cin, k, stride, pad, OH, OW, d->col);
      DT_LAUNCH_CHECK("k_im2col_v8");
    } else {
      dt_launch(ctx->pdl_on, k_im2col, ew_grid(rows * w.Ktot, ctx), 256, 0, st, in, B, H, W, Cin, k, stride, pad, OH, OW, w.Ktot, d->col);
      DT_LAUNCH_CHECK("k_im2col");
    }
    g.nseg = 1;
    g.seg[0] = GemmSeg{0, 0, 0, w.Ktot / 64, 0};
    if (fused) {
      g.a[0] = ActSrc{d->col, w.Ktot, T, 1};      // the im2col matrix as B samples of T rows
    } else {
      g.a[0] = ActSrc{d->col, w.Ktot, (int)rows, 1};
      g.B = 1;
      g.T = (int)rows;
      g.out_b_stride = 0;
    }
  }
  int rc = dt_conv_gemm(ctx, g, st);
  if (rc || fused) return rc;
  return gn2d(ctx, w, out, B, T, w.N, resid, relu, out, st);
}

// local map (B,N,N) bf16 in [-1,1] -> d->emb (B, emb_pad) f32
static int encoder_forward(dt_ctx* ctx, dt_denoiser* d, const __nv_bfloat16* lm, int64_t B, cudaStream_t st) {
  const int NM = d->NM;
  int rc;
  int H = (NM + 6 - 7) / 2 + 1;  // conv1 7x7 s2 p3, then GroupNorm + ReLU
  if ((rc = enc_conv_gn(ctx, d, d->enc_conv1, lm, B, NM, NM, 1, 7, 2, 3, H, H, nullptr, 1, d->e[1], st))) return rc;
  const int PH = (H + 2 - 3) / 2 + 1;  // maxpool 3x3 s2 p1
  dt_launch(ctx->pdl_on, k_maxpool3s2, ew_grid(B * PH * PH * 64, ctx), 256, 0, st, d->e[1], B, H, H, 64, PH, PH, d->e[0]);
  DT_LAUNCH_CHECK("k_maxpool3s2");
  __nv_bfloat16* cur = d->e[0];
  __nv_bfloat16* t1 = d->e[1];
  __nv_bfloat16* t2 = d->e[2];
  __nv_bfloat16* t3 = d->e[3];
  H = PH;
  for (int li = 0; li < 4; ++li) {
    for (int b = 0; b < 2; ++b) {
      const EncBlockW& eb = d->enc[li][b];
      const int OH = (H + 2 - 3) / eb.stride + 1;
      // identity / downsample branch: gn(conv1x1(x)) -- beside conv1 on the side stream at small batches (as the
      // U-Net's residual convs, res_block), when it takes the TMA-addressed path (no shared im2col scratch)
      const __nv_bfloat16* idt = cur;
      bool joined = false;
      if (eb.has_ds) {
        const bool fork = B <= DEN_FORK_MAXB && OH * OH <= 128 && eb.cin % 64 == 0 && eb.ds.N % 64 == 0 &&
                          eb.ds.Ktot == eb.cin && side_stream_ready(ctx, d);
        if (fork) {
          DT_CUDA(cudaEventRecord(d->ev_fork, st));
          DT_CUDA(cudaStreamWaitEvent(d->side, d->ev_fork, 0));
          if ((rc = enc_conv_gn(ctx, d, eb.ds, cur, B, H, H, eb.cin, 1, eb.stride, 0, OH, OH, nullptr, 0, t3, d->side, true, true)))
            return rc;
          DT_CUDA(cudaEventRecord(d->ev_join, d->side));
          joined = true;
        }
      }
      // y = relu(gn(conv1(x)))
      if ((rc = enc_conv_gn(ctx, d, eb.c1, cur, B, H, H, eb.cin, 3, eb.stride, 1, OH, OH, nullptr, 1, t1, st))) return rc;
      if (eb.has_ds) {
        if (joined) DT_CUDA(cudaStreamWaitEvent(st, d->ev_join, 0));
        else if ((rc = enc_conv_gn(ctx, d, eb.ds, cur, B, H, H, eb.cin, 1, eb.stride, 0, OH, OH, nullptr, 0, t3, st))) return rc;
        idt = t3;
      }
      // out = relu(gn(conv2(y)) + identity)
      if ((rc = enc_conv_gn(ctx, d, eb.c2, t1, B, OH, OH, eb.cout, 3, 1, 1, OH, OH, idt, 1, t2, st, joined))) return rc;
      __nv_bfloat16* nxt = t2;
      t2 = cur;
      cur = nxt;
      H = OH;
    }
  }
  const __nv_bfloat16* feat = cur;
  if (H * H > 1) {
    dt_launch(ctx->pdl_on, k_avgpool, ew_grid(B * 512, ctx), 256, 0, st, cur, B, H * H, 512, t1);
    DT_LAUNCH_CHECK("k_avgpool");
    feat = t1;
  }
  ConvGemm g;
  g.a[0] = ActSrc{feat, 512, (int)B, 1};
  g.n_src = 1;
  g.w = d->enc_fc.w;
  g.N = d->enc_fc.N;
  g.nseg = 1;
  g.seg[0] = GemmSeg{0, 0, 0, 512 / 64};
  g.B = 1;
  g.T = (int)B;
  g.epi = EPI_PLAIN;
  g.bias = d->enc_fc.bias;
  g.out_f32 = d->emb;
  g.ldc = d->enc_fc.N;
  return dt_conv_gemm(ctx, g, st);
}

// per-candidate FiLM part for all 12 residual blocks: one GEMM (B x kc_pad) x (kc_pad x F)
static int film_candidates(dt_ctx* ctx, dt_denoiser* d, const float* emb, int emb_ld, const float* cond, int64_t B,
                           cudaStream_t st) {
  dt_launch(ctx->pdl_on, k_film_input, ew_grid(B * d->kc_pad, ctx), 256, 0, st, emb, emb_ld, d->E, cond, d->G, B, d->kc_pad, d->cond_in);
  DT_LAUNCH_CHECK("k_film_input");
  ConvGemm g;
  g.a[0] = ActSrc{d->cond_in, d->kc_pad, (int)B, 1};
  g.n_src = 1;
  g.w = d->film_c.w;
  g.N = d->F;
  g.nseg = 1;
  g.seg[0] = GemmSeg{0, 0, 0, d->kc_pad / 64};
  g.B = 1;
  g.T = (int)B;
  g.epi = EPI_PLAIN;
  g.bias = d->film_c.bias;
  g.out_f32 = d->film_cand;
  g.ldc = d->F;
  return dt_conv_gemm(ctx, g, st);
}

static int film_times(dt_ctx* ctx, dt_denoiser* d, const float* ts_host, int K, cudaStream_t st,
                      float* private_table = nullptr) {
  // The per-step FiLM table depends only on the timesteps and the weights: an unchanged schedule (every call
  // of a planner run) reuses the table of the previous call -- stream order makes that safe on one stream;
  // a different stream recomputes.  `private_table`: compute into the caller's own [K][F] buffer instead (a
  // captured graph keeps one per schedule, filled once at capture time, so that its replays skip these launches).
  if (!private_table && d->film_time_K == K && d->film_time_stream == (void*)st &&
      memcmp(d->film_time_ts, ts_host, K * sizeof(float)) == 0)
    return DT_OK;
  TimeSteps ts;
  memset(&ts, 0, sizeof(ts));
  memcpy(ts.t, ts_host, K * sizeof(float));
  dt_launch(ctx->pdl_on, k_time_mlp1, K * 128, 256, 0, st, ts, d->t_w1, d->t_b1, d->time_h);
  DT_LAUNCH_CHECK("k_time_mlp1");
  dt_launch(ctx->pdl_on, k_time_mlp2, K * 32, 256, 0, st, d->time_h, d->t_w2, d->t_b2, d->mish_t);
  DT_LAUNCH_CHECK("k_time_mlp2");
  dt_launch(ctx->pdl_on, k_film_time, (d->F + 7) / 8, 256, 0, st, d->film_wt, d->mish_t, d->F, K,
            private_table ? private_table : d->film_time);
  DT_LAUNCH_CHECK("k_film_time");
  if (private_table) return DT_OK;
  d->film_time_K = K;
  d->film_time_stream = (void*)st;
  memcpy(d->film_time_ts, ts_host, K * sizeof(float));
  return DT_OK;
}

dt_model_cfg dt_denoiser_cfg(dt_ctx* ctx) { return ctx->den->cfg; }

#define NEED_MODEL()                                                            \
  if (!ctx) return DT_E_ARG;                                                    \
  if (!ctx->den) return dt_fail(ctx, DT_E_NOMODEL, "dt_load_denoiser has not been called")

extern "C" int dt_encode_map(dt_ctx* ctx, const void* local_map, int64_t B, float* emb_out, void* stream) {
  NEED_MODEL();
  dt_denoiser* d = ctx->den;
  if (B <= 0) return DT_OK;
  if (!local_map || !emb_out) return dt_fail(ctx, DT_E_ARG, "dt_encode_map: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  const __nv_bfloat16* lm = (const __nv_bfloat16*)local_map;
  for (int64_t b0 = 0; b0 < B; b0 += d->MB) {
    const int64_t nb = (B - b0 < d->MB) ? (B - b0) : d->MB;
    int rc = encoder_forward(ctx, d, lm + b0 * d->NM * d->NM, nb, st);
    if (rc) return rc;
    dt_launch(ctx->pdl_on, k_copy_cols, ew_grid(nb * d->E, ctx), 256, 0, st, d->emb, d->emb_pad, nb, d->E, emb_out + b0 * d->E);
    DT_LAUNCH_CHECK("k_copy_cols");
  }
  return DT_OK;
}

extern "C" int dt_unet_forward(dt_ctx* ctx, const float* sample, const float* emb, const float* cond, int64_t B,
                               float timestep, float* vel_out, void* stream) {
  NEED_MODEL();
  dt_denoiser* d = ctx->den;
  if (B <= 0) return DT_OK;
  if (!sample || !emb || !cond || !vel_out) return dt_fail(ctx, DT_E_ARG, "dt_unet_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  int rc = film_times(ctx, d, &timestep, 1, st);
  if (rc) return rc;
  for (int64_t b0 = 0; b0 < B; b0 += d->MB) {
    const int64_t nb = (B - b0 < d->MB) ? (B - b0) : d->MB;
    if ((rc = film_candidates(ctx, d, emb + b0 * d->E, d->E, cond + b0 * d->G, nb, st))) return rc;
    const int64_t rows = nb * d->T;
    dt_launch(ctx->pdl_on, k_prep_sample, ew_grid(rows * 64, ctx), 256, 0, st, sample + b0 * d->T * d->A, rows, d->A, d->X);
    DT_LAUNCH_CHECK("k_prep_sample");
    if ((rc = unet_body(ctx, d, nb, d->film_time, st))) return rc;
    if ((rc = launch_final_euler(ctx, d->F0, rows, d->C[0], d->A, d->final_w, d->final_b, 0.f, nullptr,
                                 vel_out + b0 * d->T * d->A, nullptr, st)))
      return rc;
  }
  return DT_OK;
}

// encoder once, per-candidate FiLM once, then K Euler steps of the U-Net (fm_policy.py:183-194)
static int fm_sample_body(dt_ctx* ctx, dt_denoiser* d, const float* noise, const float* cond, const __nv_bfloat16* lm,
                          int64_t B, int K, const float* ts, const float* dt, const float* d_norm, float* actions_out,
                          cudaStream_t st, const float* film_table) {
  // film_table: a ready per-step FiLM table [K][F] (a captured graph's own); null: the context's, refreshed if stale
  int rc = film_table ? DT_OK : film_times(ctx, d, ts, K, st);
  if (rc) return rc;
  if (!film_table) film_table = d->film_time;
  for (int64_t b0 = 0; b0 < B; b0 += d->MB) {
    const int64_t nb = (B - b0 < d->MB) ? (B - b0) : d->MB;
    const int64_t rows = nb * d->T;
    float* a = actions_out + b0 * d->T * d->A;
    if ((rc = encoder_forward(ctx, d, lm + b0 * d->NM * d->NM, nb, st))) return rc;
    if ((rc = film_candidates(ctx, d, d->emb, d->emb_pad, cond + b0 * d->G, nb, st))) return rc;
    DT_CUDA(cudaMemcpyAsync(a, noise + b0 * d->T * d->A, rows * d->A * sizeof(float), cudaMemcpyDeviceToDevice, st));
    for (int k = 0; k < K; ++k) {
      // (k == 0 follows the copy of the noise: a memcpy node is no programmatic-launch primary)
      dt_launch(ctx->pdl_on && k > 0, k_prep_sample, ew_grid(rows * 64, ctx), 256, 0, st, a, rows, d->A, d->X);
      DT_LAUNCH_CHECK("k_prep_sample");
      if ((rc = unet_body(ctx, d, nb, film_table + (size_t)k * d->F, st))) return rc;
      if ((rc = launch_final_euler(ctx, d->F0, rows, d->C[0], d->A, d->final_w, d->final_b, dt[k], a, nullptr,
                                   (k == K - 1) ? d_norm : nullptr, st)))
        return rc;
    }
  }
  return DT_OK;
}

extern "C" int dt_fm_sample(dt_ctx* ctx, const float* noise, const float* cond, const void* local_map, int64_t B,
                            int K, double exp_scale, const double* norm_host, float* actions_out, void* stream) {
  NEED_MODEL();
  dt_denoiser* d = ctx->den;
  if (B <= 0) return DT_OK;
  if (!noise || !cond || !local_map || !actions_out || K < 1 || K > DEN_KMAX)
    return dt_fail(ctx, DT_E_ARG, "dt_fm_sample: bad argument (1 <= K <= 64)");
  cudaStream_t st = (cudaStream_t)stream;
  // exp schedule in float32 (common/fm_utils.py:4-17); timestep = 20 * t0[k] (fm_policy.py:186-187)
  float t0[DEN_KMAX], dt[DEN_KMAX], ts[DEN_KMAX];
  {
    float sum = 0.f;
    for (int k = 0; k < K; ++k) {
      const float t = (float)k / (float)K;
      dt[k] = expf(-t * (float)exp_scale);
      sum += dt[k];
    }
    float cum = 0.f;
    for (int k = 0; k < K; ++k) {
      dt[k] = dt[k] / sum;
      t0[k] = cum;
      cum += dt[k];
      ts[k] = t0[k] * 20.0f;
    }
  }
  // un-normalisation constants: uploaded only when they change (then synchronously: the staging array is on
  // this stack frame); the steady state of a planner run enqueues without any host-device synchronisation
  int rc;
  float* d_norm = nullptr;
  if (norm_host) {
    float hn[16];
    for (int i = 0; i < 2 * d->A; ++i) hn[i] = (float)norm_host[i];
    d_norm = d->norm_dev;
    if (!d->norm_valid || memcmp(hn, d->norm_host, 2 * d->A * sizeof(float)) != 0) {
      DT_CUDA(cudaMemcpyAsync(d_norm, hn, 2 * d->A * sizeof(float), cudaMemcpyHostToDevice, st));
      DT_CUDA(cudaStreamSynchronize(st));
      memcpy(d->norm_host, hn, 2 * d->A * sizeof(float));
      d->norm_valid = true;
    }
  }
  const __nv_bfloat16* lm = (const __nv_bfloat16*)local_map;

  // ---- small batches: one CUDA graph launch instead of ~100 K kernel launches ----
  static const bool graphs_on = [] { const char* e = getenv("DITREE_GRAPHS"); return !(e && e[0] == '0'); }();
  cudaStreamCaptureStatus cap = cudaStreamCaptureStatusNone;
  // (not for nets whose widest GroupNorm takes the scratch-buffer fallback: that scratch may be re-allocated)
  if (graphs_on && B <= DEN_GRAPH_MAXB && B <= d->MB && !ctx->prof_on && d->C[2] / 8 <= 256 &&
      (st == nullptr || (cudaStreamIsCapturing(st, &cap) == cudaSuccess && cap == cudaStreamCaptureStatusNone))) {
    unsigned ebits;
    const float es = (float)exp_scale;
    memcpy(&ebits, &es, 4);
    dt_denoiser::GraphEntry* ge = nullptr;
    for (auto& g : d->graphs)
      if (g.B == (int)B && g.K == K && g.ebits == ebits && g.norm == (d_norm != nullptr)) ge = &g;
    if (!ge) {  // first call with this key: eager (runs first-use initialisation), remembered
      d->graphs.push_back(dt_denoiser::GraphEntry{(int)B, K, ebits, d_norm != nullptr, 0, nullptr, 0, nullptr});
    } else if (ge->state >= 0) {
      const size_t na = (size_t)B * d->T * d->A;
      {
        // the three inputs move into the graph's static buffers with one launch (three copy nodes cost ~3 us each)
        const size_t w0 = na, w1 = (size_t)B * d->G, w2 = ((size_t)B * d->NM * d->NM * sizeof(__nv_bfloat16) + 3) / 4;
        const size_t words = w0 + w1 + w2;
        if ((((uintptr_t)lm | (uintptr_t)d->g_lm) & 3) != 0 || ((size_t)B * d->NM * d->NM) % 2 != 0) {
          DT_CUDA(cudaMemcpyAsync(d->g_noise, noise, na * sizeof(float), cudaMemcpyDeviceToDevice, st));
          DT_CUDA(cudaMemcpyAsync(d->g_cond, cond, (size_t)B * d->G * sizeof(float), cudaMemcpyDeviceToDevice, st));
          DT_CUDA(cudaMemcpyAsync(d->g_lm, lm, (size_t)B * d->NM * d->NM * sizeof(__nv_bfloat16), cudaMemcpyDeviceToDevice, st));
        } else {
          k_stage3<<<(unsigned)((words + 255) / 256), 256, 0, st>>>(
              (const uint32_t*)noise, (uint32_t*)d->g_noise, w0, (const uint32_t*)cond, (uint32_t*)d->g_cond, w1,
              (const uint32_t*)lm, (uint32_t*)d->g_lm, w2);
          DT_LAUNCH_CHECK("k_stage3");
        }
      }
      if (ge->state == 0) {
        cudaGraph_t graph = nullptr;
        // captured on a private stream (the caller's may be the legacy default stream, which cannot be
        // captured); the instantiated graph is then launched into the caller's stream
        if (!d->cap_stream) cudaStreamCreateWithFlags(&d->cap_stream, cudaStreamNonBlocking);
        // this schedule's per-step FiLM table: computed once, here, on the caller's stream (every replay is enqueued
        // behind it), so the graph itself holds no time-MLP launches
        if (!ge->film_table && cudaMalloc(&ge->film_table, (size_t)K * d->F * sizeof(float)) != cudaSuccess) {
          cudaGetLastError();
          ge->film_table = nullptr;
        }
        bool ok = ge->film_table != nullptr && film_times(ctx, d, ts, K, st, ge->film_table) == DT_OK;
        const long long l0 = ctx->launches;   // launches of one replay = what the capture records
        ok = ok && d->cap_stream && cudaStreamBeginCapture(d->cap_stream, cudaStreamCaptureModeRelaxed) == cudaSuccess;
        if (ok) {
          rc = fm_sample_body(ctx, d, d->g_noise, d->g_cond, d->g_lm, B, K, ts, dt, d_norm, d->g_out, d->cap_stream,
                              ge->film_table);
          const cudaError_t e = cudaStreamEndCapture(d->cap_stream, &graph);
          ok = (rc == DT_OK) && (e == cudaSuccess) && graph != nullptr;
        }
        if (ok) ok = cudaGraphInstantiate(&ge->exec, graph, 0) == cudaSuccess;
        if (graph) cudaGraphDestroy(graph);
        ge->launches = ctx->launches - l0;
        ctx->launches = l0;
        if (!ok) {
          cudaGetLastError();
          ge->state = -1;  // not capturable here: stay eager
          ge->exec = nullptr;
        } else {
          ge->state = 1;
        }
      }
      if (ge->state == 1) {
        DT_CUDA(cudaGraphLaunch(ge->exec, st));
        ctx->launches += ge->launches;
        DT_CUDA(cudaMemcpyAsync(actions_out, d->g_out, na * sizeof(float), cudaMemcpyDeviceToDevice, st));
        return DT_OK;
      }
    }
  }
  return fm_sample_body(ctx, d, noise, cond, lm, B, K, ts, dt, d_norm, actions_out, st, nullptr);
}
