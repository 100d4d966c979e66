// gemm.cuh -- host-side description of one fused implicit-GEMM conv launch (see gemm.cu).
#pragma once
#include "common.cuh"

// One K-segment of the implicit GEMM: `nblk` consecutive 64-channel blocks of activation source
// `src`, read at time offset `t_off` (may be negative / past the end: TMA zero-fills = padding)
// and time phase `phase` (stride-2 convs view the time axis as (T/2, 2)).
struct GemmSeg {
  int src;    // 0 or 1 (second source = channel concat, e.g. U-Net skip connections)
  int phase;  // time phase inside the (T/P, P) view
  int t_off;  // time offset in units of the P-strided axis
  int nblk;   // number of 64-channel K blocks
  int w_gap;  // weight K blocks skipped before this segment (0 = the weights follow the previous segment's);
              // lets a launch use a subset of a packed conv's taps (2-D convs on 1 x 1 / 2 x 2 maps: the taps that
              // only ever see zero padding are left out)
};

#define GEMM_MAX_SEG 9   // a 3 x 3 conv is nine taps

// EPI_GN_RELU (the ResNet encoder): bias -> GroupNorm(groups of 16 channels) -> (+ residual) -> optional ReLU, for ANY
// number of rows per sample T <= 128 (a tile holds floor(128 / T) whole samples; statistics through shared-memory
// atomics), so the encoder's 10x10 / 5x5 / 3x3 / 2x2 / 1x1 maps need no separate normalisation kernel.
enum { EPI_PLAIN = 0, EPI_GN_MISH = 1, EPI_GN_RELU = 3 };

struct ActSrc {
  const __nv_bfloat16* ptr;  // (B, T_in, C) channel-last bf16
  int C;                     // channels (multiple of 64)
  int T_in;                  // time steps per sample in memory
  int P;                     // phase count of the time view (1, or 2 for stride-2 convs)
  // 2-D source (W2 > 0): ptr is (B, H2, W2, C) channel-last and a tile row is an output pixel (oy, ox) of an
  // oh2 x ow2 output map read with stride2; a K segment is one conv tap: GemmSeg::phase = kx - pad (x offset),
  // GemmSeg::t_off = ky - pad (y offset); out-of-range pixels are TMA zero fill = the conv's zero padding
  int W2 = 0, H2 = 0, stride2 = 1, ow2 = 0, oh2 = 0;
};

struct ConvGemm {
  // ---- operands -------------------------------------------------------------------------
  ActSrc a[2];
  int n_src = 1;
  const __nv_bfloat16* w = nullptr;  // packed weights [N][Ktot] bf16, K-major; Ktot = 64 * sum(nblk)
  int N = 0;                         // output channels (multiple of the N tile)
  int64_t w_ktot = 0;                // row length of `w` in elements when segments skip blocks (0 = 64 * sum(nblk))
  int nseg = 0;
  GemmSeg seg[GEMM_MAX_SEG];
  // ---- row geometry -----------------------------------------------------------------------
  int64_t B = 0;  // samples
  int T = 1;      // GEMM rows per sample (output time steps of this launch)
  // ---- epilogue ---------------------------------------------------------------------------
  int epi = EPI_PLAIN;
  const float* bias = nullptr;   // [N] or null
  const float* gamma = nullptr;  // GroupNorm affine (EPI_GN_MISH)
  const float* beta = nullptr;
  int group_width = 0;           // channels per GroupNorm group
  const float* film = nullptr;   // per-sample FiLM part  [B][film_ld]: scale at n, shift at N + n; null = no FiLM
  int64_t film_ld = 0;
  const float* film_t = nullptr; // per-step (batch-shared) FiLM part [2N], added to `film`
  const __nv_bfloat16* resid = nullptr;  // residual added after the activation, rows like the output
  int64_t ld_res = 0;
  int relu = 0;                  // EPI_PLAIN / EPI_GN_RELU: apply ReLU last (after the residual)
  // ---- output: row (b, t) -> out_row = b*out_b_stride + t*out_t_stride + out_off -----------
  __nv_bfloat16* out_bf16 = nullptr;
  float* out_f32 = nullptr;
  int64_t ldc = 0;
  int64_t out_b_stride = 0, out_t_stride = 1, out_off = 0;
  // ---- split-K (set by dt_conv_gemm itself for small problems; plain epilogue, fp32 partial sums) -------
  int ksplit = 1, kb_per_slice = 0;
  int64_t slice_rows = 0;
  // ---- launch ------------------------------------------------------------------------------------
  bool pdl = true;   // programmatic dependent launch (if the context allows it): off for a launch whose predecessor in
                     // its stream is not a kernel of this chain (after an event wait / on the side stream)
  int scratch = 0;   // which split-K scratch buffer (a layer running on the side stream takes its own: 1)
};

int dt_conv_gemm(dt_ctx* ctx, const ConvGemm& g, cudaStream_t st);
