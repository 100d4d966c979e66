// gemm.cu -- the tcgen05 implicit-GEMM core of the denoiser.
//
//   D[M = B*T rows, N = Cout] = sum over K-segments  A_seg[M, 64*nblk] * W[N, Kseg]^T      (bf16 in, fp32 acc)
//
// One persistent CTA per SM, warp-specialised:
//   warp 0      TMA producer: per 64-channel K block one 4-D tensor-map load of the activation
//               tile (channel, phase, time, sample) -- conv taps are time offsets of the box, the
//               zero padding is TMA's out-of-bounds fill, channel concat is a second tensor map --
//               and one 2-D load of the weight tile; 128-byte swizzle; mbarrier complete_tx.
//   warp 1      MMA issuer: tcgen05.mma.cta_group::1.kind::f16, M=128, N=BN (64/128/256), K=16,
//               operands straight from shared memory through UMMA descriptors, accumulator in
//               TMEM (two BN-column buffers so the epilogue of tile i overlaps the MMAs of i+1).
//   warps 2..9  epilogue: tcgen05.ld the accumulator (one row per thread, two warps per lane
//               quarter splitting the columns), then either
//               bias (+residual, +ReLU)  or  bias -> GroupNorm -> Mish -> FiLM (+residual);
//               GroupNorm statistics are tile-local because a tile holds whole samples (M tile =
//               128/T samples x T rows) and whole groups (BN is a multiple of the group width).
#include <cuda.h>
#include <stdlib.h>

#include <cooperative_groups.h>

#include "gemm.cuh"

#define BM 128
#define BK 64
#ifndef EPI_SPLIT
#define EPI_SPLIT 2                      // epilogue warps per TMEM lane quarter (column split)
#endif
#define GEMM_THREADS (64 + 128 * EPI_SPLIT)
#define SPIN_LIMIT (1u << 24)

// ------------------------------------------------------------------------------------------
// PTX wrappers
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_arrive_relaxed(uint32_t bar) {
  asm volatile("mbarrier.arrive.relaxed.cta.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t done = 0, spins = 0;
  while (true) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        // (a suspend-time hint -- the warp may sleep instead of spinning -- was measured: 20.50 k vs 20.56 k edges/s
        // without it, same box, two runs each; not used)
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    if (done) break;
    if (++spins > SPIN_LIMIT) __trap();  // a protocol bug must fault, not hang the device
  }
}
__device__ __forceinline__ void tma_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                            int c3) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], "
      "[%2];" ::"r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::
          "r"(dst),
      "l"(map), "r"(bar), "r"(c0), "r"(c1)
      : "memory");
}
// ---- CTA-pair (cta_group::2) variants -----------------------------------------------------------
// In a 2-CTA cluster the shared::cluster address of the even (leader) CTA's copy of a shared
// variable is the local address with bit 24 cleared.
#define DT_PEER_MASK 0xFEFFFFFFu
__device__ __forceinline__ void tma2_load_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1, int c2,
                                             int c3) {
  const unsigned long long hint = 0x1000000000000000ull;  // evict-normal
  asm volatile(
      "cp.async.bulk.tensor.4d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4, %5, %6}], [%2], %7;" ::"r"(dst),
      "l"(map), "r"(bar & DT_PEER_MASK), "r"(c0), "r"(c1), "r"(c2), "r"(c3), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tma2_load_2d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c0, int c1) {
  const unsigned long long hint = 0x14F0000000000000ull;  // evict-last: weights are re-read by every M tile
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint "
      "[%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst),
      "l"(map), "r"(bar & DT_PEER_MASK), "r"(c0), "r"(c1), "l"(hint)
      : "memory");
}
__device__ __forceinline__ void tc2_commit(uint32_t bar) {  // arrives on `bar` in BOTH CTAs of the pair
  asm volatile(
      "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
      "h"((unsigned short)3)
      : "memory");
}
__device__ __forceinline__ void tc2_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
// arrive on the copy of `bar` that lives in CTA `rank` of the cluster.  Relaxed: the arrival only hands
// TMEM back to the MMA warp (ordered by tcgen05.fence::before_thread_sync); a release would make the
// warp wait for all its outstanding global stores to drain (MEMBAR + ERRBAR) once per tile.
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t bar, uint32_t rank) {
  asm volatile(
      "{\n\t.reg .b32 ra;\n\t"
      "mapa.shared::cluster.u32 ra, %0, %1;\n\t"
      "mbarrier.arrive.relaxed.cluster.shared::cluster.b64 _, [ra];\n\t}" ::"r"(bar),
      "r"(rank)
      : "memory");
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}

__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint32_t bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void tc_mma(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t acc) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(acc)
      : "memory");
}
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr));
  asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, %0;" ::"n"(128 * EPI_SPLIT) : "memory"); }

// asynchronous 16-column TMEM load; the registers are only valid after tmem_wait() on the same array
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
      : "r"(taddr));
}
// tcgen05.wait::ld for all outstanding loads; the "+r" operands tie the loaded registers to the wait
// so the compiler cannot consume them (or move them) before it
__device__ __forceinline__ void tmem_wait(uint32_t* r) {
  asm volatile("tcgen05.wait::ld.sync.aligned;"
               : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                 "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
               :
               : "memory");
}

// UMMA shared-memory descriptor: K-major tile, 128-byte swizzle, rows of 128 B, 8-row atoms of
// 1024 B (SBO = 64 x 16 B), LBO unused for swizzled K-major (canonical value 1), version 1.
__device__ __forceinline__ uint64_t umma_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr >> 4) & 0x3FFF);
  d |= (uint64_t)1 << 16;
  d |= (uint64_t)64 << 32;
  d |= (uint64_t)1 << 46;
  d |= (uint64_t)2 << 61;
  return d;
}

// mish(x) = x * tanh(softplus(x)) = x * n / (n + 2),  n = e^x (e^x + 2).  Branch-free: clamping the
// exponent argument at 20 keeps n finite (2.4e17) and n / (n + 2) == 1 there, so mish(x) = x as it should;
// no per-element branch means the 32 MUFU chains of a chunk overlap.
__device__ __forceinline__ float mish_f(float x) {
  float e, inv;
  const float a = fminf(x, 20.0f) * 1.4426950408889634f;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e) : "f"(a));
  const float n = e * (e + 2.0f);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(inv) : "f"(n + 2.0f));
  return x * (n * inv);
}

// ---- packed fp32 (f32x2) arithmetic: one FFMA2 / FMUL2 / FADD2 does two lanes' worth of work per issue slot and
// per pass through the fma pipe, which is what bounds the fused epilogue (3-register FFMA: one warp instruction
// per two cycles per scheduler).  A u64 holds columns (j, j + 1): element j in the low half, as two consecutive
// floats loaded from memory do.
typedef unsigned long long u64;
__device__ __forceinline__ u64 pk2(float lo, float hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
__device__ __forceinline__ u64 pk2u(uint32_t lo, uint32_t hi) {
  u64 r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
  return r;
}
__device__ __forceinline__ void upk2(u64 v, float& lo, float& hi) {
  asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v));
}
__device__ __forceinline__ u64 ffma2(u64 a, u64 b, u64 c) {
  u64 r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
  return r;
}
__device__ __forceinline__ u64 fmul2(u64 a, u64 b) {
  u64 r;
  asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
__device__ __forceinline__ u64 fadd2(u64 a, u64 b) {
  u64 r;
  asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
  return r;
}
// mish on a pair: x * n / (n + 2) with n = e (e + 2), e = exp(x), written as x * (1 - 2 / (e (e + 2) + 2)): no
// clamp is needed (e = inf gives 1 / inf = 0, hence x; e = 0 gives 1 - 2 / 2 = 0) and the pair costs 5 packed
// operations + 4 MUFU instead of 2 x 9.  The subtraction loses RELATIVE accuracy only where mish itself is below
// 1e-5 in magnitude (x < -12): absolute error < 3e-7 everywhere, far below the bf16 output's resolution.
__device__ __forceinline__ u64 mish2(u64 x) {
  float a0, a1, e0, e1, d0, d1, r0, r1;
  upk2(fmul2(x, pk2(1.4426950408889634f, 1.4426950408889634f)), a0, a1);
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e0) : "f"(a0));
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(e1) : "f"(a1));
  const u64 e = pk2(e0, e1);
  upk2(ffma2(e, fadd2(e, pk2(2.0f, 2.0f)), pk2(2.0f, 2.0f)), d0, d1);
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r0) : "f"(d0));
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(r1) : "f"(d1));
  return fmul2(x, ffma2(pk2(r0, r1), pk2(-2.0f, -2.0f), pk2(1.0f, 1.0f)));
}
// two consecutive packed pairs (four floats) from shared memory in one LDS.128
__device__ __forceinline__ void lds4(const float* p, u64& a, u64& b) {
  const ulonglong2 v = *reinterpret_cast<const ulonglong2*>(p);
  a = v.x;
  b = v.y;
}

// 256-bit global accesses (sm_100: LDG.E.256 / STG.E.256): one instruction per thread moves the 32 bytes -- a whole
// sector -- a 16-column bf16 chunk of its row occupies, instead of two half-sector 128-bit ones
__device__ __forceinline__ void ldg256(const void* p, uint32_t* r) {
  asm volatile("ld.global.nc.L1::no_allocate.v8.b32 {%0, %1, %2, %3, %4, %5, %6, %7}, [%8];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7])
               : "l"(p));
}
__device__ __forceinline__ void stg256(void* p, const uint32_t* r) {
  asm volatile("st.global.v8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"l"(p), "r"(r[0]), "r"(r[1]), "r"(r[2]),
               "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
               : "memory");
}

template <int NG>
__device__ __forceinline__ float pick(const float (&a)[NG], int i) {
  float r = a[0];
#pragma unroll
  for (int k = 1; k < NG; ++k) r = (i == k) ? a[k] : r;
  return r;
}

#ifdef GEMM_TIMING
// variant build only (tools/gemm_timing.py): cycle accounting of the roles, summed over the tiles of flagged launches
//  [0] tiles seen by the MMA warp  [1] cycles waiting for a free accumulator  [2] main-loop cycles (issue to last commit)
//  [3] tiles seen by epilogue warp 0  [4] cycles waiting for the accumulator  [5] pass 1  [6] statistics + coefficients
//  [7] pass 2  [8] parameter staging + barriers before the wait  [9] kernel cycles (one CTA)  [10] CTAs
__device__ unsigned long long g_gemm_dbg[16];
extern "C" int dt_gemm_timing(unsigned long long* out16, int reset) {
  if (out16) cudaMemcpyFromSymbol(out16, g_gemm_dbg, sizeof(g_gemm_dbg));
  if (reset) {
    unsigned long long z[16] = {0};
    cudaMemcpyToSymbol(g_gemm_dbg, z, sizeof z);
  }
  return 0;
}
#define DBG_T(var) const long long var = clock64()
#define DBG_ADD(i, v) atomicAdd(&g_gemm_dbg[i], (unsigned long long)(v))
#else
#define DBG_T(var)
#define DBG_ADD(i, v)
#endif

struct GemmDev {
  int num_m_tiles, num_n_tiles, nseg, nkb_total;
  int T, rows_t, nb, tiles_per_sample;
  GemmSeg seg[GEMM_MAX_SEG];
  long long B;
  int N;
  // epilogue
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* film;
  long long film_ld;
  const float* film_t;
  const __nv_bfloat16* resid;
  long long ld_res;
  int relu;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  long long ldc, out_b_stride, out_t_stride, out_off;
  // split-K (small problems): a tile index also names a K slice of kb_per_slice 64-deep blocks; slice s writes
  // its partial sums slice_rows output rows further down (plain epilogue, fp32, reduced by k_splitk_epi)
  int ksplit, kb_per_slice;
  long long slice_rows;
  int dbg;   // GEMM_TIMING builds: account this launch
  int a_tile_bytes;  // bytes one activation TMA load delivers (rows_t * nb * 128; 16384 unless a tile is partly empty)
};

template <int BN, int CG>
struct SmemPlan {
  static constexpr int kStageA = BM * BK * 2;
  static constexpr int kStageB = (BN / CG) * BK * 2;  // with a CTA pair each CTA stages half of the weight tile
  static constexpr int kStage = kStageA + kStageB;
  static constexpr int kStages = (192 * 1024 / kStage) > 8 ? 8 : (192 * 1024 / kStage);
  static constexpr int kParamFloats = 5 * BN;                  // bias, gamma, beta, film_t scale, film_t shift
  static constexpr int kRedFloats = (4 * EPI_SPLIT) * 8 * 8 * 2;  // [warp][segment][group][sum,sq]
  static constexpr int kFilmSamples = 4;                       // tiles holding <= 4 samples (T >= 32) stage FiLM + GN coefficients
  static constexpr int kFilmFloats = 2 * kFilmSamples * 2 * BN;  // FiLM [sample][scale | shift][BN] + GN affine [sample][A | B][BN]
  static constexpr int kBytes =
      kStages * kStage + (kParamFloats + kRedFloats + kFilmFloats) * 4 + 256 /*barriers*/ + 1024 /*align*/;
};

// CG = 1: one CTA per tile (UMMA M = 128).  CG = 2: a CTA pair works on a 256-row tile with
// tcgen05.mma.cta_group::2 (UMMA M = 256): each CTA loads its own 128 activation rows and HALF of the
// weight tile, which cuts the L2 -> shared-memory traffic per MAC by a third and deepens the ring.
template <int BN, int EPI, int GW, int CG>
__global__ void __launch_bounds__(GEMM_THREADS, 1)
k_conv_gemm(const __grid_constant__ CUtensorMap mapA0, const __grid_constant__ CUtensorMap mapA1,
            const __grid_constant__ CUtensorMap mapW, const GemmDev g) {
  using P = SmemPlan<BN, CG>;
  const uint32_t rank = (CG == 2) ? cluster_ctarank() : 0u;   // 0 = leader (issues the MMAs)
  const int n_units = (CG == 2) ? (int)(gridDim.x >> 1) : (int)gridDim.x;  // tile-processing units (CTAs or pairs)
  const int unit = (CG == 2) ? (int)(blockIdx.x >> 1) : (int)blockIdx.x;
  constexpr int NG = (EPI == EPI_GN_MISH) ? BN / GW : 1;
  static_assert(NG <= 8, "at most 8 GroupNorm groups per N tile");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw = dt_smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;  // 128-byte swizzle atoms need 1024-byte alignment
  uint8_t* base_ptr = smem_raw + (base - raw);
  float* s_par = reinterpret_cast<float*>(base_ptr + P::kStages * P::kStage);
  float* s_red = s_par + P::kParamFloats;
  float* s_film = s_red + P::kRedFloats;
  float* s_coef = s_film + P::kFilmSamples * 2 * BN;
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_film + P::kFilmFloats);
  // barriers: full[kStages], empty[kStages], tmem_full[2], tmem_empty[2]; then the TMEM base word
  const uint32_t bar_full = dt_smem_u32(bars);
  const uint32_t bar_empty = bar_full + 8 * P::kStages;
  const uint32_t bar_tfull = bar_empty + 8 * P::kStages;
  const uint32_t bar_tempty = bar_tfull + 16;
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(bars + 2 * P::kStages + 4);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  constexpr uint32_t kTmemCols = 2 * BN;  // 128, 256 or 512: powers of two >= 32
  DBG_T(tk0);
  dt_pdl_launch();  // the next kernel's CTAs may take this SM's resources as soon as this CTA leaves

  if (warp == 0 && lane == 0) {
    for (int s = 0; s < P::kStages; ++s) {
      mbar_init(bar_full + 8 * s, 1);
      mbar_init(bar_empty + 8 * s, 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(bar_tfull + 8 * a, 1);
      mbar_init(bar_tempty + 8 * a, 4 * EPI_SPLIT * CG);  // one elected arrival per epilogue warp (of both CTAs)
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA0) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapA1) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&mapW) : "memory");
  }
  if (warp == 1) {
    if (CG == 2) {
      asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dt_smem_u32(s_tmem)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
    } else {
      asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dt_smem_u32(s_tmem)),
                   "r"(kTmemCols)
                   : "memory");
      asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
  }
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  // everything above (barriers, tensor-map prefetch, TMEM allocation, the pair's cluster sync) touched no global
  // memory and overlapped the previous kernel's tail; from here on its output is read
  dt_pdl_wait();

  // tiles are (unit-level M tile, N tile); a unit-level M tile is CG * 128 rows, CTA `rank` owns its 128-row slice
  const int unit_m_tiles = (g.num_m_tiles + CG - 1) / CG;
  const int total_tiles = unit_m_tiles * g.num_n_tiles * g.ksplit;

  if (warp == 0) {
    // ===================== TMA producer =====================
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units) {
      const int slice = tile % g.ksplit, tl = tile / g.ksplit;
      const int kb0 = slice * g.kb_per_slice;
      const int kb1 = (kb0 + g.kb_per_slice < g.nkb_total) ? kb0 + g.kb_per_slice : g.nkb_total;
      const int um_tile = tl / g.num_n_tiles, n_tile = tl - um_tile * g.num_n_tiles;
      const int m_tile = um_tile * CG + (int)rank;
      int b_base, t_base;
      if (g.tiles_per_sample > 0) {
        b_base = m_tile / g.tiles_per_sample;
        t_base = (m_tile - b_base * g.tiles_per_sample) * BM;
      } else {
        b_base = m_tile * g.nb;
        t_base = 0;
      }
      int kw = 0, wofs = 0;
      for (int s = 0; s < g.nseg; ++s) {
        const GemmSeg sg = g.seg[s];
        const CUtensorMap* mA = sg.src ? &mapA1 : &mapA0;
        wofs += sg.w_gap;  // weight blocks this launch leaves out
        for (int blk = 0; blk < sg.nblk; ++blk, ++kw) {
          if (kw < kb0 || kw >= kb1) continue;  // another K slice's block
          mbar_wait(bar_empty + 8 * stage, phase ^ 1);
          if (lane == 0) {
            const uint32_t sa = base + stage * P::kStage;
            if (CG == 2) {
              // both CTAs' bytes complete on the LEADER's full barrier; only the leader arms it
              if (rank == 0) mbar_expect_tx(bar_full + 8 * stage, 2 * (g.a_tile_bytes + P::kStageB));
              tma2_load_4d(sa, mA, bar_full + 8 * stage, blk * BK, sg.phase, t_base + sg.t_off, b_base);
              tma2_load_2d(sa + P::kStageA, &mapW, bar_full + 8 * stage, (kw + wofs) * BK, n_tile * BN + (int)rank * (BN / 2));
            } else {
              mbar_expect_tx(bar_full + 8 * stage, g.a_tile_bytes + P::kStageB);
              tma_load_4d(sa, mA, bar_full + 8 * stage, blk * BK, sg.phase, t_base + sg.t_off, b_base);
              tma_load_2d(sa + P::kStageA, &mapW, bar_full + 8 * stage, (kw + wofs) * BK, n_tile * BN);
            }
          }
          __syncwarp();
          if (++stage == P::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (leader CTA only when paired) =====================
    // instruction descriptor: D = f32, A = B = bf16, both K-major, N = BN, M = 128 * CG
    constexpr uint32_t idesc =
        (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(BN >> 3) << 17) | ((uint32_t)((BM * CG) >> 4) << 24);
    int stage = 0;
    uint32_t phase = 0;
    int acc = 0;
    uint32_t acc_phase = 0;
    if (rank == 0) {
      for (int tile = unit; tile < total_tiles; tile += n_units) {
        const int kb0_ = (tile % g.ksplit) * g.kb_per_slice;
        const int nkb = ((kb0_ + g.kb_per_slice < g.nkb_total) ? kb0_ + g.kb_per_slice : g.nkb_total) - kb0_;
        DBG_T(tm0);
        mbar_wait(bar_tempty + 8 * acc, acc_phase ^ 1);  // epilogue(s) have drained this accumulator
        tc_fence_after();
        DBG_T(tm1);
        const uint32_t d_addr = tmem_base + (uint32_t)(acc * BN);
        for (int kb = 0; kb < nkb; ++kb) {
          mbar_wait(bar_full + 8 * stage, phase);
          tc_fence_after();
          if (lane == 0) {
            const uint32_t sa = base + stage * P::kStage;
            const uint64_t da = umma_desc(sa), db = umma_desc(sa + P::kStageA);
#pragma unroll
            for (int k = 0; k < BK / 16; ++k) {
              // advancing K by 16 bf16 = 32 bytes inside the swizzle atom: +2 in the 16-byte address field
              if (CG == 2) tc2_mma(d_addr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
              else tc_mma(d_addr, da + (uint64_t)(2 * k), db + (uint64_t)(2 * k), idesc, (kb | k) != 0 ? 1u : 0u);
            }
            // free the smem slot (in both CTAs) when these MMAs retire; publish the accumulator after the last block
            if (CG == 2) {
              tc2_commit(bar_empty + 8 * stage);
              if (kb == nkb - 1) tc2_commit(bar_tfull + 8 * acc);
            } else {
              tc_commit(bar_empty + 8 * stage);
              if (kb == nkb - 1) tc_commit(bar_tfull + 8 * acc);
            }
          }
          __syncwarp();
          if (++stage == P::kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
#ifdef GEMM_TIMING
        if (g.dbg && lane == 0) {
          const long long tm2 = clock64();
          DBG_ADD(0, 1);
          DBG_ADD(1, tm1 - tm0);
          DBG_ADD(2, tm2 - tm1);
        }
#endif
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
      }
    }
  } else {
    // ===================== epilogue (warps 2..9) =====================
    // Two warps per TMEM lane quarter (a warp may only touch lanes 32*(warp%4)..+31): the pair splits
    // the BN accumulator columns in halves.  TMEM is read in 16-column chunks, double-buffered in
    // registers (the load of chunk c+1 is in flight while chunk c is processed); the per-column
    // parameters and FiLM rows of the NEXT tile are prefetched into registers while this tile is
    // processed, so their global-memory latency is never exposed.
    constexpr int HALF = BN / EPI_SPLIT;    // columns per epilogue warp
    constexpr int CH = 16;
    constexpr int NCH = HALF / CH;
    static_assert(HALF % CH == 0, "column split must be a multiple of the TMEM chunk");
    constexpr int NET = 128 * EPI_SPLIT;    // epilogue threads
    const int q = warp & 3;                 // TMEM lane quarter
    const int half = (warp - 2) >> 2;       // which slice of the columns (0 .. EPI_SPLIT-1)
    const int ew = warp - 2;                // 0 .. 4*EPI_SPLIT-1
    const int row = q * 32 + lane;          // row inside the M tile
    const int et = threadIdx.x - 64;        // 0 .. NET-1
    const int pcol = et % BN;               // the column whose parameters this thread stages
    const int ns = g.tiles_per_sample > 0 ? 1 : g.nb;  // samples per tile
    const bool film_smem = (EPI == EPI_GN_MISH) && g.film && ns <= P::kFilmSamples;
    constexpr int PF = (P::kFilmSamples * 2 * BN + NET - 1) / NET;  // FiLM values staged per thread (at most)
    static_assert(NG <= 8, "red[] holds 8 groups per warp");
    float pf_par[5] = {0.f, 0.f, 0.f, 0.f, 0.f};
    float pf_film[PF];
    auto prefetch = [&](int tile_k) {
      const int tile_ = tile_k / g.ksplit;  // drop the K slice
      const int um_ = tile_ / g.num_n_tiles, n0_ = (tile_ - um_ * g.num_n_tiles) * BN;
      const int m_ = um_ * CG + (int)rank;
      pf_par[0] = g.bias ? __ldg(g.bias + n0_ + pcol) : 0.f;
      if (g.resid) {
        // pull this thread's slice of the NEXT tile's residual row towards L2 a whole tile ahead: the epilogue's
        // own loads (one 16-column chunk ahead, in registers) then meet an L2 hit instead of HBM latency
        // (ncu, r02: 30 % of the stall samples of a residual layer sat on those loads)
        long long b_;
        int t_;
        if (g.tiles_per_sample > 0) {
          b_ = m_ / g.tiles_per_sample;
          t_ = (m_ - (int)b_ * g.tiles_per_sample) * BM + row;
        } else {
          b_ = (long long)m_ * g.nb + row / g.T;
          t_ = row % g.T;
        }
        if (b_ < g.B && t_ < g.T) {
          const __nv_bfloat16* rp = g.resid + (b_ * g.out_b_stride + (long long)t_ * g.out_t_stride + g.out_off) * g.ld_res +
                                    n0_ + half * HALF;
#pragma unroll
          for (int o = 0; o < HALF * 2; o += 128)
            asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(rp) + o));
        }
      }
      if (EPI == EPI_GN_RELU) {
        pf_par[1] = __ldg(g.gamma + n0_ + pcol);
        pf_par[2] = __ldg(g.beta + n0_ + pcol);
      }
      if (EPI == EPI_GN_MISH) {
        pf_par[1] = __ldg(g.gamma + n0_ + pcol);
        pf_par[2] = __ldg(g.beta + n0_ + pcol);
        pf_par[3] = g.film_t ? __ldg(g.film_t + n0_ + pcol) : 0.f;
        pf_par[4] = g.film_t ? __ldg(g.film_t + g.N + n0_ + pcol) : 0.f;
        if (film_smem) {
          const long long b_first = g.tiles_per_sample > 0 ? (long long)(m_ / g.tiles_per_sample) : (long long)m_ * g.nb;
#pragma unroll
          for (int j = 0; j < PF; ++j) {
            const int i = et + NET * j;  // index into [sample][part][BN]; its column is pcol for every j
            const int smp = i / (2 * BN), part = (i / BN) & 1;
            const long long bb = b_first + smp;
            pf_film[j] = (smp < ns && bb < g.B) ? __ldg(g.film + bb * g.film_ld + part * g.N + n0_ + pcol) : 0.f;
          }
        }
      }
    };
    if (unit < total_tiles) prefetch(unit);
    int acc = 0;
    uint32_t acc_phase = 0;
    int it = 0;
    for (int tile = unit; tile < total_tiles; tile += n_units, ++it) {
      const int slice = tile % g.ksplit, tl = tile / g.ksplit;
      const int um_tile = tl / g.num_n_tiles, n_tile = tl - um_tile * g.num_n_tiles;
      const int m_tile = um_tile * CG + (int)rank;
      const int n0 = n_tile * BN;
      long long b;
      int t;
      if (g.tiles_per_sample > 0) {
        b = m_tile / g.tiles_per_sample;
        t = (m_tile - (int)b * g.tiles_per_sample) * BM + row;
      } else {
        b = (long long)m_tile * g.nb + row / g.T;
        t = row % g.T;
      }
      // (a tile of whole samples may be partly empty when T does not divide 128: those rows belong to nobody)
      const bool valid = (b < g.B) && (t < g.T) && (g.tiles_per_sample > 0 || row < g.nb * g.T);
      DBG_T(te0);
      // publish this tile's (prefetched) parameters; the first barrier orders the previous tile's readers
      epi_bar_sync();
      if (et < BN) {
        s_par[et] = pf_par[0];
        if (EPI != EPI_PLAIN) {
          s_par[BN + et] = pf_par[1];
          s_par[2 * BN + et] = pf_par[2];
        }
      }

      if (film_smem) {
#pragma unroll
        for (int j = 0; j < PF; ++j) {
          const int i = et + NET * j;
          if (i < ns * 2 * BN) s_film[i] = pf_film[j] + (((i / BN) & 1) ? pf_par[4] : pf_par[3]);
        }
      } else if (EPI == EPI_GN_MISH && et < BN) {
        s_par[3 * BN + et] = pf_par[3];
        s_par[4 * BN + et] = pf_par[4];
      }
      epi_bar_sync();
      if (tile + n_units < total_tiles) prefetch(tile + n_units);  // in flight during this tile
      DBG_T(te1);
      mbar_wait(bar_tfull + 8 * acc, acc_phase);
      tc_fence_after();
      DBG_T(te2);
#ifdef GEMM_TIMING
      long long te3 = te2, te4 = te2;
#endif
      const uint32_t taddr = tmem_base + ((uint32_t)(q * 32) << 16) + (uint32_t)(acc * BN + half * HALF);
      const long long out_row = b * g.out_b_stride + (long long)t * g.out_t_stride + g.out_off + slice * g.slice_rows;
      const float* sp = s_par + half * HALF;  // this warp's column window of the staged parameters

      if constexpr (EPI == EPI_GN_RELU) {
        // ---- ResNet encoder: GroupNorm over (T rows x 16 channels) per sample, any T <= 128 ----
        // Deterministic (no atomics: the same inputs give the same bits whatever shares the tile): every row leaves
        // its partial sums in shared memory, then one thread per (sample, group) adds the T rows in row order.
        constexpr int NGT = BN / 16;          // groups per tile; one 16-column TMEM chunk is exactly one group
        constexpr int NCHR = HALF / 16;       // chunks (= groups) per epilogue warp
        float* s_part = s_film;               // [row][group][sum, sum of squares]   (128 * NGT * 2 floats)
        float* s_ms = s_red;                  // [sample][group][mean, rstd]          (T >= 2: <= 64 * NGT * 2 floats)
        const int smp_l = row / g.T;          // sample inside the tile (rows >= nb * T are nobody's: `valid` is false)
        const float inv_n = 1.0f / (float)(g.T * 16);
        float own_mean[NCHR], own_rstd[NCHR]; // T == 1: a row is a whole sample
        uint32_t rr[2][CH];
        tmem_ld16(taddr, rr[0]);
#pragma unroll
        for (int c = 0; c < NCHR; ++c) {
          tmem_wait(rr[c & 1]);
          if (c + 1 < NCHR) tmem_ld16(taddr + (c + 1) * CH, rr[(c + 1) & 1]);
          float sm = 0.f, sq = 0.f;
#pragma unroll
          for (int j = 0; j < CH; ++j) {
            const float v = __uint_as_float(rr[c & 1][j]) + sp[c * CH + j];
            sm += v;
            sq = fmaf(v, v, sq);
          }
          // (no divergence around the warp-aligned tcgen05.ld / wait of the next iteration: every lane stores)
          float* pp = s_part + (row * NGT + half * NCHR + c) * 2;
          pp[0] = valid ? sm : 0.f;
          pp[1] = valid ? sq : 0.f;
          own_mean[c] = sm * inv_n;
          own_rstd[c] = rsqrtf(fmaxf(sq * inv_n - own_mean[c] * own_mean[c], 0.f) + 1e-5f);
        }
        if (g.T > 1) {
          epi_bar_sync();
          for (int pr = et; pr < g.nb * NGT; pr += NET) {
            const int smp = pr / NGT, grp = pr - smp * NGT;
            const float* pp = s_part + (smp * g.T * NGT + grp) * 2;
            float a0 = 0.f, a1 = 0.f;
            for (int r = 0; r < g.T; ++r) {
              a0 += pp[r * NGT * 2];
              a1 += pp[r * NGT * 2 + 1];
            }
            const float mean = a0 * inv_n;
            s_ms[pr * 2] = mean;
            s_ms[pr * 2 + 1] = rsqrtf(fmaxf(a1 * inv_n - mean * mean, 0.f) + 1e-5f);
          }
          epi_bar_sync();
        }
        const __nv_bfloat16* res_row = (g.resid && valid) ? g.resid + out_row * g.ld_res + n0 + half * HALF : nullptr;
        __nv_bfloat16* out_b = g.out_bf16 + out_row * g.ldc + n0 + half * HALF;
        tmem_ld16(taddr, rr[0]);
#pragma unroll
        for (int c = 0; c < NCHR; ++c) {
          tmem_wait(rr[c & 1]);
          if (c + 1 < NCHR) {
            tmem_ld16(taddr + (c + 1) * CH, rr[(c + 1) & 1]);
          } else {
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
              if (CG == 2) mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
              else mbar_arrive_relaxed(bar_tempty + 8 * acc);
            }
          }
          uint32_t rs8[8] = {0u, 0u, 0u, 0u, 0u, 0u, 0u, 0u};
          if (res_row) ldg256(res_row + c * CH, rs8);
          __syncwarp();
          const float* ms = s_ms + ((valid ? smp_l : 0) * NGT + half * NCHR + c) * 2;
          const float mean = g.T > 1 ? ms[0] : own_mean[c];
          const float rstd = g.T > 1 ? ms[1] : own_rstd[c];
          uint32_t pk[CH / 2];
#pragma unroll
          for (int u = 0; u < CH / 2; ++u) {
            float y0 = __uint_as_float(rr[c & 1][2 * u]) + sp[c * CH + 2 * u];
            float y1 = __uint_as_float(rr[c & 1][2 * u + 1]) + sp[c * CH + 2 * u + 1];
            y0 = fmaf((y0 - mean) * rstd, sp[BN + c * CH + 2 * u], sp[2 * BN + c * CH + 2 * u]);
            y1 = fmaf((y1 - mean) * rstd, sp[BN + c * CH + 2 * u + 1], sp[2 * BN + c * CH + 2 * u + 1]);
            if (res_row) {
              y0 += __uint_as_float(rs8[u] << 16);
              y1 += __uint_as_float(rs8[u] & 0xFFFF0000u);
            }
            if (g.relu) {
              y0 = fmaxf(y0, 0.f);
              y1 = fmaxf(y1, 0.f);
            }
            const __nv_bfloat162 h2 = __floats2bfloat162_rn(y0, y1);
            pk[u] = *reinterpret_cast<const uint32_t*>(&h2);
          }
          if (valid) stg256(out_b + c * CH, pk);
          __syncwarp();
        }
        if (++acc == 2) {
          acc = 0;
          acc_phase ^= 1;
        }
        continue;
      }

      // GroupNorm groups seen by this warp's column slice: NLG whole groups when a group fits in the
      // slice, otherwise the slice is part of ONE group that spans SPG slices.
      constexpr int NLG = (GW <= HALF) ? HALF / GW : 1;
      constexpr int SPG = (GW <= HALF) ? 1 : GW / HALF;
      float mean[NLG], rstd[NLG];
      uint32_t rb[2][CH];
      const int smp_in_tile = g.tiles_per_sample > 0 ? 0 : row / g.T;
      if (EPI == EPI_GN_MISH) {
        float gs[NLG], gq[NLG];
        u64 gs2[NLG], gq2[NLG];   // (even column, odd column) partial sums, folded after the loop
#pragma unroll
        for (int i = 0; i < NLG; ++i) gs2[i] = gq2[i] = 0ull;
        // pass 1: per-row partial sums of (acc + bias) and its square for the groups of this slice
        tmem_ld16(taddr, rb[0]);
#pragma unroll
        for (int c = 0; c < NCH; ++c) {
          tmem_wait(rb[c & 1]);
          if (c + 1 < NCH) tmem_ld16(taddr + (c + 1) * CH, rb[(c + 1) & 1]);
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            u64 b01, b23;
            lds4(sp + c * CH + j, b01, b23);
            const u64 v01 = fadd2(pk2u(rb[c & 1][j], rb[c & 1][j + 1]), b01);
            const u64 v23 = fadd2(pk2u(rb[c & 1][j + 2], rb[c & 1][j + 3]), b23);
            const int u = (GW <= HALF) ? (c * CH + j) / GW : 0;   // GW >= 8: the four columns share a group
            gs2[u] = fadd2(gs2[u], v01);
            gq2[u] = ffma2(v01, v01, gq2[u]);
            gs2[u] = fadd2(gs2[u], v23);
            gq2[u] = ffma2(v23, v23, gq2[u]);
          }
        }
#pragma unroll
        for (int i = 0; i < NLG; ++i) {
          float lo, hi;
          upk2(gs2[i], lo, hi);
          gs[i] = lo + hi;
          upk2(gq2[i], lo, hi);
          gq[i] = lo + hi;
        }
#ifdef GEMM_TIMING
        te3 = clock64();
#endif
        // reduce over the rows of this sample held by this warp (segments of min(T,32) lanes) ...
        const int span = g.T < 32 ? g.T : 32;
#pragma unroll
        for (int i = 0; i < NLG; ++i) {
          for (int off = 1; off < span; off <<= 1) {
            gs[i] += __shfl_xor_sync(0xffffffffu, gs[i], off);
            gq[i] += __shfl_xor_sync(0xffffffffu, gq[i], off);
          }
        }
        // ... then across warps through shared memory: other lane quarters when a sample spans several
        // warps (T > 32), other column slices when a group spans several slices (GW > HALF)
        float* red = s_red;  // [warp][segment][local group][sum, sq]; reuse across tiles is ordered by the top barriers
        const int seg = lane / span;
        if ((lane % span) == 0) {
#pragma unroll
          for (int i = 0; i < NLG; ++i) {
            red[((ew * 8 + seg) * 8 + i) * 2 + 0] = gs[i];
            red[((ew * 8 + seg) * 8 + i) * 2 + 1] = gq[i];
          }
        }
        epi_bar_sync();
        const int rows_s = g.T > BM ? BM : g.T;         // rows of one sample inside this tile
        const int wps = rows_s > 32 ? rows_s / 32 : 1;  // lane quarters per sample
        const int q0 = (q / wps) * wps;
        const int h0 = (half / SPG) * SPG;
        const float inv_n = 1.0f / (float)(rows_s * GW);
#pragma unroll
        for (int i = 0; i < NLG; ++i) {
          float sa = 0.f, sq = 0.f;
          for (int qq = q0; qq < q0 + wps; ++qq) {
#pragma unroll
            for (int hh = 0; hh < SPG; ++hh) {
              const int w2 = ((qq + 2) & 3) + 4 * (h0 + hh);  // ew of the warp with lane quarter qq, column slice h0+hh
              sa += red[((w2 * 8 + seg) * 8 + i) * 2 + 0];
              sq += red[((w2 * 8 + seg) * 8 + i) * 2 + 1];
            }
          }
          mean[i] = sa * inv_n;
          const float var = fmaxf(sq * inv_n - mean[i] * mean[i], 0.f);
          rstd[i] = rsqrtf(var + 1e-5f);
        }
      }

      // Per-(sample, column) affine of the normalisation, shared by the rows of a sample:
      //   t = ((acc + bias) - mean) * rstd * gamma + beta = acc * A + B
      // computed once per tile into shared memory when a tile holds few samples (coef_smem).
      const bool coef_smem = (EPI == EPI_GN_MISH) && ns <= P::kFilmSamples;
      if (coef_smem) {
        const int rows_s = g.T > BM ? BM : g.T;
        const int idx = (g.tiles_per_sample > 0 ? row : row % g.T);  // this thread's rank among the rows of its sample
        for (int cl = idx; cl < HALF; cl += rows_s) {
          const float m_ = pick<NLG>(mean, (GW <= HALF) ? cl / GW : 0);
          const float r_ = pick<NLG>(rstd, (GW <= HALF) ? cl / GW : 0);
          const float A = r_ * sp[BN + cl];
          s_coef[(smp_in_tile * 2 + 0) * BN + half * HALF + cl] = A;
          s_coef[(smp_in_tile * 2 + 1) * BN + half * HALF + cl] = (sp[cl] - m_) * A + sp[2 * BN + cl];
        }
        epi_bar_sync();
      }

#ifdef GEMM_TIMING
      te4 = clock64();
#endif
      // pass 2: normalise / activate / modulate and store this warp's slice of the columns
      const int nh = n0 + half * HALF;
      const float* film_row = (EPI == EPI_GN_MISH && g.film && valid && !film_smem) ? g.film + b * g.film_ld + nh : nullptr;
      const float* fs = s_film + smp_in_tile * 2 * BN + half * HALF;  // staged FiLM scale row; shift row at + BN
      const float* cf = s_coef + smp_in_tile * 2 * BN + half * HALF;  // staged A row; B row at + BN
      const __nv_bfloat16* res_row = (g.resid && valid) ? g.resid + out_row * g.ld_res + nh : nullptr;
      __nv_bfloat16* out_b = g.out_bf16 ? g.out_bf16 + out_row * g.ldc + nh : nullptr;
      float* out_f = g.out_f32 ? g.out_f32 + out_row * g.ldc + nh : nullptr;
      // residual: 32 bytes (16 bf16 columns) per chunk, fetched RES_PF chunks ahead of their use
      constexpr int RES_PF = 2;
      uint32_t resq[RES_PF + 1][8];
      if (res_row) {
#pragma unroll
        for (int d = 0; d < RES_PF; ++d)
          if (d < NCH) ldg256(res_row + d * CH, resq[d]);
      }
      tmem_ld16(taddr, rb[0]);
#pragma unroll
      for (int c = 0; c < NCH; ++c) {
        const int c0 = c * CH;
        tmem_wait(rb[c & 1]);
        if (c + 1 < NCH) {
          tmem_ld16(taddr + c0 + CH, rb[(c + 1) & 1]);
        } else {
          // the last TMEM read of this tile has landed in registers: hand the accumulator back to the
          // (leader's) MMA warp now, one elected arrival per warp, and finish the math / stores afterwards
          tc_fence_before();
          __syncwarp();
          if (lane == 0) {
            if (CG == 2) mbar_arrive_cluster(bar_tempty + 8 * acc, 0);
            else mbar_arrive_relaxed(bar_tempty + 8 * acc);
          }
        }
        if (res_row && c + RES_PF < NCH) ldg256(res_row + c0 + RES_PF * CH, resq[(c + RES_PF) % (RES_PF + 1)]);
        const uint32_t* r = rb[c & 1];
        u64 y2[CH / 2];   // y2[i] = columns (c0 + 2 i, c0 + 2 i + 1)
        if (EPI == EPI_GN_MISH) {
          if (coef_smem) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              u64 A01, A23, B01, B23;
              lds4(cf + c0 + j, A01, A23);
              lds4(cf + BN + c0 + j, B01, B23);
              y2[j / 2] = mish2(ffma2(pk2u(r[j], r[j + 1]), A01, B01));
              y2[j / 2 + 1] = mish2(ffma2(pk2u(r[j + 2], r[j + 3]), A23, B23));
            }
          } else {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              u64 b01, b23, g01, g23, e01, e23;
              lds4(sp + c0 + j, b01, b23);
              lds4(sp + BN + c0 + j, g01, g23);
              lds4(sp + 2 * BN + c0 + j, e01, e23);
              const int u = (GW <= HALF) ? (c0 + j) / GW : 0;   // GW >= 8: the four columns share a group
              const u64 rs = pk2(rstd[u], rstd[u]), nm = pk2(-mean[u] * rstd[u], -mean[u] * rstd[u]);
              // ((acc + bias) - mean) * rstd * gamma + beta
              y2[j / 2] = mish2(ffma2(ffma2(fadd2(pk2u(r[j], r[j + 1]), b01), rs, nm), g01, e01));
              y2[j / 2 + 1] = mish2(ffma2(ffma2(fadd2(pk2u(r[j + 2], r[j + 3]), b23), rs, nm), g23, e23));
            }
          }
          if (film_smem) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              u64 sc01, sc23, sh01, sh23;
              lds4(fs + c0 + j, sc01, sc23);
              lds4(fs + BN + c0 + j, sh01, sh23);
              y2[j / 2] = ffma2(y2[j / 2], sc01, sh01);
              y2[j / 2 + 1] = ffma2(y2[j / 2 + 1], sc23, sh23);
            }
          } else if (film_row) {
#pragma unroll
            for (int j = 0; j < CH; j += 4) {
              const ulonglong2 sc = __ldg(reinterpret_cast<const ulonglong2*>(film_row + c0 + j));
              const ulonglong2 sh = __ldg(reinterpret_cast<const ulonglong2*>(film_row + g.N + c0 + j));
              u64 ts01, ts23, tb01, tb23;
              lds4(sp + 3 * BN + c0 + j, ts01, ts23);
              lds4(sp + 4 * BN + c0 + j, tb01, tb23);
              y2[j / 2] = ffma2(y2[j / 2], fadd2(sc.x, ts01), fadd2(sh.x, tb01));
              y2[j / 2 + 1] = ffma2(y2[j / 2 + 1], fadd2(sc.y, ts23), fadd2(sh.y, tb23));
            }
          }
        } else {
#pragma unroll
          for (int j = 0; j < CH; j += 4) {
            u64 b01, b23;
            lds4(sp + c0 + j, b01, b23);
            y2[j / 2] = fadd2(pk2u(r[j], r[j + 1]), b01);
            y2[j / 2 + 1] = fadd2(pk2u(r[j + 2], r[j + 3]), b23);
          }
        }
        if (res_row) {
          const uint32_t* w8 = resq[c % (RES_PF + 1)];
#pragma unroll
          for (int u = 0; u < CH / 2; ++u)  // bf16 -> fp32 is a 16-bit shift: word u holds columns (2u, 2u + 1)
            y2[u] = fadd2(y2[u], pk2u(w8[u] << 16, w8[u] & 0xFFFF0000u));
        }
        float y[CH];
#pragma unroll
        for (int i = 0; i < CH / 2; ++i) upk2(y2[i], y[2 * i], y[2 * i + 1]);
        if (EPI == EPI_PLAIN && g.relu) {
#pragma unroll
          for (int j = 0; j < CH; ++j) y[j] = fmaxf(y[j], 0.f);
        }
        if (valid) {
          if (out_b) {
            uint32_t pk[CH / 2];
#pragma unroll
            for (int u = 0; u < CH / 2; ++u) {
              const __nv_bfloat162 h2 = __floats2bfloat162_rn(y[2 * u], y[2 * u + 1]);
              pk[u] = *reinterpret_cast<const uint32_t*>(&h2);
            }
            stg256(out_b + c0, pk);
          }
          if (out_f) {
#pragma unroll
            for (int j = 0; j < CH; j += 4)
              *reinterpret_cast<float4*>(out_f + c0 + j) = make_float4(y[j], y[j + 1], y[j + 2], y[j + 3]);
          }
        }
      }
#ifdef GEMM_TIMING
      if (g.dbg && ew == 0 && lane == 0) {
        const long long te5 = clock64();
        DBG_ADD(3, 1);
        DBG_ADD(8, te1 - te0);
        DBG_ADD(4, te2 - te1);
        DBG_ADD(5, te3 - te2);
        DBG_ADD(6, te4 - te3);
        DBG_ADD(7, te5 - te4);
      }
#endif
      if (++acc == 2) {
        acc = 0;
        acc_phase ^= 1;
      }
    }
  }

#ifdef GEMM_TIMING
  if (g.dbg && threadIdx.x == 64) {
    DBG_ADD(9, clock64() - tk0);
    DBG_ADD(10, 1);
  }
#endif
  tc_fence_before();
  if (CG == 2) cluster_sync_all(); else __syncthreads();  // the peer's MMAs read this CTA's smem / TMEM until here
  if (warp == 1) {
    if (CG == 2)
      asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
    else
      asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(kTmemCols) : "memory");
  }
}

// ------------------------------------------------------------------------------------------
// host side: tensor maps and dispatch
// ------------------------------------------------------------------------------------------
cudaEvent_t dt_prof_event(dt_ctx* ctx);  // ctx.cu

typedef CUresult (*PFN_tmapEncodeTiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                        const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static PFN_tmapEncodeTiled get_encode() {
  static PFN_tmapEncodeTiled fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (PFN_tmapEncodeTiled)p;
  }
  return fn;
}

static int make_act_map(dt_ctx* ctx, CUtensorMap* m, const ActSrc& a, int64_t B, int rows_t, int nb) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return dt_fail(ctx, DT_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
  if (a.W2 > 0) {
    // 2-D conv source (B, H2, W2, C): the box is one tap's view of nb samples' output maps -- ow2 x oh2 pixels read
    // with the conv stride (TMA traversal stride), rows land in shared memory as (sample, oy, ox), 64 channels each
    const int s = a.stride2;
    const int bw = (a.ow2 - 1) * s + 1, bh = (a.oh2 - 1) * s + 1;
    if (a.C % 8 != 0 || s < 1 || bw > 256 || bh > 256 || nb > 256 || a.ow2 * a.oh2 != rows_t)
      return dt_fail(ctx, DT_E_UNSUPPORTED, "2-D activation shape not TMA-addressable");
    cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.W2, (cuuint64_t)a.H2, (cuuint64_t)B};
    cuuint64_t strides[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.W2 * a.C * 2, (cuuint64_t)a.H2 * a.W2 * a.C * 2};
    cuuint32_t box[4] = {BK, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)nb};
    cuuint32_t estr[4] = {1, (cuuint32_t)s, (cuuint32_t)s, 1};
    CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)a.ptr, dims, strides, box, estr,
                     CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                     CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (r != CUDA_SUCCESS) {
      char buf[200];
      snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(act2d C=%d W=%d H=%d s=%d ow=%d oh=%d nb=%d) failed: %d", a.C, a.W2,
               a.H2, s, a.ow2, a.oh2, nb, (int)r);
      return dt_fail(ctx, DT_E_CUDA, buf);
    }
    return DT_OK;
  }
  if (a.C % 8 != 0 || a.T_in % a.P != 0) return dt_fail(ctx, DT_E_UNSUPPORTED, "activation shape not TMA-addressable");
  cuuint64_t dims[4] = {(cuuint64_t)a.C, (cuuint64_t)a.P, (cuuint64_t)(a.T_in / a.P), (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)a.C * 2, (cuuint64_t)a.P * a.C * 2, (cuuint64_t)a.T_in * a.C * 2};
  cuuint32_t box[4] = {BK, 1, (cuuint32_t)rows_t, (cuuint32_t)nb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, (void*)a.ptr, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(act C=%d T=%d P=%d B=%lld) failed: %d", a.C, a.T_in, a.P,
             (long long)B, (int)r);
    return dt_fail(ctx, DT_E_CUDA, buf);
  }
  return DT_OK;
}

static int make_w_map(dt_ctx* ctx, CUtensorMap* m, const __nv_bfloat16* w, int N, int64_t Ktot, int bn) {
  PFN_tmapEncodeTiled enc = get_encode();
  if (!enc) return dt_fail(ctx, DT_E_CUDA, "cuTensorMapEncodeTiled entry point not found");
  cuuint64_t dims[2] = {(cuuint64_t)Ktot, (cuuint64_t)N};
  cuuint64_t strides[1] = {(cuuint64_t)Ktot * 2};
  cuuint32_t box[2] = {BK, (cuuint32_t)bn};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = enc(m, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, (void*)w, dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    char buf[160];
    snprintf(buf, sizeof buf, "cuTensorMapEncodeTiled(w N=%d K=%lld) failed: %d", N, (long long)Ktot, (int)r);
    return dt_fail(ctx, DT_E_CUDA, buf);
  }
  return DT_OK;
}

template <int BN, int EPI, int GW, int CG>
static int launch_gemm_cg(dt_ctx* ctx, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w,
                          const GemmDev& d, cudaStream_t st) {
  using P = SmemPlan<BN, CG>;
  static unsigned long long attr_set = 0;  // one bit per device: the attribute is per device
  const unsigned long long dev_bit = 1ull << (ctx->device & 63);
  if (!(attr_set & dev_bit)) {
    DT_CUDA(cudaFuncSetAttribute(k_conv_gemm<BN, EPI, GW, CG>, cudaFuncAttributeMaxDynamicSharedMemorySize, P::kBytes));
    attr_set |= dev_bit;
  }
  const int unit_tiles = ((d.num_m_tiles + CG - 1) / CG) * d.num_n_tiles * d.ksplit;
  const int max_units = ctx->sm_count / CG;  // persistent: one CTA (or CTA pair) per SM (pair)
  const int units = unit_tiles < max_units ? unit_tiles : max_units;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3(units * CG);
  cfg.blockDim = dim3(GEMM_THREADS);
  cfg.dynamicSmemBytes = P::kBytes;
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = CG;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl_now ? 2 : 1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->prof_on) {
    e0 = dt_prof_event(ctx);
    e1 = dt_prof_event(ctx);
    if (e0 && e1) {
      dt_ctx::ProfRec rec{BN, EPI, GW * 10 + CG, (long long)d.B * d.T, d.N, (long long)d.nkb_total * BK, 0.f, d.ksplit};
      ctx->prof_recs.push_back(rec);
      cudaEventRecord(e0, st);
    }
  }
  DT_CUDA(cudaLaunchKernelEx(&cfg, k_conv_gemm<BN, EPI, GW, CG>, a0, a1, w, d));
  if (e0 && e1) cudaEventRecord(e1, st);
  DT_LAUNCH_CHECK("k_conv_gemm");
  return DT_OK;
}

template <int BN, int EPI, int GW>
static int launch_gemm(dt_ctx* ctx, int cg, const CUtensorMap& a0, const CUtensorMap& a1, const CUtensorMap& w,
                       const GemmDev& d, cudaStream_t st) {
  if (cg == 2) return launch_gemm_cg<BN, EPI, GW, 2>(ctx, a0, a1, w, d, st);
  return launch_gemm_cg<BN, EPI, GW, 1>(ctx, a0, a1, w, d, st);
}

// 1 = one CTA per tile, 2 = CTA pairs (default); DITREE_GEMM_CG=1 forces the single-CTA kernels (debugging)
static int gemm_cta_group() {
  static int v = -1;
  if (v < 0) {
    const char* e = getenv("DITREE_GEMM_CG");
    v = (e && e[0] == '1') ? 1 : 2;
  }
  return v;
}

// ---------------------------------------------------------------------------------------------
// GroupNorm groups wider than one N tile (the `xlarge` preset: 4096 / 8 = 512 channels): the statistics of
// a group cannot be taken inside one CTA's accumulator, so the GEMM runs with the plain epilogue into an
// fp32 scratch and this kernel applies GroupNorm -> Mish -> FiLM (+ residual) per (sample, group).
// One block per (sample, group): T x GW fp32 values, two passes over them (they stay in L2 / L1).
// ---------------------------------------------------------------------------------------------
struct WideGn {
  const float* y;        // [B*T][N] fp32, bias already added
  int T, N, gw;
  const float* gamma;
  const float* beta;
  const float* film;     // [B][film_ld] or null
  long long film_ld;
  const float* film_t;   // [2N] or null
  const __nv_bfloat16* resid;
  long long ld_res;
  __nv_bfloat16* out;
  long long ldc, out_b_stride, out_t_stride, out_off;
};

__global__ void __launch_bounds__(256)
k_gn_mish_wide(WideGn p) {
  __shared__ float s_a[256], s_b[256];
  const int groups = p.N / p.gw;
  const long long b = blockIdx.x / groups;
  const int grp = blockIdx.x % groups;
  const int n0 = grp * p.gw;
  const int cnt = p.T * p.gw;
  const float* base = p.y + b * p.T * (long long)p.N + n0;
  float s = 0.f, ss = 0.f;
  for (int i = threadIdx.x; i < cnt; i += 256) {
    const int t = i / p.gw, c = i - t * p.gw;
    const float v = base[(long long)t * p.N + c];
    s += v;
    ss += v * v;
  }
  s_a[threadIdx.x] = s;
  s_b[threadIdx.x] = ss;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o) {
      s_a[threadIdx.x] += s_a[threadIdx.x + o];
      s_b[threadIdx.x] += s_b[threadIdx.x + o];
    }
    __syncthreads();
  }
  const float mean = s_a[0] / (float)cnt;
  const float var = fmaxf(s_b[0] / (float)cnt - mean * mean, 0.f);
  const float rstd = rsqrtf(var + 1e-5f);
  for (int i = threadIdx.x; i < cnt; i += 256) {
    const int t = i / p.gw, c = i - t * p.gw, n = n0 + c;
    float v = mish_f((base[(long long)t * p.N + c] - mean) * rstd * p.gamma[n] + p.beta[n]);
    if (p.film) {
      const float sc = p.film[b * p.film_ld + n] + (p.film_t ? p.film_t[n] : 0.f);
      const float sh = p.film[b * p.film_ld + p.N + n] + (p.film_t ? p.film_t[p.N + n] : 0.f);
      v = v * sc + sh;
    }
    const long long row = b * p.out_b_stride + (long long)t * p.out_t_stride + p.out_off;
    if (p.resid) v += __bfloat162float(p.resid[row * p.ld_res + n]);
    p.out[row * p.ldc + n] = __float2bfloat16(v);
  }
}

int dt_conv_gemm(dt_ctx* ctx, const ConvGemm& g, cudaStream_t st);

static int conv_gemm_wide_gn(dt_ctx* ctx, const ConvGemm& g, cudaStream_t st) {
  if (g.N % g.group_width != 0 || !g.gamma || !g.beta || !g.out_bf16)
    return dt_fail(ctx, DT_E_ARG, "dt_conv_gemm: bad wide GroupNorm problem");
  const size_t need = (size_t)g.B * g.T * g.N * sizeof(float);
  if (need > ctx->wide_bytes)   // sized by dt_load_denoiser for max_batch; callers chunk larger batches
    return dt_fail(ctx, DT_E_ARG, "dt_conv_gemm: wide GroupNorm scratch too small (batch beyond the loaded max_batch)");
  ConvGemm plain = g;
  plain.epi = EPI_PLAIN;
  plain.gamma = plain.beta = nullptr;
  plain.film = plain.film_t = nullptr;
  plain.resid = nullptr;
  plain.relu = 0;
  plain.out_bf16 = nullptr;
  plain.out_f32 = (float*)ctx->d_wide;
  plain.ldc = g.N;
  plain.out_b_stride = g.T;
  plain.out_t_stride = 1;
  plain.out_off = 0;
  int rc = dt_conv_gemm(ctx, plain, st);
  if (rc) return rc;
  WideGn p;
  p.y = (const float*)ctx->d_wide; p.T = g.T; p.N = g.N; p.gw = g.group_width;
  p.gamma = g.gamma; p.beta = g.beta; p.film = g.film; p.film_ld = g.film_ld; p.film_t = g.film_t;
  p.resid = g.resid; p.ld_res = g.ld_res; p.out = g.out_bf16; p.ldc = g.ldc;
  p.out_b_stride = g.out_b_stride; p.out_t_stride = g.out_t_stride; p.out_off = g.out_off;
  const long long blocks = g.B * (g.N / g.group_width);
  if (blocks > 0x7fffffffLL) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_conv_gemm: batch too large for the wide GroupNorm path");
  k_gn_mish_wide<<<(unsigned)blocks, 256, 0, st>>>(p);
  DT_LAUNCH_CHECK("k_gn_mish_wide");
  return DT_OK;
}

// ---------------------------------------------------------------------------------------------
// Split-K for small problems (at most 128 rows: the reference's own B = 1 planning loop).  With one M tile a
// layer keeps only N / BN = 2..8 CTAs busy and each streams its whole weight slab alone (35-50 us per layer
// although the weights could cross HBM in a few us).  The tile scheduler therefore also splits K: every
// (N tile, K slice) pair is a work item of the same tcgen05 kernel (plain epilogue, fp32 partial sums of slice s
// written slice_rows rows further down a scratch buffer), and k_splitk_epi sums the slices, adds the bias
// and applies the epilogue the fused kernel would have applied: GroupNorm -> Mish -> FiLM (+ residual) per
// (sample, group), or bias (+ residual, ReLU), with the same output addressing.
// ---------------------------------------------------------------------------------------------
struct SplitKEpi {
  const float* part;  // [S][rows][N]
  int S, rows, T, N, gw, epi, relu;
  const float* bias;
  const float* gamma;
  const float* beta;
  const float* film;
  long long film_ld;
  const float* film_t;
  const __nv_bfloat16* resid;
  long long ld_res;
  __nv_bfloat16* out_bf16;
  float* out_f32;
  long long ldc, out_b_stride, out_t_stride, out_off;
};

// one block per (sample, window of `gw` columns): a GroupNorm group (EPI_GN_MISH) or a slab of columns
#define SPLITK_EPI_THREADS 256

// A cluster of CS CTAs per (sample, window of `gw` columns) -- a GroupNorm group (EPI_GN_MISH) or a slab of
// columns -- each CTA owning T / CS rows: the slice sums are pulled by CS times more SMs than one CTA per
// window could use (the kernel is bound by each SM's L2 read rate), the group statistics are exchanged
// through distributed shared memory.
__global__ void __launch_bounds__(SPLITK_EPI_THREADS)
k_splitk_epi(SplitKEpi p) {
  namespace cgx = cooperative_groups;
  cgx::cluster_group cluster = cgx::this_cluster();
  extern __shared__ __align__(16) float s_y[];  // [T / CS][gw]
  __shared__ float s_a[SPLITK_EPI_THREADS], s_b[SPLITK_EPI_THREADS];
  __shared__ float s_part[2];
  dt_pdl_launch();
  dt_pdl_wait();  // the partial sums are the previous kernel's output
  const int CS = (int)cluster.num_blocks(), r = (int)cluster.block_rank();
  const int win = blockIdx.x / CS;
  const int windows = p.N / p.gw;
  const int b = win / windows, n0 = (win % windows) * p.gw;
  const int t_rows = p.T / CS, t_first = r * t_rows;
  const int cnt = t_rows * p.gw;
  float s = 0.f, ss = 0.f;
  const long long kstride = (long long)p.rows * p.N;
  // four elements per thread at a time, slices in ascending order per element (deterministic); the loop over
  // slices is unrolled so 16 independent loads are in flight per thread
  for (int i0 = threadIdx.x; i0 < cnt; i0 += 4 * SPLITK_EPI_THREADS) {
    const float* q[4];
    float v[4];
    bool on[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const int i = i0 + e * SPLITK_EPI_THREADS;
      on[e] = i < cnt;
      const int ii = on[e] ? i : i0;
      const int t = t_first + ii / p.gw, c = ii % p.gw, n = n0 + c;
      q[e] = p.part + (long long)(b * p.T + t) * p.N + n;
      v[e] = p.bias ? p.bias[n] : 0.f;
    }
#pragma unroll 4
    for (int k = 0; k < p.S; ++k) {
#pragma unroll
      for (int e = 0; e < 4; ++e) v[e] += __ldg(q[e] + k * kstride);
    }
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      if (on[e]) {
        s_y[i0 + e * SPLITK_EPI_THREADS] = v[e];
        s += v[e];
        ss += v[e] * v[e];
      }
    }
  }
  float mean = 0.f, rstd = 1.f;
  if (p.epi != EPI_PLAIN) {
    s_a[threadIdx.x] = s;
    s_b[threadIdx.x] = ss;
    __syncthreads();
    for (int o = SPLITK_EPI_THREADS / 2; o > 0; o >>= 1) {
      if (threadIdx.x < o) {
        s_a[threadIdx.x] += s_a[threadIdx.x + o];
        s_b[threadIdx.x] += s_b[threadIdx.x + o];
      }
      __syncthreads();
    }
    if (threadIdx.x == 0) {
      s_part[0] = s_a[0];
      s_part[1] = s_b[0];
    }
    cluster.sync();  // every CTA's partial statistics are published
    float ts = 0.f, tq = 0.f;
    for (int k = 0; k < CS; ++k) {  // same order in every CTA: identical totals
      const float* remote = cluster.map_shared_rank(s_part, k);
      ts += remote[0];
      tq += remote[1];
    }
    cluster.sync();  // nobody leaves (and frees its shared memory) while a peer may still read it
    const float total = (float)(p.T * p.gw);
    mean = ts / total;
    rstd = rsqrtf(fmaxf(tq / total - mean * mean, 0.f) + 1e-5f);
  }
  for (int i = threadIdx.x; i < cnt; i += SPLITK_EPI_THREADS) {  // a thread revisits the elements it wrote itself
    const int t = t_first + i / p.gw, c = i % p.gw, n = n0 + c;
    float v = s_y[i];
    if (p.epi == EPI_GN_RELU) v = (v - mean) * rstd * p.gamma[n] + p.beta[n];   // residual, then ReLU, below
    if (p.epi == EPI_GN_MISH) {
      v = mish_f((v - mean) * rstd * p.gamma[n] + p.beta[n]);
      if (p.film) {
        const float sc = p.film[(long long)b * p.film_ld + n] + (p.film_t ? p.film_t[n] : 0.f);
        const float sh = p.film[(long long)b * p.film_ld + p.N + n] + (p.film_t ? p.film_t[p.N + n] : 0.f);
        v = v * sc + sh;
      }
    }
    const long long row = (long long)b * p.out_b_stride + (long long)t * p.out_t_stride + p.out_off;
    if (p.resid) v += __bfloat162float(p.resid[row * p.ld_res + n]);
    if (p.epi != EPI_GN_MISH && p.relu) v = fmaxf(v, 0.f);
    if (p.out_bf16) p.out_bf16[row * p.ldc + n] = __float2bfloat16(v);
    if (p.out_f32) p.out_f32[row * p.ldc + n] = v;
  }
}

#define SPLITK_SCRATCH_BYTES (48u << 20)

int dt_conv_gemm(dt_ctx* ctx, const ConvGemm& g, cudaStream_t st);

// -> 0 = not applicable (caller continues with the fused path), 1 = done, negative = error
static int conv_gemm_splitk(dt_ctx* ctx, const ConvGemm& g, cudaStream_t st) {
  if (!ctx->splitk_on || g.ksplit > 1) return 0;
  // two shapes qualify: (a) whole samples inside ONE 128-row tile (the U-Net at a few candidates), and
  // (b) a flat [rows, N] problem -- one "sample", plain epilogue -- with a handful of 128-row tiles and a
  // long K (the encoder's last stages at planner batch sizes: 256 x 512 x 4608 is four tiles on 148 SMs)
  // (the encoder's GroupNorm+ReLU layers: any T, the samples packed into one tile)
  const bool few_rows = g.epi == EPI_GN_RELU ? (g.T <= 128 && g.B * g.T <= 128 && g.B <= 128 / g.T)
                                             : (g.T <= 64 && g.B * g.T <= 128 && 128 % g.T == 0);
  const bool flat = !few_rows && g.epi == EPI_PLAIN && g.B == 1 && g.out_t_stride == 1 && g.T <= 8192;
  if (!few_rows && !flat) return 0;
  long long nkb = 0;
  for (int s = 0; s < g.nseg; ++s) nkb += g.seg[s].nblk;
  const int rows = (int)(g.B * g.T);
  const int bn = (g.N % 256 == 0) ? 256 : ((g.N % 128 == 0) ? 128 : 64);
  const int n_tiles = (g.N / bn) * (flat ? (rows + BM - 1) / BM : 1);
  if (nkb < (flat ? 64 : 12) || n_tiles * 4 > ctx->sm_count) return 0;  // flat: measured win only from K = 4096 up (35 -> 27 us)
  int ksplit = ctx->sm_count / n_tiles;              // one work item per SM
  if (flat) {                                        // several tiles: the reduction reads ksplit x the output,
    if (ksplit > nkb / 4) ksplit = (int)(nkb / 4);   // keep the slices long and few
    if (ksplit > 16) ksplit = 16;
  } else if (ksplit > nkb / 2) {
    ksplit = (int)(nkb / 2);                         // at least two K blocks per item
  }
  if (ksplit < 2) return 0;
  const int per = (int)((nkb + ksplit - 1) / ksplit);
  ksplit = (int)((nkb + per - 1) / per);             // no empty slice
  if (g.epi != EPI_PLAIN && (g.group_width < 8 || g.N % g.group_width != 0 || g.group_width * g.T > 16384)) return 0;
  if ((size_t)ksplit * rows * g.N * sizeof(float) > SPLITK_SCRATCH_BYTES) return 0;
  void*& scratch = g.scratch ? ctx->d_splitk2 : ctx->d_splitk;
  if (!scratch) {  // fixed size, allocated once: the pointer is baked into captured graphs
    DT_CUDA(cudaMalloc(&scratch, SPLITK_SCRATCH_BYTES));
    DT_CUDA(cudaFuncSetAttribute(k_splitk_epi, cudaFuncAttributeMaxDynamicSharedMemorySize, 16384 * 4));
  }
  ConvGemm part = g;
  part.epi = EPI_PLAIN;
  part.bias = nullptr; part.gamma = part.beta = nullptr; part.film = part.film_t = nullptr;
  part.resid = nullptr; part.relu = 0;
  part.out_bf16 = nullptr;
  part.out_f32 = (float*)scratch;
  part.ldc = g.N; part.out_b_stride = g.T; part.out_t_stride = 1; part.out_off = 0;
  part.ksplit = ksplit; part.kb_per_slice = per; part.slice_rows = rows;
  int rc = dt_conv_gemm(ctx, part, st);
  if (rc) return rc;
  ctx->pdl_now = ctx->pdl_on;  // the reduction always follows its own GEMM in the stream
  SplitKEpi e;
  e.part = (const float*)scratch; e.S = ksplit; e.rows = rows; e.T = g.T; e.N = g.N; e.epi = g.epi; e.relu = g.relu;
  e.gw = (g.epi != EPI_PLAIN) ? g.group_width : 64;
  e.bias = g.bias; e.gamma = g.gamma; e.beta = g.beta; e.film = g.film; e.film_ld = g.film_ld; e.film_t = g.film_t;
  e.resid = g.resid; e.ld_res = g.ld_res; e.out_bf16 = g.out_bf16; e.out_f32 = g.out_f32;
  e.ldc = g.ldc; e.out_b_stride = g.out_b_stride; e.out_t_stride = g.out_t_stride; e.out_off = g.out_off;
  int64_t windows_b = g.B;
  int cs = 8;  // cluster size: up to 8 CTAs per window, T / CS whole rows each
  if (flat) {
    // rows are independent (no GroupNorm): re-cut the one sample into windows of up to 16 consecutive rows,
    // one CTA each (1024 elements: four per thread)
    int t2 = 16;
    while (rows % t2 != 0) t2 >>= 1;
    e.T = t2;
    e.out_b_stride = t2;  // out_t_stride == 1: row = b2 * t2 + t
    windows_b = rows / t2;
    cs = 1;
  }
  while (cs > 1 && e.T % cs != 0) cs >>= 1;
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof cfg);
  cfg.gridDim = dim3((unsigned)(windows_b * (g.N / e.gw) * cs));
  cfg.blockDim = dim3(SPLITK_EPI_THREADS);
  cfg.dynamicSmemBytes = (size_t)(e.T / cs) * e.gw * sizeof(float);
  cfg.stream = st;
  cudaLaunchAttribute attr[2];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = (unsigned)cs;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  attr[1].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[1].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = ctx->pdl_now ? 2 : 1;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->prof_on) {  // the reduction is part of the layer's cost: recorded as epi = 2 (no flops of its own)
    e0 = dt_prof_event(ctx);
    e1 = dt_prof_event(ctx);
    if (e0 && e1) {
      dt_ctx::ProfRec rec{0, 2, e.gw * 10 + cs, (long long)rows, g.N, nkb * BK, 0.f, ksplit};
      ctx->prof_recs.push_back(rec);
      cudaEventRecord(e0, st);
    }
  }
  DT_CUDA(cudaLaunchKernelEx(&cfg, k_splitk_epi, e));
  if (e0 && e1) cudaEventRecord(e1, st);
  DT_LAUNCH_CHECK("k_splitk_epi");
  return 1;
}

int dt_conv_gemm(dt_ctx* ctx, const ConvGemm& g, cudaStream_t st) {
  if (g.B <= 0) return DT_OK;
  ctx->pdl_now = ctx->pdl_on && g.pdl;
  if (g.nseg < 1 || g.nseg > GEMM_MAX_SEG || !g.w || !g.a[0].ptr || g.N % 64 != 0)
    return dt_fail(ctx, DT_E_ARG, "dt_conv_gemm: bad problem description");
  {
    const int sk = conv_gemm_splitk(ctx, g, st);
    if (sk < 0) return sk;
    if (sk == 1) return DT_OK;
  }
  if (g.epi == EPI_GN_MISH && g.group_width > 256) return conv_gemm_wide_gn(ctx, g, st);
  // tile geometry: a tile holds whole samples (T <= 128) or a 128-row slice of one sample
  int T = g.T, rows_t, nb, tps;
  if (g.epi == EPI_GN_RELU || g.a[0].W2 > 0) {
    // whole samples per tile for ANY T <= 128: floor(128 / T) samples, the rest of the tile stays empty (the
    // encoder's fused layers, and their split-K partial sums with the plain epilogue: a 2-D source's TMA box is
    // exactly one sample's output map)
    if (T < 1 || T > BM) return dt_fail(ctx, DT_E_UNSUPPORTED, "2-D conv source: at most 128 output pixels per sample");
    if (g.epi == EPI_GN_RELU && (g.group_width != 16 || !g.gamma || !g.beta || !g.out_bf16 || g.out_f32))
      return dt_fail(ctx, DT_E_UNSUPPORTED, "GroupNorm+ReLU epilogue: groups of 16 channels, bf16 output");
    if (g.epi == EPI_GN_MISH) return dt_fail(ctx, DT_E_UNSUPPORTED, "2-D conv source with the GroupNorm+Mish epilogue");
    rows_t = T;
    nb = BM / T;
    tps = 0;
  } else if (T >= BM) {
    if (T % BM != 0 && g.epi == EPI_GN_MISH) return dt_fail(ctx, DT_E_UNSUPPORTED, "GroupNorm epilogue needs T <= 128");
    rows_t = BM;
    nb = 1;
    tps = (T + BM - 1) / BM;
  } else if (BM % T != 0) {
    // odd row count: one (partially filled) tile per sample, rows >= T are masked
    if (g.epi == EPI_GN_MISH) return dt_fail(ctx, DT_E_UNSUPPORTED, "GroupNorm epilogue needs T to divide 128");
    rows_t = BM;
    nb = 1;
    tps = 1;
  } else {
    rows_t = T;
    nb = BM / T;
    tps = 0;
  }
  if (g.epi == EPI_GN_MISH && T > BM) return dt_fail(ctx, DT_E_UNSUPPORTED, "GroupNorm epilogue needs T <= 128");
  int bn = 256;
  if (g.epi == EPI_GN_RELU) {
    bn = (g.N % 128 == 0) ? 128 : 64;
  } else if (g.epi == EPI_GN_MISH) {
    const int gw = g.group_width;
    if (gw == 8) bn = 64;
    else if (gw == 16) bn = 128;
    else if (gw == 32 || gw == 64 || gw == 128 || gw == 256) bn = 256;
    else return dt_fail(ctx, DT_E_UNSUPPORTED, "GroupNorm group width must be 8..256 (power of two)");
    if (g.N % bn != 0) return dt_fail(ctx, DT_E_UNSUPPORTED, "Cout must be a multiple of the N tile");
    if (!g.gamma || !g.beta) return dt_fail(ctx, DT_E_ARG, "GroupNorm epilogue needs gamma/beta");
  } else {
    bn = (g.N % 256 == 0) ? 256 : ((g.N % 128 == 0) ? 128 : 64);
  }
  GemmDev d;
  memset(&d, 0, sizeof d);
  int64_t ktot = 0;
  for (int s = 0; s < g.nseg; ++s) {
    d.seg[s] = g.seg[s];
    ktot += (int64_t)g.seg[s].nblk * BK;
    if (g.seg[s].src < 0 || g.seg[s].src >= g.n_src) return dt_fail(ctx, DT_E_ARG, "dt_conv_gemm: bad segment source");
  }
  d.nseg = g.nseg;
  d.nkb_total = (int)(ktot / BK);
  d.T = T;
  d.rows_t = rows_t;
  d.nb = nb;
  d.tiles_per_sample = tps;
  d.num_m_tiles = tps > 0 ? (int)(g.B * tps) : (int)((g.B + nb - 1) / nb);
  d.num_n_tiles = g.N / bn;
  d.B = g.B;
  d.N = g.N;
  d.bias = g.bias; d.gamma = g.gamma; d.beta = g.beta;
  d.film = g.film; d.film_ld = g.film_ld; d.film_t = g.film_t;
  d.resid = g.resid; d.ld_res = g.ld_res; d.relu = g.relu;
  d.out_bf16 = g.out_bf16; d.out_f32 = g.out_f32;
  d.ldc = g.ldc; d.out_b_stride = g.out_b_stride; d.out_t_stride = g.out_t_stride; d.out_off = g.out_off;
  d.ksplit = g.ksplit > 1 ? g.ksplit : 1;
  d.kb_per_slice = g.ksplit > 1 ? g.kb_per_slice : d.nkb_total;
  d.slice_rows = g.slice_rows;
  d.a_tile_bytes = rows_t * nb * BK * 2;
#ifdef GEMM_TIMING
  {
    static int want_n = -2, want_k = 0, want_res = -1;
    if (want_n == -2) {
      const char* e = getenv("DITREE_GEMM_DBG");   // "N,K[,resid 0/1]"
      want_n = -1;
      if (e) sscanf(e, "%d,%d,%d", &want_n, &want_k, &want_res);
    }
    d.dbg = (g.N == want_n && ktot == want_k && g.epi == EPI_GN_MISH && (want_res < 0 || (g.resid != nullptr) == (want_res != 0)) &&
             d.num_m_tiles > 64) ? 1 : 0;
  }
#endif
  if (!d.out_bf16 && !d.out_f32) return dt_fail(ctx, DT_E_ARG, "dt_conv_gemm: no output");
  // the epilogue moves 32-byte pieces of bf16 rows with 256-bit accesses
  if ((d.out_bf16 && ((reinterpret_cast<uintptr_t>(d.out_bf16) & 31) || d.ldc % 16 != 0)) ||
      (d.resid && ((reinterpret_cast<uintptr_t>(d.resid) & 31) || d.ld_res % 16 != 0)))
    return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_conv_gemm: bf16 output / residual rows must be 32-byte aligned (pitch % 16 == 0)");

  CUtensorMap mA0, mA1, mW;
  int rc = make_act_map(ctx, &mA0, g.a[0], g.B, rows_t, nb);
  if (rc) return rc;
  if (g.n_src > 1) {
    rc = make_act_map(ctx, &mA1, g.a[1], g.B, rows_t, nb);
    if (rc) return rc;
  } else {
    mA1 = mA0;
  }
  // CTA pairs need at least two 128-row tiles; tiny problems stay on one CTA
  const int cg = (d.num_m_tiles >= 2) ? gemm_cta_group() : 1;
  rc = make_w_map(ctx, &mW, g.w, g.N, g.w_ktot > 0 ? g.w_ktot : ktot, bn / cg);
  if (rc) return rc;

  if (g.epi == EPI_GN_RELU) {
    if (bn == 128) return launch_gemm<128, EPI_GN_RELU, 16>(ctx, cg, mA0, mA1, mW, d, st);
    return launch_gemm<64, EPI_GN_RELU, 16>(ctx, cg, mA0, mA1, mW, d, st);
  }
  if (g.epi == EPI_PLAIN) {
    if (bn == 256) return launch_gemm<256, EPI_PLAIN, 256>(ctx, cg, mA0, mA1, mW, d, st);
    if (bn == 128) return launch_gemm<128, EPI_PLAIN, 128>(ctx, cg, mA0, mA1, mW, d, st);
    return launch_gemm<64, EPI_PLAIN, 64>(ctx, cg, mA0, mA1, mW, d, st);
  }
  switch (g.group_width) {
    case 8: return launch_gemm<64, EPI_GN_MISH, 8>(ctx, cg, mA0, mA1, mW, d, st);
    case 16: return launch_gemm<128, EPI_GN_MISH, 16>(ctx, cg, mA0, mA1, mW, d, st);
    case 32: return launch_gemm<256, EPI_GN_MISH, 32>(ctx, cg, mA0, mA1, mW, d, st);
    case 64: return launch_gemm<256, EPI_GN_MISH, 64>(ctx, cg, mA0, mA1, mW, d, st);
    case 128: return launch_gemm<256, EPI_GN_MISH, 128>(ctx, cg, mA0, mA1, mW, d, st);
    case 256: return launch_gemm<256, EPI_GN_MISH, 256>(ctx, cg, mA0, mA1, mW, d, st);
  }
  return dt_fail(ctx, DT_E_UNSUPPORTED, "unsupported GroupNorm group width");
}

// Test hook: C[M,N] f32 = A[M,K] bf16 row-major * W[N,K]^T bf16 (K, N multiples of 64).
extern "C" int dt_gemm_bf16(dt_ctx* ctx, const void* A, const void* W, int64_t M, int N, int K, float* C,
                            void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!A || !W || !C || M <= 0 || N % 64 != 0 || K % 64 != 0) return dt_fail(ctx, DT_E_ARG, "dt_gemm_bf16: bad argument");
  ConvGemm g;
  g.a[0].ptr = (const __nv_bfloat16*)A;
  g.a[0].C = K;
  g.a[0].T_in = (int)((M + BM - 1) / BM) * BM >= M ? (int)M : (int)M;
  g.a[0].P = 1;
  g.n_src = 1;
  g.w = (const __nv_bfloat16*)W;
  g.N = N;
  g.nseg = 1;
  g.seg[0] = GemmSeg{0, 0, 0, K / BK};
  g.B = 1;
  g.T = (int)M;   // one "sample" of M rows, tiled in 128-row slices
  g.epi = EPI_PLAIN;
  g.out_f32 = C;
  g.ldc = N;
  g.out_b_stride = 0;
  g.out_t_stride = 1;
  return dt_conv_gemm(ctx, g, (cudaStream_t)stream);
}

// Test hook for the fused 2-D conv + GroupNorm(16-channel groups) (+ residual) (+ ReLU) launch of the ResNet encoder:
// in (B, H, W, Cin) bf16 channel-last, Cin % 64 == 0; w [N][k * k * Cin] bf16 (tap-major, then channel); gamma / beta
// [N] f32; resid (B, OH, OW, N) bf16 or NULL; out (B, OH, OW, N) bf16.  OH * OW <= 128, k * k <= 9.
extern "C" int dt_conv2d_gn_bf16(dt_ctx* ctx, const void* in, int64_t B, int H, int W, int Cin, const void* w, int N, int k,
                                 int stride, int pad, const float* gamma, const float* beta, const void* resid, int relu,
                                 void* out, void* stream) {
  if (!ctx) return DT_E_ARG;
  if (!in || !w || !gamma || !beta || !out || B <= 0 || Cin % 64 != 0 || N % 64 != 0 || k < 1 || k * k > GEMM_MAX_SEG ||
      stride < 1)
    return dt_fail(ctx, DT_E_ARG, "dt_conv2d_gn_bf16: bad argument");
  const int OH = (H + 2 * pad - k) / stride + 1, OW = (W + 2 * pad - k) / stride + 1;
  if (OH < 1 || OW < 1 || OH * OW > BM) return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_conv2d_gn_bf16: at most 128 output pixels");
  ConvGemm g;
  ActSrc a{(const __nv_bfloat16*)in, Cin, 0, 1};
  a.W2 = W; a.H2 = H; a.stride2 = stride; a.ow2 = OW; a.oh2 = OH;
  g.a[0] = a;
  g.n_src = 1;
  g.w = (const __nv_bfloat16*)w;
  g.N = N;
  g.nseg = 0;
  for (int ky = 0; ky < k; ++ky)
    for (int kx = 0; kx < k; ++kx) g.seg[g.nseg++] = GemmSeg{0, kx - pad, ky - pad, Cin / 64, 0};
  g.B = B;
  g.T = OH * OW;
  g.epi = EPI_GN_RELU;
  g.gamma = gamma;
  g.beta = beta;
  g.group_width = 16;
  g.resid = (const __nv_bfloat16*)resid;
  g.ld_res = N;
  g.relu = relu;
  g.out_bf16 = (__nv_bfloat16*)out;
  g.ldc = N;
  g.out_b_stride = OH * OW;
  g.out_t_stride = 1;
  return dt_conv_gemm(ctx, g, (cudaStream_t)stream);
}
