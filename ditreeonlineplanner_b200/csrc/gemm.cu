// gemm.cu -- placeholder, replaced by the tcgen05 GEMM core.
#include "common.cuh"
extern "C" int dt_gemm_bf16(dt_ctx* ctx, const void* A, const void* W, int64_t M, int N, int K, float* C,
                            void* stream) {
  return dt_fail(ctx, DT_E_UNSUPPORTED, "dt_gemm_bf16: not built yet");
}
