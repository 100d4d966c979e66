// denoiser.cu -- placeholder, replaced by the bf16 tcgen05 denoiser.
#include "common.cuh"
void dt_denoiser_free(dt_ctx* ctx) {}
extern "C" int dt_load_denoiser(dt_ctx* ctx, const dt_tensor_desc* t, int n, const dt_model_cfg* cfg, void* stream) {
  return dt_fail(ctx, DT_E_UNSUPPORTED, "denoiser not built yet");
}
extern "C" int dt_fm_sample(dt_ctx* ctx, const float* noise, const float* cond, const void* local_map, int64_t B,
                            int K, double exp_scale, const double* norm_host, float* actions_out, void* stream) {
  return dt_fail(ctx, DT_E_NOMODEL, "denoiser not loaded");
}
extern "C" int dt_encode_map(dt_ctx* ctx, const void* local_map, int64_t B, float* emb_out, void* stream) {
  return dt_fail(ctx, DT_E_NOMODEL, "denoiser not loaded");
}
extern "C" int dt_unet_forward(dt_ctx* ctx, const float* sample, const float* emb, const float* cond, int64_t B,
                               float timestep, float* vel_out, void* stream) {
  return dt_fail(ctx, DT_E_NOMODEL, "denoiser not loaded");
}
