// mufu_bound.cu -- exhaustive error bound of dt_sincos_fast (carfast.cuh): every fp32 heading with
// |theta| <= DT_SC_MAX is compared with the float64 sin / cos.  The bound DT_SC_ERR used by the collision
// guard band must exceed the printed maxima.
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o /tmp/mufu_bound tools/mufu_bound.cu && /tmp/mufu_bound
#include <cstdio>
#include <cstdint>
#include "../ditreeonlineplanner_b200/csrc/carfast.cuh"

__global__ void k_scan(uint32_t n_bits, double* max_err, float* arg_max) {
  double worst = 0.0;
  float where = 0.f;
  for (uint64_t i = blockIdx.x * (uint64_t)blockDim.x + threadIdx.x; i <= n_bits; i += (uint64_t)gridDim.x * blockDim.x) {
    for (int sgn = 0; sgn < 2; ++sgn) {
      const float th = __uint_as_float((uint32_t)i | (sgn ? 0x80000000u : 0u));
      float sn, cs;
      dt_sincos_fast(th, sn, cs);
      double ds, dc;
      sincos((double)th, &ds, &dc);
      const double e = fmax(fabs((double)sn - ds), fabs((double)cs - dc));
      if (e > worst) { worst = e; where = th; }
    }
  }
  // block reduction through shared memory
  __shared__ double s_e[256];
  __shared__ float s_w[256];
  s_e[threadIdx.x] = worst; s_w[threadIdx.x] = where;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if (threadIdx.x < o && s_e[threadIdx.x + o] > s_e[threadIdx.x]) { s_e[threadIdx.x] = s_e[threadIdx.x + o]; s_w[threadIdx.x] = s_w[threadIdx.x + o]; }
    __syncthreads();
  }
  if (threadIdx.x == 0) { max_err[blockIdx.x] = s_e[0]; arg_max[blockIdx.x] = s_w[0]; }
}

int main() {
  const float lim = DT_SC_MAX;
  uint32_t bits;
  memcpy(&bits, &lim, 4);
  const int blocks = 148 * 8;
  double* d_e; float* d_w;
  cudaMalloc(&d_e, blocks * sizeof(double)); cudaMalloc(&d_w, blocks * sizeof(float));
  k_scan<<<blocks, 256>>>(bits, d_e, d_w);
  std::vector<double> e(blocks); std::vector<float> w(blocks);
  if (cudaMemcpy(e.data(), d_e, blocks * sizeof(double), cudaMemcpyDeviceToHost) != cudaSuccess) { printf("cuda error\n"); return 1; }
  cudaMemcpy(w.data(), d_w, blocks * sizeof(float), cudaMemcpyDeviceToHost);
  double worst = 0; float where = 0;
  for (int i = 0; i < blocks; ++i) if (e[i] > worst) { worst = e[i]; where = w[i]; }
  printf("dt_sincos_fast: max |error| over all fp32 |theta| <= %g : %.4e at theta = %.9g  (DT_SC_ERR = %.3e) %s\n", lim, worst, where,
         (double)DT_SC_ERR, worst < (double)DT_SC_ERR ? "OK" : "VIOLATED");
  return worst < (double)DT_SC_ERR ? 0 : 2;
}
