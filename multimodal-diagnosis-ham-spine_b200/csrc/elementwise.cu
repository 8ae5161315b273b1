// Small elementwise kernels used between the fused ops (activation/dropout backward, products, gates).
#include "common.cuh"
#include "../../include/mdhs_b200.h"

MDHS_DEFINE_SEED_TICK(elementwise)

extern int64_t g_mdhs_launches;

namespace {

int grid_for(int64_t items) {
  int64_t g = (items + 255) / 256;
  const int64_t cap = 148 * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// g = dy * dropout_mask(seed, idx) * act'(aux)   (bf16, 8 per thread).  aux = pre-activation (GELU) or output (ReLU).
__global__ void __launch_bounds__(256) act_dropout_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ aux,
                                                              bf16* __restrict__ g, int64_t n, int act, float drop_p,
                                                              uint64_t seed) {
  const int64_t nv = n >> 3;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float d[8], a[8];
    load8(dy + i * 8, d);
    if (act != MDHS_ACT_NONE) load8(aux + i * 8, a);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float v = d[k];
      if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t)(i * 8 + k), drop_p, inv_keep);
      if (act == MDHS_ACT_RELU) v = a[k] > 0.f ? v : 0.f;
      else if (act == MDHS_ACT_GELU) v *= gelu_erf_grad(a[k]);
      d[k] = v;
    }
    store8(g + i * 8, d);
  }
}

__global__ void relu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// c = a * b (fp32); also used for the backward (da = dc * b, db = dc * a)
__global__ void mul_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c[i] = a[i] * b[i];
}

// y = x * dropout_mask(seed, i)  (fp32; the same call implements the backward)
__global__ void dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p, uint64_t seed) {
  const float inv_keep = 1.f / (1.f - p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] * dropout_scale(seed, (uint64_t)i, p, inv_keep);
}

}  // namespace

extern "C" int mdhs_act_dropout_bwd(const void* dy, const void* aux, void* g, int64_t n, int act, float drop_p, uint64_t seed,
                                    void* stream) {
  if (!dy || !g || n <= 0 || (n % 8) || (act != MDHS_ACT_NONE && !aux)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  act_dropout_bwd_kernel<<<grid_for(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const bf16*)dy, (const bf16*)aux,
                                                                                              (bf16*)g, n, act, drop_p, seed);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_relu_bwd_f32(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  if (!dy || !y || !dx || n <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  relu_bwd_f32_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, y, dx, n);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_mul_f32(const float* a, const float* b, float* c, int64_t n, void* stream) {
  if (!a || !b || !c || n <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  mul_f32_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, c, n);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, void* stream) {
  if (!x || !y || n <= 0 || p < 0.f || p >= 1.f) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  dropout_f32_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, p, seed);
  MDHS_RETURN_LAST();
}
