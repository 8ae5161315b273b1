// Shared device/host helpers for the sm_100a kernels of the multimodal hot path.
#pragma once
#include <cuda_runtime.h>
#include <cuda_bf16.h>
#include <stdint.h>

#define MDHS_OK 0
#define MDHS_ERR_ARG 1001      // bad argument (shape / alignment / null pointer)
#define MDHS_ERR_DRIVER 1002   // driver entry point (tensor map encode) unavailable

// Returns the cudaError_t of the last launch as the C-ABI status.
#define MDHS_RETURN_LAST()                          \
  do {                                              \
    cudaError_t e__ = cudaGetLastError();           \
    return e__ == cudaSuccess ? MDHS_OK : (int)e__; \
  } while (0)

typedef __nv_bfloat16 bf16;
typedef __nv_bfloat162 bf162;

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}

// Block-wide sum for blockDim.x <= 1024; `red` is >= 32 floats of shared memory.
__device__ __forceinline__ float block_sum(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : 0.f;
  r = warp_sum(r);
  return r;
}
__device__ __forceinline__ float block_max(float v, float* red) {
  int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float r = (lane < nw) ? red[lane] : -INFINITY;
  r = warp_max(r);
  return r;
}

// erf-GELU through the complementary error function Q(|x|) = erfc(|x|/sqrt(2)) = poly5(t) * exp(-x^2/2),
// t = 1/(1 + p|x|/sqrt(2))  (Abramowitz-Stegun 7.1.26, |error| <= 1.5e-7): one ex2 + one rcp + 7 FMAs, branch-free, and
// exact in the tails because Phi(x) = Q/2 for x < 0 is formed without cancellation.  The same exponential serves the
// density term of the derivative.  (erff + expf cost ~4x as many issue slots; the GEMM epilogues are issue-bound.)
__device__ __forceinline__ float rcp_approx(float x) {
  float y;
  asm("rcp.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float ex2_approx(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ float gelu_q_(float ax, float& e) {
  const float t = rcp_approx(fmaf(0.3275911f * 0.70710678118654752440f, ax, 1.f));
  e = ex2_approx(-0.72134752044448170368f * ax * ax);   // exp(-x^2/2) = 2^(-x^2 * log2(e) / 2)
  float q = fmaf(1.061405429f, t, -1.453152027f);
  q = fmaf(q, t, 1.421413741f);
  q = fmaf(q, t, -0.284496736f);
  q = fmaf(q, t, 0.254829592f);
  return q * t * e;
}
__device__ __forceinline__ float gelu_erf(float x) {
  float e;
  const float hq = 0.5f * gelu_q_(fabsf(x), e);
  return x * (x >= 0.f ? 1.f - hq : hq);
}
// GELU and its derivative from one exponential (forward epilogue that also saves the derivative for the backward pass)
__device__ __forceinline__ void gelu_erf_both(float x, float& y, float& dy) {
  float e;
  const float hq = 0.5f * gelu_q_(fabsf(x), e);
  const float cdf = x >= 0.f ? 1.f - hq : hq;
  y = x * cdf;
  dy = fmaf(x * 0.39894228040143267794f, e, cdf);
}
__device__ __forceinline__ float gelu_erf_grad(float x) {
  float e;
  const float hq = 0.5f * gelu_q_(fabsf(x), e);
  const float cdf = x >= 0.f ? 1.f - hq : hq;
  return fmaf(x * 0.39894228040143267794f, e, cdf);
}

// 8 x bf16 <-> 8 x float through one 128-bit access.
struct alignas(16) bf16x8 {
  bf162 v[4];
};
__device__ __forceinline__ void load8(const bf16* p, float* f) {
  bf16x8 t = *reinterpret_cast<const bf16x8*>(p);
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 x = __bfloat1622float2(t.v[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}
// split form: issue the 128-bit load now (4 registers), convert when the values are consumed -- keeps many independent loads
// in flight without holding 8 fp32 registers per load
__device__ __forceinline__ bf16x8 ldraw8(const bf16* p) { return *reinterpret_cast<const bf16x8*>(p); }
__device__ __forceinline__ void cvt8(const bf16x8& t, float* f) {
#pragma unroll
  for (int i = 0; i < 4; i++) {
    float2 x = __bfloat1622float2(t.v[i]);
    f[2 * i] = x.x;
    f[2 * i + 1] = x.y;
  }
}
__device__ __forceinline__ void store8(bf16* p, const float* f) {
  bf16x8 t;
#pragma unroll
  for (int i = 0; i < 4; i++) t.v[i] = __floats2bfloat162_rn(f[2 * i], f[2 * i + 1]);
  *reinterpret_cast<bf16x8*>(p) = t;
}

// Counter-based RNG for fused dropout: one 32-bit hash per (seed, element index).
__device__ __forceinline__ uint32_t hash_u32(uint64_t seed, uint64_t idx) {
  uint64_t z = seed + idx * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  z = z ^ (z >> 31);
  return (uint32_t)(z >> 32);
}
// Per-translation-unit copy of the training-step counter folded into every dropout seed, so that a CUDA
// graph replay (which bakes the seed arguments) still draws fresh masks each step.  mdhs_seed_tick()
// advances all copies by the same amount once per step; forward and backward of one step agree.
static __device__ uint64_t g_seed_tick = 0;
#define MDHS_DEFINE_SEED_TICK(tu)                                                       \
  static __global__ void seed_tick_kernel_##tu(uint64_t inc) { g_seed_tick += inc; }    \
  void mdhs_seed_tick_##tu(uint64_t inc, cudaStream_t st) { seed_tick_kernel_##tu<<<1, 1, 0, st>>>(inc); }

// keep-mask scale: returns 0 or 1/(1-p). p == 0 -> always 1.  One 64-bit hash serves the 4 consecutive element
// indices idx & ~3 .. idx | 3 (16 random bits each), so vectorised callers pay one hash per 4 elements.
__device__ __forceinline__ uint64_t dropout_bits4(uint64_t seed, uint64_t idx4) {
  uint64_t z = seed + g_seed_tick * 0x2545F4914F6CDD1Dull + idx4 * 0x9E3779B97F4A7C15ull;
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
__device__ __forceinline__ float dropout_scale(uint64_t seed, uint64_t idx, float p, float inv_keep) {
  if (p <= 0.f) return 1.f;
  const uint32_t r16 = (uint32_t)(dropout_bits4(seed, idx >> 2) >> ((idx & 3) * 16)) & 0xffffu;
  return (float)r16 < p * 65536.f ? 0.f : inv_keep;
}
// 4 consecutive elements starting at idx (idx % 4 == 0): out[k] *= mask
__device__ __forceinline__ void dropout_apply4(uint64_t seed, uint64_t idx, float p, float inv_keep, float* v) {
  const uint64_t bits = dropout_bits4(seed, idx >> 2);
  const float thr = p * 65536.f;
#pragma unroll
  for (int k = 0; k < 4; k++) v[k] *= ((float)((uint32_t)(bits >> (16 * k)) & 0xffffu) < thr) ? 0.f : inv_keep;
}

static inline int ceil_div(int64_t a, int64_t b) { return (int)((a + b - 1) / b); }

// Number of SMs of the current device (148 on B200), queried once per translation unit: grid-size heuristics are written
// in units of it instead of a literal.
static inline int mdhs_num_sms() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    cudaGetDevice(&dev);
    if (cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || n <= 0) n = 148;
  }
  return n;
}
