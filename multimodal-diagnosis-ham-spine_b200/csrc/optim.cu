// Fused optimizer step over the flat parameter buffer (scripts/train.py:257-309 Adam/AdamW/SGD;
// mibf_net/train_resnet.py:136-139).  One launch updates every parameter: reads the (all-reduced) fp32
// gradient, updates fp32 master weights and moments, refreshes the bf16 shadow copy consumed by the
// tensor-core kernels and zeroes the gradient for the next step.  HBM-bound: 4 x float4 per thread.
#include "common.cuh"
#include "../../include/mdhs_b200.h"

extern int64_t g_mdhs_launches;
void mdhs_seed_tick_gemm_tc(uint64_t, cudaStream_t);
void mdhs_seed_tick_norm(uint64_t, cudaStream_t);
void mdhs_seed_tick_attention(uint64_t, cudaStream_t);
void mdhs_seed_tick_elementwise(uint64_t, cudaStream_t);
void mdhs_seed_tick_kan_moe(uint64_t, cudaStream_t);
void mdhs_seed_tick_convnext(uint64_t, cudaStream_t);

namespace {

// mode 0: AdamW (decoupled decay), 1: Adam (L2 decay folded into the gradient)
__global__ void __launch_bounds__(1024) adam_flat_kernel(float* __restrict__ p, float* __restrict__ g,
                                                        const bf16* __restrict__ g16, float* __restrict__ m,
                                                        float* __restrict__ v, bf16* __restrict__ shadow, int64_t n, float lr,
                                                        float beta1, float beta2, float eps, float wd, float bc1, float bc2,
                                                        float grad_scale, int mode, int zero_grad,
                                                        const float* __restrict__ lr_dev, const int* __restrict__ step_dev) {
  const int64_t nv = n >> 2;
  if (lr_dev) lr = lr_dev[0];
  if (step_dev) {  // graph-replay safe: the step count lives in device memory
    const float t = (float)step_dev[0];
    bc1 = 1.f - powf(beta1, t);
    bc2 = 1.f - powf(beta2, t);
  }
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg;
    if (g16) {   // data-parallel runs with bf16 gradient buckets: the all-reduced gradient lives in the bf16 comm buffer
      const bf162* q = reinterpret_cast<const bf162*>(g16) + i * 2;
      const float2 a = __bfloat1622float2(q[0]), b = __bfloat1622float2(q[1]);
      gg = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gg = reinterpret_cast<float4*>(g)[i];
    }
    float4 mm = reinterpret_cast<float4*>(m)[i];
    float4 vv = reinterpret_cast<float4*>(v)[i];
    float* P = reinterpret_cast<float*>(&pp);
    float* G = reinterpret_cast<float*>(&gg);
    float* M = reinterpret_cast<float*>(&mm);
    float* V = reinterpret_cast<float*>(&vv);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float gr = G[k] * grad_scale;
      if (mode == 1) gr += wd * P[k];
      else P[k] *= (1.f - lr * wd);
      M[k] = beta1 * M[k] + (1.f - beta1) * gr;
      V[k] = beta2 * V[k] + (1.f - beta2) * gr * gr;
      const float denom = sqrtf(V[k]) / sqrtf(bc2) + eps;
      P[k] -= (lr / bc1) * (M[k] / denom);
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    reinterpret_cast<float4*>(m)[i] = mm;
    reinterpret_cast<float4*>(v)[i] = vv;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
      bf162* s = reinterpret_cast<bf162*>(shadow) + i * 2;
      s[0] = __floats2bfloat162_rn(P[0], P[1]);
      s[1] = __floats2bfloat162_rn(P[2], P[3]);
    }
  }
}

__global__ void __launch_bounds__(1024) sgd_flat_kernel(float* __restrict__ p, float* __restrict__ g,
                                                       const bf16* __restrict__ g16, float* __restrict__ mom,
                                                       bf16* __restrict__ shadow, int64_t n, float lr, float momentum, float wd,
                                                       float grad_scale, int first_step, int zero_grad,
                                                       const float* __restrict__ lr_dev, const int* __restrict__ step_dev) {
  const int64_t nv = n >> 2;
  if (lr_dev) lr = lr_dev[0];
  if (step_dev) first_step = step_dev[0] <= 1;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float4 pp = reinterpret_cast<float4*>(p)[i];
    float4 gg;
    if (g16) {
      const bf162* q = reinterpret_cast<const bf162*>(g16) + i * 2;
      const float2 a = __bfloat1622float2(q[0]), b = __bfloat1622float2(q[1]);
      gg = make_float4(a.x, a.y, b.x, b.y);
    } else {
      gg = reinterpret_cast<float4*>(g)[i];
    }
    float4 bb = mom ? reinterpret_cast<float4*>(mom)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float* P = reinterpret_cast<float*>(&pp);
    float* G = reinterpret_cast<float*>(&gg);
    float* Bf = reinterpret_cast<float*>(&bb);
#pragma unroll
    for (int k = 0; k < 4; k++) {
      float gr = G[k] * grad_scale + wd * P[k];
      if (mom) {
        Bf[k] = first_step ? gr : momentum * Bf[k] + gr;  // torch.optim.SGD: buffer initialised with the gradient
        gr = Bf[k];
      }
      P[k] -= lr * gr;
    }
    reinterpret_cast<float4*>(p)[i] = pp;
    if (mom) reinterpret_cast<float4*>(mom)[i] = bb;
    if (zero_grad) reinterpret_cast<float4*>(g)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
    if (shadow) {
      bf162* s = reinterpret_cast<bf162*>(shadow) + i * 2;
      s[0] = __floats2bfloat162_rn(P[0], P[1]);
      s[1] = __floats2bfloat162_rn(P[2], P[3]);
    }
  }
}

// Launch shape.  blocks_per_sm > 0 (default 16): that many 256-thread blocks per SM.  blocks_per_sm < 0: "leave -blocks_per_sm
// SMs free" -- ONE 1024-thread block per SM on (num_sms + blocks_per_sm) SMs.  A data-parallel step runs this kernel next to
// NCCL's all-reduce of the NEXT gradient bucket; NCCL's CTAs (hundreds of threads x ~100 registers) need an SM to themselves,
// so with blocks on every SM the collective only started when the optimizer had drained (CUPTI timeline, 2 x B200: the two
// simply alternated).  With whole SMs left free they run side by side; 1024 threads x 64 B keep 64 KB of loads in flight per SM.
void launch_shape(int64_t items, int blocks_per_sm, int* grid, int* block) {
  if (blocks_per_sm < 0) {
    int sms = mdhs_num_sms() + blocks_per_sm;
    if (sms < 1) sms = 1;
    int64_t g = (items + 1023) / 1024;
    *grid = (int)(g < 1 ? 1 : (g > sms ? sms : g));
    *block = 1024;
    return;
  }
  int64_t g = (items + 255) / 256;
  const int64_t cap = (int64_t)mdhs_num_sms() * (blocks_per_sm > 0 ? blocks_per_sm : 16);
  *grid = (int)(g < 1 ? 1 : (g > cap ? cap : g));
  *block = 256;
}

}  // namespace

// n must be a multiple of 4 (the flat buffer is padded); all pointers 16-byte aligned.
extern "C" int mdhs_adam_flat(float* params, float* grads, const void* grads_bf16, float* exp_avg, float* exp_avg_sq,
                              void* shadow_bf16, int64_t n,
                              float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                              int adamw, int zero_grad, const float* lr_dev, const int* step_dev, int blocks_per_sm,
                              void* stream) {
  if (!params || !grads || !exp_avg || !exp_avg_sq || n <= 0 || (n % 4) || (step < 1 && !step_dev)) return MDHS_ERR_ARG;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  g_mdhs_launches++;
  int grid, block;
  launch_shape(n / 4, blocks_per_sm, &grid, &block);
  adam_flat_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, grads, (const bf16*)grads_bf16, exp_avg, exp_avg_sq, (bf16*)shadow_bf16, n, lr, beta1, beta2, eps, weight_decay,
      bc1, bc2, grad_scale,
      adamw ? 0 : 1, zero_grad, lr_dev, step_dev);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_sgd_flat(float* params, float* grads, const void* grads_bf16, float* momentum_buf, void* shadow_bf16,
                             int64_t n, float lr,
                             float momentum, float weight_decay, float grad_scale, int first_step, int zero_grad,
                             const float* lr_dev, const int* step_dev, int blocks_per_sm, void* stream) {
  if (!params || !grads || n <= 0 || (n % 4)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  int grid, block;
  launch_shape(n / 4, blocks_per_sm, &grid, &block);
  sgd_flat_kernel<<<grid, block, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      params, grads, (const bf16*)grads_bf16, momentum_buf, (bf16*)shadow_bf16, n, lr, momentum, weight_decay, grad_scale,
      first_step, zero_grad, lr_dev,
      step_dev);
  MDHS_RETURN_LAST();
}

namespace {
__global__ void step_begin_kernel(int* step_dev) {
  if (step_dev) step_dev[0] += 1;
}
}  // namespace

// Once per training step, before the forward pass: advances the device-side step counter (optimizer bias
// correction) and the dropout seed tick of every translation unit.  Both live in device memory so that a
// captured CUDA graph of the whole step stays correct when replayed.
extern "C" int mdhs_step_begin(int* step_dev, void* stream) {
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  g_mdhs_launches += 7;
  step_begin_kernel<<<1, 1, 0, st>>>(step_dev);
  mdhs_seed_tick_gemm_tc(1, st);
  mdhs_seed_tick_norm(1, st);
  mdhs_seed_tick_attention(1, st);
  mdhs_seed_tick_elementwise(1, st);
  mdhs_seed_tick_kan_moe(1, st);
  mdhs_seed_tick_convnext(1, st);
  MDHS_RETURN_LAST();
}
