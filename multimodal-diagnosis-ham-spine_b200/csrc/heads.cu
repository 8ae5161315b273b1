// Small-M fp32 kernels for the classification heads and losses: rows = batch (<= a few thousand),
// outputs = classes (6/7) or a gate scalar.  These are latency-bound, so each is one fused launch with
// warp-shuffle reductions instead of a tensor-core GEMM.
#include "common.cuh"
#include "../../include/mdhs_b200.h"

extern int64_t g_mdhs_launches;

namespace {

// Y[m, n] = act(sum_k X[m,k] W[n,k] + b[n]); one CTA per row, one warp per output column (strided).
__global__ void __launch_bounds__(256) linear_f32_fwd_kernel(const float* __restrict__ X, int64_t ldx, const float* __restrict__ W,
                                                             const float* __restrict__ bias, float* __restrict__ Y, int64_t ldy,
                                                             int N, int K, int act) {
  extern __shared__ float xs[];
  const int m = blockIdx.x;
  for (int k = threadIdx.x; k < K; k += blockDim.x) xs[k] = X[(int64_t)m * ldx + k];
  __syncthreads();
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  for (int n = warp; n < N; n += nw) {
    const float* w = W + (int64_t)n * K;
    float acc = 0.f;
    for (int k = lane; k < K; k += 32) acc = fmaf(xs[k], w[k], acc);
    acc = warp_sum(acc);
    if (lane == 0) {
      acc += bias ? bias[n] : 0.f;
      if (act == MDHS_ACT_RELU) acc = fmaxf(acc, 0.f);
      else if (act == MDHS_ACT_GELU) acc = gelu_erf(acc);
      Y[(int64_t)m * ldy + n] = acc;
    }
  }
}

// dX[m,k] (+)= sum_n dY[m,n] W[n,k]; one CTA per row, threads over k.
__global__ void __launch_bounds__(256) linear_f32_bwd_input_kernel(const float* __restrict__ dY, int64_t lddy,
                                                                   const float* __restrict__ W, float* __restrict__ dX,
                                                                   int64_t lddx, int N, int K, int accumulate) {
  extern __shared__ float dys[];
  const int m = blockIdx.x;
  for (int n = threadIdx.x; n < N; n += blockDim.x) dys[n] = dY[(int64_t)m * lddy + n];
  __syncthreads();
  for (int k = threadIdx.x; k < K; k += blockDim.x) {
    float acc = 0.f;
    for (int n = 0; n < N; n++) acc = fmaf(dys[n], W[(int64_t)n * K + k], acc);
    float* o = dX + (int64_t)m * lddx + k;
    *o = accumulate ? *o + acc : acc;
  }
}

// dW[n,k] += sum_m dY[m,n] X[m,k];  db[n] += sum_m dY[m,n].  grid = (ceil(K/256), N)
__global__ void __launch_bounds__(256) linear_f32_bwd_weight_kernel(const float* __restrict__ dY, int64_t lddy,
                                                                    const float* __restrict__ X, int64_t ldx,
                                                                    float* __restrict__ dW, float* __restrict__ db, int M, int N,
                                                                    int K) {
  const int n = blockIdx.y;
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  float acc = 0.f, accb = 0.f;
  for (int m = 0; m < M; m++) {
    const float g = dY[(int64_t)m * lddy + n];
    accb += g;
    if (k < K) acc = fmaf(g, X[(int64_t)m * ldx + k], acc);
  }
  if (k < K) dW[(int64_t)n * K + k] += acc;
  if (db && blockIdx.x == 0 && threadIdx.x == 0) db[n] += accb;
}

// Cross entropy (mean reduction) with optional class weights and label smoothing, torch semantics:
//   loss = sum_i [(1-eps) w[y_i] (-logp_i[y_i]) + eps/C sum_c w[c] (-logp_i[c])] / sum_i w[y_i]
// Focal variant (scripts/train.py:46-61): ce_i = w[y_i](-logp_i[y_i]); loss = mean((1-exp(-ce_i))^gamma ce_i).
// Writes loss[0] and dlogits = d loss / d logits (caller scales by the incoming gradient).  Single CTA.
__global__ void __launch_bounds__(256) ce_loss_kernel(const float* __restrict__ logits, int64_t ld, const int64_t* __restrict__ labels,
                                                      const float* __restrict__ cw, float* __restrict__ loss,
                                                      float* __restrict__ dlogits, int B, int C, float eps, int focal,
                                                      float gamma) {
  __shared__ float red[32];
  float num = 0.f, den = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float* z = logits + (int64_t)i * ld;
    float mx = -INFINITY;
    for (int c = 0; c < C; c++) mx = fmaxf(mx, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; c++) se += __expf(z[c] - mx);
    const float lse = mx + __logf(se);
    const int y = (int)labels[i];
    const float wy = cw ? cw[y] : 1.f;
    if (focal) {
      const float ce = wy * (lse - z[y]);
      const float pt = __expf(-ce);
      num += powf(fmaxf(1.f - pt, 0.f), gamma) * ce;
      den += 1.f;
    } else {
      float l = (1.f - eps) * wy * (lse - z[y]);
      if (eps > 0.f) {
        float s = 0.f;
        for (int c = 0; c < C; c++) s += (cw ? cw[c] : 1.f) * (lse - z[c]);
        l += eps / (float)C * s;
      }
      num += l;
      den += wy;
    }
  }
  num = block_sum(num, red);
  den = block_sum(den, red);
  if (threadIdx.x == 0) loss[0] = num / den;
  if (!dlogits) return;
  const float inv_den = 1.f / den;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    const float* z = logits + (int64_t)i * ld;
    float mx = -INFINITY;
    for (int c = 0; c < C; c++) mx = fmaxf(mx, z[c]);
    float se = 0.f;
    for (int c = 0; c < C; c++) se += __expf(z[c] - mx);
    const float lse = mx + __logf(se);
    const int y = (int)labels[i];
    const float wy = cw ? cw[y] : 1.f;
    if (focal) {
      const float ce = wy * (lse - z[y]);
      const float pt = __expf(-ce);
      const float om = fmaxf(1.f - pt, 0.f);
      // d/dce [(1-pt)^g ce] = (1-pt)^g + g (1-pt)^(g-1) pt ce
      const float dl = powf(om, gamma) + (om > 0.f ? gamma * powf(om, gamma - 1.f) * pt * ce : 0.f);
      for (int c = 0; c < C; c++) {
        const float pc = __expf(z[c] - lse);
        dlogits[(int64_t)i * C + c] = dl * wy * (pc - (c == y ? 1.f : 0.f)) * inv_den;
      }
    } else {
      float wsum = 0.f;
      if (eps > 0.f)
        for (int c = 0; c < C; c++) wsum += cw ? cw[c] : 1.f;
      for (int c = 0; c < C; c++) {
        const float pc = __expf(z[c] - lse);
        float g = (1.f - eps) * wy * (pc - (c == y ? 1.f : 0.f));
        if (eps > 0.f) g += eps / (float)C * (wsum * pc - (cw ? cw[c] : 1.f));
        dlogits[(int64_t)i * C + c] = g * inv_den;
      }
    }
  }
}

// y = a * x (+ b * y) elementwise fp32 (gradient scaling, accumulation).
__global__ void axpby_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, const float* __restrict__ a_dev,
                             float a, float b) {
  const float aa = a_dev ? a_dev[0] * a : a;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = aa * x[i] + (b != 0.f ? b * y[i] : 0.f);
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int mdhs_linear_f32_fwd(const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy, int M,
                                   int N, int K, int act, void* stream) {
  if (!X || !W || !Y || M <= 0 || N <= 0 || K <= 0 || K > 12000) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  linear_f32_fwd_kernel<<<M, 256, K * sizeof(float), ST(stream)>>>(X, ldx, W, bias, Y, ldy, N, K, act);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_linear_f32_bwd(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* W, float* dX,
                                   int64_t lddx, int accumulate_dx, float* dW, float* db, int M, int N, int K, void* stream) {
  if (!dY || !W || M <= 0 || N <= 0 || K <= 0 || N > 12000) return MDHS_ERR_ARG;
  if (dX) {
    g_mdhs_launches++;
    linear_f32_bwd_input_kernel<<<M, 256, N * sizeof(float), ST(stream)>>>(dY, lddy, W, dX, lddx, N, K, accumulate_dx);
  }
  if (dW) {
    if (!X) return MDHS_ERR_ARG;
    g_mdhs_launches++;
    linear_f32_bwd_weight_kernel<<<dim3(ceil_div(K, 256), N), 256, 0, ST(stream)>>>(dY, lddy, X, ldx, dW, db, M, N, K);
  }
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_ce_loss(const float* logits, int64_t ld, const int64_t* labels, const float* class_weights, float* loss,
                            float* dlogits, int B, int C, float label_smoothing, int focal, float gamma, void* stream) {
  if (!logits || !labels || !loss || B <= 0 || C <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  ce_loss_kernel<<<1, 256, 0, ST(stream)>>>(logits, ld, labels, class_weights, loss, dlogits, B, C, label_smoothing, focal, gamma);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_axpby_f32(const float* x, float* y, int64_t n, const float* a_dev, float a, float b, void* stream) {
  if (!x || !y || n <= 0) return MDHS_ERR_ARG;
  int grid = (int)((n + 255) / 256);
  if (grid > 148 * 8) grid = 148 * 8;
  g_mdhs_launches++;
  axpby_kernel<<<grid, 256, 0, ST(stream)>>>(x, y, n, a_dev, a, b);
  MDHS_RETURN_LAST();
}
