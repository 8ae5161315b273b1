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

// ------------------------------------------------------------------ supervised contrastive loss (scripts/train.py:23-44)
// f_i = x_i / max(|x_i|, 1e-12); s_ij = f_i.f_j / T; logp_ij = (s_ij - m_i) - log(sum_{k != i} exp(s_ik - m_i) + 1e-8);
// loss = -mean_i [ sum_{j in pos(i)} logp_ij / (|pos(i)| + 1e-8) ].  Three small kernels (one CTA per sample):
// normalise; loss row + G = d loss / d s (B x B, fp32); dx = normalise'( (G + G^T) f / T ).
constexpr int SUPCON_MAXB = 2048;

__global__ void __launch_bounds__(256) supcon_normalize_kernel(const float* __restrict__ x, int64_t ldx, float* __restrict__ f,
                                                               float* __restrict__ inv_norm, int D) {
  __shared__ float red[32];
  const int i = blockIdx.x;
  float q = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float v = x[(int64_t)i * ldx + d];
    q = fmaf(v, v, q);
  }
  q = block_sum(q, red);
  const float inv = 1.f / fmaxf(sqrtf(q), 1e-12f);
  if (threadIdx.x == 0) inv_norm[i] = inv;
  for (int d = threadIdx.x; d < D; d += blockDim.x) f[(int64_t)i * D + d] = x[(int64_t)i * ldx + d] * inv;
}

__global__ void __launch_bounds__(256) supcon_row_kernel(const float* __restrict__ f, const int64_t* __restrict__ labels,
                                                         float* __restrict__ loss, float* __restrict__ G, int B, int D, float inv_t) {
  __shared__ float s[SUPCON_MAXB];
  __shared__ float red[32];
  const int i = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* fi = f + (int64_t)i * D;
  for (int j = warp; j < B; j += nw) {
    const float* fj = f + (int64_t)j * D;
    float a = 0.f;
    for (int d = lane; d < D; d += 32) a = fmaf(fi[d], fj[d], a);
    a = warp_sum(a);
    if (lane == 0) s[j] = a * inv_t;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int j = threadIdx.x; j < B; j += blockDim.x) m = fmaxf(m, s[j]);
  m = block_max(m, red);
  const int64_t yi = labels[i];
  float z = 0.f, npos = 0.f, lpos = 0.f;
  for (int j = threadIdx.x; j < B; j += blockDim.x) {
    if (j == i) continue;
    const float l = s[j] - m;
    z += __expf(l);
    if (labels[j] == yi) {
      npos += 1.f;
      lpos += l;
    }
  }
  z = block_sum(z, red);
  npos = block_sum(npos, red);
  lpos = block_sum(lpos, red);
  const float Z = z + 1e-8f;
  const float wpos = 1.f / (npos + 1e-8f);
  if (threadIdx.x == 0) atomicAdd(loss, -(lpos - npos * logf(Z)) * wpos / (float)B);
  if (G != nullptr) {
    const float frac = npos * wpos;   // = sum_{j in pos} 1 / (|pos| + eps)
    for (int j = threadIdx.x; j < B; j += blockDim.x) {
      float g = 0.f;
      if (j != i) {
        const float e = __expf(s[j] - m) / Z;
        g = -(((labels[j] == yi) ? wpos : 0.f) - frac * e) / (float)B;
      }
      G[(int64_t)i * B + j] = g;
    }
  }
}

__global__ void __launch_bounds__(256) supcon_bwd_kernel(const float* __restrict__ f, const float* __restrict__ inv_norm,
                                                         const float* __restrict__ G, float* __restrict__ dx, int64_t lddx, int B,
                                                         int D, float inv_t) {
  __shared__ float w[SUPCON_MAXB];
  __shared__ float red[32];
  const int i = blockIdx.x;
  for (int j = threadIdx.x; j < B; j += blockDim.x) w[j] = (G[(int64_t)i * B + j] + G[(int64_t)j * B + i]) * inv_t;
  __syncthreads();
  // df = sum_j w_j f_j (threads over d), then the tangent projection of the normalisation
  float dot = 0.f;
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    float a = 0.f;
    for (int j = 0; j < B; j++) a = fmaf(w[j], f[(int64_t)j * D + d], a);
    dx[(int64_t)i * lddx + d] = a;          // staged, rewritten below
    dot = fmaf(a, f[(int64_t)i * D + d], dot);
  }
  dot = block_sum(dot, red);
  const float inv = inv_norm[i];
  for (int d = threadIdx.x; d < D; d += blockDim.x) {
    const float a = dx[(int64_t)i * lddx + d];
    dx[(int64_t)i * lddx + d] = (a - f[(int64_t)i * D + d] * dot) * inv;
  }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int mdhs_supcon_loss(const float* x, int64_t ldx, const int64_t* labels, float* loss, float* dx, int64_t lddx, float* f_ws,
                                float* inv_norm_ws, float* g_ws, int B, int D, float temperature, void* stream) {
  if (!x || !labels || !loss || !f_ws || !inv_norm_ws || B <= 0 || B > SUPCON_MAXB || D <= 0 || temperature <= 0.f) return MDHS_ERR_ARG;
  if (dx && !g_ws) return MDHS_ERR_ARG;
  cudaStream_t st = ST(stream);
  cudaError_t e = cudaMemsetAsync(loss, 0, sizeof(float), st);
  if (e != cudaSuccess) return (int)e;
  g_mdhs_launches += dx ? 3 : 2;
  supcon_normalize_kernel<<<B, 256, 0, st>>>(x, ldx, f_ws, inv_norm_ws, D);
  supcon_row_kernel<<<B, 256, 0, st>>>(f_ws, labels, loss, dx ? g_ws : nullptr, B, D, 1.f / temperature);
  if (dx) supcon_bwd_kernel<<<B, 256, 0, st>>>(f_ws, inv_norm_ws, g_ws, dx, lddx, B, D, 1.f / temperature);
  MDHS_RETURN_LAST();
}


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
  if (grid > (int64_t)mdhs_num_sms() * 8) grid = (int64_t)mdhs_num_sms() * 8;
  g_mdhs_launches++;
  axpby_kernel<<<grid, 256, 0, ST(stream)>>>(x, y, n, a_dev, a, b);
  MDHS_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// Adaptive level weighting of the hierarchical fusion (README.md:15 "layer-wise interaction with adaptive weighting"):
// out = sum_l softmax(logits)_l * p_l over L <= 4 pooled feature sets [n] fp32.  Backward: dp_l = w_l * dout and
// dlogits_l += w_l * (g_l - sum_k w_k g_k) with g_l = <dout, p_l> (block reduction + one atomic per block, then one thread).
// ---------------------------------------------------------------------------------------------
namespace {
struct LevelPtrs {
  const float* p[4];
  float* dp[4];
};
__device__ __forceinline__ void level_softmax(const float* logits, int L, float* w) {
  float m = -INFINITY;
  for (int l = 0; l < L; l++) m = fmaxf(m, logits[l]);
  float s = 0.f;
  for (int l = 0; l < L; l++) {
    w[l] = __expf(logits[l] - m);
    s += w[l];
  }
  for (int l = 0; l < L; l++) w[l] /= s;
}
__global__ void level_mix_fwd_kernel(LevelPtrs P, const float* __restrict__ logits, float* __restrict__ out, int64_t n, int L) {
  float w[4];
  level_softmax(logits, L, w);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    float acc = 0.f;
    for (int l = 0; l < L; l++) acc = fmaf(w[l], P.p[l][i], acc);
    out[i] = acc;
  }
}
__global__ void level_mix_bwd_kernel(LevelPtrs P, const float* __restrict__ logits, const float* __restrict__ dout, float* g_ws,
                                     int64_t n, int L) {
  __shared__ float red[32];
  float w[4], g[4] = {0.f, 0.f, 0.f, 0.f};
  level_softmax(logits, L, w);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float d = dout[i];
    for (int l = 0; l < L; l++) {
      g[l] = fmaf(d, P.p[l][i], g[l]);
      P.dp[l][i] = w[l] * d;
    }
  }
  for (int l = 0; l < L; l++) {
    const float s = block_sum(g[l], red);
    if (threadIdx.x == 0) atomicAdd(g_ws + l, s);
  }
}
__global__ void level_mix_dlogits_kernel(const float* __restrict__ logits, const float* __restrict__ g_ws, float* dlogits, int L) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float w[4];
  level_softmax(logits, L, w);
  float mean = 0.f;
  for (int l = 0; l < L; l++) mean = fmaf(w[l], g_ws[l], mean);
  for (int l = 0; l < L; l++) dlogits[l] += w[l] * (g_ws[l] - mean);
}
}  // namespace

extern "C" int mdhs_level_mix_fwd(const float* const* p, const float* logits, float* out, int64_t n, int L, void* stream) {
  if (!p || !logits || !out || n <= 0 || L < 1 || L > 4) return MDHS_ERR_ARG;
  LevelPtrs P = {};
  for (int l = 0; l < L; l++) {
    if (!p[l]) return MDHS_ERR_ARG;
    P.p[l] = p[l];
  }
  int grid = (int)((n + 255) / 256);
  if (grid > mdhs_num_sms() * 4) grid = mdhs_num_sms() * 4;
  g_mdhs_launches++;
  level_mix_fwd_kernel<<<grid, 256, 0, ST(stream)>>>(P, logits, out, n, L);
  MDHS_RETURN_LAST();
}

// g_ws: fp32 [4] workspace.  dp[l] receive w_l * dout; dlogits (may be NULL) accumulates (+=).
extern "C" int mdhs_level_mix_bwd(const float* const* p, float* const* dp, const float* logits, const float* dout, float* g_ws,
                                  float* dlogits, int64_t n, int L, void* stream) {
  if (!p || !dp || !logits || !dout || !g_ws || n <= 0 || L < 1 || L > 4) return MDHS_ERR_ARG;
  LevelPtrs P = {};
  for (int l = 0; l < L; l++) {
    if (!p[l] || !dp[l]) return MDHS_ERR_ARG;
    P.p[l] = p[l];
    P.dp[l] = dp[l];
  }
  cudaError_t e = cudaMemsetAsync(g_ws, 0, 4 * sizeof(float), ST(stream));
  if (e != cudaSuccess) return (int)e;
  int grid = (int)((n + 255) / 256);
  if (grid > mdhs_num_sms() * 4) grid = mdhs_num_sms() * 4;
  g_mdhs_launches += 2;
  level_mix_bwd_kernel<<<grid, 256, 0, ST(stream)>>>(P, logits, dout, g_ws, n, L);
  if (dlogits) level_mix_dlogits_kernel<<<1, 32, 0, ST(stream)>>>(logits, g_ws, dlogits, L);
  MDHS_RETURN_LAST();
}
