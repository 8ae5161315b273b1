// KAN (ConNexT/models/block/kan1.py) and sparsely-gated MoE (ConNexT/models/block/moe.py) kernels.
//  * kan_basis_fwd/bwd: elementwise expansion x -> [SiLU(x) | B_0..B_7(x)] (cubic B-splines on the layer's knot
//    buffer, Cox-de Boor recursion with the half-open order-0 indicator exactly like kan1.py:77-110) written as
//    the K-major bf16 operand of the tcgen05 GEMM, and its derivative.
//  * kan weight pack / gradient unpack: [base_weight | spline_weight * spline_scaler] <-> GEMM operand.
//  * moe gate forward/backward (noisy top-k gating, importance / load statistics, cv^2 balance loss) and the
//    gate-weighted combination of the expert outputs, without any host synchronisation.
#include "common.cuh"
#include "../../include/mdhs_b200.h"

extern int64_t g_mdhs_launches;

namespace {

constexpr int KAN_G = 5, KAN_K = 3;
constexpr int KAN_NB = KAN_G + KAN_K;            // 8 bases per input
constexpr int KAN_NT = KAN_G + 2 * KAN_K + 1;    // 12 knots per input

// order-0..3 bases at x; b3[8] = cubic bases, b2[9] = quadratic bases (needed for the derivative)
__device__ __forceinline__ void bspline_eval(float x, const float* __restrict__ t, float* b3, float* b2) {
  float b[KAN_NT - 1];
#pragma unroll
  for (int j = 0; j < KAN_NT - 1; j++) b[j] = (x >= t[j] && x < t[j + 1]) ? 1.f : 0.f;
#pragma unroll
  for (int k = 1; k <= KAN_K; k++) {
#pragma unroll
    for (int j = 0; j < KAN_NT - 1 - k; j++) {
      b[j] = (x - t[j]) / (t[j + k] - t[j]) * b[j] + (t[j + k + 1] - x) / (t[j + k + 1] - t[j + 1]) * b[j + 1];
    }
    if (k == KAN_K - 1) {
#pragma unroll
      for (int j = 0; j < KAN_NB + 1; j++) b2[j] = b[j];
    }
  }
#pragma unroll
  for (int j = 0; j < KAN_NB; j++) b3[j] = b[j];
}

// x fp32 [rows, in] -> op bf16 [rows, ld_op]: columns [0,in) = SiLU(x), [in + i*8 + g] = B_g(x_i); padding zeroed.
__global__ void __launch_bounds__(256) kan_basis_fwd_kernel(const float* __restrict__ x, const float* __restrict__ grid,
                                                            bf16* __restrict__ op, int64_t rows, int in, int64_t ldx,
                                                            int ld_op) {
  const int64_t total = rows * in;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % in);
    const int64_t r = idx / in;
    const float v = x[r * ldx + i];
    float t[KAN_NT], b3[KAN_NB], b2[KAN_NB + 1];
#pragma unroll
    for (int j = 0; j < KAN_NT; j++) t[j] = grid[(int64_t)i * KAN_NT + j];
    bspline_eval(v, t, b3, b2);
    bf16* o = op + r * ld_op;
    o[i] = __float2bfloat16_rn(v / (1.f + __expf(-v)));
    store8(o + in + (int64_t)i * KAN_NB, b3);
  }
}

// dx = dop[:, i] * silu'(x) + sum_g dop[:, in + i*8 + g] * B_g'(x)
__global__ void __launch_bounds__(256) kan_basis_bwd_kernel(const float* __restrict__ x, const float* __restrict__ grid,
                                                            const bf16* __restrict__ dop, float* __restrict__ dx, int64_t rows,
                                                            int in, int64_t ldx, int ld_op, int accumulate) {
  const int64_t total = rows * in;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % in);
    const int64_t r = idx / in;
    const float v = x[r * ldx + i];
    float t[KAN_NT], b3[KAN_NB], b2[KAN_NB + 1], g[8];
#pragma unroll
    for (int j = 0; j < KAN_NT; j++) t[j] = grid[(int64_t)i * KAN_NT + j];
    bspline_eval(v, t, b3, b2);
    const bf16* d = dop + r * ld_op;
    load8(d + in + (int64_t)i * KAN_NB, g);
    const float sg = 1.f / (1.f + __expf(-v));
    float acc = __bfloat162float(d[i]) * (sg * (1.f + v * (1.f - sg)));
#pragma unroll
    for (int j = 0; j < KAN_NB; j++) {
      const float db = (float)KAN_K * (b2[j] / (t[j + KAN_K] - t[j]) - b2[j + 1] / (t[j + KAN_K + 1] - t[j + 1]));
      acc += g[j] * db;
    }
    dx[idx] = accumulate ? dx[idx] + acc : acc;
  }
}

// wcat bf16 [out_pad, ld] = [base_weight | spline_weight * scaler]; rows >= out and padding columns are zero.
__global__ void kan_weight_pack_kernel(const float* __restrict__ base_w, const float* __restrict__ spline_w,
                                       const float* __restrict__ scaler, bf16* __restrict__ wcat, int out, int out_pad, int in,
                                       int ld) {
  const int64_t total = (int64_t)out_pad * ld;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(idx % ld);
    const int o = (int)(idx / ld);
    float v = 0.f;
    if (o < out) {
      if (c < in) v = base_w[(int64_t)o * in + c];
      else if (c < in * (1 + KAN_NB)) {
        const int i = (c - in) / KAN_NB;
        v = spline_w[(int64_t)o * in * KAN_NB + (c - in)] * (scaler ? scaler[(int64_t)o * in + i] : 1.f);
      }
    }
    wcat[idx] = __float2bfloat16_rn(v);
  }
}

// gcat fp32 [out_pad, ld] -> base_w.grad += , spline_w.grad += g * scaler, scaler.grad += sum_g g * spline_w
__global__ void kan_wgrad_unpack_kernel(const float* __restrict__ gcat, const float* __restrict__ spline_w,
                                        const float* __restrict__ scaler, float* __restrict__ g_base, float* __restrict__ g_spline,
                                        float* __restrict__ g_scaler, int out, int in, int ld) {
  const int64_t total = (int64_t)out * in;
  for (int64_t idx = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (int64_t)gridDim.x * blockDim.x) {
    const int i = (int)(idx % in);
    const int o = (int)(idx / in);
    const float* g = gcat + (int64_t)o * ld;
    if (g_base) g_base[idx] += g[i];
    const float sc = scaler ? scaler[idx] : 1.f;
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < KAN_NB; j++) {
      const float gj = g[in + (int64_t)i * KAN_NB + j];
      if (g_spline) g_spline[idx * KAN_NB + j] += gj * sc;
      acc += gj * spline_w[idx * KAN_NB + j];
    }
    if (g_scaler && scaler) g_scaler[idx] += acc;
  }
}

// ------------------------------------------------------------------ MoE gating
constexpr int MOE_MAXE = 8;
__device__ __forceinline__ float normal_cdf(float z) { return 0.5f * (1.f + erff(z * 0.70710678118654752440f)); }
__device__ __forceinline__ float normal_pdf(float z) { return 0.39894228040143267794f * __expf(-0.5f * z * z); }

// One warp per row.  noisy != 0: logits = clean + noise * (softplus(x.Wn) + 1e-2).  Saves everything backward needs.
__global__ void __launch_bounds__(128) moe_gate_fwd_kernel(const float* __restrict__ x, const float* __restrict__ wg,
                                                           const float* __restrict__ wn, const float* __restrict__ noise,
                                                           float* __restrict__ gates, float* __restrict__ clean,
                                                           float* __restrict__ raw, float* __restrict__ probs,
                                                           int* __restrict__ topidx, float* __restrict__ importance,
                                                           float* __restrict__ load, const float* __restrict__ nmean,
                                                           const float* __restrict__ nstd, int B, int in, int E, int k, int noisy) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float c[MOE_MAXE], r[MOE_MAXE];
#pragma unroll
  for (int e = 0; e < MOE_MAXE; e++) c[e] = r[e] = 0.f;
  for (int i = lane; i < in; i += 32) {
    const float xv = x[(int64_t)b * in + i];
#pragma unroll
    for (int e = 0; e < MOE_MAXE; e++) {
      if (e < E) {
        c[e] = fmaf(xv, wg[(int64_t)i * E + e], c[e]);
        if (noisy) r[e] = fmaf(xv, wn[(int64_t)i * E + e], r[e]);
      }
    }
  }
#pragma unroll
  for (int e = 0; e < MOE_MAXE; e++) {
    c[e] = warp_sum(c[e]);
    r[e] = warp_sum(r[e]);
  }
  if (lane != 0) return;
  float lg[MOE_MAXE], sd[MOE_MAXE], p[MOE_MAXE];
  float mx = -INFINITY;
  for (int e = 0; e < E; e++) {
    sd[e] = noisy ? (fmaxf(r[e], 0.f) + log1pf(__expf(-fabsf(r[e]))) + 1e-2f) : 0.f;  // softplus + eps
    lg[e] = c[e] + (noisy ? noise[(int64_t)b * E + e] * sd[e] : 0.f);
    mx = fmaxf(mx, lg[e]);
    clean[(int64_t)b * E + e] = c[e];
    if (raw) raw[(int64_t)b * E + e] = r[e];
  }
  float se = 0.f;
  for (int e = 0; e < E; e++) {
    p[e] = __expf(lg[e] - mx);
    se += p[e];
  }
  for (int e = 0; e < E; e++) {
    p[e] /= se;
    probs[(int64_t)b * E + e] = p[e];
  }
  // top-(k+1) by repeated selection (first maximum wins ties, like torch.topk on CUDA for tiny rows)
  const int m = min(k + 1, E);
  bool used[MOE_MAXE];
  int idx[MOE_MAXE];
  for (int e = 0; e < E; e++) used[e] = false;
  for (int j = 0; j < m; j++) {
    int best = -1;
    for (int e = 0; e < E; e++)
      if (!used[e] && (best < 0 || p[e] > p[best])) best = e;
    used[best] = true;
    idx[j] = best;
    topidx[(int64_t)b * (k + 1) + j] = best;
  }
  float s = 0.f;
  for (int j = 0; j < k; j++) s += p[idx[j]];
  for (int e = 0; e < E; e++) gates[(int64_t)b * E + e] = 0.f;
  for (int j = 0; j < k; j++) {
    const float g = p[idx[j]] / (s + 1e-6f);
    gates[(int64_t)b * E + idx[j]] = g;
    atomicAdd(importance + idx[j], g);
  }
  if (noisy && k < E) {
    // smooth load estimator (moe.py:198-229): thresholds are the (k+1)-th / k-th largest softmax values
    const float thr_in = p[idx[k]], thr_out = p[idx[k - 1]];
    const float nm = nmean ? nmean[0] : 0.f, ns = nstd ? nstd[0] : 1.f;   // Normal(self.mean, self.std) buffers, moe.py:166-167
    for (int e = 0; e < E; e++) {
      const bool is_in = lg[e] > thr_in;
      const float z = (c[e] - (is_in ? thr_in : thr_out)) / sd[e];
      atomicAdd(load + e, normal_cdf((z - nm) / ns));
    }
  } else {
    for (int j = 0; j < k; j++)
      if (p[idx[j]] / (s + 1e-6f) > 0.f) atomicAdd(load + idx[j], 1.f);
  }
}

// loss = coef * (cv2(importance) + cv2(load)); also d loss / d importance, d loss / d load (single thread)
__global__ void moe_loss_kernel(const float* __restrict__ importance, const float* __restrict__ load, float* __restrict__ loss,
                                float* __restrict__ d_imp, float* __restrict__ d_load, int E, float coef) {
  if (threadIdx.x != 0 || blockIdx.x != 0) return;
  float total = 0.f;
  for (int which = 0; which < 2; which++) {
    const float* v = which == 0 ? importance : load;
    float* d = which == 0 ? d_imp : d_load;
    if (E == 1) {
      if (d) d[0] = 0.f;
      continue;
    }
    float mean = 0.f;
    for (int e = 0; e < E; e++) mean += v[e];
    mean /= (float)E;
    float var = 0.f;
    for (int e = 0; e < E; e++) var += (v[e] - mean) * (v[e] - mean);
    var /= (float)(E - 1);
    const float den = mean * mean + 1e-10f;
    total += var / den;
    if (d)
      for (int e = 0; e < E; e++)
        d[e] = coef * (2.f * (v[e] - mean) / ((float)(E - 1) * den) - var * 2.f * mean / ((float)E * den * den));
  }
  loss[0] = coef * total;
}

// Backward of the gating.  dgates = gradient arriving at the dense gate matrix from the combine step; the balance loss
// contributes d_imp (through gates) and, in noisy training, d_load (through the normal-cdf load estimator).
// One warp per row: lane 0 does the tiny per-row algebra, then all lanes apply the rank-1 updates.
__global__ void __launch_bounds__(128) moe_gate_bwd_kernel(const float* __restrict__ x, const float* __restrict__ wg,
                                                           const float* __restrict__ wn, const float* __restrict__ noise,
                                                           const float* __restrict__ dgates, const float* __restrict__ d_imp,
                                                           const float* __restrict__ d_load, float dloss_host,
                                                           const float* __restrict__ dloss_dev,
                                                           const float* __restrict__ clean, const float* __restrict__ raw,
                                                           const float* __restrict__ probs, const int* __restrict__ topidx,
                                                           float* __restrict__ dx, float* __restrict__ dwg, float* __restrict__ dwn,
                                                           const float* __restrict__ nmean, const float* __restrict__ nstd, int B,
                                                           int in, int E, int k, int noisy) {
  const int b = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (b >= B) return;
  float dc[MOE_MAXE], dr[MOE_MAXE];
#pragma unroll
  for (int e = 0; e < MOE_MAXE; e++) dc[e] = dr[e] = 0.f;
  const float dloss = dloss_dev ? dloss_dev[0] : dloss_host;
  if (lane == 0) {
    float p[MOE_MAXE], dp[MOE_MAXE], dlg[MOE_MAXE], sd[MOE_MAXE], sig[MOE_MAXE], lg[MOE_MAXE];
    for (int e = 0; e < E; e++) {
      p[e] = probs[(int64_t)b * E + e];
      dp[e] = 0.f;
      const float r = noisy ? raw[(int64_t)b * E + e] : 0.f;
      sd[e] = noisy ? (fmaxf(r, 0.f) + log1pf(__expf(-fabsf(r))) + 1e-2f) : 1.f;
      sig[e] = 1.f / (1.f + __expf(-r));
      lg[e] = clean[(int64_t)b * E + e] + (noisy ? noise[(int64_t)b * E + e] * sd[e] : 0.f);
    }
    const int* idx = topidx + (int64_t)b * (k + 1);
    float s = 0.f;
    for (int j = 0; j < k; j++) s += p[idx[j]];
    const float inv = 1.f / (s + 1e-6f);
    float dot = 0.f;
    for (int j = 0; j < k; j++) {
      const int e = idx[j];
      const float dg = dgates[(int64_t)b * E + e] + dloss * d_imp[e];
      dot += dg * p[e];
    }
    for (int j = 0; j < k; j++) {
      const int e = idx[j];
      const float dg = dgates[(int64_t)b * E + e] + dloss * d_imp[e];
      dp[e] = dg * inv - dot * inv * inv;
    }
    if (noisy && k < E) {
      const int e_in = idx[k], e_out = idx[k - 1];
      const float thr_in = p[e_in], thr_out = p[e_out];
      for (int e = 0; e < E; e++) {
        const bool is_in = lg[e] > thr_in;
        const float thr = is_in ? thr_in : thr_out;
        const float z = (clean[(int64_t)b * E + e] - thr) / sd[e];
        const float nm = nmean ? nmean[0] : 0.f, ns = nstd ? nstd[0] : 1.f;
        const float gz = dloss * d_load[e] * normal_pdf((z - nm) / ns) / ns;
        dc[e] += gz / sd[e];                                  // through clean_values
        dr[e] += gz * (-z / sd[e]) * sig[e];                  // through noise_stddev = softplus(raw) + eps
        dp[is_in ? e_in : e_out] += -gz / sd[e];              // through the threshold (a softmax value)
      }
    }
    float pd = 0.f;
    for (int e = 0; e < E; e++) pd += p[e] * dp[e];
    for (int e = 0; e < E; e++) {
      dlg[e] = p[e] * (dp[e] - pd);
      dc[e] += dlg[e];
      if (noisy) dr[e] += dlg[e] * noise[(int64_t)b * E + e] * sig[e];
    }
  }
#pragma unroll
  for (int e = 0; e < MOE_MAXE; e++) {
    dc[e] = __shfl_sync(0xffffffffu, dc[e], 0);
    dr[e] = __shfl_sync(0xffffffffu, dr[e], 0);
  }
  for (int i = lane; i < in; i += 32) {
    const float xv = x[(int64_t)b * in + i];
    float acc = 0.f;
#pragma unroll
    for (int e = 0; e < MOE_MAXE; e++) {
      if (e < E) {
        acc += dc[e] * wg[(int64_t)i * E + e];
        if (dwg) atomicAdd(dwg + (int64_t)i * E + e, xv * dc[e]);
        if (noisy) {
          acc += dr[e] * wn[(int64_t)i * E + e];
          if (dwn) atomicAdd(dwn + (int64_t)i * E + e, xv * dr[e]);
        }
      }
    }
    if (dx) dx[(int64_t)b * in + i] += acc;
  }
}

// y[b,:] = sum_e gates[b,e] * Y[e,b,:]   and its backward
__global__ void moe_combine_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ Y, float* __restrict__ y, int B,
                                       int E, int C, int ldy) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * C) return;
  const int b = idx / C, c = idx % C;
  float acc = 0.f;
  for (int e = 0; e < E; e++) {
    const float g = gates[(int64_t)b * E + e];
    if (g != 0.f) acc += g * Y[((int64_t)e * B + b) * ldy + c];
  }
  y[idx] = acc;
}
__global__ void moe_combine_bwd_kernel(const float* __restrict__ gates, const float* __restrict__ Y, const float* __restrict__ dy,
                                       float* __restrict__ dgates, float* __restrict__ dY, int B, int E, int C, int ldy) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= B * E) return;
  const int b = idx / E, e = idx % E;
  const float g = gates[idx];
  float acc = 0.f;
  for (int c = 0; c < ldy; c++) {
    const float d = c < C ? dy[(int64_t)b * C + c] : 0.f;
    const int64_t o = ((int64_t)e * B + b) * ldy + c;
    if (g != 0.f && c < C) acc += d * Y[o];
    dY[o] = g * d;
  }
  dgates[idx] = acc;   // only selected experts matter downstream (others are masked by the top-k selection)
}

// standard normal noise for the noisy gating (moe.py:246: torch.randn_like): Box-Muller on the counter-based hash
__global__ void randn_f32_kernel(float* __restrict__ out, int64_t n, uint64_t seed) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const uint64_t bits = dropout_bits4(seed, (uint64_t)i);
    const float u1 = ((float)(uint32_t)(bits >> 40) + 1.f) * (1.f / 16777217.f);   // (0, 1)
    const float u2 = (float)(uint32_t)(bits & 0xffffffu) * (1.f / 16777216.f);
    out[i] = sqrtf(-2.f * logf(u1)) * cospif(2.f * u2);
  }
}

int grid_for(int64_t items) {
  int64_t g = (items + 255) / 256;
  const int64_t cap = (int64_t)mdhs_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

MDHS_DEFINE_SEED_TICK(kan_moe)

extern "C" int mdhs_randn_f32(float* out, int64_t n, uint64_t seed, void* stream) {
  if (!out || n <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  randn_f32_kernel<<<grid_for(n), 256, 0, ST(stream)>>>(out, n, seed);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_kan_basis_fwd(const float* x, int64_t ldx, const float* grid, void* op, int64_t rows, int in, int ld_op,
                                  void* stream) {
  if (!x || !grid || !op || rows <= 0 || in <= 0 || ld_op < in * (1 + KAN_NB) || (in % 8) || (ld_op % 8)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  if (ld_op > in * (1 + KAN_NB)) cudaMemsetAsync(op, 0, (size_t)rows * ld_op * 2, ST(stream));
  kan_basis_fwd_kernel<<<grid_for(rows * in), 256, 0, ST(stream)>>>(x, grid, (bf16*)op, rows, in, ldx, ld_op);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_kan_basis_bwd(const float* x, int64_t ldx, const float* grid, const void* dop, float* dx, int64_t rows, int in,
                                  int ld_op, int accumulate, void* stream) {
  if (!x || !grid || !dop || !dx || rows <= 0 || (in % 8)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  kan_basis_bwd_kernel<<<grid_for(rows * in), 256, 0, ST(stream)>>>(x, grid, (const bf16*)dop, dx, rows, in, ldx, ld_op, accumulate);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_kan_weight_pack(const float* base_w, const float* spline_w, const float* scaler, void* wcat, int out, int out_pad,
                                    int in, int ld, void* stream) {
  if (!base_w || !spline_w || !wcat || ld < in * (1 + KAN_NB)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  kan_weight_pack_kernel<<<grid_for((int64_t)out_pad * ld), 256, 0, ST(stream)>>>(base_w, spline_w, scaler, (bf16*)wcat, out, out_pad, in,
                                                                                  ld);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_kan_wgrad_unpack(const float* gcat, const float* spline_w, const float* scaler, float* g_base, float* g_spline,
                                     float* g_scaler, int out, int in, int ld, void* stream) {
  if (!gcat || !spline_w) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  kan_wgrad_unpack_kernel<<<grid_for((int64_t)out * in), 256, 0, ST(stream)>>>(gcat, spline_w, scaler, g_base, g_spline, g_scaler, out, in,
                                                                               ld);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_moe_gate_fwd(const float* x, const float* wg, const float* wn, const float* noise, float* gates, float* clean,
                                 float* raw, float* probs, int* topidx, float* importance, float* load, const float* normal_mean,
                                 const float* normal_std, int B, int in, int E, int k, int noisy, void* stream) {
  if (!x || !wg || !gates || !clean || !probs || !topidx || !importance || !load || E < 1 || E > MOE_MAXE || k < 1 || k > E)
    return MDHS_ERR_ARG;
  if (noisy && (!wn || !noise || !raw)) return MDHS_ERR_ARG;
  cudaMemsetAsync(importance, 0, sizeof(float) * E, ST(stream));
  cudaMemsetAsync(load, 0, sizeof(float) * E, ST(stream));
  g_mdhs_launches++;
  moe_gate_fwd_kernel<<<ceil_div(B, 4), 128, 0, ST(stream)>>>(x, wg, wn, noise, gates, clean, raw, probs, topidx, importance, load,
                                                              normal_mean, normal_std, B, in, E, k, noisy);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_moe_loss(const float* importance, const float* load, float* loss, float* d_imp, float* d_load, int E, float coef,
                             void* stream) {
  if (!importance || !load || !loss) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  moe_loss_kernel<<<1, 32, 0, ST(stream)>>>(importance, load, loss, d_imp, d_load, E, coef);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_moe_gate_bwd(const float* x, const float* wg, const float* wn, const float* noise, const float* dgates,
                                 const float* d_imp, const float* d_load, float dloss, const float* dloss_dev, const float* clean,
                                 const float* raw, const float* probs, const int* topidx, float* dx, float* dwg, float* dwn, const float* normal_mean,
                                 const float* normal_std, int B, int in, int E, int k, int noisy, void* stream) {
  if (!x || !wg || !dgates || !d_imp || !clean || !probs || !topidx) return MDHS_ERR_ARG;
  if (noisy && (!wn || !noise || !raw || !d_load)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  moe_gate_bwd_kernel<<<ceil_div(B, 4), 128, 0, ST(stream)>>>(x, wg, wn, noise, dgates, d_imp, d_load, dloss, dloss_dev, clean, raw, probs, topidx, dx,
                                                              dwg, dwn, normal_mean, normal_std, B, in, E, k, noisy);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_moe_combine_fwd(const float* gates, const float* Y, float* y, int B, int E, int C, int ldy, void* stream) {
  if (!gates || !Y || !y) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  moe_combine_fwd_kernel<<<ceil_div((int64_t)B * C, 128), 128, 0, ST(stream)>>>(gates, Y, y, B, E, C, ldy);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_moe_combine_bwd(const float* gates, const float* Y, const float* dy, float* dgates, float* dY, int B, int E, int C,
                                    int ldy, void* stream) {
  if (!gates || !Y || !dy || !dgates || !dY) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  moe_combine_bwd_kernel<<<ceil_div((int64_t)B * E, 128), 128, 0, ST(stream)>>>(gates, Y, dy, dgates, dY, B, E, C, ldy);
  MDHS_RETURN_LAST();
}
