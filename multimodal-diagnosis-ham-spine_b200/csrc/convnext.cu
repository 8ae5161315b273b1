// ConvNeXt-specific kernels (torchvision CNBlock as used by ConNexT/models/ourmodel.py:57-62 and the ConvNeXt image
// encoders of configs 4): NHWC bf16 activations [B*H*W, C].
//   * depthwise 7x7 convolution (pad 3, groups = C): forward, input gradient (same kernel, mirrored taps) and
//     weight / bias gradient.  HBM / L1-bound: each thread owns 8 channels (one 16-byte vector) x 4 adjacent output
//     pixels, so a filter row costs 10 vector loads for 28 tap applications.
//   * layer-scale + stochastic-depth + residual:  out = x + ls[c] * keep(b) * z   and its backward
//     (dz, dls; dx = dy is the identity branch and needs no kernel).
//   * single-query attention (ourmodel.py:17-31 with a 1-token query: the text -> image direction): softmax over the
//     T image positions of q.k_t (no 1/sqrt(d) in the reference), out = sum_t p_t v_t; forward + backward.
// The 1x1 / patchify convolutions, LayerNorms and MLPs of the block run on the shared GEMM / LayerNorm kernels.
#include "common.cuh"
#include "../../include/mdhs_b200.h"

MDHS_DEFINE_SEED_TICK(convnext)

extern int64_t g_mdhs_launches;

namespace {

// ------------------------------------------------------------------ depthwise 7x7
// grid = (ceil(groups / 32), ceil(C / 64)); block = 256 threads = 8 channel vectors x 32 pixel groups; a pixel group is
// 4 consecutive output pixels of one image row.  w: [C][49] fp32 (the Conv2d weight [C,1,7,7]); FLIP mirrors the taps
// (input gradient = correlation of dy with the flipped filter).
template <bool FLIP>
__global__ void __launch_bounds__(256) dwconv7_kernel(const bf16* __restrict__ x, const float* __restrict__ w,
                                                      const float* __restrict__ bias, bf16* __restrict__ y, int B, int H, int W,
                                                      int C) {
  __shared__ __align__(16) float wsm[49][64];
  const int slab = blockIdx.y * 64;
  for (int i = threadIdx.x; i < 49 * 64; i += blockDim.x) {
    const int tap = i >> 6, cl = i & 63;
    const int c = slab + cl;
    wsm[tap][cl] = c < C ? w[(int64_t)c * 49 + (FLIP ? 48 - tap : tap)] : 0.f;
  }
  __syncthreads();
  const int cv = threadIdx.x & 7;
  const int g = threadIdx.x >> 3;
  const int c0 = slab + cv * 8;
  const int wq = (W + 3) >> 2;
  const int64_t groups = (int64_t)B * H * wq;
  const int64_t G = (int64_t)blockIdx.x * 32 + g;
  if (G >= groups || c0 >= C) return;
  const int w0 = (int)(G % wq) * 4;
  const int h = (int)((G / wq) % H);
  const int b = (int)(G / ((int64_t)wq * H));
  float acc[4][8];
#pragma unroll
  for (int o = 0; o < 4; o++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc[o][k] = (!FLIP && bias != nullptr) ? bias[c0 + k] : 0.f;
#pragma unroll 1
  for (int r = 0; r < 7; r++) {
    const int hh = h + r - 3;
    if (hh < 0 || hh >= H) continue;
    const bf16* row = x + ((int64_t)(b * H + hh) * W) * C + c0;
#pragma unroll
    for (int j = 0; j < 10; j++) {
      const int ww = w0 + j - 3;
      if (ww < 0 || ww >= W) continue;
      float v[8];
      load8(row + (int64_t)ww * C, v);
#pragma unroll
      for (int o = 0; o < 4; o++) {
        const int s = j - o;
        if (s >= 0 && s < 7) {
          const float4 wa = *reinterpret_cast<const float4*>(&wsm[r * 7 + s][cv * 8]);
          const float4 wb = *reinterpret_cast<const float4*>(&wsm[r * 7 + s][cv * 8 + 4]);
          acc[o][0] = fmaf(v[0], wa.x, acc[o][0]); acc[o][1] = fmaf(v[1], wa.y, acc[o][1]);
          acc[o][2] = fmaf(v[2], wa.z, acc[o][2]); acc[o][3] = fmaf(v[3], wa.w, acc[o][3]);
          acc[o][4] = fmaf(v[4], wb.x, acc[o][4]); acc[o][5] = fmaf(v[5], wb.y, acc[o][5]);
          acc[o][6] = fmaf(v[6], wb.z, acc[o][6]); acc[o][7] = fmaf(v[7], wb.w, acc[o][7]);
        }
      }
    }
  }
  bf16* out = y + ((int64_t)(b * H + h) * W + w0) * C + c0;
#pragma unroll
  for (int o = 0; o < 4; o++)
    if (w0 + o < W) store8(out + (int64_t)o * C, acc[o]);
}

// Weight / bias gradient: dw[c][r*7+s] += sum_{b,h,w} dy[b,h,w,c] * x[b,h+r-3,w+s-3,c]; db[c] += sum dy.
// block = 224 threads = 8 channel vectors x 7 filter rows x 4 pixel lanes; grid = (pixel chunks, ceil(C / 64)).
__global__ void __launch_bounds__(224) dwconv7_wgrad_kernel(const bf16* __restrict__ x, const bf16* __restrict__ dy,
                                                            float* __restrict__ dw, float* __restrict__ db, int B, int H, int W,
                                                            int C, int64_t pix_per_block) {
  __shared__ float red[50][64];   // 49 taps + bias row, reduced across the 4 pixel lanes with shared atomics
  for (int i = threadIdx.x; i < 50 * 64; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int cv = threadIdx.x & 7;
  const int rr = (threadIdx.x >> 3) % 7;
  const int pl = threadIdx.x / 56;
  const int c0 = blockIdx.y * 64 + cv * 8;
  const int64_t total = (int64_t)B * H * W;
  const int64_t p0 = (int64_t)blockIdx.x * pix_per_block;
  const int64_t p1 = p0 + pix_per_block < total ? p0 + pix_per_block : total;
  float acc[7][8], accb[8];
#pragma unroll
  for (int s = 0; s < 7; s++)
#pragma unroll
    for (int k = 0; k < 8; k++) acc[s][k] = 0.f;
#pragma unroll
  for (int k = 0; k < 8; k++) accb[k] = 0.f;
  if (c0 < C) {
    for (int64_t p = p0 + pl; p < p1; p += 4) {
      const int wx = (int)(p % W);
      const int h = (int)((p / W) % H);
      const int b = (int)(p / ((int64_t)W * H));
      float d[8];
      load8(dy + p * C + c0, d);
      if (rr == 3) {
#pragma unroll
        for (int k = 0; k < 8; k++) accb[k] += d[k];
      }
      const int hh = h + rr - 3;
      if (hh < 0 || hh >= H) continue;
      const bf16* row = x + ((int64_t)(b * H + hh) * W) * C + c0;
#pragma unroll
      for (int s = 0; s < 7; s++) {
        const int ww = wx + s - 3;
        if (ww < 0 || ww >= W) continue;
        float v[8];
        load8(row + (int64_t)ww * C, v);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[s][k] = fmaf(d[k], v[k], acc[s][k]);
      }
    }
#pragma unroll
    for (int s = 0; s < 7; s++)
#pragma unroll
      for (int k = 0; k < 8; k++) atomicAdd(&red[rr * 7 + s][cv * 8 + k], acc[s][k]);
    if (rr == 3) {
#pragma unroll
      for (int k = 0; k < 8; k++) atomicAdd(&red[49][cv * 8 + k], accb[k]);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 50 * 64; i += blockDim.x) {
    const int tap = i >> 6, cl = i & 63;
    const int c = blockIdx.y * 64 + cl;
    if (c >= C) continue;
    const float v = red[tap][cl];
    if (tap < 49) atomicAdd(dw + (int64_t)c * 49 + tap, v);
    else if (db != nullptr) atomicAdd(db + c, v);
  }
}

// ------------------------------------------------------------------ layer scale + stochastic depth + residual
// keep(b) = 0 or 1/(1-p), one draw per sample ("row" mode of torchvision.ops.StochasticDepth), from the stateless
// generator shared with dropout: forward and backward of a step agree, graph replays draw fresh masks.
__device__ __forceinline__ float sd_keep(uint64_t seed, int b, float p, float inv_keep) {
  return p > 0.f ? dropout_scale(seed, (uint64_t)b * 4, p, inv_keep) : 1.f;
}

__global__ void __launch_bounds__(256) layer_scale_fwd_kernel(const bf16* __restrict__ x, const bf16* __restrict__ z,
                                                              const float* __restrict__ ls, bf16* __restrict__ out, int64_t rows,
                                                              int C, int rows_per_sample, float p, uint64_t seed) {
  const int cvec = C >> 3;
  const int64_t total = rows * cvec;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c0 = (int)(i % cvec) * 8;
    const int b = (int)((i / cvec) / rows_per_sample);
    const float keep = sd_keep(seed, b, p, inv_keep);
    float xv[8], zv[8], s[8];
    load8(x + i * 8, xv);
    load8(z + i * 8, zv);
    *reinterpret_cast<float4*>(s) = *reinterpret_cast<const float4*>(ls + c0);
    *reinterpret_cast<float4*>(s + 4) = *reinterpret_cast<const float4*>(ls + c0 + 4);
#pragma unroll
    for (int k = 0; k < 8; k++) xv[k] = fmaf(zv[k], s[k] * keep, xv[k]);
    store8(out + i * 8, xv);
  }
}

// dz = dy * ls[c] * keep(b);  dls[c] += sum_rows dy * z * keep(b).  Column-reduction layout of bn_bwd_reduce.
__global__ void __launch_bounds__(256) layer_scale_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ z,
                                                              const float* __restrict__ ls, bf16* __restrict__ dz,
                                                              float* __restrict__ dls, int64_t rows, int C, int rows_per_sample,
                                                              int rows_per_block, float p, uint64_t seed) {
  __shared__ float sh[8][256 + 8];
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cv * 8;
  const bool ok = c0 < C;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  float a[8], s[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    a[k] = 0.f;
    s[k] = ok ? ls[c0 + k] : 0.f;
  }
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = r0 + rows_per_block < rows ? r0 + rows_per_block : rows;
  if (ok) {
    for (int64_t r = r0 + rl; r < r1; r += 8) {
      const float keep = sd_keep(seed, (int)(r / rows_per_sample), p, inv_keep);
      float d[8], zv[8], o[8];
      load8(dy + r * C + c0, d);
      load8(z + r * C + c0, zv);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        o[k] = d[k] * s[k] * keep;
        a[k] = fmaf(d[k] * keep, zv[k], a[k]);
      }
      store8(dz + r * C + c0, o);
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) sh[rl][cv * 8 + k] = a[k];
  __syncthreads();
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < C && dls != nullptr) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; w8++) t += sh[w8][threadIdx.x];
    atomicAdd(dls + c, t);
  }
}

// ------------------------------------------------------------------ single-query attention
constexpr int SQ_MAXT = 1024;

// one CTA per sample: probs[b, t] = softmax_t(scale * q[b] . k[b, t]); out[b] = sum_t probs * v[b, t]  (fp32 out)
__global__ void __launch_bounds__(256) sq_attn_fwd_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k,
                                                          int64_t ldk, const bf16* __restrict__ v, int64_t ldv,
                                                          float* __restrict__ out, float* __restrict__ probs, int T, int D,
                                                          float scale) {
  __shared__ float sc[SQ_MAXT];
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const bf16* qb = q + (int64_t)b * ldq;
  for (int t = warp; t < T; t += nw) {
    const bf16* kt = k + ((int64_t)b * T + t) * ldk;
    float s = 0.f;
    for (int d = lane * 8; d < D; d += 256) {
      float a[8], c[8];
      load8(qb + d, a);
      load8(kt + d, c);
#pragma unroll
      for (int i = 0; i < 8; i++) s = fmaf(a[i], c[i], s);
    }
    s = warp_sum(s);
    if (lane == 0) sc[t] = s * scale;
  }
  __syncthreads();
  float m = -INFINITY;
  for (int t = threadIdx.x; t < T; t += blockDim.x) m = fmaxf(m, sc[t]);
  m = block_max(m, red);
  float e = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float w = __expf(sc[t] - m);
    sc[t] = w;
    e += w;
  }
  e = block_sum(e, red);
  __syncthreads();
  const float inv = 1.f / e;
  for (int t = threadIdx.x; t < T; t += blockDim.x) {
    const float pr = sc[t] * inv;
    sc[t] = pr;
    probs[(int64_t)b * T + t] = pr;
  }
  __syncthreads();
  for (int d = threadIdx.x * 8; d < D; d += blockDim.x * 8) {
    float o[8];
#pragma unroll
    for (int i = 0; i < 8; i++) o[i] = 0.f;
    for (int t = 0; t < T; t++) {
      float c[8];
      load8(v + ((int64_t)b * T + t) * ldv + d, c);
      const float pr = sc[t];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = fmaf(pr, c[i], o[i]);
    }
    float* op = out + (int64_t)b * D + d;
    *reinterpret_cast<float4*>(op) = make_float4(o[0], o[1], o[2], o[3]);
    *reinterpret_cast<float4*>(op + 4) = make_float4(o[4], o[5], o[6], o[7]);
  }
}

// backward: dq[b] = sum_t ds_t k_t, dk_t = ds_t q, dv_t = p_t dout, ds_t = scale * p_t (dout.v_t - sum_u p_u dout.v_u)
__global__ void __launch_bounds__(256) sq_attn_bwd_kernel(const bf16* __restrict__ q, int64_t ldq, const bf16* __restrict__ k,
                                                          int64_t ldk, const bf16* __restrict__ v, int64_t ldv,
                                                          const float* __restrict__ dout, const float* __restrict__ probs,
                                                          bf16* __restrict__ dq, int64_t lddq, bf16* __restrict__ dk, int64_t lddk,
                                                          bf16* __restrict__ dv, int64_t lddv, int T, int D, float scale) {
  __shared__ float ds[SQ_MAXT];
  __shared__ float pr[SQ_MAXT];
  __shared__ float red[32];
  const int b = blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = blockDim.x >> 5;
  const float* dob = dout + (int64_t)b * D;
  for (int t = warp; t < T; t += nw) {
    const bf16* vt = v + ((int64_t)b * T + t) * ldv;
    float s = 0.f;
    for (int d = lane * 8; d < D; d += 256) {
      float c[8];
      load8(vt + d, c);
#pragma unroll
      for (int i = 0; i < 8; i++) s = fmaf(dob[d + i], c[i], s);
    }
    s = warp_sum(s);
    if (lane == 0) {
      ds[t] = s;
      pr[t] = probs[(int64_t)b * T + t];
    }
  }
  __syncthreads();
  float dot = 0.f;
  for (int t = threadIdx.x; t < T; t += blockDim.x) dot = fmaf(pr[t], ds[t], dot);
  dot = block_sum(dot, red);
  __syncthreads();
  for (int t = threadIdx.x; t < T; t += blockDim.x) ds[t] = scale * pr[t] * (ds[t] - dot);
  __syncthreads();
  for (int d = threadIdx.x * 8; d < D; d += blockDim.x * 8) {
    float qv[8], dov[8], acc[8];
    load8(q + (int64_t)b * ldq + d, qv);
#pragma unroll
    for (int i = 0; i < 8; i++) {
      dov[i] = dob[d + i];
      acc[i] = 0.f;
    }
    for (int t = 0; t < T; t++) {
      float c[8], o1[8], o2[8];
      load8(k + ((int64_t)b * T + t) * ldk + d, c);
      const float s = ds[t], p_t = pr[t];
#pragma unroll
      for (int i = 0; i < 8; i++) {
        acc[i] = fmaf(s, c[i], acc[i]);
        o1[i] = s * qv[i];
        o2[i] = p_t * dov[i];
      }
      store8(dk + ((int64_t)b * T + t) * lddk + d, o1);
      store8(dv + ((int64_t)b * T + t) * lddv + d, o2);
    }
    store8(dq + (int64_t)b * lddq + d, acc);
  }
}

int grid_cap(int64_t items, int block) {
  int64_t g = (items + block - 1) / block;
  const int64_t cap = (int64_t)mdhs_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

extern "C" int mdhs_dwconv7_fwd(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int flip,
                                void* stream) {
  if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8)) return MDHS_ERR_ARG;
  const int64_t groups = (int64_t)B * H * ((W + 3) / 4);
  const dim3 grid((unsigned)((groups + 31) / 32), (unsigned)((C + 63) / 64));
  g_mdhs_launches++;
  if (flip) dwconv7_kernel<true><<<grid, 256, 0, ST(stream)>>>((const bf16*)x, w, nullptr, (bf16*)y, B, H, W, C);
  else dwconv7_kernel<false><<<grid, 256, 0, ST(stream)>>>((const bf16*)x, w, bias, (bf16*)y, B, H, W, C);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_dwconv7_wgrad(const void* x, const void* dy, float* dw, float* db, int B, int H, int W, int C, void* stream) {
  if (!x || !dy || !dw || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8)) return MDHS_ERR_ARG;
  const int slabs = (C + 63) / 64;
  const int64_t total = (int64_t)B * H * W;
  int chunks = (mdhs_num_sms() * 4 + slabs - 1) / slabs;
  if (chunks > total / 64) chunks = (int)(total / 64 > 0 ? total / 64 : 1);
  const int64_t ppb = (total + chunks - 1) / chunks;
  const dim3 grid((unsigned)((total + ppb - 1) / ppb), (unsigned)slabs);
  g_mdhs_launches++;
  dwconv7_wgrad_kernel<<<grid, 224, 0, ST(stream)>>>((const bf16*)x, (const bf16*)dy, dw, db, B, H, W, C, ppb);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_layer_scale_fwd(const void* x, const void* z, const float* ls, void* out, int64_t rows, int C,
                                    int rows_per_sample, float p, uint64_t seed, void* stream) {
  if (!x || !z || !ls || !out || rows <= 0 || C <= 0 || (C % 8) || rows_per_sample <= 0 || p < 0.f || p >= 1.f) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  layer_scale_fwd_kernel<<<grid_cap(rows * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)x, (const bf16*)z, ls, (bf16*)out, rows,
                                                                               C, rows_per_sample, p, seed);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_layer_scale_bwd(const void* dy, const void* z, const float* ls, void* dz, float* dls, int64_t rows, int C,
                                    int rows_per_sample, float p, uint64_t seed, void* stream) {
  if (!dy || !z || !ls || !dz || rows <= 0 || C <= 0 || (C % 8) || rows_per_sample <= 0 || p < 0.f || p >= 1.f) return MDHS_ERR_ARG;
  const int cslabs = (C + 255) / 256;
  int row_blocks = ((int64_t)mdhs_num_sms() * 8) / cslabs;
  if (row_blocks < 1) row_blocks = 1;
  int64_t rpb = (rows + row_blocks - 1) / row_blocks;
  rpb = ((rpb + 7) / 8) * 8;
  row_blocks = (int)((rows + rpb - 1) / rpb);
  g_mdhs_launches++;
  layer_scale_bwd_kernel<<<dim3(cslabs, row_blocks), 256, 0, ST(stream)>>>((const bf16*)dy, (const bf16*)z, ls, (bf16*)dz, dls, rows, C,
                                                                          rows_per_sample, (int)rpb, p, seed);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_sq_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, float* out,
                                float* probs, int B, int T, int D, float scale, void* stream) {
  if (!q || !k || !v || !out || !probs || B <= 0 || T <= 0 || T > SQ_MAXT || D <= 0 || (D % 8) || (ldq % 8) || (ldk % 8) || (ldv % 8))
    return MDHS_ERR_ARG;
  g_mdhs_launches++;
  sq_attn_fwd_kernel<<<B, 256, 0, ST(stream)>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, out, probs, T, D, scale);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_sq_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                const float* dout, const float* probs, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv,
                                int64_t lddv, int B, int T, int D, float scale, void* stream) {
  if (!q || !k || !v || !dout || !probs || !dq || !dk || !dv || B <= 0 || T <= 0 || T > SQ_MAXT || D <= 0 || (D % 8) || (ldq % 8) ||
      (ldk % 8) || (ldv % 8) || (lddq % 8) || (lddk % 8) || (lddv % 8))
    return MDHS_ERR_ARG;
  g_mdhs_launches++;
  sq_attn_bwd_kernel<<<B, 256, 0, ST(stream)>>>((const bf16*)q, ldq, (const bf16*)k, ldk, (const bf16*)v, ldv, dout, probs, (bf16*)dq,
                                               lddq, (bf16*)dk, lddk, (bf16*)dv, lddv, T, D, scale);
  MDHS_RETURN_LAST();
}
