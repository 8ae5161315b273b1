// LayerNorm and train-mode BatchNorm for the hot path: HBM-bound, 128-bit vectorised, warp-shuffle
// reductions, fp32 statistics.  Activations are bf16 [rows, C] (tokens x features / NHWC pixels x
// channels); parameters and their gradients are fp32.
#include <cstdio>
#include <cstdlib>
#include "common.cuh"
#include "../../include/mdhs_b200.h"

MDHS_DEFINE_SEED_TICK(norm)

extern int64_t g_mdhs_launches;

namespace {

// ------------------------------------------------------------------ LayerNorm
constexpr int LN_MAXCH = 8;  // C <= 8 * 256 = 2048

// One warp per row.  x may be bf16 or fp32 (XF32).  y = LN(x) * gamma + beta, optional dropout on y.
template <bool XF32, int NCH>
__global__ void __launch_bounds__(256) ln_fwd_kernel(const void* __restrict__ x_, int64_t ldx, const float* __restrict__ gamma,
                                                     const float* __restrict__ beta, bf16* __restrict__ y, int64_t ldy,
                                                     float* __restrict__ y32, float* __restrict__ mean_out,
                                                     float* __restrict__ rstd_out, int rows, int C, float eps, float drop_p,
                                                     uint64_t seed) {
  const int warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (warp >= rows) return;
  const int64_t row = warp;
  float v[NCH][8];
  const int nch = C >> 8;  // full 256-wide chunks
  const int rem = C & 255;
  float s = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    const int col = c * 256 + lane * 8;
    const bool ok = (c < nch) || (c == nch && lane * 8 < rem);
    if (ok) {
      if (XF32) {
        const float* xp = reinterpret_cast<const float*>(x_) + row * ldx + col;
        const float4 a = *reinterpret_cast<const float4*>(xp), b = *reinterpret_cast<const float4*>(xp + 4);
        v[c][0] = a.x; v[c][1] = a.y; v[c][2] = a.z; v[c][3] = a.w;
        v[c][4] = b.x; v[c][5] = b.y; v[c][6] = b.z; v[c][7] = b.w;
      } else {
        load8(reinterpret_cast<const bf16*>(x_) + row * ldx + col, v[c]);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) s += v[c][i];
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) v[c][i] = 0.f;
    }
  }
  const float mean = warp_sum(s) / (float)C;
  float q = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    const bool ok = (c < nch) || (c == nch && lane * 8 < rem);
    if (ok) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float d = v[c][i] - mean;
        q += d * d;
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / (float)C + eps);
  if (lane == 0) {
    if (mean_out) mean_out[row] = mean;
    if (rstd_out) rstd_out[row] = rstd;
  }
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    const int col = c * 256 + lane * 8;
    const bool ok = (c < nch) || (c == nch && lane * 8 < rem);
    if (ok) {
      float o[8], gm[8], bt[8];
      *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(gamma + col);
      *reinterpret_cast<float4*>(gm + 4) = *reinterpret_cast<const float4*>(gamma + col + 4);
      *reinterpret_cast<float4*>(bt) = *reinterpret_cast<const float4*>(beta + col);
      *reinterpret_cast<float4*>(bt + 4) = *reinterpret_cast<const float4*>(beta + col + 4);
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = (v[c][i] - mean) * rstd * gm[i] + bt[i];
      if (drop_p > 0.f) {
        dropout_apply4(seed, (uint64_t)row * C + col, drop_p, inv_keep, o);
        dropout_apply4(seed, (uint64_t)row * C + col + 4, drop_p, inv_keep, o + 4);
      }
      if (y) store8(y + row * ldy + col, o);
      if (y32) {
        float* yp = y32 + row * (int64_t)C + col;
        *reinterpret_cast<float4*>(yp) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(yp + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
    }
  }
}

// Backward, input gradient.  dy (bf16 or fp32) is first multiplied by the forward output-dropout mask (drop_p/seed), then
//   dx = rstd * (g*dy - mean(g*dy) - xhat * mean(g*dy*xhat)).
// dx is written as bf16 (dx) and optionally a second copy multiplied by another dropout mask
// (dx_drop: gradient of the dense branch whose output dropout used (drop2_p, seed2)).  One warp per row, no state
// carried across rows, so the kernel runs at full occupancy; the parameter gradients are a separate column reduction
// (ln_bwd_param_kernel) -- the previous single-kernel form held 4 x 24 floats per thread and ran at 12 % occupancy.
template <bool XF32, bool DYF32, int NCH>
__global__ void __launch_bounds__(256) ln_bwd_dx_kernel(const void* __restrict__ dy_, int64_t lddy, const void* __restrict__ x_,
                                                        int64_t ldx, const float* __restrict__ mean_in,
                                                        const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                                                        bf16* __restrict__ dx, int64_t lddx, bf16* __restrict__ dx_drop,
                                                        float* __restrict__ dx32, int rows, int C, float drop_p, uint64_t seed,
                                                        float drop2_p, uint64_t seed2) {
  const int lane = threadIdx.x & 31;
  const int64_t row = (blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 5;
  if (row >= rows) return;
  const int nch = C >> 8;
  const int rem = C & 255;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  const float inv_keep2 = drop2_p > 0.f ? 1.f / (1.f - drop2_p) : 1.f;
  const float mean = mean_in[row], rstd = rstd_in[row];
  float xh[NCH][8], g[NCH][8];
  float s1 = 0.f, s2 = 0.f;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    const int col = c * 256 + lane * 8;
    const bool ok = (c < nch) || (c == nch && lane * 8 < rem);
    if (ok) {
      float xv[8], dyv[8], gm[8];
      if (XF32) {
        const float* xp = reinterpret_cast<const float*>(x_) + row * ldx + col;
        *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(xp);
        *reinterpret_cast<float4*>(xv + 4) = *reinterpret_cast<const float4*>(xp + 4);
      } else {
        load8(reinterpret_cast<const bf16*>(x_) + row * ldx + col, xv);
      }
      if (DYF32) {
        const float* dp = reinterpret_cast<const float*>(dy_) + row * lddy + col;
        *reinterpret_cast<float4*>(dyv) = *reinterpret_cast<const float4*>(dp);
        *reinterpret_cast<float4*>(dyv + 4) = *reinterpret_cast<const float4*>(dp + 4);
      } else {
        load8(reinterpret_cast<const bf16*>(dy_) + row * lddy + col, dyv);
      }
      *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(gamma + col);
      *reinterpret_cast<float4*>(gm + 4) = *reinterpret_cast<const float4*>(gamma + col + 4);
      if (drop_p > 0.f) {
        dropout_apply4(seed, (uint64_t)row * C + col, drop_p, inv_keep, dyv);
        dropout_apply4(seed, (uint64_t)row * C + col + 4, drop_p, inv_keep, dyv + 4);
      }
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float h = (xv[i] - mean) * rstd;
        xh[c][i] = h;
        const float gd = dyv[i] * gm[i];
        g[c][i] = gd;
        s1 += gd;
        s2 = fmaf(gd, h, s2);
      }
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) xh[c][i] = g[c][i] = 0.f;
    }
  }
  s1 = warp_sum(s1) / (float)C;
  s2 = warp_sum(s2) / (float)C;
#pragma unroll
  for (int c = 0; c < NCH; c++) {
    const int col = c * 256 + lane * 8;
    const bool ok = (c < nch) || (c == nch && lane * 8 < rem);
    if (ok) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = rstd * (g[c][i] - s1 - xh[c][i] * s2);
      if (dx) store8(dx + row * lddx + col, o);
      if (dx32) {
        float* p = dx32 + row * (int64_t)C + col;
        *reinterpret_cast<float4*>(p) = make_float4(o[0], o[1], o[2], o[3]);
        *reinterpret_cast<float4*>(p + 4) = make_float4(o[4], o[5], o[6], o[7]);
      }
      if (dx_drop) {
        dropout_apply4(seed2, (uint64_t)row * C + col, drop2_p, inv_keep2, o);
        dropout_apply4(seed2, (uint64_t)row * C + col + 4, drop2_p, inv_keep2, o + 4);
        store8(dx_drop + row * lddx + col, o);
      }
    }
  }
}

// ---- narrow rows (C <= 128, bf16, no dropout: the ConvNeXt stage-1 / stem LayerNorms with C = 96 / 128 over 401408 rows).
// A warp per 192-byte row leaves 20 of 32 lanes idle and too few bytes in flight (ncu-less arithmetic: 154 MB in 95 us =
// 1.6 TB/s).  Here half a warp owns a row and keeps two rows in flight.
__device__ __forceinline__ float half_sum(float v, unsigned mask) {
#pragma unroll
  for (int o = 8; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
  return v;
}

__global__ void __launch_bounds__(256) ln_fwd_small_kernel(const bf16* __restrict__ x, int64_t ldx, const float* __restrict__ gamma,
                                                           const float* __restrict__ beta, bf16* __restrict__ y, int64_t ldy,
                                                           float* __restrict__ mean_out, float* __restrict__ rstd_out, int rows,
                                                           int C, float eps) {
  const int sub = threadIdx.x & 15;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);
  const int64_t row0 = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4) * 2;
  if (row0 >= rows) return;   // whole half-warps leave together
  const int col = sub * 8;
  const bool ok = col < C;
  const bool two = row0 + 1 < rows;
  float v[2][8];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    if (ok && (r == 0 || two)) load8(x + (row0 + r) * ldx + col, v[r]);
    else {
#pragma unroll
      for (int i = 0; i < 8; i++) v[r][i] = 0.f;
    }
  }
  float gm[8], bt[8];
  if (ok) {
    *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(gamma + col);
    *reinterpret_cast<float4*>(gm + 4) = *reinterpret_cast<const float4*>(gamma + col + 4);
    *reinterpret_cast<float4*>(bt) = *reinterpret_cast<const float4*>(beta + col);
    *reinterpret_cast<float4*>(bt + 4) = *reinterpret_cast<const float4*>(beta + col + 4);
  }
  const float inv_c = 1.f / (float)C;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    float s = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) s += v[r][i];
    const float mean = half_sum(s, hmask) * inv_c;
    float q = 0.f;
    if (ok) {
#pragma unroll
      for (int i = 0; i < 8; i++) {
        const float d = v[r][i] - mean;
        q += d * d;
      }
    }
    const float rstd = rsqrtf(half_sum(q, hmask) * inv_c + eps);
    if (r == 1 && !two) break;
    if (sub == 0) {
      if (mean_out) mean_out[row0 + r] = mean;
      if (rstd_out) rstd_out[row0 + r] = rstd;
    }
    if (ok) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = (v[r][i] - mean) * rstd * gm[i] + bt[i];
      store8(y + (row0 + r) * ldy + col, o);
    }
  }
}

__global__ void __launch_bounds__(256) ln_bwd_dx_small_kernel(const bf16* __restrict__ dy, int64_t lddy, const bf16* __restrict__ x,
                                                              int64_t ldx, const float* __restrict__ mean_in,
                                                              const float* __restrict__ rstd_in, const float* __restrict__ gamma,
                                                              bf16* __restrict__ dx, int64_t lddx, int rows, int C) {
  const int sub = threadIdx.x & 15;
  const unsigned hmask = 0xffffu << (threadIdx.x & 16);
  const int64_t row0 = ((blockIdx.x * (int64_t)blockDim.x + threadIdx.x) >> 4) * 2;
  if (row0 >= rows) return;
  const int col = sub * 8;
  const bool ok = col < C;
  const bool two = row0 + 1 < rows;
  float xv[2][8], dv[2][8];
#pragma unroll
  for (int r = 0; r < 2; r++) {
    if (ok && (r == 0 || two)) {
      load8(x + (row0 + r) * ldx + col, xv[r]);
      load8(dy + (row0 + r) * lddy + col, dv[r]);
    } else {
#pragma unroll
      for (int i = 0; i < 8; i++) xv[r][i] = dv[r][i] = 0.f;
    }
  }
  float gm[8];
  if (ok) {
    *reinterpret_cast<float4*>(gm) = *reinterpret_cast<const float4*>(gamma + col);
    *reinterpret_cast<float4*>(gm + 4) = *reinterpret_cast<const float4*>(gamma + col + 4);
  } else {
#pragma unroll
    for (int i = 0; i < 8; i++) gm[i] = 0.f;
  }
  const float inv_c = 1.f / (float)C;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const bool live = r == 0 || two;
    const float mean = live ? mean_in[row0 + r] : 0.f, rstd = live ? rstd_in[row0 + r] : 0.f;
    float xh[8], g[8], s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int i = 0; i < 8; i++) {
      xh[i] = ok ? (xv[r][i] - mean) * rstd : 0.f;
      g[i] = dv[r][i] * gm[i];
      s1 += g[i];
      s2 = fmaf(g[i], xh[i], s2);
    }
    s1 = half_sum(s1, hmask) * inv_c;
    s2 = half_sum(s2, hmask) * inv_c;
    if (live && ok) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; i++) o[i] = rstd * (g[i] - s1 - xh[i] * s2);
      store8(dx + (row0 + r) * lddx + col, o);
    }
  }
}

// Backward, parameter gradients: dgamma[c] += sum_r dy'[r,c] * xhat[r,c], dbeta[c] += sum_r dy'[r,c] (dy' = dy times the
// forward output-dropout mask).  Column reduction with the bn_bwd_reduce thread layout: block = 32 channel vectors(8) x 8
// row lanes over a 256-column slab, grid = (slabs, row blocks), fp32 atomics at the end.
template <bool XF32, bool DYF32, int CVW>
__global__ void __launch_bounds__(256) ln_bwd_param_kernel(const void* __restrict__ dy_, int64_t lddy, const void* __restrict__ x_,
                                                           int64_t ldx, const float* __restrict__ mean_in,
                                                           const float* __restrict__ rstd_in, float* __restrict__ dgamma,
                                                           float* __restrict__ dbeta, int rows, int C, int rows_per_block,
                                                           float drop_p, uint64_t seed, const bf16* __restrict__ bsrc,
                                                           int64_t ldb, float* __restrict__ dbias) {
  // bsrc / dbias (optional): dbias[c] += sum_r bsrc[r, c] -- the bias gradient of the dense layer that fed this LayerNorm's
  // residual sum (its output gradient is the dx / dx_drop that ln_bwd_dx just wrote): one more bf16 read here instead of a
  // separate column-sum launch (24 of the 61 col_stats launches of a BERT-base step).
  // CVW channel vectors x (256 / CVW) row lanes: 32 x 8 in general, 16 x 16 for C <= 128 (C = 96: 12 of 16 lanes busy
  // instead of 12 of 32)
  constexpr int RL = 256 / CVW, SLAB = CVW * 8;
  __shared__ float sh[3][RL][SLAB + 8];
  const int cv = threadIdx.x % CVW, rl = threadIdx.x / CVW;
  const int c0 = blockIdx.x * SLAB + cv * 8;
  const bool ok = c0 < C;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  float a[8], b[8], cb[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = b[k] = cb[k] = 0.f;
  const int r0 = blockIdx.y * rows_per_block;
  const int r1 = min(rows, r0 + rows_per_block);
  if (ok) {
#pragma unroll 2
    for (int r = r0 + rl; r < r1; r += RL) {
      float xv[8], dyv[8];
      if (bsrc) {
        float bv[8];
        load8(bsrc + (int64_t)r * ldb + c0, bv);
#pragma unroll
        for (int k = 0; k < 8; k++) cb[k] += bv[k];
      }
      if (XF32) {
        const float* xp = reinterpret_cast<const float*>(x_) + (int64_t)r * ldx + c0;
        *reinterpret_cast<float4*>(xv) = *reinterpret_cast<const float4*>(xp);
        *reinterpret_cast<float4*>(xv + 4) = *reinterpret_cast<const float4*>(xp + 4);
      } else {
        load8(reinterpret_cast<const bf16*>(x_) + (int64_t)r * ldx + c0, xv);
      }
      if (DYF32) {
        const float* dp = reinterpret_cast<const float*>(dy_) + (int64_t)r * lddy + c0;
        *reinterpret_cast<float4*>(dyv) = *reinterpret_cast<const float4*>(dp);
        *reinterpret_cast<float4*>(dyv + 4) = *reinterpret_cast<const float4*>(dp + 4);
      } else {
        load8(reinterpret_cast<const bf16*>(dy_) + (int64_t)r * lddy + c0, dyv);
      }
      if (drop_p > 0.f) {
        dropout_apply4(seed, (uint64_t)r * C + c0, drop_p, inv_keep, dyv);
        dropout_apply4(seed, (uint64_t)r * C + c0 + 4, drop_p, inv_keep, dyv + 4);
      }
      const float mu = mean_in[r], rs = rstd_in[r];
#pragma unroll
      for (int k = 0; k < 8; k++) {
        a[k] = fmaf(dyv[k], (xv[k] - mu) * rs, a[k]);
        b[k] += dyv[k];
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    sh[0][rl][cv * 8 + k] = a[k];
    sh[1][rl][cv * 8 + k] = b[k];
    sh[2][rl][cv * 8 + k] = cb[k];
  }
  __syncthreads();
  const int c = blockIdx.x * SLAB + threadIdx.x;
  if (threadIdx.x < SLAB && c < C) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int w = 0; w < RL; w++) {
      s0 += sh[0][w][threadIdx.x];
      s1 += sh[1][w][threadIdx.x];
      s2 += sh[2][w][threadIdx.x];
    }
    if (dgamma) atomicAdd(dgamma + c, s0);
    if (dbeta) atomicAdd(dbeta + c, s1);
    if (dbias) atomicAdd(dbias + c, s2);
  }
}

// ------------------------------------------------------------------ BatchNorm (NHWC, train mode)
// Finalise batch statistics from fp64 column sums (written by the GEMM epilogue or col_stats):
// mean / invstd, running-stat update (momentum, unbiased variance) and the fused scale/shift pair.
__global__ void bn_finalize_kernel(const double* __restrict__ colsum, const double* __restrict__ colsumsq, int64_t count,
                                   const float* __restrict__ gamma, const float* __restrict__ beta,
                                   float* __restrict__ running_mean, float* __restrict__ running_var, float momentum,
                                   float eps, float* __restrict__ mean, float* __restrict__ invstd, float* __restrict__ scale,
                                   float* __restrict__ shift, int C, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  float mu, var;
  if (training) {
    const double inv_count = 1.0 / (double)count;
    const double m = colsum[c] * inv_count;
    double v = colsumsq[c] * inv_count - m * m;
    if (v < 0.0) v = 0.0;
    mu = (float)m;
    var = (float)v;
    if (running_mean) {
      const double unbiased = count > 1 ? v * (double)count / (double)(count - 1) : v;
      // explicit fma form: identical rounding in bn_finalize_kernel and bn_fwd_kernel
      running_mean[c] = fmaf(momentum, mu, (1.f - momentum) * running_mean[c]);
      running_var[c] = fmaf(momentum, (float)unbiased, (1.f - momentum) * running_var[c]);
    }
  } else {
    mu = running_mean[c];
    var = running_var[c];
  }
  const float is = rsqrtf(var + eps);
  if (mean) mean[c] = mu;
  if (invstd) invstd[c] = is;
  const float sc = gamma[c] * is;
  scale[c] = sc;
  shift[c] = beta[c] - mu * sc;
}

// 8 consecutive per-channel values with 128-bit loads (lanes of a warp read consecutive 32-byte / 64-byte pieces: fully
// coalesced).  The scalar form -- 8 separate 4-byte loads at a 32-byte lane stride per array -- cost 8 L1 wavefronts per
// load instruction and made the per-thread coefficient prologue, not HBM, the floor of the fused kernels (30 us on a 6 MB layer).
__device__ __forceinline__ void ldvec8(const float* __restrict__ p, float* out) {
  *reinterpret_cast<float4*>(out) = __ldg(reinterpret_cast<const float4*>(p));
  *reinterpret_cast<float4*>(out + 4) = __ldg(reinterpret_cast<const float4*>(p + 4));
}
__device__ __forceinline__ void ldvec8(const double* __restrict__ p, double* out) {
#pragma unroll
  for (int i = 0; i < 4; i++) *reinterpret_cast<double2*>(out + 2 * i) = __ldg(reinterpret_cast<const double2*>(p + 2 * i));
}
__device__ __forceinline__ void stvec8(float* __restrict__ p, const float* v) {
  *reinterpret_cast<float4*>(p) = *reinterpret_cast<const float4*>(v);
  *reinterpret_cast<float4*>(p + 4) = *reinterpret_cast<const float4*>(v + 4);
}

// Shared streaming body of the two forward kernels: y = act(x * sc + sh (+ residual)), 8 channels per thread held in
// registers (the grid stride is a multiple of C/8), two independent 16-byte streams in flight per thread.
// mask_out (optional, relu only): one byte per 8-channel vector, bit k = (output k > 0) -- the backward of a layer WITH a
// residual input reads this 1-bit/element mask instead of the whole bf16 output (2 bytes/element, twice).
__device__ __forceinline__ uint8_t relu_bits(const float* v) {
  uint32_t m = 0;
#pragma unroll
  for (int k = 0; k < 8; k++) m |= (v[k] > 0.f ? 1u : 0u) << k;
  return (uint8_t)m;
}
template <bool MASK>
__device__ __forceinline__ void bn_apply_stream(const bf16* __restrict__ x, const bf16* __restrict__ residual,
                                                bf16* __restrict__ y, const float* sc, const float* sh, int64_t i0,
                                                int64_t stride, int64_t total_vec, int relu, uint8_t* __restrict__ mask_out) {
  int64_t i = i0;
  for (; i + stride < total_vec; i += 2 * stride) {
    float v0[8], v1[8], r0[8], r1[8];
    const bf16x8 q0 = ldraw8(x + i * 8), q1 = ldraw8(x + (i + stride) * 8);
    bf16x8 p0, p1;
    if (residual) {
      p0 = ldraw8(residual + i * 8);
      p1 = ldraw8(residual + (i + stride) * 8);
    }
    cvt8(q0, v0);
    cvt8(q1, v1);
    if (residual) {
      cvt8(p0, r0);
      cvt8(p1, r1);
    }
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float o0 = fmaf(v0[k], sc[k], sh[k]);   // the backward recomputes the ReLU mask from exactly this expression
      float o1 = fmaf(v1[k], sc[k], sh[k]);
      if (residual) {
        o0 += r0[k];
        o1 += r1[k];
      }
      if (relu) {
        o0 = fmaxf(o0, 0.f);
        o1 = fmaxf(o1, 0.f);
      }
      v0[k] = o0;
      v1[k] = o1;
    }
    store8(y + i * 8, v0);
    store8(y + (i + stride) * 8, v1);
    if (MASK) {
      mask_out[i] = relu_bits(v0);
      mask_out[i + stride] = relu_bits(v1);
    }
  }
  if (i < total_vec) {
    float v[8], r[8];
    load8(x + i * 8, v);
    if (residual) load8(residual + i * 8, r);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float o = fmaf(v[k], sc[k], sh[k]);
      if (residual) o += r[k];
      if (relu) o = fmaxf(o, 0.f);
      v[k] = o;
    }
    store8(y + i * 8, v);
    if (MASK) mask_out[i] = relu_bits(v);
  }
}

// y = act(x * scale[c] + shift[c] (+ residual)); 8 channels per thread, grid-stride over rows*C/8.  The host sizes the
// grid so that the stride is a multiple of C/8: a thread then stays on ONE channel vector and its scale / shift live in
// registers (per-iteration parameter loads made this kernel L1-bound: ncu l1tex 89 % at 65 % of HBM peak).
template <bool MASK>
__global__ void __launch_bounds__(256, 3) bn_apply_kernel(const bf16* __restrict__ x, const float* __restrict__ scale,
                                                       const float* __restrict__ shift, const bf16* __restrict__ residual,
                                                       bf16* __restrict__ y, int64_t total_vec, int C, int relu,
                                                       uint8_t* __restrict__ mask_out) {
  const int cvec = C >> 3;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int c0 = (int)(i0 % cvec) * 8;
  float sc[8], sh[8];
  *reinterpret_cast<float4*>(sc) = *reinterpret_cast<const float4*>(scale + c0);
  *reinterpret_cast<float4*>(sc + 4) = *reinterpret_cast<const float4*>(scale + c0 + 4);
  *reinterpret_cast<float4*>(sh) = *reinterpret_cast<const float4*>(shift + c0);
  *reinterpret_cast<float4*>(sh + 4) = *reinterpret_cast<const float4*>(shift + c0 + 4);
  bn_apply_stream<MASK>(x, residual, y, sc, sh, i0, stride, total_vec, relu, mask_out);
}

// Per-channel reductions for BN backward: sum(dy') and sum(dy' * (x - mean)), dy' = dy * [y > 0] when relu.  The ReLU
// mask comes from y when the layer had a residual input, otherwise it is recomputed from x (fmaf(x, scale, shift) > 0 is
// exactly what the forward evaluated), which saves one full read of y.
// Block = 32 channel-vectors(8) x RL row lanes over a `Cw`-wide view of the matrix: Cw = C, or 256 when C in {64, 128}
// (the contiguous [rows, C] matrix re-read as [rows*C/256, 256]; column j holds channel j % C), so narrow layers keep all
// 32 vector lanes busy.  grid = (Cw/256 ceil, row blocks).
// Round 2: (i) four rows per thread are requested before any is consumed (8-12 independent 16-byte loads in flight),
// (ii) the grid is ~2 fat blocks per SM instead of 8 thin ones -- every block ends with 512 fp64 atomics on the same
// few cache lines, and with 1046 blocks those 535 k same-address atomics, not HBM, paced the small layers (27 us for a
// 25 MB layer).  The (x - mean) centring happens here; the 1/sigma factor is applied by the consumer.
template <int RL>
__global__ void __launch_bounds__(32 * RL, 512 / (32 * RL)) bn_bwd_reduce_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                            const bf16* __restrict__ y, const uint8_t* __restrict__ mbits,
                                                            const float* __restrict__ mean,
                                                            const float* __restrict__ scale, const float* __restrict__ shift,
                                                            double* __restrict__ sum_dy, double* __restrict__ sum_dy_xc,
                                                            int64_t rows, int C, int Cw, int rows_per_block, int relu) {
  __shared__ float sh[2][8][256 + 8];
  const int cv = threadIdx.x & 31;   // which 8-channel vector inside the 256-column slab
  const int rl = threadIdx.x >> 5;   // row lane 0..RL-1
  const int c0 = blockIdx.x * 256 + cv * 8;
  const bool ok = c0 < Cw;
  const int ch0 = c0 % C;
  const bool use_bits = relu && (mbits != nullptr);          // 1-bit/element ReLU mask written by the forward
  const bool remask = relu && !use_bits && (y == nullptr);
  const bool use_y = relu && !use_bits && !remask;
  const int cw8 = Cw >> 3, cv0 = c0 >> 3;
  float a[8], b[8], mu[8], sc[8], sf[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    a[k] = b[k] = 0.f;
    mu[k] = ok ? mean[ch0 + k] : 0.f;
    sc[k] = (ok && remask) ? scale[ch0 + k] : 0.f;
    sf[k] = (ok && remask) ? shift[ch0 + k] : 0.f;
  }
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  auto accum = [&](const float* d, const float* xv, const float* yv, uint32_t bits) {
#pragma unroll
    for (int k = 0; k < 8; k++) {
      bool on = true;
      if (use_bits) on = (bits >> k) & 1u;
      else if (relu) on = (remask ? fmaf(xv[k], sc[k], sf[k]) : yv[k]) > 0.f;
      const float dd = on ? d[k] : 0.f;
      a[k] += dd;
      b[k] = fmaf(dd, xv[k] - mu[k], b[k]);
    }
  };
  if (ok) {
    constexpr int U = 4;
    int64_t r = r0 + rl;
    for (; r + (U - 1) * RL < r1; r += U * RL) {
      bf16x8 rd[U], rx[U], ry[U];
      uint32_t mb[U];
#pragma unroll
      for (int u = 0; u < U; u++) {
        rd[u] = ldraw8(dy + (r + u * RL) * Cw + c0);
        rx[u] = ldraw8(x + (r + u * RL) * Cw + c0);
        if (use_y) ry[u] = ldraw8(y + (r + u * RL) * Cw + c0);
        mb[u] = use_bits ? mbits[(r + u * RL) * cw8 + cv0] : 0u;
      }
#pragma unroll
      for (int u = 0; u < U; u++) {
        float d[8], xv[8], yv[8];
        cvt8(rd[u], d);
        cvt8(rx[u], xv);
        if (use_y) cvt8(ry[u], yv);
        accum(d, xv, yv, mb[u]);
      }
    }
    for (; r < r1; r += RL) {
      float d[8], xv[8], yv[8];
      load8(dy + r * Cw + c0, d);
      load8(x + r * Cw + c0, xv);
      if (use_y) load8(y + r * Cw + c0, yv);
      accum(d, xv, yv, use_bits ? mbits[r * cw8 + cv0] : 0u);
    }
  }
  // cross-lane reduction through shared memory, 8 row lanes per round
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int round = 0; round < RL / 8; round++) {
    if ((rl >> 3) == round) {
#pragma unroll
      for (int k = 0; k < 8; k++) {
        sh[0][rl & 7][cv * 8 + k] = a[k];
        sh[1][rl & 7][cv * 8 + k] = b[k];
      }
    }
    __syncthreads();
    if (threadIdx.x < 256) {
#pragma unroll
      for (int w = 0; w < 8; w++) {
        s0 += sh[0][w][threadIdx.x];
        s1 += sh[1][w][threadIdx.x];
      }
    }
    __syncthreads();
  }
  if (Cw != C) {   // folded view: columns t, t + C, t + 2C, ... belong to channel t
    if (threadIdx.x < 256) {
      sh[0][0][threadIdx.x] = s0;
      sh[1][0][threadIdx.x] = s1;
    }
    __syncthreads();
    if (threadIdx.x < C) {
      s0 = s1 = 0.f;
      for (int j = threadIdx.x; j < 256; j += C) {
        s0 += sh[0][0][j];
        s1 += sh[1][0][j];
      }
    }
  }
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (threadIdx.x < 256 && c < C) {
    atomicAdd(sum_dy + c, (double)s0);
    atomicAdd(sum_dy_xc + c, (double)s1);
  }
}

// Per-channel coefficients of the BN input gradient, from the fp64 reductions sum(dy') and sum(dy' * (x - mean)):
//   train: dx = gamma*invstd * (dy' - sum_dy/M - xhat * sum_dy_xhat/M) = ca * dy' + cb * x + cc
//   eval : dx = gamma*invstd * dy'   (running statistics are constants: F.batch_norm(training=False) backward)
// Also accumulates dgamma += sum(dy' * xhat), dbeta += sum(dy') (one thread per channel).
// (Round 2 tried deriving these inside the apply kernel, per thread for its 8 channels: the fp64 prologue of ~600 k threads
//  cost 20-30 us per launch -- profiles/r02_bn_microbench.md -- against ~4 us for this launch, so it stays separate.)
__global__ void bn_bwd_coeff_kernel(const float* __restrict__ mean, const float* __restrict__ invstd,
                                    const float* __restrict__ gamma, const double* __restrict__ sum_dy,
                                    const double* __restrict__ sum_dy_xc, const float* __restrict__ scale,
                                    const float* __restrict__ shift, float* __restrict__ coef, float* __restrict__ dgamma,
                                    float* __restrict__ dbeta, int64_t rows, int C, int training) {
  const int c = blockIdx.x * blockDim.x + threadIdx.x;
  if (c >= C) return;
  const double inv_m = 1.0 / (double)rows;
  const double is = invstd[c], mu = mean[c], g = gamma[c];
  const double s1 = sum_dy[c], s2 = sum_dy_xc[c] * is;    // s2 = sum(dy' * xhat)
  const double ca = g * is;
  const double cb = training ? -g * is * is * s2 * inv_m : 0.0;
  coef[c] = (float)ca;
  coef[C + c] = (float)cb;
  coef[2 * C + c] = training ? (float)(-ca * s1 * inv_m - cb * mu) : 0.f;
  if (scale) {
    coef[3 * C + c] = scale[c];
    coef[4 * C + c] = shift[c];
  }
  if (dgamma) {
    dgamma[c] += (float)s2;
    dbeta[c] += (float)s1;
  }
}

// dx = ca[c] * dy' + cb[c] * x + cc[c]; optionally also writes dy' (the ReLU-masked incoming gradient) for the
// identity branch.  Pure streaming, 8 channels per thread; the grid stride is a multiple of C/8 (see bn_apply), so the
// five coefficient vectors of a thread's channel group are loaded once (128-bit loads).  coef rows 3 and 4 hold the forward
// scale / shift when the ReLU mask is recomputed from x (y == nullptr).
// MODE: where the ReLU mask comes from -- 0 no ReLU, 1 recomputed from (x, scale, shift), 2 the bf16 output y, 3 the forward's
// bit mask.  (Compile-time: with run-time flags every launch carried the code and registers of all four paths.)
template <int MODE, bool DZ>
__global__ void __launch_bounds__(256, 3) bn_bwd_apply_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ x,
                                                              const bf16* __restrict__ y, const uint8_t* __restrict__ mbits,
                                                              const float* __restrict__ coef, bf16* __restrict__ dx,
                                                              bf16* __restrict__ dz_, int64_t rows, int C) {
  const int cvec = C >> 3;
  const int64_t total_vec = rows * cvec;
  constexpr bool relu = MODE != 0, use_bits = MODE == 3, remask = MODE == 1, use_y = MODE == 2;
  bf16* dz = DZ ? dz_ : nullptr;
  const int64_t i0 = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t stride = (int64_t)gridDim.x * blockDim.x;
  const int c0 = (int)(i0 % cvec) * 8;
  float ca[8], cb[8], cc[8], sc[8], sf[8];
  ldvec8(coef + c0, ca);
  ldvec8(coef + C + c0, cb);
  ldvec8(coef + 2 * C + c0, cc);
  if (remask) {
    ldvec8(coef + 3 * C + c0, sc);
    ldvec8(coef + 4 * C + c0, sf);
  }
  int64_t i = i0;
  for (; i + stride < total_vec; i += 2 * stride) {   // two independent row groups in flight per thread
    bf16x8 rd[2], rx[2], ry[2];
    uint32_t mb[2];
#pragma unroll
    for (int u = 0; u < 2; u++) {
      rd[u] = ldraw8(dy + (i + u * stride) * 8);
      rx[u] = ldraw8(x + (i + u * stride) * 8);
      if (use_y) ry[u] = ldraw8(y + (i + u * stride) * 8);
      mb[u] = use_bits ? mbits[i + u * stride] : 0u;
    }
#pragma unroll
    for (int u = 0; u < 2; u++) {
      float d[8], xv[8], yv[8], o[8];
      cvt8(rd[u], d);
      cvt8(rx[u], xv);
      if (use_y) cvt8(ry[u], yv);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        bool on = true;
        if (use_bits) on = (mb[u] >> k) & 1u;
        else if (relu) on = (remask ? fmaf(xv[k], sc[k], sf[k]) : yv[k]) > 0.f;
        const float dd = on ? d[k] : 0.f;
        d[k] = dd;
        o[k] = fmaf(ca[k], dd, fmaf(cb[k], xv[k], cc[k]));
      }
      store8(dx + (i + u * stride) * 8, o);
      if (dz) store8(dz + (i + u * stride) * 8, d);
    }
  }
  if (i < total_vec) {
    float d[8], xv[8], yv[8], o[8];
    load8(dy + i * 8, d);
    load8(x + i * 8, xv);
    if (use_y) load8(y + i * 8, yv);
    const uint32_t bits = use_bits ? mbits[i] : 0u;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      bool on = true;
      if (use_bits) on = (bits >> k) & 1u;
      else if (relu) on = (remask ? fmaf(xv[k], sc[k], sf[k]) : yv[k]) > 0.f;
      const float dd = on ? d[k] : 0.f;
      d[k] = dd;
      o[k] = fmaf(ca[k], dd, fmaf(cb[k], xv[k], cc[k]));
    }
    store8(dx + i * 8, o);
    if (dz) store8(dz + i * 8, d);
  }
}

// Per-column sum / sum of squares of a bf16 [rows, C] matrix into fp64 (standalone BN statistics) or
// fp32 (+=, bias gradients).  Same thread layout (and the same folded Cw-wide view for C in {64, 128}) as bn_bwd_reduce.
__global__ void __launch_bounds__(256) col_stats_kernel(const bf16* __restrict__ x, int64_t ldx, double* __restrict__ sum64,
                                                        double* __restrict__ sumsq64, float* __restrict__ sum32,
                                                        int64_t rows, int C, int Cw, int rows_per_block) {
  __shared__ float sh[2][8][256 + 8];
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cv * 8;
  const bool ok = c0 < Cw;
  float a[8], b[8];
#pragma unroll
  for (int k = 0; k < 8; k++) a[k] = b[k] = 0.f;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(rows, r0 + rows_per_block);
  if (ok) {
    int64_t r = r0 + rl;
    for (; r + 24 < r1; r += 32) {
      float v[4][8];
#pragma unroll
      for (int u = 0; u < 4; u++) load8(x + (r + 8 * u) * ldx + c0, v[u]);
#pragma unroll
      for (int u = 0; u < 4; u++)
#pragma unroll
        for (int k = 0; k < 8; k++) {
          a[k] += v[u][k];
          b[k] = fmaf(v[u][k], v[u][k], b[k]);
        }
    }
    for (; r < r1; r += 8) {
      float v[8];
      load8(x + r * ldx + c0, v);
#pragma unroll
      for (int k = 0; k < 8; k++) {
        a[k] += v[k];
        b[k] = fmaf(v[k], v[k], b[k]);
      }
    }
  }
#pragma unroll
  for (int k = 0; k < 8; k++) {
    sh[0][rl][cv * 8 + k] = a[k];
    sh[1][rl][cv * 8 + k] = b[k];
  }
  __syncthreads();
  float s0 = 0.f, s1 = 0.f;
#pragma unroll
  for (int w = 0; w < 8; w++) {
    s0 += sh[0][w][threadIdx.x];
    s1 += sh[1][w][threadIdx.x];
  }
  if (Cw != C) {
    __syncthreads();
    sh[0][0][threadIdx.x] = s0;
    sh[1][0][threadIdx.x] = s1;
    __syncthreads();
    if (threadIdx.x < C) {
      s0 = s1 = 0.f;
      for (int j = threadIdx.x; j < 256; j += C) {
        s0 += sh[0][0][j];
        s1 += sh[1][0][j];
      }
    }
  }
  const int c = blockIdx.x * 256 + threadIdx.x;
  if (c < C) {
    if (sum64) atomicAdd(sum64 + c, (double)s0);
    if (sumsq64) atomicAdd(sumsq64 + c, (double)s1);
    if (sum32) atomicAdd(sum32 + c, s0);
  }
}

// Folded view for narrow matrices: [rows, C] contiguous with 256 % C == 0 is re-read as [rows*C/256, 256].
static inline bool fold_view(int64_t rows, int C, int64_t ld, int64_t* rows_w, int* Cw) {
  if (C < 256 && (256 % C) == 0 && ld == C && (rows % (256 / C)) == 0) {
    *rows_w = rows / (256 / C);
    *Cw = 256;
    return true;
  }
  *rows_w = rows;
  *Cw = C;
  return false;
}

// Grid for the channel-vector streaming kernels: (grid * 256) % (C / 8) == 0, close to 16 blocks per SM.
int grid_for_channels(int64_t total_vec, int C, int vec_per_thread = 1) {
  const int cvec = C >> 3;
  int a = cvec, b = 256;
  while (b) {
    const int t = a % b;
    a = b;
    b = t;
  }
  const int unit = cvec / a;                     // grid must be a multiple of cvec / gcd(cvec, 256)
  // vec_per_thread > 1: kernels with a per-thread coefficient prologue want it amortised over a few vectors
  int64_t g = (total_vec + 256 * vec_per_thread - 1) / (256 * vec_per_thread);
  const int64_t cap = (int64_t)mdhs_num_sms() * 16;
  if (g > cap) g = cap;
  g = (g / unit) * unit;
  if (g < unit) g = unit;
  return (int)g;
}

int grid_for(int64_t work_items, int block) {
  int64_t g = (work_items + block - 1) / block;
  const int64_t cap = (int64_t)mdhs_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

extern "C" int mdhs_layernorm_fwd(const void* x, int x_f32, int64_t ldx, const float* gamma, const float* beta, void* y_bf16,
                                  int64_t ldy, float* y_f32, float* mean, float* rstd, int rows, int C, float eps,
                                  float drop_p, uint64_t seed, void* stream) {
  if (!x || !gamma || !beta || rows <= 0 || C <= 0 || (C % 8) || C > LN_MAXCH * 256) return MDHS_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = ceil_div(rows, 8);
  g_mdhs_launches++;
#define LNF(XF, N) ln_fwd_kernel<XF, N><<<blocks, 256, 0, st>>>(x, ldx, gamma, beta, (bf16*)y_bf16, ldy, y_f32, mean, rstd, rows, C, eps, drop_p, seed)
#define LNF_N(XF)                      \
  do {                                 \
    if (C <= 256) LNF(XF, 1);          \
    else if (C <= 512) LNF(XF, 2);     \
    else if (C <= 768) LNF(XF, 3);     \
    else if (C <= 1024) LNF(XF, 4);    \
    else LNF(XF, 8);                   \
  } while (0)
  if (!x_f32 && C <= 128 && y_bf16 && !y_f32 && drop_p <= 0.f) {
    ln_fwd_small_kernel<<<ceil_div(rows, 32), 256, 0, st>>>((const bf16*)x, ldx, gamma, beta, (bf16*)y_bf16, ldy, mean, rstd, rows, C,
                                                            eps);
    MDHS_RETURN_LAST();
  }
  if (x_f32) LNF_N(true); else LNF_N(false);
#undef LNF_N
#undef LNF
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_layernorm_bwd(const void* dy, int dy_f32, int64_t lddy, const void* x, int x_f32, int64_t ldx,
                                  const float* mean, const float* rstd, const float* gamma, void* dx_bf16, int64_t lddx,
                                  void* dx_drop_bf16, float* dx_f32, float* dgamma, float* dbeta, float* dbias, int rows, int C,
                                  float drop_p, uint64_t seed, float drop2_p, uint64_t seed2, void* stream) {
  if (!dy || !x || !mean || !rstd || !gamma || rows <= 0 || (C % 8) || C > LN_MAXCH * 256) return MDHS_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int blocks = ceil_div(rows, 8);
  // dbias: column sums of the gradient handed to the dense branch (dx_drop when given, else dx)
  const bf16* bsrc = dbias ? (const bf16*)(dx_drop_bf16 ? dx_drop_bf16 : dx_bf16) : nullptr;
  if (dbias && !bsrc) return MDHS_ERR_ARG;
  const bool params = (dgamma != nullptr) || (dbeta != nullptr) || (dbias != nullptr);
  g_mdhs_launches += params ? 2 : 1;
#define LNB1(XF, DF, N)                                                                                                       \
  ln_bwd_dx_kernel<XF, DF, N><<<blocks, 256, 0, st>>>(dy, lddy, x, ldx, mean, rstd, gamma, (bf16*)dx_bf16, lddx,               \
                                                      (bf16*)dx_drop_bf16, dx_f32, rows, C, drop_p, seed, drop2_p, seed2)
#define LNB(XF, DF)                     \
  do {                                  \
    if (C <= 256) LNB1(XF, DF, 1);      \
    else if (C <= 512) LNB1(XF, DF, 2); \
    else if (C <= 768) LNB1(XF, DF, 3); \
    else if (C <= 1024) LNB1(XF, DF, 4);\
    else LNB1(XF, DF, 8);               \
  } while (0)
  const bool narrow = !x_f32 && !dy_f32 && C <= 128;
  if (narrow && dx_bf16 && !dx_drop_bf16 && !dx_f32 && drop_p <= 0.f) {
    ln_bwd_dx_small_kernel<<<ceil_div(rows, 32), 256, 0, st>>>((const bf16*)dy, lddy, (const bf16*)x, ldx, mean, rstd, gamma,
                                                               (bf16*)dx_bf16, lddx, rows, C);
  } else if (x_f32) {
    if (dy_f32) LNB(true, true); else LNB(true, false);
  } else {
    if (dy_f32) LNB(false, true); else LNB(false, false);
  }
#undef LNB
#undef LNB1
  if (params) {
    const int slab = narrow ? 128 : 256;
    const int cslabs = ceil_div(C, slab);
    int row_blocks = ((int64_t)mdhs_num_sms() * 4) / cslabs;
    if (row_blocks < 1) row_blocks = 1;
    int rpb = ceil_div(rows, row_blocks);
    rpb = ((rpb + 15) / 16) * 16;
    row_blocks = ceil_div(rows, rpb);
    const dim3 grid(cslabs, row_blocks);
#define LNP(XF, DF) ln_bwd_param_kernel<XF, DF, 32><<<grid, 256, 0, st>>>(dy, lddy, x, ldx, mean, rstd, dgamma, dbeta, rows, C, rpb, drop_p, seed, bsrc, lddx, dbias)
    if (narrow) {
      ln_bwd_param_kernel<false, false, 16><<<grid, 256, 0, st>>>(dy, lddy, x, ldx, mean, rstd, dgamma, dbeta, rows, C, rpb, drop_p,
                                                                  seed, bsrc, lddx, dbias);
    } else if (x_f32) {
      if (dy_f32) LNP(true, true); else LNP(true, false);
    } else {
      if (dy_f32) LNP(false, true); else LNP(false, false);
    }
#undef LNP
  }
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_bn_finalize(const double* colsum, const double* colsumsq, int64_t count, const float* gamma,
                                const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                                float* mean, float* invstd, float* scale, float* shift, int C, int training, void* stream) {
  if (!gamma || !beta || !scale || !shift || C <= 0) return MDHS_ERR_ARG;
  if (training && (!colsum || !colsumsq || count <= 0)) return MDHS_ERR_ARG;
  if (!training && (!running_mean || !running_var)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      colsum, colsumsq, count, gamma, beta, running_mean, running_var, momentum, eps, mean, invstd, scale, shift, C, training);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_bn_apply(const void* x, const float* scale, const float* shift, const void* residual, void* y,
                             int64_t rows, int C, int relu, void* stream) {
  if (!x || !scale || !shift || !y || rows <= 0 || (C % 8)) return MDHS_ERR_ARG;
  const int64_t total_vec = rows * (C / 8);
  g_mdhs_launches++;
  bn_apply_kernel<false><<<grid_for_channels(total_vec, C), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)x, scale, shift, (const bf16*)residual, (bf16*)y, total_vec, C, relu, nullptr);
  MDHS_RETURN_LAST();
}

// finalize + apply as one C-ABI call (two launches: the per-channel fp64 work runs once per channel, not once per thread)
extern "C" int mdhs_bn_fwd(const void* x, const double* colsum, const double* colsumsq, const float* gamma, const float* beta,
                           float* running_mean, float* running_var, float momentum, float eps, const void* residual, void* y,
                           float* mean, float* invstd, float* scale, float* shift, void* relu_mask, int64_t rows, int C,
                           int relu, int training, void* stream) {
  if (!x || !y || !gamma || !beta || !scale || !shift || rows <= 0 || C <= 0 || (C % 8)) return MDHS_ERR_ARG;
  if (relu_mask && !relu) return MDHS_ERR_ARG;
  if (training && (!colsum || !colsumsq)) return MDHS_ERR_ARG;
  if (!training && (!running_mean || !running_var)) return MDHS_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  const int64_t total_vec = rows * (C / 8);
  g_mdhs_launches += 2;
  bn_finalize_kernel<<<ceil_div(C, 128), 128, 0, st>>>(colsum, colsumsq, rows, gamma, beta, running_mean, running_var, momentum,
                                                       eps, mean, invstd, scale, shift, C, training);
  if (relu_mask)
    bn_apply_kernel<true><<<grid_for_channels(total_vec, C), 256, 0, st>>>((const bf16*)x, scale, shift, (const bf16*)residual,
                                                                          (bf16*)y, total_vec, C, relu, (uint8_t*)relu_mask);
  else
    bn_apply_kernel<false><<<grid_for_channels(total_vec, C), 256, 0, st>>>((const bf16*)x, scale, shift, (const bf16*)residual,
                                                                           (bf16*)y, total_vec, C, relu, nullptr);
  MDHS_RETURN_LAST();
}

// MDHS_BN_REDUCE="RL,blocks_per_sm" (RL in {8, 16}): tuning knob for the backward reduction grid (default 8,2)
static void bn_reduce_cfg(int* rl, int* bps) {
  static int s_rl = 0, s_bps = 0;
  if (s_rl == 0) {
    s_rl = 8;
    s_bps = 2;
    const char* e = getenv("MDHS_BN_REDUCE");
    if (e) {
      int a = 0, b = 0;
      if (sscanf(e, "%d,%d", &a, &b) == 2 && (a == 8 || a == 16) && b >= 1 && b <= 16) {
        s_rl = a;
        s_bps = b;
      }
    }
  }
  *rl = s_rl;
  *bps = s_bps;
}

extern "C" int mdhs_bn_bwd(const void* dy, const void* x, const void* y, const void* relu_mask, const float* mean,
                           const float* invstd,
                           const float* gamma, const float* scale, const float* shift, double* sum_dy, double* sum_dy_xc,
                           float* coef, void* dx, void* dz, float* dgamma, float* dbeta, int64_t rows, int C, int relu,
                           int training, int sums_ready, void* stream) {
  if (!dy || !x || !mean || !invstd || !gamma || !sum_dy || !sum_dy_xc || !coef || !dx || rows <= 0 || (C % 8)) return MDHS_ERR_ARG;
  // the ReLU mask comes from the forward's bit mask, from y, or is recomputed from (x, scale, shift)
  if (relu && !y && !relu_mask && (!scale || !shift)) return MDHS_ERR_ARG;
  if ((dgamma == nullptr) != (dbeta == nullptr)) return MDHS_ERR_ARG;
  cudaStream_t st = reinterpret_cast<cudaStream_t>(stream);
  if (!sums_ready) {
    cudaError_t e;
    if (sum_dy_xc == sum_dy + C) {   // the usual [2, C] workspace: one memset node instead of two
      e = cudaMemsetAsync(sum_dy, 0, sizeof(double) * 2 * C, st);
      if (e != cudaSuccess) return (int)e;
    } else {
      e = cudaMemsetAsync(sum_dy, 0, sizeof(double) * C, st);
      if (e != cudaSuccess) return (int)e;
      e = cudaMemsetAsync(sum_dy_xc, 0, sizeof(double) * C, st);
      if (e != cudaSuccess) return (int)e;
    }
    int64_t rows_w;
    int Cw;
    fold_view(rows, C, C, &rows_w, &Cw);
    const int cslabs = ceil_div(Cw, 256);
    int rl, bps;
    bn_reduce_cfg(&rl, &bps);
    int row_blocks = (mdhs_num_sms() * bps) / cslabs;
    if (row_blocks < 1) row_blocks = 1;
    int rpb = ceil_div(rows_w, row_blocks);
    rpb = ((rpb + 4 * rl - 1) / (4 * rl)) * (4 * rl);
    row_blocks = ceil_div(rows_w, rpb);
    g_mdhs_launches++;
    const dim3 grid(cslabs, row_blocks);
    if (rl == 16)
      bn_bwd_reduce_kernel<16><<<grid, 512, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)y, (const uint8_t*)relu_mask, mean, scale, shift, sum_dy,
                                                     sum_dy_xc, rows_w, C, Cw, rpb, relu);
    else
      bn_bwd_reduce_kernel<8><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)y, (const uint8_t*)relu_mask, mean, scale, shift, sum_dy,
                                                    sum_dy_xc, rows_w, C, Cw, rpb, relu);
  }
  g_mdhs_launches += 2;
  const bool remask = relu && !y && !relu_mask;
  bn_bwd_coeff_kernel<<<ceil_div(C, 128), 128, 0, st>>>(mean, invstd, gamma, sum_dy, sum_dy_xc, remask ? scale : nullptr,
                                                        remask ? shift : nullptr, coef, dgamma, dbeta, rows, C, training);
  const int64_t total_vec = rows * (C / 8);
  const int mode = !relu ? 0 : (relu_mask ? 3 : (y ? 2 : 1));
  const int grid = grid_for_channels(total_vec, C);
#define BN_APPLY(MODE, DZ)                                                                                                  \
  bn_bwd_apply_kernel<MODE, DZ><<<grid, 256, 0, st>>>((const bf16*)dy, (const bf16*)x, (const bf16*)y, (const uint8_t*)relu_mask, \
                                                     coef, (bf16*)dx, (bf16*)dz, rows, C)
#define BN_APPLY_DZ(MODE)              \
  do {                                 \
    if (dz) BN_APPLY(MODE, true);      \
    else BN_APPLY(MODE, false);        \
  } while (0)
  switch (mode) {
    case 0: BN_APPLY_DZ(0); break;
    case 1: BN_APPLY_DZ(1); break;
    case 2: BN_APPLY_DZ(2); break;
    default: BN_APPLY_DZ(3); break;
  }
#undef BN_APPLY_DZ
#undef BN_APPLY
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_col_stats(const void* x, int64_t ldx, double* sum64, double* sumsq64, float* sum32, int64_t rows, int C,
                              void* stream) {
  if (!x || rows <= 0 || (C % 8) || (ldx % 8)) return MDHS_ERR_ARG;
  int64_t rows_w;
  int Cw;
  if (fold_view(rows, C, ldx, &rows_w, &Cw)) ldx = Cw;
  const int cslabs = ceil_div(Cw, 256);
  int row_blocks = ((int64_t)mdhs_num_sms() * 8) / cslabs;
  // short matrices (bias gradients of [8192, N] token matrices): at least 128 rows per block, otherwise the kernel is all
  // prologue + atomics (394 blocks of 24 rows each for N = 768)
  if (row_blocks > rows_w / 128) row_blocks = (int)(rows_w / 128);
  if (row_blocks < 1) row_blocks = 1;
  int rpb = ceil_div(rows_w, row_blocks);
  rpb = ((rpb + 31) / 32) * 32;
  row_blocks = ceil_div(rows_w, rpb);
  g_mdhs_launches++;
  col_stats_kernel<<<dim3(cslabs, row_blocks), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)x, ldx, sum64, sumsq64, sum32, rows_w, C, Cw, rpb);
  MDHS_RETURN_LAST();
}
