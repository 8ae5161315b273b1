// Data-movement kernels around the implicit-GEMM convolutions of the image encoder (NHWC bf16):
// patch gather (im2col) / scatter-free col2im, 3x3/2 max-pool, token pooling, weight re-layout.
// All are HBM-bound: 16-byte vectors along the channel axis, grid sized from the element count.
#include "common.cuh"
#include "../../include/mdhs_b200.h"

extern int64_t g_mdhs_launches;

namespace {

int grid_for(int64_t items, int block) {
  int64_t g = (items + block - 1) / block;
  const int64_t cap = (int64_t)mdhs_num_sms() * 32;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// NCHW fp32 image -> im2col rows for the stem (C_in = 3 is too narrow for channel vectors).
// col[(b,ho,wo), (r*S+s)*C + c] ; columns >= R*S*C are zero padding up to ldc.
// A block assembles IM2COL_PIX consecutive patch rows in shared memory -- one task per (pixel, filter row, channel)
// reads the S consecutive input floats of that filter row and scatters them as bf16 -- and then streams the finished
// rows (one contiguous IM2COL_PIX * ldc * 2-byte chunk of the patch matrix) out with 16-byte stores.  The previous
// one-thread-per-8-columns form spent ~60 integer instructions per element on index arithmetic (ncu: 79 % issue-bound,
// 0.9 ms for the ResNet stem at B = 128).
constexpr int IM2COL_PIX = 32;
// Test-time augmentation folded into the addressing (scripts/predict.py:33-42): with V > 1 the output holds V variants of the
// Bsrc source images stacked on the batch axis (row block v = transform `codes >> 4v & 15` of every image: 0 identity,
// 1 hflip, 2 vflip, 3 rot90), read straight from the un-expanded batch -- the V x B x 3 x H x W fp32 copy never exists.
__device__ __forceinline__ void tta_src(int code, int H, int W, int h, int w, int& sh, int& sw) {
  sh = h;
  sw = w;
  if (code == 1) sw = W - 1 - w;
  else if (code == 2) sh = H - 1 - h;
  else if (code == 3) { sh = w; sw = W - 1 - h; }   // rot90(k=1): out[i][j] = x[j][W-1-i] (square images)
}
__global__ void __launch_bounds__(256) im2col_nchw_f32_kernel(const float* __restrict__ x, bf16* __restrict__ col, int B, int C,
                                                              int H, int W, int R, int S, int stride, int pad, int Ho, int Wo,
                                                              int ldc, int Bsrc, uint32_t codes) {
  extern __shared__ __align__(16) uint8_t im2col_smem[];
  bf16* tile = reinterpret_cast<bf16*>(im2col_smem);
  const int64_t total_rows = (int64_t)B * Ho * Wo;
  const int64_t row0 = (int64_t)blockIdx.x * IM2COL_PIX;
  const int K = R * S * C;
  // zero the padding columns once (they are never written by the tasks)
  for (int i = threadIdx.x; i < IM2COL_PIX * (ldc - K); i += blockDim.x) {
    const int p = i / (ldc - K), k = K + i % (ldc - K);
    tile[p * ldc + k] = __float2bfloat16(0.f);
  }
  const int tasks = IM2COL_PIX * R * C;
  for (int t = threadIdx.x; t < tasks; t += blockDim.x) {
    const int p = t % IM2COL_PIX;          // consecutive threads -> consecutive output pixels (coalesced-ish reads)
    const int rc = t / IM2COL_PIX;
    const int c = rc % C, r = rc / C;
    const int64_t row = row0 + p;
    bf16* dst = tile + p * ldc + (r * S) * C + c;
    if (row >= total_rows) continue;
    const int wo = (int)(row % Wo);
    const int ho = (int)((row / Wo) % Ho);
    const int b = (int)(row / ((int64_t)Wo * Ho));
    const int h = ho * stride - pad + r;
    const int w0 = wo * stride - pad;
    const bool h_ok = (h >= 0 && h < H);
    const int code = (codes >> (4 * (b / Bsrc))) & 15;     // B == Bsrc and codes == 0 without TTA
    if (code == 0) {
      const float* src = x + (((int64_t)(b % Bsrc) * C + c) * H + (h_ok ? h : 0)) * W;
      for (int sx = 0; sx < S; sx++) {
        const int w = w0 + sx;
        const float v = (h_ok && w >= 0 && w < W) ? __ldg(src + w) : 0.f;
        dst[sx * C] = __float2bfloat16(v);
      }
    } else {
      const float* plane = x + ((int64_t)(b % Bsrc) * C + c) * H * W;
      for (int sx = 0; sx < S; sx++) {
        const int w = w0 + sx;
        float v = 0.f;
        if (h_ok && w >= 0 && w < W) {
          int sh, sw;
          tta_src(code, H, W, h, w, sh, sw);
          v = __ldg(plane + (int64_t)sh * W + sw);
        }
        dst[sx * C] = __float2bfloat16(v);
      }
    }
  }
  __syncthreads();
  const int64_t rows_here = total_rows - row0 < IM2COL_PIX ? total_rows - row0 : IM2COL_PIX;
  const int nvec = (int)(rows_here * ldc / 8);       // ldc % 8 == 0: the block's rows are one contiguous chunk
  const uint4* tv = reinterpret_cast<const uint4*>(tile);
  uint4* gv = reinterpret_cast<uint4*>(col + row0 * ldc);
  for (int i = threadIdx.x; i < nvec; i += blockDim.x) gv[i] = tv[i];
}

// Row-strip form of the kernel above (the default): a block owns ONE output row (b, ho).  It first stages the R input rows
// x C channels the strip reads -- coalesced fp32 reads, zero padding and the TTA source mapping applied here -- into shared
// memory, then every thread assembles whole 16-byte vectors of the patch matrix (8 consecutive k = (r, s, c) entries of one
// pixel, looked up through a k -> patch-offset table) and stores them to the strip's contiguous Wo * ldc chunk.  The
// task-per-(pixel, r, c) kernel above scattered 2-byte values into a shared tile at a 304-byte pixel stride (4-way bank
// conflicts) behind stride-2 global reads and ran at 1.4 TB/s (393 us for the 488 MB stem matrix of a 128-image batch).
__global__ void __launch_bounds__(256) im2col_nchw_rows_kernel(const float* __restrict__ x, bf16* __restrict__ col, int C, int H,
                                                               int W, int R, int S, int stride, int pad, int Ho, int Wo, int ldc,
                                                               int Bsrc, uint32_t codes, int PW) {
  extern __shared__ __align__(16) uint8_t im2col_smem[];
  float* patch = reinterpret_cast<float*>(im2col_smem);            // [C][R][PW]
  int* koff = reinterpret_cast<int*>(patch + (C * R * PW + 3) / 4 * 4);   // [ldc]: patch offset of column k, -1 = padding
  const int b = blockIdx.x / Ho, ho = blockIdx.x % Ho;
  const int K = R * S * C;
  for (int k = threadIdx.x; k < ldc; k += blockDim.x) {
    const int c = k % C, tap = k / C;
    koff[k] = k < K ? (c * R + tap / S) * PW + tap % S : -1;
  }
  const int code = (codes >> (4 * (b / Bsrc))) & 15;     // 0 without TTA
  const float* img = x + (int64_t)(b % Bsrc) * C * H * W;
  const int h0 = ho * stride - pad;
  // one warp per (channel, filter row) line of the strip; up to 8 independent loads per lane in flight (PW <= 256 is the
  // common case: a per-element loop with its index divisions kept one load per thread in flight and ran at 1.3 TB/s)
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  for (int cr = warp; cr < C * R; cr += 8) {
    const int r = cr % R, c = cr / R;
    const int h = h0 + r;
    const bool h_ok = h >= 0 && h < H;
    float* line = patch + cr * PW;
    const float* src = img + (int64_t)c * H * W;
    for (int j0 = 0; j0 < PW; j0 += 256) {
      float v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int j = j0 + u * 32 + lane, w = j - pad;
        v[u] = 0.f;
        if (j < PW && h_ok && w >= 0 && w < W) {
          int sh = h, sw = w;
          if (code != 0) tta_src(code, H, W, h, w, sh, sw);
          v[u] = __ldg(src + (int64_t)sh * W + sw);
        }
      }
#pragma unroll
      for (int u = 0; u < 8; u++) {
        const int j = j0 + u * 32 + lane;
        if (j < PW) line[j] = v[u];
      }
    }
  }
  __syncthreads();
  const int vpr = ldc >> 3;                                // 16-byte vectors per patch-matrix row
  uint4* out = reinterpret_cast<uint4*>(col + ((int64_t)b * Ho + ho) * Wo * ldc);
  for (int i = threadIdx.x; i < Wo * vpr; i += blockDim.x) {
    const int wo = i / vpr, k0 = (i - wo * vpr) * 8;
    const float* base = patch + wo * stride;
    int o[8];
    *reinterpret_cast<int4*>(o) = *reinterpret_cast<const int4*>(koff + k0);
    *reinterpret_cast<int4*>(o + 4) = *reinterpret_cast<const int4*>(koff + k0 + 4);
    float v[8];
#pragma unroll
    for (int e = 0; e < 8; e++) v[e] = o[e] >= 0 ? base[o[e]] : 0.f;
    store8(reinterpret_cast<bf16*>(out + i), v);
  }
}

// NHWC bf16 -> im2col rows, 8 channels per thread.  col[(b,ho,wo), (r*S+s)*C + c].
__global__ void __launch_bounds__(256) im2col_nhwc_kernel(const bf16* __restrict__ x, bf16* __restrict__ col, int B, int H, int W,
                                                          int C, int R, int S, int stride, int pad, int Ho, int Wo) {
  const int cvec = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * R * S * cvec;
  const uint4 zero = make_uint4(0, 0, 0, 0);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t t = i / cvec;
    const int rs = (int)(t % (R * S));
    t /= (R * S);
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    const int r = rs / S, s = rs % S;
    const int h = ho * stride - pad + r, w = wo * stride - pad + s;
    uint4 v = zero;
    if (h >= 0 && h < H && w >= 0 && w < W)
      v = *reinterpret_cast<const uint4*>(x + (((int64_t)b * H + h) * W + w) * C + cv * 8);
    *reinterpret_cast<uint4*>(col + i * 8) = v;
  }
}

// Gradient of im2col as a gather (no atomics): dx[b,h,w,c] = sum over the taps that read this pixel.
// If `add` is given (same shape as dx) it is summed in (residual-branch gradient).
__global__ void __launch_bounds__(256) col2im_nhwc_kernel(const bf16* __restrict__ dcol, const bf16* __restrict__ add,
                                                          bf16* __restrict__ dx, int B, int H, int W, int C, int R, int S,
                                                          int stride, int pad, int Ho, int Wo) {
  const int cvec = C >> 3;
  const int64_t total = (int64_t)B * H * W * cvec;
  const int64_t ldc = (int64_t)R * S * C;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t t = i / cvec;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.f;
    if (add) load8(add + i * 8, acc);
    for (int r = 0; r < R; r++) {
      const int hn = h + pad - r;
      if (hn < 0 || (hn % stride) != 0) continue;
      const int ho = hn / stride;
      if (ho >= Ho) continue;
      for (int s = 0; s < S; s++) {
        const int wn = w + pad - s;
        if (wn < 0 || (wn % stride) != 0) continue;
        const int wo = wn / stride;
        if (wo >= Wo) continue;
        float v[8];
        load8(dcol + (((int64_t)b * Ho + ho) * Wo + wo) * ldc + (int64_t)(r * S + s) * C + cv * 8, v);
#pragma unroll
        for (int k = 0; k < 8; k++) acc[k] += v[k];
      }
    }
    store8(dx + i * 8, acc);
  }
}

// 3x3 stride-2 pad-1 max-pool (torchvision ResNet stem), NHWC bf16; records the winning tap (0..8, first
// maximum in row-major window order like ATen) so that the backward pass is a deterministic gather.
__global__ void __launch_bounds__(256) maxpool3x3s2_fwd_kernel(const bf16* __restrict__ x, bf16* __restrict__ y,
                                                               uint8_t* __restrict__ idx, int B, int H, int W, int C, int Ho,
                                                               int Wo) {
  const int cvec = C >> 3;
  const int64_t total = (int64_t)B * Ho * Wo * cvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t t = i / cvec;
    const int wo = (int)(t % Wo);
    t /= Wo;
    const int ho = (int)(t % Ho);
    const int b = (int)(t / Ho);
    float best[8];
    uint8_t bi[8];
#pragma unroll
    for (int k = 0; k < 8; k++) {
      best[k] = -INFINITY;
      bi[k] = 0;
    }
    for (int r = 0; r < 3; r++) {
      const int h = ho * 2 - 1 + r;
      if (h < 0 || h >= H) continue;
      for (int s = 0; s < 3; s++) {
        const int w = wo * 2 - 1 + s;
        if (w < 0 || w >= W) continue;
        float v[8];
        load8(x + (((int64_t)b * H + h) * W + w) * C + cv * 8, v);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          if (v[k] > best[k]) {
            best[k] = v[k];
            bi[k] = (uint8_t)(r * 3 + s);
          }
        }
      }
    }
    store8(y + i * 8, best);
    uint2 packed;
    packed.x = bi[0] | (bi[1] << 8) | (bi[2] << 16) | ((uint32_t)bi[3] << 24);
    packed.y = bi[4] | (bi[5] << 8) | (bi[6] << 16) | ((uint32_t)bi[7] << 24);
    *reinterpret_cast<uint2*>(idx + i * 8) = packed;
  }
}

__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                               bf16* __restrict__ dx, int B, int H, int W, int C, int Ho,
                                                               int Wo) {
  const int cvec = C >> 3;
  const int64_t total = (int64_t)B * H * W * cvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    int64_t t = i / cvec;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.f;
    for (int r = 0; r < 3; r++) {
      const int hn = h + 1 - r;
      if (hn < 0 || (hn & 1)) continue;
      const int ho = hn >> 1;
      if (ho >= Ho) continue;
      for (int s = 0; s < 3; s++) {
        const int wn = w + 1 - s;
        if (wn < 0 || (wn & 1)) continue;
        const int wo = wn >> 1;
        if (wo >= Wo) continue;
        const int64_t o = ((((int64_t)b * Ho + ho) * Wo + wo) * cvec + cv) * 8;
        const uint2 packed = *reinterpret_cast<const uint2*>(idx + o);
        float v[8];
        load8(dy + o, v);
        const uint32_t tap = (uint32_t)(r * 3 + s);
#pragma unroll
        for (int k = 0; k < 8; k++) {
          const uint32_t word = k < 4 ? packed.x : packed.y;
          if (((word >> ((k & 3) * 8)) & 0xffu) == tap) acc[k] += v[k];
        }
      }
    }
    store8(dx + i * 8, acc);
  }
}

// Even H and W (the 112 x 112 stem map): a thread owns a 2 x 2 input block and 8 channels.  The block's pixels are reached
// only from the four windows (i, j), (i, j+1), (i+1, j), (i+1, j+1), through nine (window, tap) pairs in total -- four dy / idx
// loads serve four input pixels (the per-pixel gather above loads 2.25 windows per pixel on average).
__global__ void __launch_bounds__(256) maxpool3x3s2_bwd_block_kernel(const bf16* __restrict__ dy, const uint8_t* __restrict__ idx,
                                                                     bf16* __restrict__ dx, int B, int H, int W, int C, int Ho,
                                                                     int Wo) {
  const int cvec = C >> 3, H2 = H >> 1, W2 = W >> 1;
  const int64_t total = (int64_t)B * H2 * W2 * cvec;
  for (int64_t t = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(t % cvec);
    int64_t q = t / cvec;
    const int j = (int)(q % W2);
    q /= W2;
    const int i = (int)(q % H2);
    const int b = (int)(q / H2);
    float g[4][8];       // dy of windows (i,j), (i,j+1), (i+1,j), (i+1,j+1)
    uint2 tp[4];         // their winning taps (one byte per channel)
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const int ho = i + (k >> 1), wo = j + (k & 1);
      if (ho < Ho && wo < Wo) {
        const int64_t o = ((((int64_t)b * Ho + ho) * Wo + wo) * cvec + cv) * 8;
        load8(dy + o, g[k]);
        tp[k] = *reinterpret_cast<const uint2*>(idx + o);
      } else {
#pragma unroll
        for (int c = 0; c < 8; c++) g[k][c] = 0.f;
        tp[k] = make_uint2(0xffffffffu, 0xffffffffu);
      }
    }
    float o00[8], o01[8], o10[8], o11[8];
#pragma unroll
    for (int c = 0; c < 8; c++) {
      const int sh = (c & 3) * 8;
      uint32_t t0 = ((c < 4 ? tp[0].x : tp[0].y) >> sh) & 0xffu, t1 = ((c < 4 ? tp[1].x : tp[1].y) >> sh) & 0xffu;
      uint32_t t2 = ((c < 4 ? tp[2].x : tp[2].y) >> sh) & 0xffu, t3 = ((c < 4 ? tp[3].x : tp[3].y) >> sh) & 0xffu;
      // tap = r * 3 + s of the input pixel (2 ho - 1 + r, 2 wo - 1 + s) inside window (ho, wo)
      o00[c] = t0 == 4u ? g[0][c] : 0.f;
      o01[c] = (t0 == 5u ? g[0][c] : 0.f) + (t1 == 3u ? g[1][c] : 0.f);
      o10[c] = (t0 == 7u ? g[0][c] : 0.f) + (t2 == 1u ? g[2][c] : 0.f);
      o11[c] = (t0 == 8u ? g[0][c] : 0.f) + (t1 == 6u ? g[1][c] : 0.f) + (t2 == 2u ? g[2][c] : 0.f) + (t3 == 0u ? g[3][c] : 0.f);
    }
    bf16* p = dx + ((((int64_t)b * H + 2 * i) * W + 2 * j) * cvec + cv) * 8;
    store8(p, o00);
    store8(p + (int64_t)cvec * 8, o01);
    store8(p + (int64_t)W * cvec * 8, o10);
    store8(p + (int64_t)(W + 1) * cvec * 8, o11);
  }
}

// Mean over the token axis: x [B, T, C] bf16 -> y [B, C] (fp32 and/or bf16), scaled by `scale`
// (1/T for a mean; multi-scale fusion averages three pooled vectors with an extra 1/3).
__global__ void __launch_bounds__(256) mean_tokens_fwd_kernel(const bf16* __restrict__ x, float* __restrict__ y32,
                                                              bf16* __restrict__ y16, int B, int T, int C, float scale,
                                                              int accumulate) {
  const int cvec = C >> 3;
  const int64_t total = (int64_t)B * cvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const int b = (int)(i / cvec);
    float acc[8];
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] = 0.f;
    const bf16* p = x + (int64_t)b * T * C + cv * 8;
    for (int t = 0; t < T; t++) {
      float v[8];
      load8(p + (int64_t)t * C, v);
#pragma unroll
      for (int k = 0; k < 8; k++) acc[k] += v[k];
    }
#pragma unroll
    for (int k = 0; k < 8; k++) acc[k] *= scale;
    if (y32) {
      float* o = y32 + (int64_t)b * C + cv * 8;
#pragma unroll
      for (int k = 0; k < 8; k++) o[k] = accumulate ? o[k] + acc[k] : acc[k];
    }
    if (y16) store8(y16 + (int64_t)b * C + cv * 8, acc);
  }
}

// dx[b,t,c] = scale * dy[b,c]  (dy fp32 or bf16), optionally added to an existing gradient.
__global__ void __launch_bounds__(256) mean_tokens_bwd_kernel(const float* __restrict__ dy32, const bf16* __restrict__ dy16,
                                                              bf16* __restrict__ dx, int B, int T, int C, float scale) {
  const int cvec = C >> 3;
  const int64_t total = (int64_t)B * T * cvec;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int cv = (int)(i % cvec);
    const int b = (int)(i / ((int64_t)cvec * T));
    float v[8];
    if (dy32) {
      const float* p = dy32 + (int64_t)b * C + cv * 8;
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] = p[k] * scale;
    } else {
      load8(dy16 + (int64_t)b * C + cv * 8, v);
#pragma unroll
      for (int k = 0; k < 8; k++) v[k] *= scale;
    }
    store8(dx + i * 8, v);
  }
}

// Conv weight OIHW fp32 -> [O, (r*S+s)*I + i] bf16 with row stride ldk (zero padded), and the inverse
// accumulation of the GEMM-layout fp32 gradient back into the OIHW fp32 grad.
__global__ void conv_weight_pack_kernel(const float* __restrict__ w, bf16* __restrict__ wp, int O, int I, int R, int S, int ldk) {
  const int64_t total = (int64_t)O * ldk;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int k = (int)(i % ldk);
    const int o = (int)(i / ldk);
    float v = 0.f;
    if (k < R * S * I) {
      const int c = k % I, rs = k / I;
      v = w[((int64_t)o * I + c) * R * S + rs];
    }
    wp[i] = __float2bfloat16_rn(v);
  }
}
// dgrad operand for stride-1 "same" convolutions: wt[c_in, ((R-1-r)*S + (S-1-s))*O + o] = w[o, c_in, r, s]
__global__ void conv_weight_pack_dgrad_kernel(const float* __restrict__ w, bf16* __restrict__ wt, int O, int I, int R, int S) {
  const int64_t total = (int64_t)O * I * R * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int o = (int)(i % O);
    int64_t t = i / O;
    const int rs = (int)(t % (R * S));
    const int c = (int)(t / (R * S));
    const int r = R - 1 - rs / S, s = S - 1 - rs % S;
    wt[i] = __float2bfloat16_rn(w[(((int64_t)o * I + c) * R + r) * S + s]);
  }
}
__global__ void conv_wgrad_unpack_kernel(const float* __restrict__ gp, float* __restrict__ g, int O, int I, int R, int S, int ldk) {
  const int64_t total = (int64_t)O * I * R * S;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int rs = (int)(i % (R * S));
    const int64_t oc = i / (R * S);
    const int c = (int)(oc % I);
    const int o = (int)(oc / I);
    g[i] += gp[(int64_t)o * ldk + (int64_t)rs * I + c];
  }
}

// fp32 -> bf16 (weights after an optimizer step) and bf16 -> fp32.
__global__ void __launch_bounds__(256) cast_f32_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int64_t n) {
  const int64_t nv = n >> 3;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    const float4 a = *reinterpret_cast<const float4*>(x + i * 8), b = *reinterpret_cast<const float4*>(x + i * 8 + 4);
    const float v[8] = {a.x, a.y, a.z, a.w, b.x, b.y, b.z, b.w};
    store8(y + i * 8, v);
  }
  if (blockIdx.x == 0 && threadIdx.x < (n & 7)) {
    const int64_t j = (nv << 3) + threadIdx.x;
    y[j] = __float2bfloat16_rn(x[j]);
  }
}
__global__ void __launch_bounds__(256) cast_bf16_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = __bfloat162float(x[i]);
}

// NHWC bf16 [B,H,W,C] -> NCHW fp32 (module boundary for hooks / Grad-CAM) and back.
__global__ void nhwc_bf16_to_nchw_f32_kernel(const bf16* __restrict__ x, float* __restrict__ y, int B, int H, int W, int C) {
  const int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int w = (int)(i % W);
    int64_t t = i / W;
    const int h = (int)(t % H);
    t /= H;
    const int c = (int)(t % C);
    const int b = (int)(t / C);
    y[i] = __bfloat162float(x[(((int64_t)b * H + h) * W + w) * C + c]);
  }
}
__global__ void nchw_f32_to_nhwc_bf16_kernel(const float* __restrict__ x, bf16* __restrict__ y, int B, int H, int W, int C) {
  const int64_t total = (int64_t)B * C * H * W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int c = (int)(i % C);
    int64_t t = i / C;
    const int w = (int)(t % W);
    t /= W;
    const int h = (int)(t % H);
    const int b = (int)(t / H);
    y[i] = __float2bfloat16_rn(x[(((int64_t)b * C + c) * H + h) * W + w]);
  }
}

// Test-time-augmentation variants stacked on the batch axis (scripts/predict.py:33-42: identity, hflip = flip(-1),
// vflip = flip(-2), rot90 = torch.rot90(k=1, dims=(-2,-1)); square images): y[v, b, c, i, j] = x[b, c, src(i, j)].
// `codes` packs one 4-bit transform id per variant (0 identity, 1 hflip, 2 vflip, 3 rot90).  One pass writes all variants.
__global__ void __launch_bounds__(256) tta_expand_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t per_variant,
                                                         int H, int W, int V, uint32_t codes) {
  const int64_t total = per_variant * V;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int v = (int)(i / per_variant);
    const int64_t e = i - (int64_t)v * per_variant;
    const int j = (int)(e % W);
    const int r = (int)((e / W) % H);
    const int64_t plane = e / ((int64_t)W * H);
    const int code = (codes >> (4 * v)) & 15;
    int sr = r, sj = j;
    if (code == 1) sj = W - 1 - j;
    else if (code == 2) sr = H - 1 - r;
    else if (code == 3) { sr = j; sj = W - 1 - r; }   // rot90(k=1): out[i][j] = x[j][W-1-i]
    y[i] = x[(plane * H + sr) * W + sj];
  }
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

static int im2col_nchw_launch(const float* x, void* col, int Bsrc, int V, uint32_t codes, int C, int H, int W, int R, int S,
                              int stride, int pad, int ldc, void* stream) {
  if (!x || !col || ldc < R * S * C || (ldc % 8) || V < 1 || V > 8 || Bsrc <= 0) return MDHS_ERR_ARG;
  for (int v = 0; v < V; v++) {
    const int code = (codes >> (4 * v)) & 15;
    if (code > 3 || (code == 3 && H != W)) return MDHS_ERR_ARG;
  }
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  g_mdhs_launches++;
  const int B = Bsrc * V;
  const int64_t rows = (int64_t)B * Ho * Wo;
  if ((uintptr_t)col & 15) return MDHS_ERR_ARG;
  const int PW = (Wo - 1) * stride + S;
  const size_t strip = (size_t)((C * R * PW + 3) / 4 * 4) * sizeof(float) + (size_t)ldc * sizeof(int);
  if (strip <= 48 * 1024 && (int64_t)B * Ho < (1ll << 31)) {
    im2col_nchw_rows_kernel<<<(unsigned)(B * Ho), 256, strip, ST(stream)>>>(x, (bf16*)col, C, H, W, R, S, stride, pad, Ho, Wo, ldc,
                                                                          Bsrc, codes, PW);
    MDHS_RETURN_LAST();
  }
  const size_t smem = (size_t)IM2COL_PIX * ldc * sizeof(bf16);
  if (smem > 48 * 1024) return MDHS_ERR_ARG;
  im2col_nchw_f32_kernel<<<(unsigned)((rows + IM2COL_PIX - 1) / IM2COL_PIX), 256, smem, ST(stream)>>>(
      x, (bf16*)col, B, C, H, W, R, S, stride, pad, Ho, Wo, ldc, Bsrc, codes);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_im2col_nchw_f32(const float* x, void* col, int B, int C, int H, int W, int R, int S, int stride, int pad,
                                    int ldc, void* stream) {
  return im2col_nchw_launch(x, col, B, 1, 0u, C, H, W, R, S, stride, pad, ldc, stream);
}

// V test-time-augmentation variants of the B source images, stacked on the batch axis of the patch matrix (V * B * Ho * Wo rows)
extern "C" int mdhs_im2col_nchw_f32_tta(const float* x, void* col, int B, int C, int H, int W, int R, int S, int stride, int pad,
                                        int ldc, int V, int codes, void* stream) {
  return im2col_nchw_launch(x, col, B, V, (uint32_t)codes, C, H, W, R, S, stride, pad, ldc, stream);
}

extern "C" int mdhs_im2col_nhwc(const void* x, void* col, int B, int H, int W, int C, int R, int S, int stride, int pad,
                                void* stream) {
  if (!x || !col || (C % 8)) return MDHS_ERR_ARG;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  g_mdhs_launches++;
  im2col_nhwc_kernel<<<grid_for((int64_t)B * Ho * Wo * R * S * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)col, B, H,
                                                                                                    W, C, R, S, stride, pad, Ho, Wo);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_col2im_nhwc(const void* dcol, const void* add, void* dx, int B, int H, int W, int C, int R, int S, int stride,
                                int pad, void* stream) {
  if (!dcol || !dx || (C % 8)) return MDHS_ERR_ARG;
  const int Ho = (H + 2 * pad - R) / stride + 1, Wo = (W + 2 * pad - S) / stride + 1;
  g_mdhs_launches++;
  col2im_nhwc_kernel<<<grid_for((int64_t)B * H * W * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)dcol, (const bf16*)add,
                                                                                          (bf16*)dx, B, H, W, C, R, S, stride, pad,
                                                                                          Ho, Wo);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_maxpool3x3s2_fwd(const void* x, void* y, void* idx, int B, int H, int W, int C, void* stream) {
  if (!x || !y || !idx || (C % 8)) return MDHS_ERR_ARG;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  g_mdhs_launches++;
  maxpool3x3s2_fwd_kernel<<<grid_for((int64_t)B * Ho * Wo * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)x, (bf16*)y,
                                                                                                 (uint8_t*)idx, B, H, W, C, Ho, Wo);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int B, int H, int W, int C, void* stream) {
  if (!dy || !dx || !idx || (C % 8)) return MDHS_ERR_ARG;
  const int Ho = (H + 2 - 3) / 2 + 1, Wo = (W + 2 - 3) / 2 + 1;
  g_mdhs_launches++;
  if ((H % 2) == 0 && (W % 2) == 0) {
    maxpool3x3s2_bwd_block_kernel<<<grid_for((int64_t)B * (H / 2) * (W / 2) * (C / 8), 256), 256, 0, ST(stream)>>>(
        (const bf16*)dy, (const uint8_t*)idx, (bf16*)dx, B, H, W, C, Ho, Wo);
    MDHS_RETURN_LAST();
  }
  maxpool3x3s2_bwd_kernel<<<grid_for((int64_t)B * H * W * (C / 8), 256), 256, 0, ST(stream)>>>((const bf16*)dy, (const uint8_t*)idx,
                                                                                               (bf16*)dx, B, H, W, C, Ho, Wo);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_mean_tokens_fwd(const void* x, float* y32, void* y16, int B, int T, int C, float scale, int accumulate,
                                    void* stream) {
  if (!x || (!y32 && !y16) || (C % 8)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  mean_tokens_fwd_kernel<<<grid_for((int64_t)B * (C / 8), 128), 128, 0, ST(stream)>>>((const bf16*)x, y32, (bf16*)y16, B, T, C, scale,
                                                                                      accumulate);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_mean_tokens_bwd(const float* dy32, const void* dy16, void* dx, int B, int T, int C, float scale, void* stream) {
  if ((!dy32 && !dy16) || !dx || (C % 8)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  mean_tokens_bwd_kernel<<<grid_for((int64_t)B * T * (C / 8), 256), 256, 0, ST(stream)>>>(dy32, (const bf16*)dy16, (bf16*)dx, B, T, C,
                                                                                          scale);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_conv_weight_pack(const float* w, void* wp, int O, int I, int R, int S, int ldk, void* stream) {
  if (!w || !wp || ldk < R * S * I) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  conv_weight_pack_kernel<<<grid_for((int64_t)O * ldk, 256), 256, 0, ST(stream)>>>(w, (bf16*)wp, O, I, R, S, ldk);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_conv_weight_pack_dgrad(const float* w, void* wt, int O, int I, int R, int S, void* stream) {
  if (!w || !wt) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  conv_weight_pack_dgrad_kernel<<<grid_for((int64_t)O * I * R * S, 256), 256, 0, ST(stream)>>>(w, (bf16*)wt, O, I, R, S);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_conv_wgrad_unpack(const float* gp, float* g, int O, int I, int R, int S, int ldk, void* stream) {
  if (!gp || !g || ldk < R * S * I) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  conv_wgrad_unpack_kernel<<<grid_for((int64_t)O * I * R * S, 256), 256, 0, ST(stream)>>>(gp, g, O, I, R, S, ldk);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream) {
  if (!x || !y || n <= 0 || ((uintptr_t)x & 15) || ((uintptr_t)y & 15)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  cast_f32_bf16_kernel<<<grid_for(n / 8 + 1, 256), 256, 0, ST(stream)>>>(x, (bf16*)y, n);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream) {
  if (!x || !y || n <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  cast_bf16_f32_kernel<<<grid_for(n, 256), 256, 0, ST(stream)>>>((const bf16*)x, y, n);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int H, int W, int C, void* stream) {
  if (!x || !y) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  nhwc_bf16_to_nchw_f32_kernel<<<grid_for((int64_t)B * H * W * C, 256), 256, 0, ST(stream)>>>((const bf16*)x, y, B, H, W, C);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int H, int W, int C, void* stream) {
  if (!x || !y) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  nchw_f32_to_nhwc_bf16_kernel<<<grid_for((int64_t)B * H * W * C, 256), 256, 0, ST(stream)>>>(x, (bf16*)y, B, H, W, C);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_tta_expand(const float* x, float* y, int B, int C, int H, int W, int V, uint32_t codes, void* stream) {
  if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0 || V <= 0 || V > 8) return MDHS_ERR_ARG;
  for (int v = 0; v < V; v++) {
    const int code = (codes >> (4 * v)) & 15;
    if (code > 3 || (code == 3 && H != W)) return MDHS_ERR_ARG;
  }
  const int64_t per_variant = (int64_t)B * C * H * W;
  g_mdhs_launches++;
  tta_expand_kernel<<<grid_for(per_variant * V, 256), 256, 0, ST(stream)>>>(x, y, per_variant, H, W, V, codes);
  MDHS_RETURN_LAST();
}
