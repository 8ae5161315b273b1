// ConvNeXt-specific kernels (torchvision CNBlock as used by ConNexT/models/ourmodel.py:57-62 and the ConvNeXt image
// encoders of configs 4): NHWC bf16 activations [B*H*W, C].
//   * depthwise 7x7 convolution (pad 3, groups = C): forward, input gradient (same kernel, mirrored taps) and
//     weight / bias gradient.  C % 32 == 0 (every ConvNeXt width): TMA-staged halo boxes, a channel pair and a 4 x 7
//     output tile per thread, the 49 taps in registers, packed fp32 FMAs (dwconv7_tma_kernel / dwconv7_wgrad_tma_kernel).
//     Other widths: the older kernels (8 channels x 4 pixels per thread, taps read from shared memory).
//   * layer-scale + stochastic-depth + residual:  out = x + ls[c] * keep(b) * z   and its backward
//     (dz, dls and the bias gradient of the Linear that produced z; dx = dy is the identity branch and needs no kernel).
//   * single-query attention (ourmodel.py:17-31 with a 1-token query: the text -> image direction): softmax over the
//     T image positions of q.k_t (no 1/sqrt(d) in the reference), out = sum_t p_t v_t; forward + backward.
// The 1x1 / patchify convolutions, LayerNorms and MLPs of the block run on the shared GEMM / LayerNorm kernels.
#include <cuda.h>
#include <cstdlib>
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

// ------------------------------------------------------------------ depthwise 7x7, TMA-staged register tiles (C % 32 == 0)
// The kernels above fetch a filter tap from shared memory for every 8 multiply-adds and are bound by the L1 / shared
// pipe (ncu: l1tex 95 %, issue 32 %).  Here a thread owns ONE channel pair and a 4 x 7 output tile: the pair's 49 taps
// live in registers as packed fp32 pairs and every multiply-add is an fma.rn.f32x2 over both channels (1372 per tile).
// A block owns a 32-channel slab and walks boxes of NB images x NH*7 rows x NW*4 columns; the box and its 3-pixel halo
// arrive by one 4-D TMA load (zero fill outside the image: no predicates, no address arithmetic) into a double-buffered
// shared tile that the threads read with constant offsets.  The box is 4*NW + 7 pixels wide: an odd pixel count makes
// the two half-warps (tiles 7 rows or one image apart, 64 bytes each) hit disjoint banks.
template <int NW_, int NH_, int NB_>
struct DwCfg {
  static constexpr int NW = NW_, NH = NH_, NB = NB_;
  static constexpr int WB = 4 * NW + 7, HB = 7 * NH + 6;           // halo box (pixels)
  static constexpr int DWB = 4 * NW + 1, DHB = 7 * NH;             // dy box of the weight gradient
  static constexpr int SLOTS = NW * NH * NB, THREADS = 16 * SLOTS;
  static constexpr int XBYTES = NB * HB * WB * 64, DYBYTES = NB * DHB * DWB * 64;
  static constexpr int XBUF = (XBYTES + 127) / 128 * 128, DYBUF = (DYBYTES + 127) / 128 * 128;
  static_assert(SLOTS % 2 == 0 && (NH == 2 || NB % 2 == 0), "half-warps pair over rows or images");
};

__device__ __forceinline__ uint64_t fma2(uint64_t a, uint64_t b, uint64_t c) {
  uint64_t d;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
  return d;
}
__device__ __forceinline__ uint64_t pack2(float lo, float hi) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "f"(lo), "f"(hi));
  return d;
}
__device__ __forceinline__ float2 unpack2(uint64_t v) {
  float2 f;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(f.x), "=f"(f.y) : "l"(v));
  return f;
}
// two packed bf16 -> packed fp32 pair (bf16 is the high half of an fp32)
__device__ __forceinline__ uint64_t bf2_to_f2(uint32_t u) {
  uint64_t d;
  asm("mov.b64 %0, {%1, %2};" : "=l"(d) : "r"(u << 16), "r"(u & 0xffff0000u));
  return d;
}
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.b32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void dw_mbar_init(uint32_t bar) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], 1;" ::"r"(bar) : "memory");
}
__device__ __forceinline__ void dw_mbar_expect(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}
__device__ __forceinline__ void dw_mbar_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  do {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(bar), "r"(parity)
        : "memory");
  } while (!ok);
}
__device__ __forceinline__ void dw_tma_4d(uint32_t dst, const CUtensorMap* map, uint32_t bar, int c, int w, int h, int b) {
  asm volatile(
      "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(dst),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(bar), "r"(c), "r"(w), "r"(h), "r"(b)
      : "memory");
}

struct DwSlot {
  int nw, nh, nb;
};
// slot -> tile inside the box; the two half-warps of a warp differ in nh (NH == 2) or in nb
template <class Cfg>
__device__ __forceinline__ DwSlot dw_slot(int sl) {
  DwSlot s;
  if (Cfg::NH == 2) {
    s.nh = sl & 1;
    const int r = sl >> 1;
    s.nw = r % Cfg::NW;
    s.nb = r / Cfg::NW;
  } else {
    s.nh = 0;
    s.nb = sl % Cfg::NB;
    s.nw = sl / Cfg::NB;
  }
  return s;
}

struct DwBox {
  int w, h, b;   // first output column / row / image of the box
};
template <class Cfg>
__device__ __forceinline__ DwBox dw_box(int bi, int nbw, int nbh) {
  DwBox o;
  o.w = (bi % nbw) * 4 * Cfg::NW;
  o.h = ((bi / nbw) % nbh) * 7 * Cfg::NH;
  o.b = (bi / (nbw * nbh)) * Cfg::NB;
  return o;
}

template <class Cfg, bool FLIP>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
    dwconv7_tma_kernel(const __grid_constant__ CUtensorMap tmX, const float* __restrict__ w, const float* __restrict__ bias,
                       bf16* __restrict__ y, int B, int H, int W, int C, int nbw, int nbh, int nboxes) {
  extern __shared__ __align__(128) uint8_t dw_smem[];
  __shared__ float wsm[49][32];
  __shared__ __align__(8) uint64_t bars[2];
  const int slab = blockIdx.y * 32;
  const uint32_t xs = (uint32_t)__cvta_generic_to_shared(dw_smem);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    dw_mbar_init(bar0);
    dw_mbar_init(bar0 + 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int k = 0; k < 2; k++) {
      const int bi = blockIdx.x + k * gridDim.x;
      if (bi < nboxes) {
        const DwBox o = dw_box<Cfg>(bi, nbw, nbh);
        dw_mbar_expect(bar0 + 8 * k, Cfg::XBYTES);
        dw_tma_4d(xs + k * Cfg::XBUF, &tmX, bar0 + 8 * k, slab, o.w - 3, o.h - 3, o.b);
      }
    }
  }
  for (int i = threadIdx.x; i < 49 * 32; i += blockDim.x) {
    const int cl = i / 49, tap = i - cl * 49;     // coalesced over the [C][49] weight
    wsm[FLIP ? 48 - tap : tap][cl] = w[(int64_t)(slab + cl) * 49 + tap];
  }
  __syncthreads();
  const int cp = threadIdx.x & 15;
  const int c0 = slab + cp * 2;
  uint64_t wt[49];
#pragma unroll
  for (int k = 0; k < 49; k++) wt[k] = *reinterpret_cast<const uint64_t*>(&wsm[k][cp * 2]);
  const uint64_t init = (!FLIP && bias != nullptr) ? pack2(bias[c0], bias[c0 + 1]) : 0ull;
  const DwSlot sl = dw_slot<Cfg>(threadIdx.x >> 4);
  const uint32_t tile_off = (uint32_t)(((sl.nb * Cfg::HB + sl.nh * 7) * Cfg::WB + sl.nw * 4) * 64 + cp * 4);
  int it = 0;
  for (int bi = blockIdx.x; bi < nboxes; bi += gridDim.x, it++) {
    const int buf = it & 1;
    const DwBox o = dw_box<Cfg>(bi, nbw, nbh);
    dw_mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);
    const uint32_t base = xs + buf * Cfg::XBUF + tile_off;
    uint64_t acc[7][4];
#pragma unroll
    for (int oh = 0; oh < 7; oh++)
#pragma unroll
      for (int ow = 0; ow < 4; ow++) acc[oh][ow] = init;
#pragma unroll
    for (int i = 0; i < 13; i++) {
      uint64_t xin[10];
#pragma unroll
      for (int j = 0; j < 10; j++) xin[j] = bf2_to_f2(lds32(base + (i * Cfg::WB + j) * 64));
#pragma unroll
      for (int oh = 0; oh < 7; oh++) {
        const int r = i - oh;
        if (r < 0 || r > 6) continue;
#pragma unroll
        for (int s = 0; s < 7; s++)
#pragma unroll
          for (int ow = 0; ow < 4; ow++) acc[oh][ow] = fma2(xin[ow + s], wt[r * 7 + s], acc[oh][ow]);
      }
    }
    __syncthreads();   // every thread has read this buffer: refill it with the box two iterations ahead
    if (threadIdx.x == 0) {
      const int nb = bi + 2 * gridDim.x;
      if (nb < nboxes) {
        const DwBox n = dw_box<Cfg>(nb, nbw, nbh);
        dw_mbar_expect(bar0 + 8 * buf, Cfg::XBYTES);
        dw_tma_4d(xs + buf * Cfg::XBUF, &tmX, bar0 + 8 * buf, slab, n.w - 3, n.h - 3, n.b);
      }
    }
    const int b = o.b + sl.nb, h0 = o.h + sl.nh * 7, w0 = o.w + sl.nw * 4;
    if (b < B) {
#pragma unroll
      for (int oh = 0; oh < 7; oh++) {
        const int h = h0 + oh;
        if (h >= H) break;
        bf16* out = y + ((int64_t)(b * H + h) * W + w0) * C + c0;
#pragma unroll
        for (int ow = 0; ow < 4; ow++) {
          if (w0 + ow >= W) break;
          const float2 f = unpack2(acc[oh][ow]);
          *reinterpret_cast<__nv_bfloat162*>(out + (int64_t)ow * C) = __floats2bfloat162_rn(f.x, f.y);
        }
      }
    }
  }
}

// Weight / bias gradient, same tiling: 49 packed accumulators per thread over every box the block walks; partial sums
// meet in shared memory (one shuffle + shared atomics), then one global atomic per tap and channel per block.
template <class Cfg>
__global__ void __launch_bounds__(Cfg::THREADS, 1)
    dwconv7_wgrad_tma_kernel(const __grid_constant__ CUtensorMap tmX, const __grid_constant__ CUtensorMap tmDy,
                             float* __restrict__ dw, float* __restrict__ db, int nbw, int nbh, int nboxes) {
  extern __shared__ __align__(128) uint8_t dw_smem[];
  __shared__ float red[50][32];
  __shared__ __align__(8) uint64_t bars[2];
  constexpr int STAGE = Cfg::XBUF + Cfg::DYBUF;
  const int slab = blockIdx.y * 32;
  const uint32_t xs = (uint32_t)__cvta_generic_to_shared(dw_smem);
  const uint32_t bar0 = (uint32_t)__cvta_generic_to_shared(bars);
  if (threadIdx.x == 0) {
    dw_mbar_init(bar0);
    dw_mbar_init(bar0 + 8);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    for (int k = 0; k < 2; k++) {
      const int bi = blockIdx.x + k * gridDim.x;
      if (bi < nboxes) {
        const DwBox o = dw_box<Cfg>(bi, nbw, nbh);
        dw_mbar_expect(bar0 + 8 * k, Cfg::XBYTES + Cfg::DYBYTES);
        dw_tma_4d(xs + k * STAGE, &tmX, bar0 + 8 * k, slab, o.w - 3, o.h - 3, o.b);
        dw_tma_4d(xs + k * STAGE + Cfg::XBUF, &tmDy, bar0 + 8 * k, slab, o.w, o.h, o.b);
      }
    }
  }
  for (int i = threadIdx.x; i < 50 * 32; i += blockDim.x) (&red[0][0])[i] = 0.f;
  __syncthreads();
  const int cp = threadIdx.x & 15;
  const DwSlot sl = dw_slot<Cfg>(threadIdx.x >> 4);
  const uint32_t x_off = (uint32_t)(((sl.nb * Cfg::HB + sl.nh * 7) * Cfg::WB + sl.nw * 4) * 64 + cp * 4);
  const uint32_t d_off = (uint32_t)(Cfg::XBUF + ((sl.nb * Cfg::DHB + sl.nh * 7) * Cfg::DWB + sl.nw * 4) * 64 + cp * 4);
  uint64_t acc[49], accb = 0ull;
#pragma unroll
  for (int k = 0; k < 49; k++) acc[k] = 0ull;
  const uint64_t one2 = pack2(1.f, 1.f);
  int it = 0;
  for (int bi = blockIdx.x; bi < nboxes; bi += gridDim.x, it++) {
    const int buf = it & 1;
    dw_mbar_wait(bar0 + 8 * buf, (it >> 1) & 1);
    const uint32_t xb = xs + buf * STAGE + x_off, dyb = xs + buf * STAGE + d_off;
    uint32_t d[7][4];     // dy outside the image is zero-filled by TMA, so ragged tiles add nothing
#pragma unroll
    for (int oh = 0; oh < 7; oh++)
#pragma unroll
      for (int ow = 0; ow < 4; ow++) {
        d[oh][ow] = lds32(dyb + (oh * Cfg::DWB + ow) * 64);
        accb = fma2(bf2_to_f2(d[oh][ow]), one2, accb);
      }
#pragma unroll
    for (int i = 0; i < 13; i++) {
      uint64_t xin[10];
#pragma unroll
      for (int j = 0; j < 10; j++) xin[j] = bf2_to_f2(lds32(xb + (i * Cfg::WB + j) * 64));
#pragma unroll
      for (int oh = 0; oh < 7; oh++) {
        const int r = i - oh;
        if (r < 0 || r > 6) continue;
#pragma unroll
        for (int ow = 0; ow < 4; ow++) {
          const uint64_t g = bf2_to_f2(d[oh][ow]);
#pragma unroll
          for (int s = 0; s < 7; s++) acc[r * 7 + s] = fma2(g, xin[ow + s], acc[r * 7 + s]);
        }
      }
    }
    __syncthreads();
    if (threadIdx.x == 0) {
      const int nb = bi + 2 * gridDim.x;
      if (nb < nboxes) {
        const DwBox n = dw_box<Cfg>(nb, nbw, nbh);
        dw_mbar_expect(bar0 + 8 * buf, Cfg::XBYTES + Cfg::DYBYTES);
        dw_tma_4d(xs + buf * STAGE, &tmX, bar0 + 8 * buf, slab, n.w - 3, n.h - 3, n.b);
        dw_tma_4d(xs + buf * STAGE + Cfg::XBUF, &tmDy, bar0 + 8 * buf, slab, n.w, n.h, n.b);
      }
    }
  }
  // lanes l and l + 16 hold the same channel pair
#pragma unroll
  for (int k = 0; k < 50; k++) {
    float2 f = unpack2(k < 49 ? acc[k] : accb);
    f.x += __shfl_xor_sync(0xffffffffu, f.x, 16);
    f.y += __shfl_xor_sync(0xffffffffu, f.y, 16);
    if ((threadIdx.x & 16) == 0) {
      atomicAdd(&red[k][cp * 2], f.x);
      atomicAdd(&red[k][cp * 2 + 1], f.y);
    }
  }
  __syncthreads();
  for (int i = threadIdx.x; i < 50 * 32; i += blockDim.x) {
    const int cl = i / 50, tap = i - cl * 50;
    const float v = red[tap][cl];
    if (tap < 49) atomicAdd(dw + (int64_t)(slab + cl) * 49 + tap, v);
    else if (db != nullptr) atomicAdd(db + slab + cl, v);
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

// dz = dy * ls[c] * keep(b);  dls[c] += sum_rows dy * z * keep(b);  dbias[c] += sum_rows dz (the bias gradient of the
// Linear that produced z).  Column-reduction layout of bn_bwd_reduce.
__global__ void __launch_bounds__(256) layer_scale_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ z,
                                                              const float* __restrict__ ls, bf16* __restrict__ dz,
                                                              float* __restrict__ dls, float* __restrict__ dbias, int64_t rows,
                                                              int C, int rows_per_sample, int rows_per_block, float p,
                                                              uint64_t seed) {
  __shared__ float sh[8][256 + 8];
  const int cv = threadIdx.x & 31, rl = threadIdx.x >> 5;
  const int c0 = blockIdx.x * 256 + cv * 8;
  const bool ok = c0 < C;
  const float inv_keep = p > 0.f ? 1.f / (1.f - p) : 1.f;
  float a[8], s[8], bsum[8];
#pragma unroll
  for (int k = 0; k < 8; k++) {
    a[k] = 0.f;
    bsum[k] = 0.f;
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
        bsum[k] += o[k];
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
  if (dbias == nullptr) return;   // uniform
  __syncthreads();
#pragma unroll
  for (int k = 0; k < 8; k++) sh[rl][cv * 8 + k] = bsum[k];
  __syncthreads();
  if (c < C) {
    float t = 0.f;
#pragma unroll
    for (int w8 = 0; w8 < 8; w8++) t += sh[w8][threadIdx.x];
    atomicAdd(dbias + c, t);
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

// MDHS_DWCONV_LEGACY=1 keeps the shared-memory-tap kernels (A/B timing); C % 32 != 0 always uses them.
bool dw_legacy() {
  static const bool v = [] { const char* e = getenv("MDHS_DWCONV_LEGACY"); return e && atoi(e) != 0; }();
  return v;
}

int grid_cap(int64_t items, int block) {
  int64_t g = (items + block - 1) / block;
  const int64_t cap = (int64_t)mdhs_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

}  // namespace

#define ST(s) reinterpret_cast<cudaStream_t>(s)

// ---- host side of the TMA-staged depthwise kernels
typedef CUresult (*DwEncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                               const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                               CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static DwEncodeFn dw_encode() {
  static DwEncodeFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<DwEncodeFn>(f);
  }();
  return fn;
}

// [B, H, W, C] bf16 as a 4-D tensor (C innermost); box = 32 channels x bw x bh x bb pixels, zeros outside the tensor.
static int dw_map(CUtensorMap* map, const void* ptr, int B, int H, int W, int C, int bw, int bh, int bb) {
  DwEncodeFn enc = dw_encode();
  if (!enc) return MDHS_ERR_DRIVER;
  cuuint64_t dims[4] = {(cuuint64_t)C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)B};
  cuuint64_t strides[3] = {(cuuint64_t)C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
  cuuint32_t box[4] = {32, (cuuint32_t)bw, (cuuint32_t)bh, (cuuint32_t)bb};
  cuuint32_t estr[4] = {1, 1, 1, 1};
  CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 4, const_cast<void*>(ptr), dims, strides, box, estr,
                   CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                   CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? MDHS_OK : MDHS_ERR_ARG;
}

// blocks per slab: about one block per SM over all slabs, every block of a slab walking the same number of boxes
static int dw_grid_x(int nboxes, int slabs) {
  int64_t gx = (mdhs_num_sms() + slabs - 1) / slabs;
  if (gx > nboxes) gx = nboxes;
  const int64_t iters = (nboxes + gx - 1) / gx;
  return (int)((nboxes + iters - 1) / iters);
}

template <class Cfg>
static int dw_fwd_launch(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int flip,
                         cudaStream_t st) {
  CUtensorMap tm;
  const int rc = dw_map(&tm, x, B, H, W, C, Cfg::WB, Cfg::HB, Cfg::NB);
  if (rc != MDHS_OK) return rc;
  const int nbw = (W + 4 * Cfg::NW - 1) / (4 * Cfg::NW), nbh = (H + 7 * Cfg::NH - 1) / (7 * Cfg::NH);
  const int64_t nboxes = (int64_t)nbw * nbh * ((B + Cfg::NB - 1) / Cfg::NB);
  if (nboxes > (1 << 30)) return MDHS_ERR_ARG;
  const int slabs = C / 32;
  const dim3 grid((unsigned)dw_grid_x((int)nboxes, slabs), (unsigned)slabs);
  constexpr int smem = 2 * Cfg::XBUF;
  static bool once[2] = {false, false};
  auto kf = dwconv7_tma_kernel<Cfg, false>;
  auto kt = dwconv7_tma_kernel<Cfg, true>;
  if (!once[flip ? 1 : 0]) {
    if (cudaFuncSetAttribute(flip ? kt : kf, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return (int)cudaGetLastError();
    once[flip ? 1 : 0] = true;
  }
  if (flip) kt<<<grid, Cfg::THREADS, smem, st>>>(tm, w, nullptr, (bf16*)y, B, H, W, C, nbw, nbh, (int)nboxes);
  else kf<<<grid, Cfg::THREADS, smem, st>>>(tm, w, bias, (bf16*)y, B, H, W, C, nbw, nbh, (int)nboxes);
  MDHS_RETURN_LAST();
}

template <class Cfg>
static int dw_wgrad_launch(const void* x, const void* dy, float* dw, float* db, int B, int H, int W, int C, cudaStream_t st) {
  CUtensorMap tx, td;
  int rc = dw_map(&tx, x, B, H, W, C, Cfg::WB, Cfg::HB, Cfg::NB);
  if (rc == MDHS_OK) rc = dw_map(&td, dy, B, H, W, C, Cfg::DWB, Cfg::DHB, Cfg::NB);
  if (rc != MDHS_OK) return rc;
  const int nbw = (W + 4 * Cfg::NW - 1) / (4 * Cfg::NW), nbh = (H + 7 * Cfg::NH - 1) / (7 * Cfg::NH);
  const int64_t nboxes = (int64_t)nbw * nbh * ((B + Cfg::NB - 1) / Cfg::NB);
  if (nboxes > (1 << 30)) return MDHS_ERR_ARG;
  const int slabs = C / 32;
  const dim3 grid((unsigned)dw_grid_x((int)nboxes, slabs), (unsigned)slabs);
  constexpr int smem = 2 * (Cfg::XBUF + Cfg::DYBUF);
  static bool once = false;
  auto k = dwconv7_wgrad_tma_kernel<Cfg>;
  if (!once) {
    if (cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, smem) != cudaSuccess) return (int)cudaGetLastError();
    once = true;
  }
  k<<<grid, Cfg::THREADS, smem, st>>>(tx, td, dw, db, nbw, nbh, (int)nboxes);
  MDHS_RETURN_LAST();
}

// box shapes by image width: 28 columns x 14 rows (56 / 28-pixel maps), 16 x 14 x 2 images (14), 8 x 7 x 8 images (7)
typedef DwCfg<7, 2, 1> DwWide;
typedef DwCfg<4, 2, 2> DwMid;
typedef DwCfg<2, 1, 8> DwSmall;
typedef DwCfg<2, 1, 4> DwSmallWgrad;   // x + dy boxes of 8 images would not fit twice in shared memory

extern "C" int mdhs_dwconv7_fwd(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int flip,
                                void* stream) {
  if (!x || !w || !y || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  if (C % 32 == 0 && !dw_legacy()) {
    if (W > 16) return dw_fwd_launch<DwWide>(x, w, bias, y, B, H, W, C, flip, ST(stream));
    if (W > 8) return dw_fwd_launch<DwMid>(x, w, bias, y, B, H, W, C, flip, ST(stream));
    return dw_fwd_launch<DwSmall>(x, w, bias, y, B, H, W, C, flip, ST(stream));
  }
  const int64_t groups = (int64_t)B * H * ((W + 3) / 4);
  const dim3 grid((unsigned)((groups + 31) / 32), (unsigned)((C + 63) / 64));
  if (flip) dwconv7_kernel<true><<<grid, 256, 0, ST(stream)>>>((const bf16*)x, w, nullptr, (bf16*)y, B, H, W, C);
  else dwconv7_kernel<false><<<grid, 256, 0, ST(stream)>>>((const bf16*)x, w, bias, (bf16*)y, B, H, W, C);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_dwconv7_wgrad(const void* x, const void* dy, float* dw, float* db, int B, int H, int W, int C, void* stream) {
  if (!x || !dy || !dw || B <= 0 || H <= 0 || W <= 0 || C <= 0 || (C % 8)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  if (C % 32 == 0 && !dw_legacy()) {
    if (W > 16) return dw_wgrad_launch<DwWide>(x, dy, dw, db, B, H, W, C, ST(stream));
    if (W > 8) return dw_wgrad_launch<DwMid>(x, dy, dw, db, B, H, W, C, ST(stream));
    return dw_wgrad_launch<DwSmallWgrad>(x, dy, dw, db, B, H, W, C, ST(stream));
  }
  const int slabs = (C + 63) / 64;
  const int64_t total = (int64_t)B * H * W;
  int chunks = (mdhs_num_sms() * 4 + slabs - 1) / slabs;
  if (chunks > total / 64) chunks = (int)(total / 64 > 0 ? total / 64 : 1);
  const int64_t ppb = (total + chunks - 1) / chunks;
  const dim3 grid((unsigned)((total + ppb - 1) / ppb), (unsigned)slabs);
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

extern "C" int mdhs_layer_scale_bwd(const void* dy, const void* z, const float* ls, void* dz, float* dls, float* dbias, int64_t rows, int C,
                                    int rows_per_sample, float p, uint64_t seed, void* stream) {
  if (!dy || !z || !ls || !dz || rows <= 0 || C <= 0 || (C % 8) || rows_per_sample <= 0 || p < 0.f || p >= 1.f) return MDHS_ERR_ARG;
  const int cslabs = (C + 255) / 256;
  int row_blocks = ((int64_t)mdhs_num_sms() * 8) / cslabs;
  if (row_blocks < 1) row_blocks = 1;
  int64_t rpb = (rows + row_blocks - 1) / row_blocks;
  rpb = ((rpb + 7) / 8) * 8;
  row_blocks = (int)((rows + rpb - 1) / rpb);
  g_mdhs_launches++;
  layer_scale_bwd_kernel<<<dim3(cslabs, row_blocks), 256, 0, ST(stream)>>>((const bf16*)dy, (const bf16*)z, ls, (bf16*)dz, dls, dbias, rows,
                                                                          C, rows_per_sample, (int)rpb, p, seed);
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
