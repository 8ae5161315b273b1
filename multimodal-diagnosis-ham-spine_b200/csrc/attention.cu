// Fused small-sequence attention (forward + backward) for every multi-head attention on the path:
// BERT self-attention (S <= 512, d = 64), nn.MultiheadAttention self/cross attention of the fusion
// blocks (d = 32; 49..784 queries, <= 512 keys).  Sequences are short, so one CTA keeps the whole
// K/V of one (batch, head) in shared memory and scale + key mask + softmax + dropout + P.V happen in
// one pass; probabilities are never written to HBM (the backward pass recomputes them from the saved
// log-sum-exp).  Buffers are token-major [B*S, ld] bf16 with head h in columns [h*D, (h+1)*D).
#include "common.cuh"
#include "../../include/mdhs_b200.h"

MDHS_DEFINE_SEED_TICK(attention)

extern int64_t g_mdhs_launches;

namespace {

constexpr int QT = 64;       // queries per CTA
constexpr int NWARPS = 4;

struct AttnParams {
  const bf16 *q, *k, *v, *o, *d_o;
  bf16 *out, *dq, *dk, *dv;
  float *dk32, *dv32;  // fp32 accumulation targets when several query tiles share one K/V
  int64_t ldq, ldk, ldv, ldo;
  const uint8_t* key_mask;
  float* lse;
  int B, H, Sq, Sk;
  float scale, drop_p;
  uint64_t seed;
};

template <int D>
__device__ __forceinline__ void load_row_f32(const bf16* p, float* r) {
#pragma unroll
  for (int v = 0; v < D / 8; v++) load8(p + v * 8, r + v * 8);
}

// padded shared rows are only 4-byte aligned: read them as bf16 pairs
template <int D>
__device__ __forceinline__ void load_row_smem(const bf16* p, float* r) {
  const bf162* pp = reinterpret_cast<const bf162*>(p);
#pragma unroll
  for (int w = 0; w < D / 2; w++) {
    const float2 v = __bfloat1622float2(pp[w]);
    r[2 * w] = v.x;
    r[2 * w + 1] = v.y;
  }
}

// cooperative copy of `rows` rows of D bf16 (global row stride ld) into padded shared rows of D+2
template <int D>
__device__ __forceinline__ void stage_rows(const bf16* g, int64_t ld, int rows, bf16* s) {
  constexpr int VPR = D / 8;
  for (int i = threadIdx.x; i < rows * VPR; i += blockDim.x) {
    const int r = i / VPR, v = i % VPR;
    const uint4 val = *reinterpret_cast<const uint4*>(g + (int64_t)r * ld + v * 8);
    uint32_t* dst = reinterpret_cast<uint32_t*>(s + r * (D + 2) + v * 8);  // (D+2)*2 bytes per row: 4-byte aligned
    dst[0] = val.x; dst[1] = val.y; dst[2] = val.z; dst[3] = val.w;
  }
}

template <int D>
__device__ __forceinline__ float dot_row(const float* qreg, const bf16* srow) {
  float acc = 0.f;
  const bf162* kp = reinterpret_cast<const bf162*>(srow);
#pragma unroll
  for (int w = 0; w < D / 2; w++) {
    const float2 kv = __bfloat1622float2(kp[w]);
    acc = fmaf(qreg[2 * w], kv.x, acc);
    acc = fmaf(qreg[2 * w + 1], kv.y, acc);
  }
  return acc;
}

template <int D, int KCH>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = sK + (size_t)p.Sk * (D + 2);
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.y * QT;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int DPL = D / 32;

  stage_rows<D>(p.k + (int64_t)b * p.Sk * p.ldk + h * D, p.ldk, p.Sk, sK);
  stage_rows<D>(p.v + (int64_t)b * p.Sk * p.ldv + h * D, p.ldv, p.Sk, sV);
  __syncthreads();

  bool valid[KCH];
#pragma unroll
  for (int c = 0; c < KCH; c++) {
    const int j = lane + 32 * c;
    valid[c] = j < p.Sk && (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.Sk + j] != 0);
  }
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const int q1 = min(p.Sq, q0 + QT);
  for (int i = q0 + warp; i < q1; i += NWARPS) {
    float qreg[D];
    load_row_f32<D>(p.q + ((int64_t)b * p.Sq + i) * p.ldq + h * D, qreg);
    float s[KCH];
    float mx = -INFINITY;
#pragma unroll
    for (int c = 0; c < KCH; c++) {
      s[c] = -INFINITY;
      if (valid[c]) s[c] = p.scale * dot_row<D>(qreg, sK + (lane + 32 * c) * (D + 2));
      mx = fmaxf(mx, s[c]);
    }
    mx = warp_max(mx);
    float sum = 0.f;
#pragma unroll
    for (int c = 0; c < KCH; c++) {
      s[c] = valid[c] ? __expf(s[c] - mx) : 0.f;
      sum += s[c];
    }
    sum = warp_sum(sum);
    const float inv = 1.f / sum;
    if (lane == 0 && p.lse) p.lse[(int64_t)bh * p.Sq + i] = mx + __logf(sum);
#pragma unroll
    for (int c = 0; c < KCH; c++) {
      s[c] *= inv;
      if (p.drop_p > 0.f)
        s[c] *= dropout_scale(p.seed, ((uint64_t)bh * p.Sq + i) * p.Sk + lane + 32 * c, p.drop_p, inv_keep);
    }
    float acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; d++) acc[d] = 0.f;
#pragma unroll
    for (int c = 0; c < KCH; c++) {
      const int jn = min(32, p.Sk - 32 * c);
      for (int l = 0; l < jn; l++) {
        const float pj = __shfl_sync(0xffffffffu, s[c], l);
        const bf16* vr = sV + (l + 32 * c) * (D + 2) + lane * DPL;
        if (DPL == 2) {
          const float2 vv = __bfloat1622float2(*reinterpret_cast<const bf162*>(vr));
          acc[0] = fmaf(pj, vv.x, acc[0]);
          acc[DPL - 1] = fmaf(pj, vv.y, acc[DPL - 1]);
        } else {
          acc[0] = fmaf(pj, __bfloat162float(vr[0]), acc[0]);
        }
      }
    }
    bf16* op = p.out + ((int64_t)b * p.Sq + i) * p.ldo + h * D + lane * DPL;
    if (DPL == 2) *reinterpret_cast<bf162*>(op) = __floats2bfloat162_rn(acc[0], acc[DPL - 1]);
    else op[0] = __float2bfloat16_rn(acc[0]);
  }
}

template <int D, int KCH>
__global__ void __launch_bounds__(NWARPS * 32) attn_bwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = sK + (size_t)p.Sk * (D + 2);
  bf16* sQ = sV + (size_t)p.Sk * (D + 2);
  bf16* sdO = sQ + QT * (D + 2);
  float* sLse = reinterpret_cast<float*>(sdO + QT * (D + 2));
  float* sDelta = sLse + QT;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.y * QT;
  const int nq = min(QT, p.Sq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int DPL = D / 32;
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;

  stage_rows<D>(p.k + (int64_t)b * p.Sk * p.ldk + h * D, p.ldk, p.Sk, sK);
  stage_rows<D>(p.v + (int64_t)b * p.Sk * p.ldv + h * D, p.ldv, p.Sk, sV);
  stage_rows<D>(p.q + ((int64_t)b * p.Sq + q0) * p.ldq + h * D, p.ldq, nq, sQ);
  stage_rows<D>(p.d_o + ((int64_t)b * p.Sq + q0) * p.ldo + h * D, p.ldo, nq, sdO);
  // delta_i = dO_i . O_i  (equals sum_j P_ij dP_ij, also with dropout)
  for (int i = warp; i < nq; i += NWARPS) {
    const bf16* op = p.o + ((int64_t)b * p.Sq + q0 + i) * p.ldo + h * D + lane * DPL;
    const bf16* dp = p.d_o + ((int64_t)b * p.Sq + q0 + i) * p.ldo + h * D + lane * DPL;
    float a = 0.f;
#pragma unroll
    for (int d = 0; d < DPL; d++) a += __bfloat162float(op[d]) * __bfloat162float(dp[d]);
    a = warp_sum(a);
    if (lane == 0) {
      sDelta[i] = a;
      sLse[i] = p.lse[(int64_t)bh * p.Sq + q0 + i];
    }
  }
  __syncthreads();

  bool valid[KCH];
#pragma unroll
  for (int c = 0; c < KCH; c++) {
    const int j = lane + 32 * c;
    valid[c] = j < p.Sk && (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.Sk + j] != 0);
  }

  // ---- phase A: one warp per query -> dQ
  for (int i = warp; i < nq; i += NWARPS) {
    float reg[D];
    load_row_smem<D>(sQ + i * (D + 2), reg);
    float pr[KCH], ds[KCH];
    const float lse = sLse[i], delta = sDelta[i];
#pragma unroll
    for (int c = 0; c < KCH; c++) {
      pr[c] = 0.f;
      if (valid[c]) pr[c] = __expf(p.scale * dot_row<D>(reg, sK + (lane + 32 * c) * (D + 2)) - lse);
    }
    load_row_smem<D>(sdO + i * (D + 2), reg);
#pragma unroll
    for (int c = 0; c < KCH; c++) {
      ds[c] = 0.f;
      if (valid[c]) {
        float dpj = dot_row<D>(reg, sV + (lane + 32 * c) * (D + 2));
        if (p.drop_p > 0.f)
          dpj *= dropout_scale(p.seed, ((uint64_t)bh * p.Sq + q0 + i) * p.Sk + lane + 32 * c, p.drop_p, inv_keep);
        ds[c] = pr[c] * (dpj - delta);
      }
    }
    float acc[DPL];
#pragma unroll
    for (int d = 0; d < DPL; d++) acc[d] = 0.f;
#pragma unroll
    for (int c = 0; c < KCH; c++) {
      const int jn = min(32, p.Sk - 32 * c);
      for (int l = 0; l < jn; l++) {
        const float dj = __shfl_sync(0xffffffffu, ds[c], l);
        const bf16* kr = sK + (l + 32 * c) * (D + 2) + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; d++) acc[d] = fmaf(dj, __bfloat162float(kr[d]), acc[d]);
      }
    }
    bf16* qp = p.dq + ((int64_t)b * p.Sq + q0 + i) * p.ldq + h * D + lane * DPL;
#pragma unroll
    for (int d = 0; d < DPL; d++) qp[d] = __float2bfloat16_rn(acc[d] * p.scale);
  }

  // ---- phase B: one warp per key, lanes over the tile's queries -> dK, dV
  const bool multi_tile = gridDim.y > 1;
  for (int j = warp; j < p.Sk; j += NWARPS) {
    const bool key_ok = (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.Sk + j] != 0);
    float pd[QT / 32], ds[QT / 32];
    float kreg[D];
    load_row_smem<D>(sK + j * (D + 2), kreg);
#pragma unroll
    for (int c = 0; c < QT / 32; c++) {
      const int i = lane + 32 * c;
      pd[c] = 0.f;
      ds[c] = 0.f;
      if (key_ok && i < nq) pd[c] = __expf(p.scale * dot_row<D>(kreg, sQ + i * (D + 2)) - sLse[i]);
    }
    load_row_smem<D>(sV + j * (D + 2), kreg);
#pragma unroll
    for (int c = 0; c < QT / 32; c++) {
      const int i = lane + 32 * c;
      if (key_ok && i < nq) {
        float dpj = dot_row<D>(kreg, sdO + i * (D + 2));
        float m = 1.f;
        if (p.drop_p > 0.f) m = dropout_scale(p.seed, ((uint64_t)bh * p.Sq + q0 + i) * p.Sk + j, p.drop_p, inv_keep);
        ds[c] = pd[c] * (dpj * m - sDelta[i]);
        pd[c] *= m;
      }
    }
    float accv[DPL], acck[DPL];
#pragma unroll
    for (int d = 0; d < DPL; d++) accv[d] = acck[d] = 0.f;
#pragma unroll
    for (int c = 0; c < QT / 32; c++) {
      const int in = min(32, nq - 32 * c);
      for (int l = 0; l < in; l++) {
        const float pv = __shfl_sync(0xffffffffu, pd[c], l);
        const float dsv = __shfl_sync(0xffffffffu, ds[c], l);
        const bf16* dor = sdO + (l + 32 * c) * (D + 2) + lane * DPL;
        const bf16* qr = sQ + (l + 32 * c) * (D + 2) + lane * DPL;
#pragma unroll
        for (int d = 0; d < DPL; d++) {
          accv[d] = fmaf(pv, __bfloat162float(dor[d]), accv[d]);
          acck[d] = fmaf(dsv, __bfloat162float(qr[d]), acck[d]);
        }
      }
    }
    const int64_t krow = (int64_t)b * p.Sk + j;
    if (multi_tile) {
#pragma unroll
      for (int d = 0; d < DPL; d++) {
        atomicAdd(p.dk32 + krow * p.ldk + h * D + lane * DPL + d, acck[d] * p.scale);
        atomicAdd(p.dv32 + krow * p.ldv + h * D + lane * DPL + d, accv[d]);
      }
    } else {
#pragma unroll
      for (int d = 0; d < DPL; d++) {
        p.dk[krow * p.ldk + h * D + lane * DPL + d] = __float2bfloat16_rn(acck[d] * p.scale);
        p.dv[krow * p.ldv + h * D + lane * DPL + d] = __float2bfloat16_rn(accv[d]);
      }
    }
  }
}

template <int D>
size_t fwd_smem(int Sk) { return (size_t)2 * Sk * (D + 2) * 2; }
template <int D>
size_t bwd_smem(int Sk) { return (size_t)2 * Sk * (D + 2) * 2 + (size_t)2 * QT * (D + 2) * 2 + 2 * QT * 4; }

template <int D, int KCH>
int launch_fwd(const AttnParams& p, cudaStream_t st) {
  const size_t sm = fwd_smem<D>(p.Sk);
  auto kern = attn_fwd_kernel<D, KCH>;
  static size_t configured = 0;
  if (sm > 48 * 1024 && sm > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured = sm;
  }
  kern<<<dim3(p.B * p.H, ceil_div(p.Sq, QT)), NWARPS * 32, sm, st>>>(p);
  MDHS_RETURN_LAST();
}
template <int D, int KCH>
int launch_bwd(const AttnParams& p, cudaStream_t st) {
  const size_t sm = bwd_smem<D>(p.Sk);
  auto kern = attn_bwd_kernel<D, KCH>;
  static size_t configured = 0;
  if (sm > 48 * 1024 && sm > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured = sm;
  }
  kern<<<dim3(p.B * p.H, ceil_div(p.Sq, QT)), NWARPS * 32, sm, st>>>(p);
  MDHS_RETURN_LAST();
}

template <int D>
int dispatch(const AttnParams& p, bool bwd, cudaStream_t st) {
  const int kch = ceil_div(p.Sk, 32);
  if (kch <= 2) return bwd ? launch_bwd<D, 2>(p, st) : launch_fwd<D, 2>(p, st);
  if (kch <= 4) return bwd ? launch_bwd<D, 4>(p, st) : launch_fwd<D, 4>(p, st);
  if (kch <= 8) return bwd ? launch_bwd<D, 8>(p, st) : launch_fwd<D, 8>(p, st);
  return bwd ? launch_bwd<D, 16>(p, st) : launch_fwd<D, 16>(p, st);
}

}  // namespace

extern "C" int mdhs_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                                  int64_t ldo, const uint8_t* key_mask, float* lse, int B, int H, int Sq, int Sk, int D,
                                  float scale, float drop_p, uint64_t seed, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0 || Sk > 512) return MDHS_ERR_ARG;
  if ((D != 32 && D != 64) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 2)) return MDHS_ERR_ARG;
  AttnParams p{};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.out = (bf16*)out;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.key_mask = key_mask; p.lse = lse;
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.drop_p = drop_p; p.seed = seed;
  g_mdhs_launches++;
  return D == 64 ? dispatch<64>(p, false, reinterpret_cast<cudaStream_t>(stream))
                 : dispatch<32>(p, false, reinterpret_cast<cudaStream_t>(stream));
}

// dk/dv: bf16 outputs (same strides as k/v).  When Sq > 64 several query tiles contribute to each key, so
// the caller must also pass zero-initialised fp32 workspaces dk32/dv32 ([B*Sk, ldk] / [B*Sk, ldv]); the
// bf16 dk/dv are then produced by mdhs_cast_f32_bf16 on the host side.
extern "C" int mdhs_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  const void* o, const void* d_o, int64_t ldo, const uint8_t* key_mask, const float* lse,
                                  void* dq, void* dk, void* dv, float* dk32, float* dv32, int B, int H, int Sq, int Sk, int D,
                                  float scale, float drop_p, uint64_t seed, void* stream) {
  if (!q || !k || !v || !o || !d_o || !lse || !dq || B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0 || Sk > 512) return MDHS_ERR_ARG;
  if ((D != 32 && D != 64) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return MDHS_ERR_ARG;
  const bool multi = Sq > QT;
  if (multi ? (!dk32 || !dv32) : (!dk || !dv)) return MDHS_ERR_ARG;
  AttnParams p{};
  p.q = (const bf16*)q; p.k = (const bf16*)k; p.v = (const bf16*)v; p.o = (const bf16*)o; p.d_o = (const bf16*)d_o;
  p.dq = (bf16*)dq; p.dk = (bf16*)dk; p.dv = (bf16*)dv; p.dk32 = dk32; p.dv32 = dv32;
  p.ldq = ldq; p.ldk = ldk; p.ldv = ldv; p.ldo = ldo;
  p.key_mask = key_mask; p.lse = const_cast<float*>(lse);
  p.B = B; p.H = H; p.Sq = Sq; p.Sk = Sk; p.scale = scale; p.drop_p = drop_p; p.seed = seed;
  g_mdhs_launches++;
  return D == 64 ? dispatch<64>(p, true, reinterpret_cast<cudaStream_t>(stream))
                 : dispatch<32>(p, true, reinterpret_cast<cudaStream_t>(stream));
}
