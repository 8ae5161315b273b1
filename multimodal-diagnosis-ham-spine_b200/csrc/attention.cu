// Fused small-sequence attention (forward + backward) on the tensor cores for every multi-head attention on the
// path: BERT self-attention (S <= 512, d = 64), nn.MultiheadAttention self/cross attention of the fusion blocks
// (d = 32; 49..784 queries, <= 512 keys).  One CTA = one (batch, head, 64-query tile); the whole K/V of that head
// stays in shared memory; each warp owns 16 query rows: S = Q K^T and O = P V (and the five products of the
// backward) are mma.sync m16n8k16 bf16 tiles fed by ldmatrix, with scale + key mask + online softmax + dropout in
// registers between them.  Probabilities never reach HBM (the backward recomputes them from the saved
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

constexpr int KB = 64;       // keys per inner block
constexpr int PSTR = KB + 8; // row stride (elements) of the P / dS exchange tiles

// ---------------------------------------------------------------- warp-level tensor-core primitives
// mma.sync m16n8k16 (bf16 x bf16 -> fp32).  The per-(batch, head) problems here are 64 x 64 x {32,64} blocks:
// far below one tcgen05 128 x N tile, and latency- rather than throughput-bound, so the warp-synchronous MMA with
// register-resident accumulators (no TMEM round trip for the softmax) is the right instrument.
__device__ __forceinline__ void mma16816(float* c, const uint32_t* a, uint32_t b0, uint32_t b1) {
  asm volatile(
      "mma.sync.aligned.m16n8k16.row.col.f32.bf16.bf16.f32 {%0, %1, %2, %3}, {%4, %5, %6, %7}, {%8, %9}, {%0, %1, %2, %3};"
      : "+f"(c[0]), "+f"(c[1]), "+f"(c[2]), "+f"(c[3])
      : "r"(a[0]), "r"(a[1]), "r"(a[2]), "r"(a[3]), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void ldsm4(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ void ldsm4t(uint32_t* r, uint32_t addr) {
  asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0, %1, %2, %3}, [%4];"
               : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
               : "r"(addr));
}
__device__ __forceinline__ uint32_t sm_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t pack2(float a, float b) {
  bf162 v = __floats2bfloat162_rn(a, b);
  return *reinterpret_cast<uint32_t*>(&v);
}

// A fragment (16 rows x 16 k) of a row-major smem matrix X[m][k] with row stride `ld` elements
__device__ __forceinline__ void lda_rowmajor(uint32_t* a, const bf16* X, int ld, int m0, int k0, int lane) {
  ldsm4(a, sm_u32(X + (m0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + k0 + (lane >> 4) * 8));
}
// B fragments of two adjacent n-tiles (n0..n0+15) for C = A . X^T with X[n][k] row-major: r0,r1 -> tile n0; r2,r3 -> n0+8
__device__ __forceinline__ void ldb_nk(uint32_t* r, const bf16* X, int ld, int n0, int k0, int lane) {
  ldsm4(r, sm_u32(X + (n0 + (lane & 7) + (lane >> 4) * 8) * ld + k0 + ((lane >> 3) & 1) * 8));
}
// B fragments of two adjacent n-tiles for C = A . X with X[k][n] row-major (transposing load)
__device__ __forceinline__ void ldb_kn(uint32_t* r, const bf16* X, int ld, int k0, int n0, int lane) {
  ldsm4t(r, sm_u32(X + (k0 + (lane & 7) + ((lane >> 3) & 1) * 8) * ld + n0 + (lane >> 4) * 8));
}
// A fragment (16 m x 16 k) of X^T with X[k][m] row-major (transposing load)
__device__ __forceinline__ void lda_trans(uint32_t* a, const bf16* X, int ld, int k0, int m0, int lane) {
  ldsm4t(a, sm_u32(X + (k0 + (lane & 7) + (lane >> 4) * 8) * ld + m0 + ((lane >> 3) & 1) * 8));
}

// cooperative ASYNCHRONOUS copy (cp.async, 16 bytes each, zero-fill for rows [rows, rows_pad)) of `rows` rows of D bf16
// (global row stride ld) into smem rows of D+8.  All tiles of a CTA (K, V, Q, dO) are requested back to back and waited for
// once (stage_wait): one memory latency per CTA instead of one per tensor -- these kernels are latency-bound (ncu: 17-22 %
// occupancy, 1 TB/s).
template <int D>
__device__ __forceinline__ void stage_rows(const bf16* g, int64_t ld, int rows, int rows_pad, bf16* s) {
  constexpr int VPR = D / 8;
  for (int i = threadIdx.x; i < rows_pad * VPR; i += blockDim.x) {
    const int r = i / VPR, v = i % VPR;
    const bool ok = r < rows;
    const bf16* src = ok ? g + (int64_t)r * ld + v * 8 : g;
    const uint32_t dst = sm_u32(s + r * (D + 8) + v * 8);
    const int nbytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
}
__device__ __forceinline__ void stage_wait() {
  asm volatile("cp.async.commit_group;" ::: "memory");
  asm volatile("cp.async.wait_group 0;" ::: "memory");
}

// dropout keep-scales of two consecutive elements (idx, idx + 1)
__device__ __forceinline__ void dropout_pair(uint64_t seed, uint64_t idx, float p, float inv_keep, float& m0, float& m1) {
  const float thr = p * 65536.f;
  const uint64_t b = dropout_bits4(seed, idx >> 2);
  const int sh = (int)(idx & 3) * 16;
  m0 = ((float)((uint32_t)(b >> sh) & 0xffffu) < thr) ? 0.f : inv_keep;
  if ((idx & 3) != 3) {
    m1 = ((float)((uint32_t)(b >> (sh + 16)) & 0xffffu) < thr) ? 0.f : inv_keep;
  } else {
    m1 = ((float)((uint32_t)dropout_bits4(seed, (idx + 1) >> 2) & 0xffffu) < thr) ? 0.f : inv_keep;
  }
}

// additive key mask (0 / -inf) for the padded key range
__device__ __forceinline__ void stage_mask(const AttnParams& p, int b, int skp, float* sMask) {
  for (int j = threadIdx.x; j < skp; j += blockDim.x) {
    const bool ok = j < p.Sk && (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.Sk + j] != 0);
    sMask[j] = ok ? 0.f : -INFINITY;
  }
}

template <int D>
__global__ void __launch_bounds__(NWARPS * 32) attn_fwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int LD = D + 8;
  const int skp = (p.Sk + KB - 1) / KB * KB;
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = sK + (size_t)skp * LD;
  bf16* sQ = sV + (size_t)skp * LD;
  float* sMask = reinterpret_cast<float*>(sQ + QT * LD);
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.y * QT;
  const int nq = min(QT, p.Sq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;

  stage_rows<D>(p.k + (int64_t)b * p.Sk * p.ldk + h * D, p.ldk, p.Sk, skp, sK);
  stage_rows<D>(p.v + (int64_t)b * p.Sk * p.ldv + h * D, p.ldv, p.Sk, skp, sV);
  stage_rows<D>(p.q + ((int64_t)b * p.Sq + q0) * p.ldq + h * D, p.ldq, nq, QT, sQ);
  stage_mask(p, b, skp, sMask);
  stage_wait();
  __syncthreads();

  const int r0 = warp * 16;
  if (r0 >= nq) return;
  uint32_t qa[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; kk++) lda_rowmajor(qa[kk], sQ, LD, r0, kk * 16, lane);
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;
  const uint64_t rowbase0 = ((uint64_t)bh * p.Sq + q0 + r0 + g) * p.Sk;
  const uint64_t rowbase1 = rowbase0 + (uint64_t)8 * p.Sk;

  for (int kb = 0; kb < skp; kb += KB) {
    float s[KB / 8][4];
#pragma unroll
    for (int i = 0; i < KB / 8; i++) s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
#pragma unroll
    for (int kk = 0; kk < D / 16; kk++) {
#pragma unroll
      for (int np = 0; np < KB / 16; np++) {
        uint32_t bb[4];
        ldb_nk(bb, sK, LD, kb + np * 16, kk * 16, lane);
        mma16816(s[2 * np], qa[kk], bb[0], bb[1]);
        mma16816(s[2 * np + 1], qa[kk], bb[2], bb[3]);
      }
    }
    float mx[2] = {-INFINITY, -INFINITY};
#pragma unroll
    for (int nt = 0; nt < KB / 8; nt++) {
      const float2 mk = *reinterpret_cast<const float2*>(sMask + kb + nt * 8 + 2 * tig);
      s[nt][0] = s[nt][0] * p.scale + mk.x;
      s[nt][1] = s[nt][1] * p.scale + mk.y;
      s[nt][2] = s[nt][2] * p.scale + mk.x;
      s[nt][3] = s[nt][3] * p.scale + mk.y;
      mx[0] = fmaxf(mx[0], fmaxf(s[nt][0], s[nt][1]));
      mx[1] = fmaxf(mx[1], fmaxf(s[nt][2], s[nt][3]));
    }
    float corr[2], muse[2];
#pragma unroll
    for (int r = 0; r < 2; r++) {
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 1));
      mx[r] = fmaxf(mx[r], __shfl_xor_sync(0xffffffffu, mx[r], 2));
      const float mnew = fmaxf(mrow[r], mx[r]);
      muse[r] = (mnew == -INFINITY) ? 0.f : mnew;
      corr[r] = __expf(mrow[r] - muse[r]);   // exp(-inf) = 0 on the first block
      mrow[r] = mnew;
      lrow[r] *= corr[r];
    }
#pragma unroll
    for (int i = 0; i < D / 8; i++) {
      o[i][0] *= corr[0];
      o[i][1] *= corr[0];
      o[i][2] *= corr[1];
      o[i][3] *= corr[1];
    }
    uint32_t pa[KB / 16][4];
#pragma unroll
    for (int nt = 0; nt < KB / 8; nt++) {
      float e0 = __expf(s[nt][0] - muse[0]), e1 = __expf(s[nt][1] - muse[0]);
      float e2 = __expf(s[nt][2] - muse[1]), e3 = __expf(s[nt][3] - muse[1]);
      lrow[0] += e0 + e1;
      lrow[1] += e2 + e3;
      if (p.drop_p > 0.f) {
        const uint64_t col = (uint64_t)(kb + nt * 8 + 2 * tig);
        float m0, m1;
        dropout_pair(p.seed, rowbase0 + col, p.drop_p, inv_keep, m0, m1);
        e0 *= m0;
        e1 *= m1;
        dropout_pair(p.seed, rowbase1 + col, p.drop_p, inv_keep, m0, m1);
        e2 *= m0;
        e3 *= m1;
      }
      pa[nt >> 1][(nt & 1) * 2] = pack2(e0, e1);
      pa[nt >> 1][(nt & 1) * 2 + 1] = pack2(e2, e3);
    }
#pragma unroll
    for (int kk = 0; kk < KB / 16; kk++) {
#pragma unroll
      for (int dp = 0; dp < D / 16; dp++) {
        uint32_t bb[4];
        ldb_kn(bb, sV, LD, kb + kk * 16, dp * 16, lane);
        mma16816(o[2 * dp], pa[kk], bb[0], bb[1]);
        mma16816(o[2 * dp + 1], pa[kk], bb[2], bb[3]);
      }
    }
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 1);
    lrow[r] += __shfl_xor_sync(0xffffffffu, lrow[r], 2);
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int row = r0 + g + r * 8;
    if (row < nq) {
      const float inv = 1.f / lrow[r];
      bf16* op = p.out + ((int64_t)b * p.Sq + q0 + row) * p.ldo + h * D + 2 * tig;
#pragma unroll
      for (int i = 0; i < D / 8; i++)
        *reinterpret_cast<bf162*>(op + i * 8) = __floats2bfloat162_rn(o[i][2 * r] * inv, o[i][2 * r + 1] * inv);
      if (tig == 0 && p.lse) p.lse[(int64_t)bh * p.Sq + q0 + row] = mrow[r] + __logf(lrow[r]);
    }
  }
}

template <int D>
__global__ void __launch_bounds__(NWARPS * 32) attn_bwd_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int LD = D + 8;
  const int skp = (p.Sk + KB - 1) / KB * KB;
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = sK + (size_t)skp * LD;
  bf16* sQ = sV + (size_t)skp * LD;
  bf16* sdO = sQ + QT * LD;
  bf16* sP = sdO + QT * LD;
  bf16* sdS = sP + QT * PSTR;
  float* sMask = reinterpret_cast<float*>(sdS + QT * PSTR);
  float* sLse = sMask + skp;
  float* sDelta = sLse + QT;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.y * QT;
  const int nq = min(QT, p.Sq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const float inv_keep = p.drop_p > 0.f ? 1.f / (1.f - p.drop_p) : 1.f;

  stage_rows<D>(p.k + (int64_t)b * p.Sk * p.ldk + h * D, p.ldk, p.Sk, skp, sK);
  stage_rows<D>(p.v + (int64_t)b * p.Sk * p.ldv + h * D, p.ldv, p.Sk, skp, sV);
  stage_rows<D>(p.q + ((int64_t)b * p.Sq + q0) * p.ldq + h * D, p.ldq, nq, QT, sQ);
  stage_rows<D>(p.d_o + ((int64_t)b * p.Sq + q0) * p.ldo + h * D, p.ldo, nq, QT, sdO);
  stage_mask(p, b, skp, sMask);
  // delta_i = dO_i . O_i (= sum_j P_ij dP_ij, also under dropout); rows past the tile end get P = 0 through lse = +inf
  {
    constexpr int LPR = D / 8;            // lanes per row (8 elements each)
    constexpr int RPW = 32 / LPR;         // rows per warp pass
    for (int i = warp * RPW + lane / LPR; i < QT; i += NWARPS * RPW) {
      float a = 0.f;
      if (i < nq) {
        float x[8], y[8];
        const int64_t off = ((int64_t)b * p.Sq + q0 + i) * p.ldo + h * D + (lane % LPR) * 8;
        load8(p.o + off, x);
        load8(p.d_o + off, y);
#pragma unroll
        for (int d = 0; d < 8; d++) a = fmaf(x[d], y[d], a);
      }
#pragma unroll
      for (int o = LPR / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
      if ((lane % LPR) == 0) {
        sDelta[i] = a;
        sLse[i] = i < nq ? p.lse[(int64_t)bh * p.Sq + q0 + i] : INFINITY;
      }
    }
  }
  stage_wait();
  __syncthreads();

  const int r0 = warp * 16;
  uint32_t qa[D / 16][4], doa[D / 16][4];
#pragma unroll
  for (int kk = 0; kk < D / 16; kk++) {
    lda_rowmajor(qa[kk], sQ, LD, r0, kk * 16, lane);
    lda_rowmajor(doa[kk], sdO, LD, r0, kk * 16, lane);
  }
  const float lse0 = sLse[r0 + g], lse1 = sLse[r0 + g + 8];
  const float dl0 = sDelta[r0 + g], dl1 = sDelta[r0 + g + 8];
  float dq[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; i++) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const uint64_t rowbase0 = ((uint64_t)bh * p.Sq + q0 + r0 + g) * p.Sk;
  const uint64_t rowbase1 = rowbase0 + (uint64_t)8 * p.Sk;
  const bool multi_tile = gridDim.y > 1;

  for (int kb = 0; kb < skp; kb += KB) {
    float s[KB / 8][4], dp[KB / 8][4];
#pragma unroll
    for (int i = 0; i < KB / 8; i++) {
      s[i][0] = s[i][1] = s[i][2] = s[i][3] = 0.f;
      dp[i][0] = dp[i][1] = dp[i][2] = dp[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < D / 16; kk++) {
#pragma unroll
      for (int np = 0; np < KB / 16; np++) {
        uint32_t bb[4];
        ldb_nk(bb, sK, LD, kb + np * 16, kk * 16, lane);
        mma16816(s[2 * np], qa[kk], bb[0], bb[1]);
        mma16816(s[2 * np + 1], qa[kk], bb[2], bb[3]);
        ldb_nk(bb, sV, LD, kb + np * 16, kk * 16, lane);
        mma16816(dp[2 * np], doa[kk], bb[0], bb[1]);
        mma16816(dp[2 * np + 1], doa[kk], bb[2], bb[3]);
      }
    }
    uint32_t dsa[KB / 16][4];
#pragma unroll
    for (int nt = 0; nt < KB / 8; nt++) {
      const float2 mk = *reinterpret_cast<const float2*>(sMask + kb + nt * 8 + 2 * tig);
      float p0 = __expf(s[nt][0] * p.scale + mk.x - lse0), p1 = __expf(s[nt][1] * p.scale + mk.y - lse0);
      float p2 = __expf(s[nt][2] * p.scale + mk.x - lse1), p3 = __expf(s[nt][3] * p.scale + mk.y - lse1);
      float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
      if (p.drop_p > 0.f) {
        const uint64_t col = (uint64_t)(kb + nt * 8 + 2 * tig);
        dropout_pair(p.seed, rowbase0 + col, p.drop_p, inv_keep, m0, m1);
        dropout_pair(p.seed, rowbase1 + col, p.drop_p, inv_keep, m2, m3);
      }
      const float d0 = p0 * (dp[nt][0] * m0 - dl0), d1 = p1 * (dp[nt][1] * m1 - dl0);
      const float d2 = p2 * (dp[nt][2] * m2 - dl1), d3 = p3 * (dp[nt][3] * m3 - dl1);
      const uint32_t ds01 = pack2(d0, d1), ds23 = pack2(d2, d3);
      dsa[nt >> 1][(nt & 1) * 2] = ds01;
      dsa[nt >> 1][(nt & 1) * 2 + 1] = ds23;
      const int c = nt * 8 + 2 * tig;
      *reinterpret_cast<uint32_t*>(sP + (r0 + g) * PSTR + c) = pack2(p0 * m0, p1 * m1);
      *reinterpret_cast<uint32_t*>(sP + (r0 + g + 8) * PSTR + c) = pack2(p2 * m2, p3 * m3);
      *reinterpret_cast<uint32_t*>(sdS + (r0 + g) * PSTR + c) = ds01;
      *reinterpret_cast<uint32_t*>(sdS + (r0 + g + 8) * PSTR + c) = ds23;
    }
    // dQ += dS . K_blk
#pragma unroll
    for (int kk = 0; kk < KB / 16; kk++) {
#pragma unroll
      for (int dd = 0; dd < D / 16; dd++) {
        uint32_t bb[4];
        ldb_kn(bb, sK, LD, kb + kk * 16, dd * 16, lane);
        mma16816(dq[2 * dd], dsa[kk], bb[0], bb[1]);
        mma16816(dq[2 * dd + 1], dsa[kk], bb[2], bb[3]);
      }
    }
    __syncthreads();
    // this warp's 16 keys of the block: dV = Pd^T . dO, dK = dS^T . Q (reduction over the tile's 64 queries)
    float dv[D / 8][4], dk[D / 8][4];
#pragma unroll
    for (int i = 0; i < D / 8; i++) {
      dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
      dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
    }
#pragma unroll
    for (int kk = 0; kk < QT / 16; kk++) {
      uint32_t ap[4], ads[4];
      lda_trans(ap, sP, PSTR, kk * 16, r0, lane);
      lda_trans(ads, sdS, PSTR, kk * 16, r0, lane);
#pragma unroll
      for (int dd = 0; dd < D / 16; dd++) {
        uint32_t bb[4];
        ldb_kn(bb, sdO, LD, kk * 16, dd * 16, lane);
        mma16816(dv[2 * dd], ap, bb[0], bb[1]);
        mma16816(dv[2 * dd + 1], ap, bb[2], bb[3]);
        ldb_kn(bb, sQ, LD, kk * 16, dd * 16, lane);
        mma16816(dk[2 * dd], ads, bb[0], bb[1]);
        mma16816(dk[2 * dd + 1], ads, bb[2], bb[3]);
      }
    }
#pragma unroll
    for (int r = 0; r < 2; r++) {
      const int key = kb + r0 + g + r * 8;
      if (key < p.Sk) {
        const int64_t krow = (int64_t)b * p.Sk + key;
#pragma unroll
        for (int i = 0; i < D / 8; i++) {
          const int c = h * D + i * 8 + 2 * tig;
          if (multi_tile) {
            atomicAdd(p.dk32 + krow * p.ldk + c, dk[i][2 * r] * p.scale);
            atomicAdd(p.dk32 + krow * p.ldk + c + 1, dk[i][2 * r + 1] * p.scale);
            atomicAdd(p.dv32 + krow * p.ldv + c, dv[i][2 * r]);
            atomicAdd(p.dv32 + krow * p.ldv + c + 1, dv[i][2 * r + 1]);
          } else {
            *reinterpret_cast<bf162*>(p.dk + krow * p.ldk + c) = __floats2bfloat162_rn(dk[i][2 * r] * p.scale, dk[i][2 * r + 1] * p.scale);
            *reinterpret_cast<bf162*>(p.dv + krow * p.ldv + c) = __floats2bfloat162_rn(dv[i][2 * r], dv[i][2 * r + 1]);
          }
        }
      }
    }
    __syncthreads();
  }
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int row = r0 + g + r * 8;
    if (row < nq) {
      bf16* qp = p.dq + ((int64_t)b * p.Sq + q0 + row) * p.ldq + h * D + 2 * tig;
#pragma unroll
      for (int i = 0; i < D / 8; i++)
        *reinterpret_cast<bf162*>(qp + i * 8) = __floats2bfloat162_rn(dq[i][2 * r] * p.scale, dq[i][2 * r + 1] * p.scale);
    }
  }
}

template <int D>
size_t fwd_smem(int Sk) {
  const int skp = (Sk + KB - 1) / KB * KB;
  return (size_t)2 * skp * (D + 8) * 2 + (size_t)QT * (D + 8) * 2 + (size_t)skp * 4;
}
template <int D>
size_t bwd_smem(int Sk) {
  const int skp = (Sk + KB - 1) / KB * KB;
  return (size_t)2 * skp * (D + 8) * 2 + (size_t)2 * QT * (D + 8) * 2 + (size_t)2 * QT * PSTR * 2 + (size_t)skp * 4 + 2 * QT * 4;
}

template <int D>
int launch_fwd(const AttnParams& p, cudaStream_t st) {
  const size_t sm = fwd_smem<D>(p.Sk);
  auto kern = attn_fwd_kernel<D>;
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
int launch_bwd(const AttnParams& p, cudaStream_t st) {
  const size_t sm = bwd_smem<D>(p.Sk);
  auto kern = attn_bwd_kernel<D>;
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
  return bwd ? launch_bwd<D>(p, st) : launch_fwd<D>(p, st);
}

}  // namespace

extern "C" int mdhs_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                                  int64_t ldo, const uint8_t* key_mask, float* lse, int B, int H, int Sq, int Sk, int D,
                                  float scale, float drop_p, uint64_t seed, void* stream) {
  if (!q || !k || !v || !out || B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0 || Sk > 512) return MDHS_ERR_ARG;
  if ((D != 32 && D != 64) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return MDHS_ERR_ARG;
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
