// Fused small-sequence attention (forward + backward) on the tensor cores for every multi-head attention on the
// path: BERT self-attention (S <= 512, d = 64), nn.MultiheadAttention self/cross attention of the fusion blocks
// (d = 32; 49..784 queries, <= 512 keys).  Up to 64 queries: one CTA = one (batch, head); the whole K/V of that head
// stays in shared memory; each warp owns 16 query rows: S = Q K^T and O = P V (and the five products of the
// backward) are mma.sync m16n8k16 bf16 tiles fed by ldmatrix, with scale + key mask + online softmax + dropout in
// registers between them.  More than 64 queries: 8-warp forward CTAs and a two-pass backward (dQ pass, then a dK/dV pass
// whose warps own key rows) -- see "long sequences" below.  Probabilities never reach HBM (the backward recomputes them
// from the saved log-sum-exp).  Buffers are token-major [B*S, ld] bf16 with head h in columns [h*D, (h+1)*D).
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
// 2^x straight on the MUFU (ftz; 2^-inf = 0).  The softmax works in the log2 domain: scores are scaled by scale * log2(e)
// once (one FFMA with the additive mask) and every exponential is a single FADD + MUFU.EX2 instead of __expf's
// multiply / range test / fix-up sequence.
__device__ __forceinline__ float ex2f(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
constexpr float LOG2E = 1.4426950408889634f, LN2 = 0.6931471805599453f;
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

// Attention-probability dropout: the keep mask of element (query row, key) is a pure function of (seed, launch tick, row,
// key), identical in the forward and in every backward pass.  One 64-bit hash per query ROW gives a row key; a 32-bit mix
// per PAIR of adjacent keys gives two 16-bit uniforms (ncu on the first long-sequence build: the 64-bit hash per element
// pair was > half of all instructions and the ALU pipe, not the tensor pipe, was the busiest unit).
__device__ __forceinline__ uint32_t attn_row_key(uint64_t seed, uint64_t row) {
  const uint64_t z = dropout_bits4(seed, row);
  return (uint32_t)(z ^ (z >> 32));
}
__device__ __forceinline__ uint32_t attn_pair_bits(uint32_t rowkey, uint32_t key_pair) {
  uint32_t h = rowkey ^ (key_pair * 0x9E3779B1u);
  h *= 0x85EBCA6Bu;
  h ^= h >> 15;
  h *= 0xC2B2AE35u;
  h ^= h >> 16;
  return h;
}
// keep-scales (0 or 1/(1-p)) of keys 2 * key_pair and 2 * key_pair + 1; thr = p * 65536
__device__ __forceinline__ void attn_drop_pair(uint32_t rowkey, uint32_t key_pair, uint32_t thr, float inv_keep, float& m0,
                                               float& m1) {
  const uint32_t h = attn_pair_bits(rowkey, key_pair);
  m0 = (h & 0xffffu) < thr ? 0.f : inv_keep;
  m1 = (h >> 16) < thr ? 0.f : inv_keep;
}
__device__ __forceinline__ float attn_drop_one(uint32_t rowkey, uint32_t key, uint32_t thr, float inv_keep) {
  const uint32_t h = attn_pair_bits(rowkey, key >> 1);
  return ((key & 1u) ? (h >> 16) : (h & 0xffffu)) < thr ? 0.f : inv_keep;
}

// additive key mask (0 / -inf) for the padded key range
__device__ __forceinline__ void stage_mask(const AttnParams& p, int b, int skp, float* sMask) {
  for (int j = threadIdx.x; j < skp; j += blockDim.x) {
    const bool ok = j < p.Sk && (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.Sk + j] != 0);
    sMask[j] = ok ? 0.f : -INFINITY;
  }
}

template <int D, bool DROP>
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
  const float inv_keep = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const float c2 = p.scale * LOG2E;   // scores in the log2 domain
  const uint32_t drop_thr = (uint32_t)(p.drop_p * 65536.f);
  const uint32_t rowkey0 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g) : 0u;
  const uint32_t rowkey1 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g + 8) : 0u;

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
      s[nt][0] = fmaf(s[nt][0], c2, mk.x);
      s[nt][1] = fmaf(s[nt][1], c2, mk.y);
      s[nt][2] = fmaf(s[nt][2], c2, mk.x);
      s[nt][3] = fmaf(s[nt][3], c2, mk.y);
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
      corr[r] = ex2f(mrow[r] - muse[r]);   // exp(-inf) = 0 on the first block
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
      float e0 = ex2f(s[nt][0] - muse[0]), e1 = ex2f(s[nt][1] - muse[0]);
      float e2 = ex2f(s[nt][2] - muse[1]), e3 = ex2f(s[nt][3] - muse[1]);
      lrow[0] += e0 + e1;
      lrow[1] += e2 + e3;
      if (DROP) {
        const uint32_t kp = (uint32_t)(kb + nt * 8 + 2 * tig) >> 1;
        float m0, m1;
        attn_drop_pair(rowkey0, kp, drop_thr, inv_keep, m0, m1);
        e0 *= m0;
        e1 *= m1;
        attn_drop_pair(rowkey1, kp, drop_thr, inv_keep, m0, m1);
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
      if (tig == 0 && p.lse) p.lse[(int64_t)bh * p.Sq + q0 + row] = mrow[r] * LN2 + __logf(lrow[r]);
    }
  }
}

template <int D, bool DROP>
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
  const float inv_keep = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const float c2 = p.scale * LOG2E;   // scores in the log2 domain

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
        sLse[i] = i < nq ? p.lse[(int64_t)bh * p.Sq + q0 + i] * LOG2E : INFINITY;
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
  const uint32_t drop_thr = (uint32_t)(p.drop_p * 65536.f);
  const uint32_t rowkey0 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g) : 0u;
  const uint32_t rowkey1 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g + 8) : 0u;
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
      float p0 = ex2f(fmaf(s[nt][0], c2, mk.x - lse0)), p1 = ex2f(fmaf(s[nt][1], c2, mk.y - lse0));
      float p2 = ex2f(fmaf(s[nt][2], c2, mk.x - lse1)), p3 = ex2f(fmaf(s[nt][3], c2, mk.y - lse1));
      float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
      if (DROP) {
        const uint32_t kp = (uint32_t)(kb + nt * 8 + 2 * tig) >> 1;
        attn_drop_pair(rowkey0, kp, drop_thr, inv_keep, m0, m1);
        attn_drop_pair(rowkey1, kp, drop_thr, inv_keep, m2, m3);
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

// ------------------------------------------------------------------ long sequences (Sq > 64)
// The kernels above give a CTA 64 queries and 4 warps and make every CTA wait for the whole K/V of its head: fine for the
// 64-token BERT / 49-token image problems, latency-bound beyond (S = 256: 4 CTAs per head each staging all K/V, 8 warps
// per SM).  The long path uses 8 warps per CTA, stages K/V as one cp.async group per 64-key block so the first block's
// math starts while the rest is in flight, and splits the backward into two exchange-free passes:
//   dQ pass : warps own 16 query rows (forward layout): S, dP, dS in registers, dQ += dS K.
//   dKV pass: warps own 16 KEY rows and walk the query blocks with the transposed products S^T = K Q^T, dP^T = V dO^T,
//             so P^T / dS^T come out of the accumulators already in A-fragment layout: dV += P^T dO, dK += dS^T Q stay in
//             registers over all queries and are written once as bf16 -- no fp32 atomics, no workspaces, no shared-memory
//             exchange.  S and dP are computed twice (7 products instead of 5); these kernels are far from MMA-bound.
constexpr int LNW = 8;   // warps of the long-path CTAs (128 query or key rows)

template <int D>
__device__ __forceinline__ void stage_rows_range(const bf16* g, int64_t ld, int r_begin, int r_end, int rows, bf16* s) {
  constexpr int VPR = D / 8;
  for (int i = threadIdx.x + r_begin * VPR; i < r_end * VPR; i += blockDim.x) {
    const int r = i / VPR, v = i % VPR;
    const bool ok = r < rows;
    const bf16* src = ok ? g + (int64_t)r * ld + v * 8 : g;
    const uint32_t dst = sm_u32(s + r * (D + 8) + v * 8);
    const int nbytes = ok ? 16 : 0;
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(nbytes) : "memory");
  }
}
__device__ __forceinline__ void cp_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
// wait until at most `pending` of the most recent groups are still in flight (pending <= 7)
__device__ __forceinline__ void cp_wait_pending(int pending) {
  switch (pending) {
    case 0: asm volatile("cp.async.wait_group 0;" ::: "memory"); break;
    case 1: asm volatile("cp.async.wait_group 1;" ::: "memory"); break;
    case 2: asm volatile("cp.async.wait_group 2;" ::: "memory"); break;
    case 3: asm volatile("cp.async.wait_group 3;" ::: "memory"); break;
    case 4: asm volatile("cp.async.wait_group 4;" ::: "memory"); break;
    case 5: asm volatile("cp.async.wait_group 5;" ::: "memory"); break;
    case 6: asm volatile("cp.async.wait_group 6;" ::: "memory"); break;
    default: asm volatile("cp.async.wait_group 7;" ::: "memory"); break;
  }
}

// delta_i = dO_i . O_i and lse_i for `count` rows starting at global query row `q_first` of head (b, h); rows >= Sq get
// delta 0 / lse +inf (P = 0).
template <int D>
__device__ __forceinline__ void stage_delta_lse(const AttnParams& p, int b, int h, int bh, int q_first, int count, float* sDelta,
                                                float* sLse) {
  constexpr int LPR = D / 8;
  constexpr int RPW = 32 / LPR;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  for (int i = warp * RPW + lane / LPR; i < count; i += nwarps * RPW) {
    const bool ok = q_first + i < p.Sq;
    float a = 0.f;
    if (ok) {
      float x[8], y[8];
      const int64_t off = ((int64_t)b * p.Sq + q_first + i) * p.ldo + h * D + (lane % LPR) * 8;
      load8(p.o + off, x);
      load8(p.d_o + off, y);
#pragma unroll
      for (int d = 0; d < 8; d++) a = fmaf(x[d], y[d], a);
    }
#pragma unroll
    for (int o = LPR / 2; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
    if ((lane % LPR) == 0) {
      sDelta[i] = a;
      sLse[i] = ok ? p.lse[(int64_t)bh * p.Sq + q_first + i] * LOG2E : INFINITY;
    }
  }
}

template <int D, bool DROP>
__global__ void __launch_bounds__(LNW * 32, 2) attn_fwd_long_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int LD = D + 8, QTL = LNW * 16;
  const int skp = (p.Sk + KB - 1) / KB * KB, nkb = skp / KB;
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = sK + (size_t)skp * LD;
  bf16* sQ = sV + (size_t)skp * LD;
  float* sMask = reinterpret_cast<float*>(sQ + QTL * LD);
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.y * QTL;
  const int nq = min(QTL, p.Sq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const bf16* kg = p.k + (int64_t)b * p.Sk * p.ldk + h * D;
  const bf16* vg = p.v + (int64_t)b * p.Sk * p.ldv + h * D;

  stage_rows_range<D>(p.q + ((int64_t)b * p.Sq + q0) * p.ldq + h * D, p.ldq, 0, QTL, nq, sQ);
  for (int j = 0; j < nkb; j++) {
    stage_rows_range<D>(kg, p.ldk, j * KB, (j + 1) * KB, p.Sk, sK);
    stage_rows_range<D>(vg, p.ldv, j * KB, (j + 1) * KB, p.Sk, sV);
    cp_commit();
  }
  stage_mask(p, b, skp, sMask);

  const int r0 = warp * 16;
  uint32_t qa[D / 16][4];
  float o[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; i++) o[i][0] = o[i][1] = o[i][2] = o[i][3] = 0.f;
  float mrow[2] = {-INFINITY, -INFINITY}, lrow[2] = {0.f, 0.f};
  const float inv_keep = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const float c2 = p.scale * LOG2E;   // scores in the log2 domain
  const uint32_t drop_thr = (uint32_t)(p.drop_p * 65536.f);
  const uint32_t rowkey0 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g) : 0u;
  const uint32_t rowkey1 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g + 8) : 0u;

  for (int j = 0; j < nkb; j++) {
    const int kb = j * KB;
    cp_wait_pending(nkb - 1 - j);
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int kk = 0; kk < D / 16; kk++) lda_rowmajor(qa[kk], sQ, LD, r0, kk * 16, lane);
    }
    if (r0 >= nq) continue;   // warp-uniform; the warp still takes part in the barriers above
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
      s[nt][0] = fmaf(s[nt][0], c2, mk.x);
      s[nt][1] = fmaf(s[nt][1], c2, mk.y);
      s[nt][2] = fmaf(s[nt][2], c2, mk.x);
      s[nt][3] = fmaf(s[nt][3], c2, mk.y);
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
      corr[r] = ex2f(mrow[r] - muse[r]);
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
      float e0 = ex2f(s[nt][0] - muse[0]), e1 = ex2f(s[nt][1] - muse[0]);
      float e2 = ex2f(s[nt][2] - muse[1]), e3 = ex2f(s[nt][3] - muse[1]);
      lrow[0] += e0 + e1;
      lrow[1] += e2 + e3;
      if (DROP) {
        const uint32_t kp = (uint32_t)(kb + nt * 8 + 2 * tig) >> 1;
        float m0, m1;
        attn_drop_pair(rowkey0, kp, drop_thr, inv_keep, m0, m1);
        e0 *= m0;
        e1 *= m1;
        attn_drop_pair(rowkey1, kp, drop_thr, inv_keep, m0, m1);
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
  if (r0 >= nq) return;
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
      if (tig == 0 && p.lse) p.lse[(int64_t)bh * p.Sq + q0 + row] = mrow[r] * LN2 + __logf(lrow[r]);
    }
  }
}

// dQ pass: one CTA = (batch, head, 128 queries); warps own 16 query rows.
template <int D, bool DROP>
__global__ void __launch_bounds__(LNW * 32, 2) attn_bwd_dq_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int LD = D + 8, QTL = LNW * 16;
  const int skp = (p.Sk + KB - 1) / KB * KB, nkb = skp / KB;
  bf16* sK = reinterpret_cast<bf16*>(smem);
  bf16* sV = sK + (size_t)skp * LD;
  bf16* sQ = sV + (size_t)skp * LD;
  bf16* sdO = sQ + QTL * LD;
  float* sMask = reinterpret_cast<float*>(sdO + QTL * LD);
  float* sLse = sMask + skp;
  float* sDelta = sLse + QTL;
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int q0 = blockIdx.y * QTL;
  const int nq = min(QTL, p.Sq - q0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const float inv_keep = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const float c2 = p.scale * LOG2E;   // scores in the log2 domain
  const bf16* kg = p.k + (int64_t)b * p.Sk * p.ldk + h * D;
  const bf16* vg = p.v + (int64_t)b * p.Sk * p.ldv + h * D;

  stage_rows_range<D>(p.q + ((int64_t)b * p.Sq + q0) * p.ldq + h * D, p.ldq, 0, QTL, nq, sQ);
  stage_rows_range<D>(p.d_o + ((int64_t)b * p.Sq + q0) * p.ldo + h * D, p.ldo, 0, QTL, nq, sdO);
  for (int j = 0; j < nkb; j++) {
    stage_rows_range<D>(kg, p.ldk, j * KB, (j + 1) * KB, p.Sk, sK);
    stage_rows_range<D>(vg, p.ldv, j * KB, (j + 1) * KB, p.Sk, sV);
    cp_commit();
  }
  stage_mask(p, b, skp, sMask);
  stage_delta_lse<D>(p, b, h, bh, q0, QTL, sDelta, sLse);
  if (p.dk32 != nullptr) {   // scratch given: publish delta for the dK/dV pass (same thread that wrote sDelta[i])
    __syncthreads();
    for (int i = threadIdx.x; i < nq; i += blockDim.x) p.dk32[(int64_t)bh * p.Sq + q0 + i] = sDelta[i];
  }

  const int r0 = warp * 16;
  uint32_t qa[D / 16][4], doa[D / 16][4];
  float lse0 = 0.f, lse1 = 0.f, dl0 = 0.f, dl1 = 0.f;
  float dq[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; i++) dq[i][0] = dq[i][1] = dq[i][2] = dq[i][3] = 0.f;
  const uint32_t drop_thr = (uint32_t)(p.drop_p * 65536.f);
  const uint32_t rowkey0 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g) : 0u;
  const uint32_t rowkey1 = DROP ? attn_row_key(p.seed, (uint64_t)bh * p.Sq + q0 + r0 + g + 8) : 0u;

  for (int j = 0; j < nkb; j++) {
    const int kb = j * KB;
    cp_wait_pending(nkb - 1 - j);
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int kk = 0; kk < D / 16; kk++) {
        lda_rowmajor(qa[kk], sQ, LD, r0, kk * 16, lane);
        lda_rowmajor(doa[kk], sdO, LD, r0, kk * 16, lane);
      }
      lse0 = sLse[r0 + g];
      lse1 = sLse[r0 + g + 8];
      dl0 = sDelta[r0 + g];
      dl1 = sDelta[r0 + g + 8];
    }
    if (r0 >= nq) continue;
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
      const float p0 = ex2f(fmaf(s[nt][0], c2, mk.x - lse0)), p1 = ex2f(fmaf(s[nt][1], c2, mk.y - lse0));
      const float p2 = ex2f(fmaf(s[nt][2], c2, mk.x - lse1)), p3 = ex2f(fmaf(s[nt][3], c2, mk.y - lse1));
      float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
      if (DROP) {
        const uint32_t kp = (uint32_t)(kb + nt * 8 + 2 * tig) >> 1;
        attn_drop_pair(rowkey0, kp, drop_thr, inv_keep, m0, m1);
        attn_drop_pair(rowkey1, kp, drop_thr, inv_keep, m2, m3);
      }
      dsa[nt >> 1][(nt & 1) * 2] = pack2(p0 * (dp[nt][0] * m0 - dl0), p1 * (dp[nt][1] * m1 - dl0));
      dsa[nt >> 1][(nt & 1) * 2 + 1] = pack2(p2 * (dp[nt][2] * m2 - dl1), p3 * (dp[nt][3] * m3 - dl1));
    }
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
  }
  if (r0 >= nq) return;
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

// dK / dV pass: one CTA = (batch, head, 64 keys), 4 warps, warps own 16 key rows.  Q / dO stream through a two-stage
// cp.async ring of 64-query blocks (so three CTAs fit an SM: ~58 KB of shared memory, <= 168 registers); each block is
// consumed as two 32-query halves to keep the S^T / dP^T accumulators small.  delta / lse / dropout row keys of every
// query of the head sit in shared memory (delta comes from the dQ pass through `p.dk32` when the caller gave scratch).
constexpr int DKV_NW = 4, DKV_KT = DKV_NW * 16, DKV_QH = 32;

template <int D, bool DROP>
__global__ void __launch_bounds__(DKV_NW * 32, 3) attn_bwd_dkv_kernel(const AttnParams p) {
  extern __shared__ __align__(16) uint8_t smem[];
  constexpr int LD = D + 8, KT = DKV_KT;
  const int sqp = (p.Sq + KB - 1) / KB * KB, nqb = sqp / KB;
  bf16* sQ = reinterpret_cast<bf16*>(smem);          // [2][KB][LD]
  bf16* sdO = sQ + 2 * KB * LD;                      // [2][KB][LD]
  bf16* sK = sdO + 2 * KB * LD;
  bf16* sV = sK + KT * LD;
  float* sLse = reinterpret_cast<float*>(sV + KT * LD);
  float* sDelta = sLse + sqp;
  uint32_t* sRowKey = reinterpret_cast<uint32_t*>(sDelta + sqp);
  const int bh = blockIdx.x, b = bh / p.H, h = bh % p.H;
  const int k0 = blockIdx.y * KT;
  const int nk = min(KT, p.Sk - k0);
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int g = lane >> 2, tig = lane & 3;
  const float inv_keep = DROP ? 1.f / (1.f - p.drop_p) : 1.f;
  const float c2 = p.scale * LOG2E;   // scores in the log2 domain
  const bf16* qg = p.q + (int64_t)b * p.Sq * p.ldq + h * D;
  const bf16* og = p.d_o + (int64_t)b * p.Sq * p.ldo + h * D;

  stage_rows_range<D>(p.k + ((int64_t)b * p.Sk + k0) * p.ldk + h * D, p.ldk, 0, KT, nk, sK);
  stage_rows_range<D>(p.v + ((int64_t)b * p.Sk + k0) * p.ldv + h * D, p.ldv, 0, KT, nk, sV);
  for (int j = 0; j < 2; j++) {
    if (j < nqb) {
      stage_rows_range<D>(qg + (int64_t)j * KB * p.ldq, p.ldq, 0, KB, p.Sq - j * KB, sQ + j * KB * LD);
      stage_rows_range<D>(og + (int64_t)j * KB * p.ldo, p.ldo, 0, KB, p.Sq - j * KB, sdO + j * KB * LD);
    }
    cp_commit();
  }
  if (p.dk32 != nullptr) {   // delta from the dQ pass
    for (int i = threadIdx.x; i < sqp; i += blockDim.x) {
      const bool ok = i < p.Sq;
      sDelta[i] = ok ? p.dk32[(int64_t)bh * p.Sq + i] : 0.f;
      sLse[i] = ok ? p.lse[(int64_t)bh * p.Sq + i] * LOG2E : INFINITY;
    }
  } else {
    stage_delta_lse<D>(p, b, h, bh, 0, sqp, sDelta, sLse);
  }
  if (DROP)
    for (int i = threadIdx.x; i < sqp; i += blockDim.x) sRowKey[i] = attn_row_key(p.seed, (uint64_t)bh * p.Sq + i);

  const int r0 = warp * 16;
  const bool active = r0 < nk;
  // additive key mask of this thread's two key rows (padded / masked keys: -inf -> P = 0)
  float mk0, mk1;
  {
    const int key0 = k0 + r0 + g, key1 = key0 + 8;
    mk0 = (key0 < p.Sk && (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.Sk + key0] != 0)) ? 0.f : -INFINITY;
    mk1 = (key1 < p.Sk && (p.key_mask == nullptr || p.key_mask[(int64_t)b * p.Sk + key1] != 0)) ? 0.f : -INFINITY;
  }
  uint32_t ka[D / 16][4], va[D / 16][4];
  float dv[D / 8][4], dk[D / 8][4];
#pragma unroll
  for (int i = 0; i < D / 8; i++) {
    dv[i][0] = dv[i][1] = dv[i][2] = dv[i][3] = 0.f;
    dk[i][0] = dk[i][1] = dk[i][2] = dk[i][3] = 0.f;
  }
  const uint32_t drop_thr = (uint32_t)(p.drop_p * 65536.f);
  const uint32_t key0 = (uint32_t)(k0 + r0 + g);

  for (int j = 0; j < nqb; j++) {
    const int buf = j & 1;
    const bf16* bQ = sQ + buf * KB * LD;
    const bf16* bdO = sdO + buf * KB * LD;
    if (j + 1 < nqb) cp_wait_pending(1);
    else cp_wait_pending(0);
    __syncthreads();
    if (j == 0) {
#pragma unroll
      for (int kk = 0; kk < D / 16; kk++) {
        lda_rowmajor(ka[kk], sK, LD, r0, kk * 16, lane);
        lda_rowmajor(va[kk], sV, LD, r0, kk * 16, lane);
      }
    }
    if (active) {
#pragma unroll
      for (int half = 0; half < KB / DKV_QH; half++) {
        const int ql = half * DKV_QH;        // first query row of this half inside the stage
        const int qb = j * KB + ql;          // ... and inside the head
        float st[DKV_QH / 8][4], dpt[DKV_QH / 8][4];
#pragma unroll
        for (int i = 0; i < DKV_QH / 8; i++) {
          st[i][0] = st[i][1] = st[i][2] = st[i][3] = 0.f;
          dpt[i][0] = dpt[i][1] = dpt[i][2] = dpt[i][3] = 0.f;
        }
#pragma unroll
        for (int kk = 0; kk < D / 16; kk++) {
#pragma unroll
          for (int np = 0; np < DKV_QH / 16; np++) {
            uint32_t bb[4];
            ldb_nk(bb, bQ, LD, ql + np * 16, kk * 16, lane);
            mma16816(st[2 * np], ka[kk], bb[0], bb[1]);
            mma16816(st[2 * np + 1], ka[kk], bb[2], bb[3]);
            ldb_nk(bb, bdO, LD, ql + np * 16, kk * 16, lane);
            mma16816(dpt[2 * np], va[kk], bb[0], bb[1]);
            mma16816(dpt[2 * np + 1], va[kk], bb[2], bb[3]);
          }
        }
        uint32_t pta[DKV_QH / 16][4], dsta[DKV_QH / 16][4];
#pragma unroll
        for (int nt = 0; nt < DKV_QH / 8; nt++) {
          const int qi = qb + nt * 8 + 2 * tig;
          const float2 ls = *reinterpret_cast<const float2*>(sLse + qi);
          const float2 dl = *reinterpret_cast<const float2*>(sDelta + qi);
          float p0 = ex2f(fmaf(st[nt][0], c2, mk0 - ls.x)), p1 = ex2f(fmaf(st[nt][1], c2, mk0 - ls.y));
          float p2 = ex2f(fmaf(st[nt][2], c2, mk1 - ls.x)), p3 = ex2f(fmaf(st[nt][3], c2, mk1 - ls.y));
          float m0 = 1.f, m1 = 1.f, m2 = 1.f, m3 = 1.f;
          if (DROP) {
            const uint2 rk = *reinterpret_cast<const uint2*>(sRowKey + qi);
            m0 = attn_drop_one(rk.x, key0, drop_thr, inv_keep);
            m1 = attn_drop_one(rk.y, key0, drop_thr, inv_keep);
            m2 = attn_drop_one(rk.x, key0 + 8, drop_thr, inv_keep);
            m3 = attn_drop_one(rk.y, key0 + 8, drop_thr, inv_keep);
          }
          dsta[nt >> 1][(nt & 1) * 2] = pack2(p0 * (dpt[nt][0] * m0 - dl.x), p1 * (dpt[nt][1] * m1 - dl.y));
          dsta[nt >> 1][(nt & 1) * 2 + 1] = pack2(p2 * (dpt[nt][2] * m2 - dl.x), p3 * (dpt[nt][3] * m3 - dl.y));
          pta[nt >> 1][(nt & 1) * 2] = pack2(p0 * m0, p1 * m1);
          pta[nt >> 1][(nt & 1) * 2 + 1] = pack2(p2 * m2, p3 * m3);
        }
#pragma unroll
        for (int kk = 0; kk < DKV_QH / 16; kk++) {
#pragma unroll
          for (int dd = 0; dd < D / 16; dd++) {
            uint32_t bb[4];
            ldb_kn(bb, bdO, LD, ql + kk * 16, dd * 16, lane);
            mma16816(dv[2 * dd], pta[kk], bb[0], bb[1]);
            mma16816(dv[2 * dd + 1], pta[kk], bb[2], bb[3]);
            ldb_kn(bb, bQ, LD, ql + kk * 16, dd * 16, lane);
            mma16816(dk[2 * dd], dsta[kk], bb[0], bb[1]);
            mma16816(dk[2 * dd + 1], dsta[kk], bb[2], bb[3]);
          }
        }
      }
    }
    __syncthreads();   // every warp is done with this stage: refill it with the block two ahead
    if (j + 2 < nqb) {
      stage_rows_range<D>(qg + (int64_t)(j + 2) * KB * p.ldq, p.ldq, 0, KB, p.Sq - (j + 2) * KB, sQ + buf * KB * LD);
      stage_rows_range<D>(og + (int64_t)(j + 2) * KB * p.ldo, p.ldo, 0, KB, p.Sq - (j + 2) * KB, sdO + buf * KB * LD);
    }
    cp_commit();
  }
  if (!active) return;
#pragma unroll
  for (int r = 0; r < 2; r++) {
    const int key = k0 + r0 + g + r * 8;
    if (key < p.Sk) {
      const int64_t krow = (int64_t)b * p.Sk + key;
#pragma unroll
      for (int i = 0; i < D / 8; i++) {
        const int c = h * D + i * 8 + 2 * tig;
        *reinterpret_cast<bf162*>(p.dk + krow * p.ldk + c) = __floats2bfloat162_rn(dk[i][2 * r] * p.scale, dk[i][2 * r + 1] * p.scale);
        *reinterpret_cast<bf162*>(p.dv + krow * p.ldv + c) = __floats2bfloat162_rn(dv[i][2 * r], dv[i][2 * r + 1]);
      }
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
  const bool drop = p.drop_p > 0.f;
  auto kern = drop ? attn_fwd_kernel<D, true> : attn_fwd_kernel<D, false>;
  static size_t configured[2] = {0, 0};
  if (sm > 48 * 1024 && sm > configured[drop]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured[drop] = sm;
  }
  kern<<<dim3(p.B * p.H, ceil_div(p.Sq, QT)), NWARPS * 32, sm, st>>>(p);
  MDHS_RETURN_LAST();
}
template <int D>
int launch_bwd(const AttnParams& p, cudaStream_t st) {
  const size_t sm = bwd_smem<D>(p.Sk);
  const bool drop = p.drop_p > 0.f;
  auto kern = drop ? attn_bwd_kernel<D, true> : attn_bwd_kernel<D, false>;
  static size_t configured[2] = {0, 0};
  if (sm > 48 * 1024 && sm > configured[drop]) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured[drop] = sm;
  }
  kern<<<dim3(p.B * p.H, ceil_div(p.Sq, QT)), NWARPS * 32, sm, st>>>(p);
  MDHS_RETURN_LAST();
}

constexpr size_t SMEM_LIMIT = 227 * 1024;
template <int D>
size_t fwd_long_smem(int Sk) {
  const int skp = (Sk + KB - 1) / KB * KB;
  return (size_t)2 * skp * (D + 8) * 2 + (size_t)LNW * 16 * (D + 8) * 2 + (size_t)skp * 4;
}
template <int D>
size_t dq_smem(int Sk) {
  const int skp = (Sk + KB - 1) / KB * KB;
  return (size_t)2 * skp * (D + 8) * 2 + (size_t)2 * LNW * 16 * (D + 8) * 2 + (size_t)skp * 4 + (size_t)2 * LNW * 16 * 4;
}
template <int D>
size_t dkv_smem(int Sq) {
  const int sqp = (Sq + KB - 1) / KB * KB;
  return (size_t)4 * KB * (D + 8) * 2 + (size_t)2 * DKV_KT * (D + 8) * 2 + (size_t)3 * sqp * 4;
}

template <class K>
int set_smem(K kern, size_t sm, size_t& configured) {
  if (sm > 48 * 1024 && sm > configured) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm);
    if (e != cudaSuccess) return (int)e;
    configured = sm;
  }
  return 0;
}

template <int D>
bool long_fwd_ok(int Sq, int Sk) { return Sq > QT && fwd_long_smem<D>(Sk) <= SMEM_LIMIT; }
template <int D>
bool long_bwd_ok(int Sq, int Sk) {
  return Sq > QT && dq_smem<D>(Sk) <= SMEM_LIMIT && dkv_smem<D>(Sq) <= SMEM_LIMIT;
}

template <int D>
int launch_fwd_long(const AttnParams& p, cudaStream_t st) {
  const size_t sm = fwd_long_smem<D>(p.Sk);
  static size_t configured[2] = {0, 0};
  const bool drop = p.drop_p > 0.f;
  auto kern = drop ? attn_fwd_long_kernel<D, true> : attn_fwd_long_kernel<D, false>;
  if (int rc = set_smem(kern, sm, configured[drop])) return rc;
  kern<<<dim3(p.B * p.H, ceil_div(p.Sq, LNW * 16)), LNW * 32, sm, st>>>(p);
  MDHS_RETURN_LAST();
}
template <int D>
int launch_bwd_long(const AttnParams& p, cudaStream_t st) {
  const size_t sm = dq_smem<D>(p.Sk);
  static size_t c0[2] = {0, 0}, c1[2] = {0, 0};
  const bool drop = p.drop_p > 0.f;
  auto kq = drop ? attn_bwd_dq_kernel<D, true> : attn_bwd_dq_kernel<D, false>;
  if (int rc = set_smem(kq, sm, c0[drop])) return rc;
  kq<<<dim3(p.B * p.H, ceil_div(p.Sq, LNW * 16)), LNW * 32, sm, st>>>(p);
  const size_t sk = dkv_smem<D>(p.Sq);
  g_mdhs_launches++;
  auto kk = drop ? attn_bwd_dkv_kernel<D, true> : attn_bwd_dkv_kernel<D, false>;
  if (int rc = set_smem(kk, sk, c1[drop])) return rc;
  kk<<<dim3(p.B * p.H, ceil_div(p.Sk, DKV_KT)), DKV_NW * 32, sk, st>>>(p);
  MDHS_RETURN_LAST();
}

template <int D>
int dispatch(const AttnParams& p, bool bwd, cudaStream_t st) {
  if (bwd) return long_bwd_ok<D>(p.Sq, p.Sk) ? launch_bwd_long<D>(p, st) : launch_bwd<D>(p, st);
  return long_fwd_ok<D>(p.Sq, p.Sk) ? launch_fwd_long<D>(p, st) : launch_fwd<D>(p, st);
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

// 1 when mdhs_attention_bwd needs the fp32 workspaces for this shape (Sq > 64 and the two-pass long-sequence kernels do
// not fit in shared memory), else 0.
extern "C" int mdhs_attention_bwd_workspace(int Sq, int Sk, int D) {
  if (Sq <= QT) return 0;
  return (D == 64 ? long_bwd_ok<64>(Sq, Sk) : long_bwd_ok<32>(Sq, Sk)) ? 0 : 1;
}

// dk/dv: bf16 outputs (same strides as k/v).  Sq <= 64: one query tile per head writes them directly.  Sq > 64: the dQ pass
// and the dK/dV pass (keys own the accumulators) write bf16 directly as well; only when mdhs_attention_bwd_workspace() says
// so (very long query sets) several query tiles accumulate into zero-initialised fp32 workspaces dk32/dv32
// ([B*Sk, ldk] / [B*Sk, ldv]) and the caller casts them to bf16.
extern "C" int mdhs_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                                  const void* o, const void* d_o, int64_t ldo, const uint8_t* key_mask, const float* lse,
                                  void* dq, void* dk, void* dv, float* dk32, float* dv32, int B, int H, int Sq, int Sk, int D,
                                  float scale, float drop_p, uint64_t seed, void* stream) {
  if (!q || !k || !v || !o || !d_o || !lse || !dq || B <= 0 || H <= 0 || Sq <= 0 || Sk <= 0 || Sk > 512) return MDHS_ERR_ARG;
  if ((D != 32 && D != 64) || (ldq % 8) || (ldk % 8) || (ldv % 8) || (ldo % 8)) return MDHS_ERR_ARG;
  const bool multi = mdhs_attention_bwd_workspace(Sq, Sk, D) != 0;
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
