// MIBF-Net specific kernels (mibf_net/attention.py:47-70 IBFA with one token per modality; mibf_net/model_resnet.py:76-94
// MP-Loss).  Rows = batch; latency-bound: one warp (IBFA) or one thread (loss) per sample, fp32 math.
#include "common.cuh"
#include "../../include/mdhs_b200.h"

extern int64_t g_mdhs_launches;

namespace {

// IBFA with seq_len_x = seq_len_y = 1: per (sample, head) two keys {Kx, Ky}, two values {Vx, Vy}:
//   a = softmax([Q.Kx, Q.Ky] / sqrt(D)),  out = a0 Vx + a1 Vy.
// kqv_x is the fused projection [K_x | Q_x | V_x] (row stride ldx), kv_y = [K_y | V_y] (row stride ldy); C = H*D.
__global__ void __launch_bounds__(128) ibfa_fwd_kernel(const bf16* __restrict__ kqv_x, int64_t ldx, const bf16* __restrict__ kv_y,
                                                       int64_t ldy, bf16* __restrict__ out, float* __restrict__ probs, int B, int H,
                                                       int D) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * H) return;
  const int b = w / H, h = w % H, C = H * D;
  const bf16* kx = kqv_x + (int64_t)b * ldx + h * D;
  const bf16* q = kx + C;
  const bf16* vx = kx + 2 * C;
  const bf16* ky = kv_y + (int64_t)b * ldy + h * D;
  const bf16* vy = ky + C;
  float sx = 0.f, sy = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float qq = __bfloat162float(q[d]);
    sx += qq * __bfloat162float(kx[d]);
    sy += qq * __bfloat162float(ky[d]);
  }
  const float scale = rsqrtf((float)D);
  sx = warp_sum(sx) * scale;
  sy = warp_sum(sy) * scale;
  const float mx = fmaxf(sx, sy);
  const float ex = __expf(sx - mx), ey = __expf(sy - mx);
  const float a0 = ex / (ex + ey), a1 = ey / (ex + ey);
  if (lane == 0) {
    probs[2 * w] = a0;
    probs[2 * w + 1] = a1;
  }
  bf16* o = out + (int64_t)b * C + h * D;
  for (int d = lane; d < D; d += 32) o[d] = __float2bfloat16_rn(a0 * __bfloat162float(vx[d]) + a1 * __bfloat162float(vy[d]));
}

__global__ void __launch_bounds__(128) ibfa_bwd_kernel(const bf16* __restrict__ kqv_x, int64_t ldx, const bf16* __restrict__ kv_y,
                                                       int64_t ldy, const bf16* __restrict__ dout, const float* __restrict__ probs,
                                                       bf16* __restrict__ dkqv_x, bf16* __restrict__ dkv_y, int B, int H, int D) {
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  if (w >= B * H) return;
  const int b = w / H, h = w % H, C = H * D;
  const bf16* kx = kqv_x + (int64_t)b * ldx + h * D;
  const bf16* q = kx + C;
  const bf16* vx = kx + 2 * C;
  const bf16* ky = kv_y + (int64_t)b * ldy + h * D;
  const bf16* vy = ky + C;
  const bf16* go = dout + (int64_t)b * C + h * D;
  const float a0 = probs[2 * w], a1 = probs[2 * w + 1];
  float da0 = 0.f, da1 = 0.f;
  for (int d = lane; d < D; d += 32) {
    const float g = __bfloat162float(go[d]);
    da0 += g * __bfloat162float(vx[d]);
    da1 += g * __bfloat162float(vy[d]);
  }
  da0 = warp_sum(da0);
  da1 = warp_sum(da1);
  const float dot = a0 * da0 + a1 * da1;
  const float scale = rsqrtf((float)D);
  const float ds0 = a0 * (da0 - dot) * scale, ds1 = a1 * (da1 - dot) * scale;
  bf16* gkx = dkqv_x + (int64_t)b * ldx + h * D;
  bf16* gq = gkx + C;
  bf16* gvx = gkx + 2 * C;
  bf16* gky = dkv_y + (int64_t)b * ldy + h * D;
  bf16* gvy = gky + C;
  for (int d = lane; d < D; d += 32) {
    const float g = __bfloat162float(go[d]), qq = __bfloat162float(q[d]);
    gq[d] = __float2bfloat16_rn(ds0 * __bfloat162float(kx[d]) + ds1 * __bfloat162float(ky[d]));
    gkx[d] = __float2bfloat16_rn(ds0 * qq);
    gky[d] = __float2bfloat16_rn(ds1 * qq);
    gvx[d] = __float2bfloat16_rn(a0 * g);
    gvy[d] = __float2bfloat16_rn(a1 * g);
  }
}

// MP-Loss, forward + backward in one launch (single CTA, B <= a few thousand, C <= 32):
//   kl_i = clamp(nan_to_num((KL(p||q) + KL(q||p)) / 2), 0, 10), p = softmax(img), q = softmax(txt) clamped to [1e-8, 1]
//   loss = 0.3 CE(img) + 0.6 CE(txt) + 1.1 * mean_i(exp(kl_i)) * CE(fused)
constexpr int MPL_MAXC = 32;
__device__ __forceinline__ float row_softmax(const float* z, int C, float* p) {
  float mx = -INFINITY;
  for (int c = 0; c < C; c++) mx = fmaxf(mx, z[c]);
  float se = 0.f;
  for (int c = 0; c < C; c++) {
    p[c] = __expf(z[c] - mx);
    se += p[c];
  }
  const float inv = 1.f / se;
  for (int c = 0; c < C; c++) p[c] *= inv;
  return mx + __logf(se);
}
__global__ void __launch_bounds__(256) mp_loss_kernel(const float* __restrict__ zi, const float* __restrict__ zt,
                                                      const float* __restrict__ zf, const int64_t* __restrict__ labels,
                                                      float* __restrict__ loss, float* __restrict__ gi, float* __restrict__ gt,
                                                      float* __restrict__ gf, int B, int C) {
  __shared__ float red[32];
  float ce_i = 0.f, ce_t = 0.f, ce_f = 0.f, ekl = 0.f;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float p[MPL_MAXC], q[MPL_MAXC], r[MPL_MAXC];
    const int y = (int)labels[i];
    ce_i += row_softmax(zi + (int64_t)i * C, C, p) - zi[(int64_t)i * C + y];
    ce_t += row_softmax(zt + (int64_t)i * C, C, q) - zt[(int64_t)i * C + y];
    ce_f += row_softmax(zf + (int64_t)i * C, C, r) - zf[(int64_t)i * C + y];
    float kl = 0.f;
    for (int c = 0; c < C; c++) {
      const float pc = fminf(fmaxf(p[c], 1e-8f), 1.f), qc = fminf(fmaxf(q[c], 1e-8f), 1.f);
      kl += 0.5f * (pc - qc) * (__logf(pc) - __logf(qc));
    }
    if (!(kl == kl)) kl = 0.f;
    kl = fminf(fmaxf(kl, 0.f), 10.f);
    ekl += __expf(kl);
  }
  const float invB = 1.f / (float)B;
  ce_i = block_sum(ce_i, red) * invB;
  ce_t = block_sum(ce_t, red) * invB;
  ce_f = block_sum(ce_f, red) * invB;
  const float w = block_sum(ekl, red) * invB;  // mean(exp(kl))
  if (threadIdx.x == 0) loss[0] = 0.3f * ce_i + 0.6f * ce_t + 1.1f * w * ce_f;
  if (!gi) return;
  for (int i = threadIdx.x; i < B; i += blockDim.x) {
    float p[MPL_MAXC], q[MPL_MAXC], r[MPL_MAXC], gp[MPL_MAXC], gq[MPL_MAXC];
    const int y = (int)labels[i];
    row_softmax(zi + (int64_t)i * C, C, p);
    row_softmax(zt + (int64_t)i * C, C, q);
    row_softmax(zf + (int64_t)i * C, C, r);
    float kl = 0.f;
    for (int c = 0; c < C; c++) {
      const bool pin = p[c] >= 1e-8f && p[c] <= 1.f, qin = q[c] >= 1e-8f && q[c] <= 1.f;
      const float pc = fminf(fmaxf(p[c], 1e-8f), 1.f), qc = fminf(fmaxf(q[c], 1e-8f), 1.f);
      const float lp = __logf(pc), lq = __logf(qc);
      kl += 0.5f * (pc - qc) * (lp - lq);
      gp[c] = pin ? 0.5f * ((lp - lq) + 1.f - qc / pc) : 0.f;
      gq[c] = qin ? 0.5f * ((lq - lp) + 1.f - pc / qc) : 0.f;
    }
    const bool live = (kl == kl) && kl > 0.f && kl < 10.f;   // clamp / nan_to_num pass no gradient outside
    const float klc = (kl == kl) ? fminf(fmaxf(kl, 0.f), 10.f) : 0.f;
    const float coef = live ? 1.1f * ce_f * invB * __expf(klc) : 0.f;
    float dp = 0.f, dq = 0.f;
    for (int c = 0; c < C; c++) {
      dp += p[c] * gp[c];
      dq += q[c] * gq[c];
    }
    for (int c = 0; c < C; c++) {
      const float oh = (c == y) ? 1.f : 0.f;
      gi[(int64_t)i * C + c] = 0.3f * invB * (p[c] - oh) + coef * p[c] * (gp[c] - dp);
      gt[(int64_t)i * C + c] = 0.6f * invB * (q[c] - oh) + coef * q[c] * (gq[c] - dq);
      gf[(int64_t)i * C + c] = 1.1f * w * invB * (r[c] - oh);
    }
  }
}

}  // namespace

extern "C" int mdhs_ibfa_fwd(const void* kqv_x, int64_t ldx, const void* kv_y, int64_t ldy, void* out, float* probs, int B, int H,
                             int D, void* stream) {
  if (!kqv_x || !kv_y || !out || !probs || B <= 0 || H <= 0 || D <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  ibfa_fwd_kernel<<<ceil_div((int64_t)B * H, 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)kqv_x, ldx, (const bf16*)kv_y, ldy, (bf16*)out, probs, B, H, D);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_ibfa_bwd(const void* kqv_x, int64_t ldx, const void* kv_y, int64_t ldy, const void* dout, const float* probs,
                             void* dkqv_x, void* dkv_y, int B, int H, int D, void* stream) {
  if (!kqv_x || !kv_y || !dout || !probs || !dkqv_x || !dkv_y || B <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  ibfa_bwd_kernel<<<ceil_div((int64_t)B * H, 4), 128, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      (const bf16*)kqv_x, ldx, (const bf16*)kv_y, ldy, (const bf16*)dout, probs, (bf16*)dkqv_x, (bf16*)dkv_y, B, H, D);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_mp_loss(const float* img_logits, const float* txt_logits, const float* fused_logits, const int64_t* labels,
                            float* loss, float* g_img, float* g_txt, float* g_fused, int B, int C, void* stream) {
  if (!img_logits || !txt_logits || !fused_logits || !labels || !loss || B <= 0 || C <= 0 || C > MPL_MAXC) return MDHS_ERR_ARG;
  if (g_img && (!g_txt || !g_fused)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  mp_loss_kernel<<<1, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(img_logits, txt_logits, fused_logits, labels, loss, g_img, g_txt,
                                                                        g_fused, B, C);
  MDHS_RETURN_LAST();
}
