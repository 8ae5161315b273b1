// Small elementwise kernels used between the fused ops (activation/dropout backward, products, gates).
#include "common.cuh"
#include "../../include/mdhs_b200.h"

MDHS_DEFINE_SEED_TICK(elementwise)

extern int64_t g_mdhs_launches;

namespace {

int grid_for(int64_t items) {
  int64_t g = (items + 255) / 256;
  const int64_t cap = (int64_t)mdhs_num_sms() * 16;
  return (int)(g < 1 ? 1 : (g > cap ? cap : g));
}

// g = dy * dropout_mask(seed, idx) * act'(aux)   (bf16, 8 per thread).  aux = pre-activation (GELU) or output (ReLU).
__global__ void __launch_bounds__(256) act_dropout_bwd_kernel(const bf16* __restrict__ dy, const bf16* __restrict__ aux,
                                                              bf16* __restrict__ g, int64_t n, int act, float drop_p,
                                                              uint64_t seed) {
  const int64_t nv = n >> 3;
  const float inv_keep = drop_p > 0.f ? 1.f / (1.f - drop_p) : 1.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nv; i += (int64_t)gridDim.x * blockDim.x) {
    float d[8], a[8];
    load8(dy + i * 8, d);
    if (act != MDHS_ACT_NONE) load8(aux + i * 8, a);
#pragma unroll
    for (int k = 0; k < 8; k++) {
      float v = d[k];
      if (drop_p > 0.f) v *= dropout_scale(seed, (uint64_t)(i * 8 + k), drop_p, inv_keep);
      if (act == MDHS_ACT_RELU) v = a[k] > 0.f ? v : 0.f;
      else if (act == MDHS_ACT_GELU) v *= gelu_erf_grad(a[k]);
      d[k] = v;
    }
    store8(g + i * 8, d);
  }
}

__global__ void relu_bwd_f32_kernel(const float* __restrict__ dy, const float* __restrict__ y, float* __restrict__ dx, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    dx[i] = y[i] > 0.f ? dy[i] : 0.f;
}

// c = a * b (fp32); also used for the backward (da = dc * b, db = dc * a)
__global__ void mul_f32_kernel(const float* __restrict__ a, const float* __restrict__ b, float* __restrict__ c, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) c[i] = a[i] * b[i];
}

// y = x * dropout_mask(seed, i)  (fp32; the same call implements the backward)
__global__ void dropout_f32_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t n, float p, uint64_t seed) {
  const float inv_keep = 1.f / (1.f - p);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x)
    y[i] = x[i] * dropout_scale(seed, (uint64_t)i, p, inv_keep);
}

// grad[i] += (float)sum64[i]; the fp64 workspaces (column sums written by a GEMM epilogue) are cleared for their next use
__global__ void sum64_to_grad_kernel(double* __restrict__ sum64, double* __restrict__ sumsq64, float* __restrict__ grad, int n) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  grad[i] += (float)sum64[i];
  sum64[i] = 0.0;
  if (sumsq64 != nullptr) sumsq64[i] = 0.0;
}

}  // namespace

namespace {
// out = a * x (+ b * y) on bf16 vectors of 8 (token-tensor averages of the global / local crops, model.py:303-315)
__global__ void __launch_bounds__(256) axpby_bf16_kernel(const bf16* __restrict__ x, const bf16* __restrict__ y, bf16* __restrict__ out,
                                                         int64_t nvec, float a, float b) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nvec; i += (int64_t)gridDim.x * blockDim.x) {
    float xv[8], yv[8];
    load8(x + i * 8, xv);
    if (y != nullptr) load8(y + i * 8, yv);
#pragma unroll
    for (int k = 0; k < 8; k++) xv[k] = y != nullptr ? fmaf(a, xv[k], b * yv[k]) : a * xv[k];
    store8(out + i * 8, xv);
  }
}

// Global + local views of an NCHW fp32 batch (model.py:292-301): y[0:B] = x, y[B:2B] = the centre crop [y0:y0+ch, x0:x0+cw]
// resized back to (H, W) with F.interpolate(mode="bilinear", align_corners=False) semantics.
__global__ void __launch_bounds__(256) global_local_kernel(const float* __restrict__ x, float* __restrict__ y, int64_t planes, int H,
                                                           int W, int y0, int x0, int ch, int cw) {
  const int64_t per = planes * H * W;
  const float sh = (float)ch / (float)H, sw = (float)cw / (float)W;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < 2 * per; i += (int64_t)gridDim.x * blockDim.x) {
    if (i < per) {
      y[i] = x[i];
      continue;
    }
    const int64_t e = i - per;
    const int j = (int)(e % W);
    const int r = (int)((e / W) % H);
    const int64_t plane = e / ((int64_t)W * H);
    float fr = ((float)r + 0.5f) * sh - 0.5f, fc = ((float)j + 0.5f) * sw - 0.5f;
    fr = fr < 0.f ? 0.f : fr;
    fc = fc < 0.f ? 0.f : fc;
    const int r0 = (int)fr, c0 = (int)fc;
    const int r1 = r0 + 1 < ch ? r0 + 1 : ch - 1, c1 = c0 + 1 < cw ? c0 + 1 : cw - 1;
    const float lr = fr - (float)r0, lc = fc - (float)c0;
    const float* pl = x + plane * H * W;
    const float v00 = pl[(int64_t)(y0 + r0) * W + x0 + c0], v01 = pl[(int64_t)(y0 + r0) * W + x0 + c1];
    const float v10 = pl[(int64_t)(y0 + r1) * W + x0 + c0], v11 = pl[(int64_t)(y0 + r1) * W + x0 + c1];
    y[i] = (1.f - lr) * ((1.f - lc) * v00 + lc * v01) + lr * ((1.f - lc) * v10 + lc * v11);
  }
}
}  // namespace

namespace {
__device__ __forceinline__ float sigmoidf_(float x) { return 1.f / (1.f + __expf(-x)); }

// LSTM cell, pointwise part (torch nn.LSTM gate order i, f, g, o; modules/sequence_blocks.py:21-33): gates [B, 4H] fp32 are
// the summed input / hidden projections incl. both biases.  Saves the gate activations for the backward pass.
__global__ void lstm_cell_fwd_kernel(const float* __restrict__ gates, const float* __restrict__ c_prev, float* __restrict__ h,
                                     float* __restrict__ c, float* __restrict__ act, int B, int H) {
  const int64_t n = (int64_t)B * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / H;
    const int j = (int)(e - b * H);
    const float* g4 = gates + b * 4 * H;
    const float i = sigmoidf_(g4[j]), f = sigmoidf_(g4[H + j]), g = tanhf(g4[2 * H + j]), o = sigmoidf_(g4[3 * H + j]);
    const float cp = c_prev ? c_prev[e] : 0.f;
    const float cn = fmaf(f, cp, i * g);
    c[e] = cn;
    h[e] = o * tanhf(cn);
    float* a4 = act + b * 4 * H;
    a4[j] = i; a4[H + j] = f; a4[2 * H + j] = g; a4[3 * H + j] = o;
  }
}
// dh, dc (either may be NULL = zero) -> dgates [B, 4H], dc_prev [B, H]
__global__ void lstm_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ dc, const float* __restrict__ act,
                                     const float* __restrict__ c_prev, const float* __restrict__ c, float* __restrict__ dgates,
                                     float* __restrict__ dc_prev, int B, int H) {
  const int64_t n = (int64_t)B * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / H;
    const int j = (int)(e - b * H);
    const float* a4 = act + b * 4 * H;
    const float i = a4[j], f = a4[H + j], g = a4[2 * H + j], o = a4[3 * H + j];
    const float tc = tanhf(c[e]);
    const float dhe = dh ? dh[e] : 0.f;
    const float dct = (dc ? dc[e] : 0.f) + dhe * o * (1.f - tc * tc);
    const float cp = c_prev ? c_prev[e] : 0.f;
    float* d4 = dgates + b * 4 * H;
    d4[j] = dct * g * i * (1.f - i);
    d4[H + j] = dct * cp * f * (1.f - f);
    d4[2 * H + j] = dct * i * (1.f - g * g);
    d4[3 * H + j] = dhe * tc * o * (1.f - o);
    dc_prev[e] = dct * f;
  }
}
}  // namespace

namespace {
// GRU cell, pointwise part (torch nn.GRU gate order r, z, n): gi = W_i x + b_i, gh = W_h h + b_h, both [B, 3H];
// n = tanh(gi_n + r * gh_n), h' = (1 - z) n + z h.  act keeps r, z, n for the backward.
__global__ void gru_cell_fwd_kernel(const float* __restrict__ gi, const float* __restrict__ gh, const float* __restrict__ h_prev,
                                    float* __restrict__ h, float* __restrict__ act, int B, int H) {
  const int64_t n_el = (int64_t)B * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / H;
    const int j = (int)(e - b * H);
    const float* a = gi + b * 3 * H;
    const float* c = gh + b * 3 * H;
    const float r = sigmoidf_(a[j] + c[j]), z = sigmoidf_(a[H + j] + c[H + j]);
    const float n = tanhf(a[2 * H + j] + r * c[2 * H + j]);
    h[e] = (1.f - z) * n + z * h_prev[e];
    float* s3 = act + b * 3 * H;
    s3[j] = r; s3[H + j] = z; s3[2 * H + j] = n;
  }
}
__global__ void gru_cell_bwd_kernel(const float* __restrict__ dh, const float* __restrict__ act, const float* __restrict__ gh,
                                    const float* __restrict__ h_prev, float* __restrict__ dgi, float* __restrict__ dgh,
                                    float* __restrict__ dh_prev, int B, int H) {
  const int64_t n_el = (int64_t)B * H;
  for (int64_t e = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; e < n_el; e += (int64_t)gridDim.x * blockDim.x) {
    const int64_t b = e / H;
    const int j = (int)(e - b * H);
    const float* s3 = act + b * 3 * H;
    const float r = s3[j], z = s3[H + j], n = s3[2 * H + j];
    const float d = dh[e];
    const float dn_pre = d * (1.f - z) * (1.f - n * n);
    const float dz_pre = d * (h_prev[e] - n) * z * (1.f - z);
    const float dr_pre = dn_pre * gh[b * 3 * H + 2 * H + j] * r * (1.f - r);
    float* a = dgi + b * 3 * H;
    float* c = dgh + b * 3 * H;
    a[j] = dr_pre; a[H + j] = dz_pre; a[2 * H + j] = dn_pre;
    c[j] = dr_pre; c[H + j] = dz_pre; c[2 * H + j] = dn_pre * r;
    dh_prev[e] = d * z;
  }
}
}  // namespace

extern "C" int mdhs_gru_cell_fwd(const float* gi, const float* gh, const float* h_prev, float* h, float* act, int B, int H,
                                 void* stream) {
  if (!gi || !gh || !h_prev || !h || !act || B <= 0 || H <= 0) return MDHS_ERR_ARG;
  int64_t g = ((int64_t)B * H + 255) / 256;
  if (g > (int64_t)mdhs_num_sms() * 8) g = (int64_t)mdhs_num_sms() * 8;
  g_mdhs_launches++;
  gru_cell_fwd_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gi, gh, h_prev, h, act, B, H);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_gru_cell_bwd(const float* dh, const float* act, const float* gh, const float* h_prev, float* dgi, float* dgh,
                                 float* dh_prev, int B, int H, void* stream) {
  if (!dh || !act || !gh || !h_prev || !dgi || !dgh || !dh_prev || B <= 0 || H <= 0) return MDHS_ERR_ARG;
  int64_t g = ((int64_t)B * H + 255) / 256;
  if (g > (int64_t)mdhs_num_sms() * 8) g = (int64_t)mdhs_num_sms() * 8;
  g_mdhs_launches++;
  gru_cell_bwd_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dh, act, gh, h_prev, dgi, dgh, dh_prev, B, H);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_lstm_cell_fwd(const float* gates, const float* c_prev, float* h, float* c, float* act, int B, int H, void* stream) {
  if (!gates || !h || !c || !act || B <= 0 || H <= 0) return MDHS_ERR_ARG;
  int64_t g = ((int64_t)B * H + 255) / 256;
  if (g > (int64_t)mdhs_num_sms() * 8) g = (int64_t)mdhs_num_sms() * 8;
  g_mdhs_launches++;
  lstm_cell_fwd_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(gates, c_prev, h, c, act, B, H);
  MDHS_RETURN_LAST();
}
extern "C" int mdhs_lstm_cell_bwd(const float* dh, const float* dc, const float* act, const float* c_prev, const float* c,
                                  float* dgates, float* dc_prev, int B, int H, void* stream) {
  if (!act || !c || !dgates || !dc_prev || B <= 0 || H <= 0) return MDHS_ERR_ARG;
  int64_t g = ((int64_t)B * H + 255) / 256;
  if (g > (int64_t)mdhs_num_sms() * 8) g = (int64_t)mdhs_num_sms() * 8;
  g_mdhs_launches++;
  lstm_cell_bwd_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dh, dc, act, c_prev, c, dgates, dc_prev, B, H);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_axpby_bf16(const void* x, const void* y, void* out, int64_t n, float a, float b, void* stream) {
  if (!x || !out || n <= 0 || (n % 8)) return MDHS_ERR_ARG;
  int64_t g = (n / 8 + 255) / 256;
  if (g > (int64_t)mdhs_num_sms() * 16) g = (int64_t)mdhs_num_sms() * 16;
  g_mdhs_launches++;
  axpby_bf16_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const bf16*)x, (const bf16*)y, (bf16*)out, n / 8, a, b);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_global_local(const float* x, float* y, int B, int C, int H, int W, float crop_ratio, void* stream) {
  if (!x || !y || B <= 0 || C <= 0 || H <= 0 || W <= 0 || !(crop_ratio > 0.f) || crop_ratio > 1.f) return MDHS_ERR_ARG;
  int ch = (int)((float)H * crop_ratio), cw = (int)((float)W * crop_ratio);   // int(h * ratio), model.py:294-295
  ch = ch < 1 ? 1 : ch;
  cw = cw < 1 ? 1 : cw;
  const int y0 = (H - ch) / 2 > 0 ? (H - ch) / 2 : 0, x0 = (W - cw) / 2 > 0 ? (W - cw) / 2 : 0;
  const int64_t planes = (int64_t)B * C;
  int64_t g = (2 * planes * H * W + 255) / 256;
  if (g > (int64_t)mdhs_num_sms() * 16) g = (int64_t)mdhs_num_sms() * 16;
  g_mdhs_launches++;
  global_local_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, planes, H, W, y0, x0, ch, cw);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_act_dropout_bwd(const void* dy, const void* aux, void* g, int64_t n, int act, float drop_p, uint64_t seed,
                                    void* stream) {
  if (!dy || !g || n <= 0 || (n % 8) || (act != MDHS_ACT_NONE && !aux)) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  act_dropout_bwd_kernel<<<grid_for(n / 8), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>((const bf16*)dy, (const bf16*)aux,
                                                                                              (bf16*)g, n, act, drop_p, seed);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_relu_bwd_f32(const float* dy, const float* y, float* dx, int64_t n, void* stream) {
  if (!dy || !y || !dx || n <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  relu_bwd_f32_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(dy, y, dx, n);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_sum64_to_grad(double* sum64, double* sumsq64, float* grad, int n, void* stream) {
  if (!sum64 || !grad || n <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  sum64_to_grad_kernel<<<(n + 255) / 256, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(sum64, sumsq64, grad, n);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_mul_f32(const float* a, const float* b, float* c, int64_t n, void* stream) {
  if (!a || !b || !c || n <= 0) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  mul_f32_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(a, b, c, n);
  MDHS_RETURN_LAST();
}

extern "C" int mdhs_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, void* stream) {
  if (!x || !y || n <= 0 || p < 0.f || p >= 1.f) return MDHS_ERR_ARG;
  g_mdhs_launches++;
  dropout_f32_kernel<<<grid_for(n), 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(x, y, n, p, seed);
  MDHS_RETURN_LAST();
}

// ---------------------------------------------------------------------------------------------
// GPU-side input pipeline (replaces the per-sample torchvision transforms of data_loader.py:343-372 after the
// JPEG decode): uint8 HWC batch -> per-sample crop box (RandomResizedCrop / Resize + CenterCrop geometry chosen by the
// host) -> bilinear resample to (out_h, out_w) (F.interpolate(align_corners=False) sampling, like `_center_crop` +
// interpolate of model.py:292-301) -> optional horizontal / vertical flip -> ToTensor (/255) -> Normalize(mean, std)
// -> fp32 NCHW, the layout ImageEncoder.forward receives.  One thread = one output pixel (3 channels).
// boxes: [B, 4] fp32 (y0, x0, h, w) in source pixels; flips: [B] uint8, bit 0 = hflip, bit 1 = vflip (both may be NULL).
// ---------------------------------------------------------------------------------------------
namespace {
__global__ void preprocess_u8_kernel(const uint8_t* __restrict__ src, float* __restrict__ dst, const float* __restrict__ boxes,
                                     const uint8_t* __restrict__ flips, int B, int Hs, int Ws, int Ho, int Wo, float3 mean,
                                     float3 inv_std) {
  const int64_t total = (int64_t)B * Ho * Wo;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (int64_t)gridDim.x * blockDim.x) {
    const int b = (int)(i / ((int64_t)Ho * Wo));
    const int rem = (int)(i - (int64_t)b * Ho * Wo);
    int oy = rem / Wo, ox = rem - oy * Wo;
    const uint8_t f = flips ? flips[b] : 0;
    const int sy_o = (f & 2) ? Ho - 1 - oy : oy, sx_o = (f & 1) ? Wo - 1 - ox : ox;   // flip = read the mirrored output pixel
    float y0 = 0.f, x0 = 0.f, bh = (float)Hs, bw = (float)Ws;
    if (boxes) {
      y0 = boxes[b * 4 + 0];
      x0 = boxes[b * 4 + 1];
      bh = boxes[b * 4 + 2];
      bw = boxes[b * 4 + 3];
    }
    // align_corners = False: source coordinate of the output pixel centre, clamped at 0 like ATen's area_pixel_compute
    float fy = ((float)sy_o + 0.5f) * (bh / (float)Ho) - 0.5f;
    float fx = ((float)sx_o + 0.5f) * (bw / (float)Wo) - 0.5f;
    fy = fy < 0.f ? 0.f : fy;
    fx = fx < 0.f ? 0.f : fx;
    int iy = (int)fy, ix = (int)fx;
    const float wy = fy - (float)iy, wx = fx - (float)ix;
    const int bhi = (int)bh, bwi = (int)bw;
    const int iy1 = iy + (iy < bhi - 1 ? 1 : 0), ix1 = ix + (ix < bwi - 1 ? 1 : 0);
    const int by = (int)y0, bx = (int)x0;
    auto at = [&](int yy, int xx, int c) -> float {
      int Y = by + yy, X = bx + xx;
      Y = Y < 0 ? 0 : (Y >= Hs ? Hs - 1 : Y);
      X = X < 0 ? 0 : (X >= Ws ? Ws - 1 : X);
      return (float)src[(((int64_t)b * Hs + Y) * Ws + X) * 3 + c];
    };
    const float m[3] = {mean.x, mean.y, mean.z}, is[3] = {inv_std.x, inv_std.y, inv_std.z};
#pragma unroll
    for (int c = 0; c < 3; c++) {
      const float top = at(iy, ix, c) * (1.f - wx) + at(iy, ix1, c) * wx;
      const float bot = at(iy1, ix, c) * (1.f - wx) + at(iy1, ix1, c) * wx;
      const float v = (top * (1.f - wy) + bot * wy) * (1.f / 255.f);
      dst[(((int64_t)b * 3 + c) * Ho + oy) * Wo + ox] = (v - m[c]) * is[c];
    }
  }
}
}  // namespace

extern "C" int mdhs_preprocess_u8(const uint8_t* src, float* dst, const float* boxes, const uint8_t* flips, int B, int Hs, int Ws,
                                  int Ho, int Wo, const float* mean3, const float* std3, void* stream) {
  if (!src || !dst || !mean3 || !std3 || B <= 0 || Hs <= 0 || Ws <= 0 || Ho <= 0 || Wo <= 0) return MDHS_ERR_ARG;
  const int64_t total = (int64_t)B * Ho * Wo;
  int64_t g = (total + 255) / 256;
  if (g > (int64_t)mdhs_num_sms() * 16) g = (int64_t)mdhs_num_sms() * 16;
  g_mdhs_launches++;
  preprocess_u8_kernel<<<(int)g, 256, 0, reinterpret_cast<cudaStream_t>(stream)>>>(
      src, dst, boxes, flips, B, Hs, Ws, Ho, Wo, make_float3(mean3[0], mean3[1], mean3[2]),
      make_float3(1.f / std3[0], 1.f / std3[1], 1.f / std3[2]));
  MDHS_RETURN_LAST();
}
