/*
 * mdhs_b200.h -- C ABI of the B200 (sm_100a) kernels behind the multimodal hot path of
 * IamJerryXu/Multimodal-Diagnosis-HAM-Spine.
 *
 * The reference has no FFI / operator layer of its own: every entry point
 * below replaces a torch / torchvision / transformers call made by the reference module named in
 * its comment (paths relative to the reference root).  The Python host side
 * (multimodal-diagnosis-ham-spine_b200/) keeps the reference's nn.Module constructors and calls
 * these through ctypes.
 *
 * Conventions: all pointers are DEVICE pointers unless said otherwise; row-major; `ld*` are row
 * strides in elements; bf16 = __nv_bfloat16; every function enqueues work on `stream` (a
 * cudaStream_t passed as void*) and returns 0 on success, a cudaError_t value or MDHS_ERR_* (>=1000)
 * otherwise.  Nothing allocates, frees or synchronises; workspaces come from the caller.
 */
#ifndef MDHS_B200_H
#define MDHS_B200_H
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

#define MDHS_ABI_VERSION 4
int mdhs_abi_version(void);
/* Number of kernels launched through this library since load (for bench.py's gpu_launches). */
int64_t mdhs_launch_count(void);

/* ---------------------------------------------------------------------------------------------
 * Dense contraction on the 5th-gen tensor cores (tcgen05.mma, TMEM accumulators, TMA operands).
 *   D[M,N] (+)= epilogue( sum_k A(m,k) * B(n,k) )
 * A is K-major ([M,lda], k contiguous) or MN-major (stored [K,lda], m contiguous); same for B
 * ([N,ldb] or stored [K,ldb]).  This one kernel therefore serves forward (X.W^T), dgrad (dY.W) and
 * wgrad (dY^T.X) of every nn.Linear / 1x1 conv / im2col'ed conv on the path without transposes:
 *   encoder.py:80-86 (proj), transformers BertSelfAttention/BertIntermediate/BertOutput Linear
 *   (via encoder.py:130-134, mibf_net/bert.py:11-13), modules/fusion_blocks.py:17-40,106-113
 *   (nn.MultiheadAttention projections, FFN), torchvision Bottleneck convs (encoder.py:61-68,
 *   mibf_net/model_resnet.py:15), ConNexT/models/block/kan1.py:156-160 (base+spline GEMMs).
 * Epilogue order: +bias[n] -> (store pre-activation to aux_out) -> act -> *act'(aux_in) -> +residual
 *   -> store bf16 / fp32 / atomic fp32 add (split-K and gradient accumulation).
 * Epilogue order: +bias -> aux_out -> act -> *act'(aux_in) -> dropout -> +residual -> store (-> colsum).
 * Requirements: N % 8 == 0, all leading dimensions % 8 == 0, 16-byte aligned pointers (K is free).
 */
enum {
  MDHS_ACT_NONE = 0, MDHS_ACT_RELU = 1, MDHS_ACT_GELU = 2,
  MDHS_ACT_GELU_DERIV = 3,   /* act: erf-GELU whose aux_out receives GELU'(pre-activation) instead of the pre-activation */
  MDHS_ACT_MUL = 4           /* dact: multiply by aux_in itself (the derivative saved by MDHS_ACT_GELU_DERIV) */
};
enum { MDHS_DT_BF16 = 0, MDHS_DT_F32 = 1 };
typedef struct {
  const void* A; int64_t lda; int32_t a_mn_major;
  const void* B; int64_t ldb; int32_t b_mn_major;
  void* D; int64_t ldd; int32_t d_dtype;      /* MDHS_DT_* */
  int32_t accumulate;                         /* 1: atomic fp32 add into D (d_dtype must be F32) */
  int32_t M, N, K;
  const float* bias;                          /* fp32 [N] or NULL */
  void* aux_out; int64_t ld_aux_out;          /* bf16 pre-activation copy or NULL */
  const void* aux_in; int64_t ld_aux_in;      /* bf16 tensor feeding act' (dact) or NULL */
  int32_t act;                                /* MDHS_ACT_* applied in the epilogue */
  int32_t dact;                               /* MDHS_ACT_*: multiply by act'(aux_in) */
  const void* residual; int64_t ldr; int32_t r_dtype;
  int32_t split_k;                            /* 0/1 = none; >1 needs accumulate=1; <0 = auto (with accumulate) */
  int32_t bn_hint;                            /* 0 = auto; else 64/128/256 tile width */
  double* colsum; double* colsumsq;           /* optional fp64 [N] atomics: per-column sum / sum of
                                                 squares of the stored value (train-mode BN stats) */
  float dropout_p; uint64_t dropout_seed;     /* inverted dropout after act/act' (mask = hash(seed, m*N+n)) */
  /* Implicit-GEMM convolution (torchvision Conv2d 3x3 / strided 1x1 of the ResNet trunk, encoder.py:61-72): one operand
   * is the NHWC bf16 activation X[cN, cH, cW, cC] read through a TMA im2col descriptor, no patch matrix in HBM.
   *   conv_mode 1: A(m, k) = X[n, p*stride - pad + r, q*stride - pad + s, c], m = (n, p, q) output pixel, k = (r, s, c)
   *                (A = X, lda ignored, M = cN*Ho*Wo, K = cR*cS*cC): fprop with B = packed weights [O, (r,s,c)], and
   *                stride-1 dgrad with X = dY, pad' = R-1-pad and B = mdhs_conv_weight_pack_dgrad weights;
   *   conv_mode 2: B(n, k) = the same matrix with n = (r, s, c), k = pixel (B = X, b_mn_major = 1, ldb ignored): wgrad
   *                dW[o, (r,s,c)] = sum_pixels dY[pixel, o] * im2col(X)[pixel, (r,s,c)] with A = dY (a_mn_major = 1).
   * cC % 64 == 0, cR == cS, 1 <= c_stride <= 8. */
  int32_t conv_mode; int32_t cN, cH, cW, cC, cR, cS, c_stride, c_pad;
  /* BatchNorm-backward reduction fused into the epilogue of the GEMM that PRODUCES dy (a dgrad GEMM whose output is the
   * gradient wrt a BN(+ReLU) output; torchvision Bottleneck conv1/conv2 via encoder.py:61-72): with stat_x = the raw
   * (pre-BN) activation [M, N] bf16, colsum receives sum_m dy'[m,n] and colsumsq sum_m dy'[m,n] * (stat_x[m,n] -
   * stat_mean[n]), dy' = dy * [fmaf(stat_x, stat_scale, stat_shift) > 0] when stat_relu (else dy' = dy).  D itself stays
   * the unmasked dy.  Needs colsum/colsumsq, bf16 D, no aux_in / aux_out / bf16 residual, no split-K, tile width <= 128. */
  const void* stat_x; int64_t ld_stat_x; const float* stat_mean; const float* stat_scale; const float* stat_shift;
  int32_t stat_relu;
} mdhs_gemm_args;
int mdhs_gemm_bf16(const mdhs_gemm_args* args, void* stream);

/* ---------------------------------------------------------------------------------------------
 * LayerNorm (nn.LayerNorm in modules/fusion_blocks.py:18,22,34,113,211,240; modules/heads.py:35;
 * transformers BertEmbeddings/BertSelfOutput/BertOutput LayerNorm, eps 1e-12).  x is bf16 or fp32
 * [rows, C]; statistics fp32; optional inverted dropout on the output (BERT embeddings).
 * Backward accumulates (+=) dgamma/dbeta and can emit a second dx copy masked by the dropout of the
 * dense branch that fed the residual sum; dbias (optional) += column sums of that copy (dx_drop, else dx) = the bias
 * gradient of that dense layer.  C % 8 == 0, C <= 2048.
 */
int mdhs_layernorm_fwd(const void* x, int x_f32, int64_t ldx, const float* gamma, const float* beta, void* y_bf16,
                       int64_t ldy, float* y_f32, float* mean, float* rstd, int rows, int C, float eps,
                       float drop_p, uint64_t seed, void* stream);
int mdhs_layernorm_bwd(const void* dy, int dy_f32, int64_t lddy, const void* x, int x_f32, int64_t ldx,
                       const float* mean, const float* rstd, const float* gamma, void* dx_bf16, int64_t lddx,
                       void* dx_drop_bf16, float* dx_f32, float* dgamma, float* dbeta, float* dbias, int rows, int C,
                       float drop_p, uint64_t seed, float drop2_p, uint64_t seed2, void* stream);

/* ---------------------------------------------------------------------------------------------
 * BatchNorm2d on NHWC bf16 [rows = B*H*W, C] (torchvision ResNet: encoder.py:61-68,
 * mibf_net/model_resnet.py:15; eps 1e-5, momentum 0.1).  Train mode: the conv GEMM epilogue (or
 * mdhs_col_stats) produces fp64 per-channel sums; finalize turns them into mean/invstd, updates the
 * running statistics and emits scale/shift; apply fuses normalise + residual add + ReLU.
 * mdhs_bn_fwd = finalize + apply as one call (scale / shift from the fp64 sums or, training == 0, from the running
 * statistics; mean / invstd / scale / shift are published for the backward).
 * mdhs_bn_bwd = reduce, coefficients, apply; relu != 0 masks dy with (y > 0); dz optionally receives the
 * masked dy (gradient of the identity branch); dgamma/dbeta accumulate (+=).  Workspaces: sum_dy = sum(dy'),
 * sum_dy_xc = sum(dy' * (x - mean)) (fp64 [C] each) and coef (fp32 [5*C]).
 * With relu != 0 and y == NULL the mask is recomputed from x as fmaf(x, scale, shift) > 0 (bit-identical to what the
 * forward evaluated; only valid for layers without a residual input), which saves one full read of y in both passes.
 * relu_mask (optional, uint8 [rows, C/8], bit k of byte j = channel 8j+k): mdhs_bn_fwd writes (output > 0) per element and
 * mdhs_bn_bwd reads it instead of y -- 1 bit instead of 16 per element in both backward passes of the residual layers.
 * training == 0: eval-mode backward, dx = gamma * invstd * dy' (running statistics are constants).
 * sums_ready != 0: the two sums were already produced (by the epilogue of the GEMM that wrote dy, see
 * mdhs_gemm_args.stat_x); the reduce launch is skipped.
 */
int mdhs_bn_finalize(const double* colsum, const double* colsumsq, int64_t count, const float* gamma,
                     const float* beta, float* running_mean, float* running_var, float momentum, float eps,
                     float* mean, float* invstd, float* scale, float* shift, int C, int training, void* stream);
int mdhs_bn_apply(const void* x, const float* scale, const float* shift, const void* residual, void* y,
                  int64_t rows, int C, int relu, void* stream);
int mdhs_bn_fwd(const void* x, const double* colsum, const double* colsumsq, const float* gamma, const float* beta,
                float* running_mean, float* running_var, float momentum, float eps, const void* residual, void* y,
                float* mean, float* invstd, float* scale, float* shift, void* relu_mask, int64_t rows, int C, int relu,
                int training, void* stream);
int mdhs_bn_bwd(const void* dy, const void* x, const void* y, const void* relu_mask, const float* mean, const float* invstd,
                const float* gamma, const float* scale, const float* shift, double* sum_dy, double* sum_dy_xc,
                float* coef, void* dx, void* dz, float* dgamma, float* dbeta, int64_t rows, int C, int relu, int training,
                int sums_ready, void* stream);
/* column sums of a bf16 [rows, C] matrix: fp64 sum / sum of squares (BN statistics) and/or fp32 += (bias grads) */
int mdhs_col_stats(const void* x, int64_t ldx, double* sum64, double* sumsq64, float* sum32, int64_t rows, int C,
                   void* stream);

/* ---------------------------------------------------------------------------------------------
 * Convolution plumbing (NHWC bf16).  Non-1x1 convolutions are lowered to mdhs_gemm_bf16 on patch
 * matrices: col[(b,ho,wo), (r*S+s)*C + c].  col2im is the gather-form gradient (optionally + `add`).
 * 3x3/2 max-pool keeps the winning tap index for a deterministic backward.  (torchvision ResNet stem /
 * Bottleneck convs used by encoder.py:61-72 and mibf_net/model_resnet.py:15.)
 */
int mdhs_im2col_nchw_f32(const float* x, void* col, int B, int C, int H, int W, int R, int S, int stride, int pad,
                         int ldc, void* stream);
/* stem patch matrix of V test-time-augmentation variants (codes: 4 bits per variant, 0 identity 1 hflip 2 vflip 3 rot90;
 * scripts/predict.py:33-42) read straight from the un-expanded batch: V * B * Ho * Wo rows */
int mdhs_im2col_nchw_f32_tta(const float* x, void* col, int B, int C, int H, int W, int R, int S, int stride, int pad,
                             int ldc, int V, int codes, void* stream);
int mdhs_im2col_nhwc(const void* x, void* col, int B, int H, int W, int C, int R, int S, int stride, int pad,
                     void* stream);
int mdhs_col2im_nhwc(const void* dcol, const void* add, void* dx, int B, int H, int W, int C, int R, int S,
                     int stride, int pad, void* stream);
int mdhs_maxpool3x3s2_fwd(const void* x, void* y, void* idx, int B, int H, int W, int C, void* stream);
int mdhs_maxpool3x3s2_bwd(const void* dy, const void* idx, void* dx, int B, int H, int W, int C, void* stream);
/* mean over tokens (encoder.py tokens -> model.py:283-290 pooling, fusion_blocks.py:96-98,143-145) */
int mdhs_mean_tokens_fwd(const void* x, float* y32, void* y16, int B, int T, int C, float scale, int accumulate,
                         void* stream);
int mdhs_mean_tokens_bwd(const float* dy32, const void* dy16, void* dx, int B, int T, int C, float scale,
                         void* stream);
/* weight re-layout OIHW fp32 <-> GEMM operand [O, (r*S+s)*I + i] (row stride ldk) */
int mdhs_conv_weight_pack(const float* w, void* wp, int O, int I, int R, int S, int ldk, void* stream);
int mdhs_conv_weight_pack_dgrad(const float* w, void* wt, int O, int I, int R, int S, void* stream);
int mdhs_conv_wgrad_unpack(const float* gp, float* g, int O, int I, int R, int S, int ldk, void* stream);
int mdhs_cast_f32_bf16(const float* x, void* y, int64_t n, void* stream);
int mdhs_cast_bf16_f32(const void* x, float* y, int64_t n, void* stream);
/* GPU-side input pipeline (data_loader.py:343-372 after the decode): uint8 [B,Hs,Ws,3] -> per-sample crop box boxes[B,4]
 * (y0,x0,h,w; NULL = whole image) -> bilinear (align_corners=False) resample to Ho x Wo -> flips[B] (bit0 hflip, bit1 vflip;
 * NULL = none) -> /255 -> (x - mean) / std -> fp32 [B,3,Ho,Wo].  mean3 / std3 are HOST pointers to 3 floats. */
int mdhs_preprocess_u8(const uint8_t* src, float* dst, const float* boxes, const uint8_t* flips, int B, int Hs, int Ws, int Ho,
                       int Wo, const float* mean3, const float* std3, void* stream);
int mdhs_nhwc_bf16_to_nchw_f32(const void* x, float* y, int B, int H, int W, int C, void* stream);
int mdhs_nchw_f32_to_nhwc_bf16(const float* x, void* y, int B, int H, int W, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Fused multi-head attention, forward and backward, for short sequences (Sk <= 512, D in {32, 64}):
 * softmax(scale * Q K^T + key mask) (dropout) V.  Replaces the eager attention of transformers
 * BertSelfAttention (encoder.py:130-134) and nn.MultiheadAttention (fusion_blocks.py:19-32,107-112).
 * Token-major bf16 buffers [B*S, ld], head h in columns [h*D,(h+1)*D); key_mask [B,Sk] 1 = attend.
 */
int mdhs_attention_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, void* out,
                       int64_t ldo, const uint8_t* key_mask, float* lse, int B, int H, int Sq, int Sk, int D,
                       float scale, float drop_p, uint64_t seed, void* stream);
int mdhs_attention_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv,
                       const void* o, const void* d_o, int64_t ldo, const uint8_t* key_mask, const float* lse,
                       void* dq, void* dk, void* dv, float* dk32, float* dv32, int B, int H, int Sq, int Sk, int D,
                       float scale, float drop_p, uint64_t seed, void* stream);
/* dq / dk / dv are bf16 outputs with the strides of q / k / v.  Sq <= 64: one CTA per head.  Sq > 64: a dQ pass (CTAs own
 * 128 queries) and a dK/dV pass (CTAs own 64 keys and stream every query block past them) -- no atomics; dk32, if given,
 * is a [B*H*Sq] fp32 scratch through which the dQ pass hands delta = rowsum(dO * O) to the dK/dV pass (NULL: recomputed),
 * dv32 is unused.  Only when mdhs_attention_bwd_workspace(Sq, Sk, D) returns 1 (shapes whose tiles do not fit in shared
 * memory) must the caller pass zeroed fp32 workspaces dk32 [B*Sk, ldk] / dv32 [B*Sk, ldv], which then receive dK / dV
 * instead of dk / dv. */
int mdhs_attention_bwd_workspace(int Sq, int Sk, int D);

/* BERT embeddings (transformers BertEmbeddings via encoder.py:130-134): gather + gradient scatter */
int mdhs_embed_gather(const int64_t* ids, const int64_t* type_ids, const float* word, const float* pos,
                      const float* type, float* e, int rows, int S, int C, int vocab, void* stream);
int mdhs_embed_scatter(const float* de, const int64_t* ids, const int64_t* type_ids, float* gword, float* gpos,
                       float* gtype, int rows, int S, int C, int vocab, void* stream);

/* test-time augmentation (scripts/predict.py:33-42): V variants of a square NCHW fp32 batch stacked on the batch axis;
 * codes = 4 bits per variant: 0 identity, 1 hflip (flip(-1)), 2 vflip (flip(-2)), 3 rot90 (torch.rot90 k=1, dims (-2,-1)) */
int mdhs_tta_expand(const float* x, float* y, int B, int C, int H, int W, int V, uint32_t codes, void* stream);

/* ---------------------------------------------------------------------------------------------
 * Small fp32 head kernels: nn.Linear with few outputs (model.py:195-200 classifier, gating.py:10-14),
 * cross entropy with label smoothing / class weights / focal form (scripts/train.py:46-61,252-254).
 */
int mdhs_linear_f32_fwd(const float* X, int64_t ldx, const float* W, const float* bias, float* Y, int64_t ldy, int M,
                        int N, int K, int act, void* stream);
int mdhs_linear_f32_bwd(const float* dY, int64_t lddy, const float* X, int64_t ldx, const float* W, float* dX,
                        int64_t lddx, int accumulate_dx, float* dW, float* db, int M, int N, int K, void* stream);
int mdhs_ce_loss(const float* logits, int64_t ld, const int64_t* labels, const float* class_weights, float* loss,
                 float* dlogits, int B, int C, float label_smoothing, int focal, float gamma, void* stream);
int mdhs_axpby_f32(const float* x, float* y, int64_t n, const float* a_dev, float a, float b, void* stream);
/* supervised contrastive loss on (B, D) fp32 features (scripts/train.py:23-44 SupConLoss, temperature 0.07): loss[0] and,
 * when dx != NULL, d loss / d x.  Workspaces: f_ws [B*D], inv_norm_ws [B], g_ws [B*B] (fp32). */
int mdhs_supcon_loss(const float* x, int64_t ldx, const int64_t* labels, float* loss, float* dx, int64_t lddx, float* f_ws,
                     float* inv_norm_ws, float* g_ws, int B, int D, float temperature, void* stream);

/* elementwise helpers between fused ops: g = dy * dropout_mask * act'(aux) (bf16), ReLU backward, product (fp32) */
int mdhs_act_dropout_bwd(const void* dy, const void* aux, void* g, int64_t n, int act, float drop_p, uint64_t seed,
                         void* stream);
int mdhs_relu_bwd_f32(const float* dy, const float* y, float* dx, int64_t n, void* stream);
int mdhs_mul_f32(const float* a, const float* b, float* c, int64_t n, void* stream);
/* grad[i] += sum64[i], then sum64 / sumsq64 (optional) are zeroed: turns the fp64 column sums a GEMM epilogue produced
 * (mdhs_gemm_args.colsum on the dgrad that writes the pre-activation gradient) into a Linear bias gradient */
int mdhs_sum64_to_grad(double* sum64, double* sumsq64, float* grad, int n, void* stream);
/* adaptive level weighting of the hierarchical fusion (README.md:15; fusion_type "hierarchical"): out = sum_l softmax(logits)_l
 * p[l] over L <= 4 fp32 vectors of n elements (p / dp are HOST arrays of device pointers); backward writes dp[l] = w_l * dout
 * and accumulates dlogits (+=); g_ws = fp32 [4] workspace. */
int mdhs_level_mix_fwd(const float* const* p, const float* logits, float* out, int64_t n, int L, void* stream);
int mdhs_level_mix_bwd(const float* const* p, float* const* dp, const float* logits, const float* dout, float* g_ws,
                       float* dlogits, int64_t n, int L, void* stream);
int mdhs_dropout_f32(const float* x, float* y, int64_t n, float p, uint64_t seed, void* stream);
/* global-local branch (model.py:292-315): out = a*x (+ b*y) on bf16 token tensors (0.5 * (global + local)); the
 * [global | centre-crop resized bilinearly, align_corners = False] image pair stacked on the batch axis */
int mdhs_axpby_bf16(const void* x, const void* y, void* out, int64_t n, float a, float b, void* stream);
int mdhs_global_local(const float* x, float* y, int B, int C, int H, int W, float crop_ratio, void* stream);
/* sequence branch (modules/sequence_blocks.py:21-33,60-64, nn.LSTM): pointwise LSTM cell on the summed gate projections
 * [B, 4H] (gate order i, f, g, o); act keeps the gate activations for the backward; NULL c_prev / dh / dc mean zeros */
int mdhs_lstm_cell_fwd(const float* gates, const float* c_prev, float* h, float* c, float* act, int B, int H, void* stream);
int mdhs_lstm_cell_bwd(const float* dh, const float* dc, const float* act, const float* c_prev, const float* c, float* dgates,
                       float* dc_prev, int B, int H, void* stream);
/* nn.GRU cell (gate order r, z, n): gi = W_i x + b_i and gh = W_h h + b_h, both [B, 3H]; act keeps r, z, n */
int mdhs_gru_cell_fwd(const float* gi, const float* gh, const float* h_prev, float* h, float* act, int B, int H, void* stream);
int mdhs_gru_cell_bwd(const float* dh, const float* act, const float* gh, const float* h_prev, float* dgi, float* dgh,
                      float* dh_prev, int B, int H, void* stream);

/* ---------------------------------------------------------------------------------------------
 * MIBF-Net: IBFA cross-attention with one token per modality (mibf_net/attention.py:47-70; keys/values of x and
 * y concatenated -> softmax over 2 keys per head) on the fused projections [K_x|Q_x|V_x] and [K_y|V_y];
 * MP-Loss forward + backward (mibf_net/model_resnet.py:76-94, attention.py:25-28).
 */
int mdhs_ibfa_fwd(const void* kqv_x, int64_t ldx, const void* kv_y, int64_t ldy, void* out, float* probs, int B, int H,
                  int D, void* stream);
int mdhs_ibfa_bwd(const void* kqv_x, int64_t ldx, const void* kv_y, int64_t ldy, const void* dout, const float* probs,
                  void* dkqv_x, void* dkv_y, int B, int H, int D, void* stream);
int mdhs_mp_loss(const float* img_logits, const float* txt_logits, const float* fused_logits, const int64_t* labels,
                 float* loss, float* g_img, float* g_txt, float* g_fused, int B, int C, void* stream);

/* ---------------------------------------------------------------------------------------------
 * KAN (ConNexT/models/block/kan1.py:77-165, grid_size 5 / spline_order 3) and sparsely-gated MoE
 * (ConNexT/models/block/moe.py:171-291).  kan_basis_*: x -> bf16 GEMM operand [SiLU(x) | 8 cubic B-spline bases per
 * input] and its derivative; weight pack/unpack between (base_weight, spline_weight, spline_scaler) and the fused
 * operand; MoE noisy top-k gating forward/backward, cv^2 balance loss and gate-weighted expert combination.
 */
int mdhs_kan_basis_fwd(const float* x, int64_t ldx, const float* grid, void* op, int64_t rows, int in, int ld_op,
                       void* stream);
int mdhs_kan_basis_bwd(const float* x, int64_t ldx, const float* grid, const void* dop, float* dx, int64_t rows, int in,
                       int ld_op, int accumulate, void* stream);
int mdhs_kan_weight_pack(const float* base_w, const float* spline_w, const float* scaler, void* wcat, int out, int out_pad,
                         int in, int ld, void* stream);
int mdhs_kan_wgrad_unpack(const float* gcat, const float* spline_w, const float* scaler, float* g_base, float* g_spline,
                          float* g_scaler, int out, int in, int ld, void* stream);
int mdhs_moe_gate_fwd(const float* x, const float* wg, const float* wn, const float* noise, float* gates, float* clean,
                      float* raw, float* probs, int* topidx, float* importance, float* load, const float* normal_mean,
                      const float* normal_std, int B, int in, int E, int k, int noisy, void* stream);
int mdhs_moe_loss(const float* importance, const float* load, float* loss, float* d_imp, float* d_load, int E, float coef,
                  void* stream);
int mdhs_moe_gate_bwd(const float* x, const float* wg, const float* wn, const float* noise, const float* dgates,
                      const float* d_imp, const float* d_load, float dloss, const float* dloss_dev, const float* clean,
                      const float* raw, const float* probs, const int* topidx, float* dx, float* dwg, float* dwn,
                      const float* normal_mean, const float* normal_std, int B, int in, int E, int k, int noisy, void* stream);
/* standard-normal noise (moe.py:246 randn_like) from the counter-based generator shared with dropout */
int mdhs_randn_f32(float* out, int64_t n, uint64_t seed, void* stream);
int mdhs_moe_combine_fwd(const float* gates, const float* Y, float* y, int B, int E, int C, int ldy, void* stream);
int mdhs_moe_combine_bwd(const float* gates, const float* Y, const float* dy, float* dgates, float* dY, int B, int E, int C,
                         int ldy, void* stream);

/* ---------------------------------------------------------------------------------------------
 * ConvNeXt (torchvision CNBlock behind ConNexT/models/ourmodel.py:57-62 and ConNexT/models/image_encoder.py): depthwise
 * 7x7 convolution (pad 3) on NHWC bf16 -- forward (flip = 0, + bias) / input gradient (flip = 1: mirrored taps on dy)
 * and weight + bias gradient (dw [C][49], db [C]: fp32 +=); layer-scale + row-mode stochastic depth + residual
 * out = x + ls[c] * keep(sample) * z with keep in {0, 1/(1-p)} drawn per sample, and its backward (dz, dls +=).
 * Single-query attention of the text -> image CrossAttention (ourmodel.py:17-31, a 1-token query over T image
 * positions; the reference applies no 1/sqrt(d)): probs = softmax_t(scale * q.k_t), out = sum_t probs_t v_t.
 */
int mdhs_dwconv7_fwd(const void* x, const float* w, const float* bias, void* y, int B, int H, int W, int C, int flip,
                     void* stream);
int mdhs_dwconv7_wgrad(const void* x, const void* dy, float* dw, float* db, int B, int H, int W, int C, void* stream);
int mdhs_layer_scale_fwd(const void* x, const void* z, const float* ls, void* out, int64_t rows, int C, int rows_per_sample,
                         float p, uint64_t seed, void* stream);
/* dbias (optional): += column sums of dz, the bias gradient of the Linear that produced z (CNBlock block.5) */
int mdhs_layer_scale_bwd(const void* dy, const void* z, const float* ls, void* dz, float* dls, float* dbias, int64_t rows,
                         int C, int rows_per_sample, float p, uint64_t seed, void* stream);
int mdhs_sq_attn_fwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, float* out,
                     float* probs, int B, int T, int D, float scale, void* stream);
int mdhs_sq_attn_bwd(const void* q, int64_t ldq, const void* k, int64_t ldk, const void* v, int64_t ldv, const float* dout,
                     const float* probs, void* dq, int64_t lddq, void* dk, int64_t lddk, void* dv, int64_t lddv, int B, int T,
                     int D, float scale, void* stream);

/* Fused optimizer step on the flat parameter buffer (scripts/train.py:257-309).  grads_bf16 (optional): read the gradient
 * from this bf16 buffer instead of `grads` (data-parallel runs all-reduce bf16 buckets; `grads` is still zeroed).
 * blocks_per_sm: > 0 (0 = 16) resident 256-thread blocks per SM; < 0: leave that many SMs free (one 1024-thread block on each
 * of the others) so that the collective of the next gradient bucket can run beside the optimizer of the current one. */
int mdhs_adam_flat(float* params, float* grads, const void* grads_bf16, float* exp_avg, float* exp_avg_sq, void* shadow_bf16,
                   int64_t n, float lr, float beta1, float beta2, float eps, float weight_decay, int step, float grad_scale,
                   int adamw, int zero_grad, const float* lr_dev, const int* step_dev, int blocks_per_sm, void* stream);
int mdhs_sgd_flat(float* params, float* grads, const void* grads_bf16, float* momentum_buf, void* shadow_bf16, int64_t n,
                  float lr, float momentum, float weight_decay, float grad_scale, int first_step, int zero_grad,
                  const float* lr_dev, const int* step_dev, int blocks_per_sm, void* stream);
/* Persistent GEMM grids leave `n` SMs free (0 = use all): set while a collective kernel (NCCL all-reduce of gradient
 * buckets, mibf_net/train_resnet.py:84-88's DDP) runs concurrently, so that the GEMM's CTAs are all co-resident instead of
 * spilling a second wave behind the collective's CTAs. */
int mdhs_set_sm_reserve(int n);
/* Work distribution of the persistent GEMM grids: 1 (default; MDHS_GEMM_DYNAMIC=0 in the environment starts with 0) = CTAs
 * draw work items from a global counter, so a grid that does not get all of its SMs at once (a collective's CTAs, another
 * stream's kernels) loses nothing but those SMs; 0 = static round-robin.  GEMMs with column statistics in the epilogue and
 * GEMMs of at most one round of tiles always use the static schedule.  The counter slots (32 KB) are allocated on the device
 * that is current at the first large GEMM issued outside stream capture: one device per process, like the reference's
 * one-process-per-GPU DDP launch. */
int mdhs_set_gemm_dynamic(int on);
/* lr_dev / step_dev (optional device scalars) override lr / step so a captured CUDA graph of the step can be
 * replayed with a changing learning rate and step count.  mdhs_step_begin: once per step before the forward:
 * ++*step_dev and advance the dropout seed tick folded into every stateless dropout mask. */
int mdhs_step_begin(int* step_dev, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDHS_B200_H */
