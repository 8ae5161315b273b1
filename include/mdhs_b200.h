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

#define MDHS_ABI_VERSION 1
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
 * Requirements: K % 8 == 0, N % 2 == 0, lda/ldb % 8 == 0, 16-byte aligned A/B.
 */
enum { MDHS_ACT_NONE = 0, MDHS_ACT_RELU = 1, MDHS_ACT_GELU = 2 };
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
  int32_t split_k;                            /* 0/1 = none; >1 needs accumulate=1 */
  int32_t bn_hint;                            /* 0 = auto; else 64/128/256 tile width */
  double* colsum; double* colsumsq;           /* optional fp64 [N] atomics: per-column sum / sum of
                                                 squares of the stored value (train-mode BN stats) */
} mdhs_gemm_args;
int mdhs_gemm_bf16(const mdhs_gemm_args* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* MDHS_B200_H */
