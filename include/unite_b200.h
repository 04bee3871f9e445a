/*
 * unite_b200 — C ABI of the B200-native UNITE training-step hot path.
 *
 * The reference (reddyav1/unite) is pure Python/PyTorch and has NO FFI layer; its
 * boundary is the Python API (model factories, forward signatures, state_dict keys, train_one_epoch).
 * These entry points sit *behind* that Python surface: each one replaces the library call sequence the
 * reference issues at the cited file:line.  Conventions (all functions):
 *   - plain pointers and sizes; every buffer is device memory owned by the caller (PyTorch's allocator),
 *     borrowed for the call; nothing is allocated, freed or retained by the library;
 *   - stream-ordered on `stream` (a cudaStream_t passed as void*), no host synchronisation;
 *   - returns 0 when the work was enqueued, non-zero on argument / launch error with a thread-local
 *     message available from ub_last_error();
 *   - bf16 = raw 16-bit brain-float, row-major, leading dimensions in ELEMENTS.
 */
#ifndef UNITE_B200_H
#define UNITE_B200_H
#include <stdint.h>
#if defined(__GNUC__)
#define UB_API __attribute__((visibility("default")))
#else
#define UB_API
#endif
#ifdef __cplusplus
extern "C" {
#endif

UB_API int ub_version(void);
UB_API const char* ub_last_error(void);
UB_API int ub_sm_count(void);

/* ------------------------------------------------------------------------------------------------
 * GEMM (tcgen05/TMEM/TMA):  C[M,N] = epilogue( A[M,K] * B[N,K]^T )      bf16 x bf16 -> fp32 accumulate
 * Replaces every nn.Linear / F.linear / Conv3d(stride==kernel) / `x @ proj` on the path:
 *   teacher  clip.py:55-64 (in_proj, out_proj, c_fc, c_proj), clip.py:146 (conv1), clip.py:170 (x @ proj)
 *   student  modeling_finetune.py:108 (qkv), :117 (proj), :67-71 (fc1/fc2), :174 (PatchEmbed.proj),
 *            modeling_adaptation.py:204 (Linear_Decoder.head), and their autograd backward (dgrad / wgrad).
 * ---------------------------------------------------------------------------------------------- */
enum { UB_ACT_NONE = 0, UB_ACT_QUICKGELU = 1, UB_ACT_GELU = 2, UB_ACT_DGELU = 3 };

typedef struct ub_gemm_epilogue {
  const float* bias;      /* [N] added to the accumulator, or NULL                                         */
  const float* residual;  /* fp32 [M, ldr] added last, or NULL (may alias C when C is fp32)                 */
  const float* row_scale; /* per-sample scale (DropPath keep/(1-p)), indexed row / rows_per_scale, or NULL */
  const void* aux_in;     /* bf16 [M, ld_aux]: UB_ACT_DGELU multiplies by gelu'(aux_in)                    */
  void* aux_out;          /* bf16 [M, ld_aux]: UB_ACT_GELU also stores the pre-activation here, or NULL    */
  int64_t ldr;
  int64_t ld_aux;
  int32_t rows_per_scale;
  int32_t act;      /* UB_ACT_*                                                                           */
  int32_t out_fp32; /* 0: C is bf16, 1: C is fp32                                                         */
  int32_t accumulate; /* 1: C (fp32) += result with red.add (required when split_k > 1)                   */
} ub_gemm_epilogue;

/* a_mn_major / b_mn_major = 1: the operand is stored transposed, i.e. A is [K, lda>=M] / B is [K, ldb>=N]
 * row-major (used by weight-gradient GEMMs, where the contraction runs over tokens).                      */
UB_API int ub_gemm_bf16(const void* A, int64_t lda, int a_mn_major, const void* B, int64_t ldb, int b_mn_major,
                 void* C, int64_t ldc, int M, int N, int K, const ub_gemm_epilogue* ep, int split_k,
                 void* stream);

#ifdef __cplusplus
}
#endif
#endif
