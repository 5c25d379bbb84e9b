/* mmsum_b200 — C-ABI of the B200-native MultimodalSum training-step kernels.
 *
 * The reference (nc-ai/MultimodalSum) has no FFI layer: its hot path reaches native code only
 * through torch (ATen/cuBLAS/apex).  Each entry point below therefore replaces a torch call site
 * of the reference; the file:line cited is relative to /root/reference.
 *
 * Conventions
 *   - every pointer is a DEVICE pointer owned by the caller; the library never allocates or frees
 *   - `stream` is a cudaStream_t passed as void*; all calls are asynchronous, no host sync
 *   - return value: 0 ok; < 0 invalid argument / driver problem detected on the host before launch;
 *     > 0 a cudaError_t from the launch.  No C++ exception crosses this boundary.
 *   - activations / weights are bf16 (uint16 storage), statistics / master gradients fp32,
 *     token ids int64, masks uint8 (1 = valid)
 */
#ifndef MMSUM_B200_H_
#define MMSUM_B200_H_
#include <stdint.h>
#ifdef __cplusplus
extern "C" {
#endif

/* ---- dense contraction (tcgen05/TMEM/TMA) -----------------------------------------------------
 * D[M,N] (+)= epi(alpha * sum_k A(m,k) B(n,k)).  Replaces nn.Linear/F.linear forward and the dgrad /
 * wgrad GEMMs autograd derives from them: src/transformer/modeling_multimodalsum.py:272-273 (fc1/fc2),
 * :695-704 (q/k/v/out/alpha/beta_proj), :2281 (LM head), src/img_encoder.py:40, src/table_encoder.py:70-72. */
typedef struct MmsumGemmArgs {
  const void* A;      /* bf16. K-major: [M,K] row-major (lda = row pitch in elements); MN-major: [K,M] row-major */
  const void* B;      /* bf16. K-major: [N,K] row-major; MN-major: [K,N] row-major */
  void* D;            /* [M,N] row-major, bf16 or fp32 (out_f32) */
  int64_t lda, ldb, ldd;
  int32_t M, N, K;
  int32_t a_mn_major, b_mn_major;
  int32_t out_f32;    /* 1: D is fp32 */
  int32_t accumulate; /* 1: D += (fp32 only; TMA reduce-add) */
  int32_t splits;     /* split-K factor; 0 = auto (only >1 when accumulate) */
  int32_t block_n;    /* 0 = auto, 128 or 256 */
  int32_t raster_m_fast;
  float alpha;
  const float* bias;  /* fp32 [N] or NULL, added after alpha */
  int32_t act;        /* 0 none, 1 GELU(erf), 2 ReLU */
  int32_t aux_mode;   /* 0 none, 1 store pre-activation (bf16) to aux, 2 multiply result by act'(aux) */
  void* aux;          /* bf16 [M,N] */
  int64_t ld_aux;
} MmsumGemmArgs;
int mmsum_gemm_bf16(const MmsumGemmArgs* args, void* stream);

#ifdef __cplusplus
}
#endif
#endif
