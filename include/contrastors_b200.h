/* contrastors_b200 -- C ABI of the B200-native contrastive hot path.
 *
 * The reference (nomic-ai/contrastors) is pure Python and has no FFI: its "operator boundary" is the set of
 * third-party native ops it calls.  Each entry point below replaces one of those
 * call sites; the Python shim in contrastors_b200/ binds them with ctypes and re-exposes the reference's
 * Python names (clip_loss, cache_loss, grad_cache_loss, gather_with_grad, LogitScale, BiEncoder, ...).
 *
 * Conventions: plain device pointers + sizes, row-major, caller owns every buffer (workspaces included),
 * every call takes a cudaStream_t (as void*) and never synchronises; returns 0 on success, non-zero on error with
 * a thread-local message from cx_last_error().  Re-entrant: may be called from PyTorch's autograd thread.
 * There is NO CPU fallback: without a sm_100 device the compute calls fail with an error.
 */
#ifndef CONTRASTORS_B200_H
#define CONTRASTORS_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef void* cx_stream_t; /* cudaStream_t */

#if defined(__GNUC__)
#define CX_API __attribute__((visibility("default")))
#else
#define CX_API
#endif

enum { CX_OK = 0, CX_ERR_INVALID = 1, CX_ERR_CUDA = 2, CX_ERR_UNSUPPORTED = 3 };
enum { CX_BF16 = 0, CX_F32 = 1 };
enum { CX_MAJOR_K = 0, CX_MAJOR_MN = 1 };

CX_API const char* cx_last_error(void);
CX_API int cx_version(void);
/* number of kernels this library has launched in the calling process (bench.py's gpu_launches evidence) */
CX_API unsigned long long cx_launch_count(void);

/* ---- dense contraction on tcgen05 (replaces torch.matmul / FusedDense = cuBLASLt:
 *      layers/attention.py:82-85,112,243  layers/mlp.py:24,61,68-83  loss.py:109)
 * C[M,N] (+)= alpha * A (x) B, bf16 operands, fp32 accumulate in TMEM.
 *   a_major = CX_MAJOR_K : A is [M,K] row-major (lda elements between rows); CX_MAJOR_MN: A is stored [K,M].
 *   b_major = CX_MAJOR_K : B is [N,K] row-major (i.e. an nn.Linear weight);  CX_MAJOR_MN: B is stored [K,N].
 *   c_dtype CX_BF16 or CX_F32; accumulate != 0 (fp32 only) adds into C with TMA reduce-add. */
CX_API int cx_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int a_major, int b_major, int64_t lda,
                 int64_t ldb, int64_t ldc, int c_dtype, int accumulate, float alpha, cx_stream_t stream);

/* ---- fused InfoNCE (replaces loss.py:105-130 = matmul + LogitScale + F.cross_entropy + argmax, and its autograd
 *      backward; modeling_dual_encoder.py:54-65; the Matryoshka loop text_text.py:352-369 via k_dim/row scales)
 * q [n, ldq] bf16, d [m, ldd] bf16 (first k_dim columns are used), logits s_ij = scale*rq_i*rd_j*<q_i,d_j>,
 * rq/rd optional fp32 per-row inverse norms (NULL = 1).  label_i = (i + label_offset) * label_stride.
 * scale_dev / coef_dev: optional DEVICE scalars multiplied into scale / coef, so a trainable LogitScale parameter and
 * autograd's grad_output never need a host read (the reference syncs on neither).
 * Forward writes: lse[n] (fp32), argmax[n] (int32, first max wins), label_logit[n], and
 *   stats[0] = sum_i (lse_i - s_i,label_i)  (caller divides by n and applies the world-size factor)
 *   stats[1] = number of rows whose argmax == label.
 * Backward: dS_ij = coef * (exp(s_ij - lse_i) - [j == label_i]) in bf16 (never leaves L2-sized workspace), then
 *   dq[n,k_dim] = scale * rq_i * sum_j dS_ij rd_j d_j   (fp32, ldq_out)
 *   dd[m,k_dim] = scale * rd_j * sum_i dS_ij rq_i q_i   (fp32, ldd_out; accumulate_dd != 0 adds into dd)
 *   stats[2]  = sum_ij dS_ij * s_ij   (= d loss / d log(scale))
 * When rq/rd are given the caller finishes the chain rule through F.normalize (cx_l2norm_bwd, g_prescaled = 1).
 * workspace: cx_infonce_workspace_bytes(n, m) bytes of device memory, reusable across calls on one stream. */
CX_API size_t cx_infonce_workspace_bytes(int n, int m);
CX_API int cx_infonce_fwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int k_dim, float scale,
                   const float* scale_dev, const float* rq, const float* rd, int label_offset, int label_stride, float* lse, int32_t* argmax,
                   float* label_logit, float* stats, void* workspace, cx_stream_t stream);
CX_API int cx_infonce_bwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int k_dim, float scale,
                   const float* scale_dev, const float* rq, const float* rd, int label_offset, int label_stride, const float* lse, float coef,
                   const float* coef_dev,
                   float* dq, int64_t lddq, float* dd, int64_t lddd, int accumulate_dd, float* stats, void* workspace,
                   cx_stream_t stream);

/* ---- row utilities for the loss path (F.normalize modeling_biencoder.py:317, text_text.py:355-356; dtype casts) */
/* y_bf16[rows,k] = bf16(x[rows,:k] * (normalize ? 1/max(||x[:k]||,eps) : 1)); inv_norm[rows] optional out */
CX_API int cx_rows_to_bf16(const float* x, int64_t ldx, void* y_bf16, int64_t ldy, float* inv_norm, int rows, int k,
                    int normalize, cx_stream_t stream);
/* backward of F.normalize over the first k columns: gx = g' - y*(g'.y), y = x*inv_norm,
 * g' = g * (g_prescaled ? 1 : inv_norm); accumulate != 0 adds into gx */
CX_API int cx_l2norm_bwd(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* inv_norm, float* gx,
                  int64_t ldgx, int rows, int k, int g_prescaled, int accumulate, cx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTRASTORS_B200_H */
