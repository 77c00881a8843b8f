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

/* A/B switch of the GEMM launch mode, process-wide: the largest cluster the launcher may use (0 = default = 2: CTA pairs,
 * cta_group::2; 4 = two CTA pairs sharing their B tile by TMA multicast when M % 512 == 0 -- measured slower, opt-in;
 * 1 = one CTA per tile). */
CX_API int cx_gemm_select_cluster(int max_cluster_ctas);

/* ---- dense contraction on tcgen05 (replaces torch.matmul / FusedDense = cuBLASLt:
 *      layers/attention.py:82-85,112,243  layers/mlp.py:24,61,68-83  loss.py:109)
 * C[M,N] (+)= alpha * A (x) B, bf16 operands, fp32 accumulate in TMEM.
 *   a_major = CX_MAJOR_K : A is [M,K] row-major (lda elements between rows); CX_MAJOR_MN: A is stored [K,M].
 *   b_major = CX_MAJOR_K : B is [N,K] row-major (i.e. an nn.Linear weight);  CX_MAJOR_MN: B is stored [K,N].
 *   c_dtype CX_BF16 or CX_F32; accumulate != 0 (fp32 only) adds into C with TMA reduce-add. */
CX_API int cx_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int a_major, int b_major, int64_t lda,
                 int64_t ldb, int64_t ldc, int c_dtype, int accumulate, float alpha, cx_stream_t stream);

/* Gated-MLP backward through the activation, fused into the fc2 dgrad GEMM: dyg[M, 2I] = [ da * silu(g) | da * y * silu'(g) ]
 * with da = dout[M, K] x w2[K, I] kept in TMEM (never written to HBM) and yg = [y | g] the pre-activations saved by
 * cx_gemm_swiglu.  Replaces the autograd of flash-attn's swiglu + the fc2 input-gradient GEMM
 * (/root/reference/src/contrastors/layers/mlp.py:68-83).  I % 256 == 0, M >= 256. */
CX_API int cx_gemm_swiglu_bwd(const void* dout, const void* w2, const void* yg, void* dyg, int M, int I, int K, int64_t ld_dout,
                       int64_t ld_w2, int64_t ld_yg, int64_t ld_dyg, cx_stream_t stream);

/* ---- QKV projection with the rotary embedding fused into the GEMM epilogue (layers/attention.py:112-133 =
 * Wqkv GEMM + apply_rotary_emb x2 + torch.stack): qkv[T, n_out] = x w^T; heads (64 columns) inside [0, rope_cols) are
 * rotated NeoX-style by the angles pos[t] * inv_freq[j] (inv_freq: 32 fp32 values base^(-2j/64); the angle is formed in fp32 as
 * the reference's cos/sin cache does, cos/sin evaluated in the epilogue). */
CX_API int cx_gemm_qkv_rope(const void* x, const void* w, void* qkv, int T, int n_out, int K, int64_t ldx, int64_t ldw,
                     int64_t ldo, const int32_t* pos, const float* inv_freq, int rope_cols, cx_stream_t stream);

/* ---- gated-MLP first layer with the SwiGLU fused into the GEMM epilogue (layers/mlp.py:68-75: fc11, fc12, swiglu)
 * w1 [2I, K] = [fc11; fc12] (nn.Linear layout); act_out[M, I] = (x fc11^T) * silu(x fc12^T), bf16;
 * yg_out [M, 2I] = [x fc11^T | x fc12^T] kept for the backward, or NULL (no-grad forward). I % 128 == 0. */
CX_API int cx_gemm_swiglu(const void* x, const void* w1, void* act_out, void* yg_out, int M, int I, int K, int64_t ldx,
                   int64_t ldw, int64_t ld_act, int64_t ld_yg, cx_stream_t stream);

/* ---- fused InfoNCE (replaces loss.py:105-130 = matmul + LogitScale + F.cross_entropy + argmax, and its autograd
 *      backward; modeling_dual_encoder.py:54-65; the Matryoshka loop text_text.py:352-369 via k_dim/row scales)
 * q [n, ldq] bf16, d [m, ldd] bf16 (first k_dim columns are used), logits s_ij = scale*rq_i*rd_j*<q_i,d_j>,
 * rq/rd optional fp32 per-row inverse norms (NULL = 1).  label_i = (i + label_offset) * label_stride.
 * scale_dev / coef_dev: optional DEVICE scalars multiplied into scale / coef, so a trainable LogitScale parameter and
 * autograd's grad_output never need a host read (the reference syncs on neither).
 * Forward writes: lse[n] (fp32), argmax[n] (int32, first max wins), label_logit[n], and
 *   stats[0] = sum_i (lse_i - s_i,label_i)  (caller divides by n and applies the world-size factor)
 *   stats[1] = number of rows whose argmax == label.
 * Backward: dS_ij = exp(s_ij - lse_i) - [j == label_i], UNSCALED in [-1, 1], as fp16 in the workspace; then
 *   dq[n,k_dim] = coef * scale * sum_j dS_ij (rd_j d_j)   (fp32, ldq_out)  = d loss / d (rq_i q_i)
 *   dd[m,k_dim] = coef * scale * sum_i dS_ij (rq_i q_i)   (fp32, ldd_out; accumulate_dd != 0 adds into dd)
 *   stats[2]  = coef * sum_ij dS_ij * s_ij   (= d loss / d log(scale))
 * i.e. the gradients with respect to the (normalised) rows the logits were formed from; when rq/rd are given the caller
 * finishes the chain rule through F.normalize (cx_l2norm_bwd, g_prescaled = 0).  The contractions read fp16 copies of
 * rq_i q_i / rd_j d_j (exact for bf16 values in [2^-14, 65504], saturating beyond: the label_stride >= 1 and finite-operand
 * contract of this entry point; the reference's all-zero labels for ungathered documents at world size > 1 are rejected).
 * workspace: cx_infonce_workspace_bytes(n, m, k_dim) bytes of device memory, reusable across calls on one stream. */
CX_API size_t cx_infonce_workspace_bytes(int n, int m, int k_dim);
CX_API int cx_infonce_fwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int k_dim, float scale,
                   const float* scale_dev, const float* rq, const float* rd, int label_offset, int label_stride, float* lse, int32_t* argmax,
                   float* label_logit, float* stats, void* workspace, cx_stream_t stream);
/* Matryoshka forward in ONE accumulation over K = dims[n_dims-1] (replaces the per-dim loop of trainers/text_text.py:352-369):
 * dims ascending multiples of 64 (<= 8 of them); rq [n_dims][n], rd [n_dims][m] = inverse L2 norms of the row prefixes; outputs per
 * prefix: lse / argmax / label_logit [n_dims][n], stats [n_dims][4] (stats[s][0] = sum_i (lse - label logit), [s][1] = hits).
 * 2*n*m*K FLOPs instead of 2*n*m*sum(dims); each prefix's logits are the running sum of segment products kept in tensor memory. */
CX_API size_t cx_infonce_mat_workspace_bytes(int n, int m, int n_dims);
CX_API int cx_infonce_mat_fwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int n_dims, const int32_t* dims,
                       float scale, const float* scale_dev, const float* rq, const float* rd, int label_offset, int label_stride,
                       float* lse, int32_t* argmax, float* label_logit, float* stats, void* workspace, cx_stream_t stream);
/* Matryoshka backward from ONE more accumulation (2..4 prefix dims): emits, per SEGMENT t of the columns, the raw gradient
 * contribution dq_raw[:, seg_t] = a * T_t d[:, seg_t], dd_raw[:, seg_t] = a * T_t^T q[:, seg_t] with a = scale * coef * (*coef_gamma_dev)
 * and T_t = sum_{s>=t} wrel_s diag(rq_s)(softmax_s - onehot) diag(rd_s) * (*inv_gamma_dev)  (fp16 workspace), plus the per-prefix
 * scalars alpha[s][i] = sum_j wrel_s (softmax_s - onehot) S_s, beta[s][j] (column sums).  The caller finishes
 * dq = dq_raw - q * sum_{s>=t} coef' rq_s^2 alpha_s (see loss.py::_matryoshka_backward).  4*n*m*K FLOPs for the contractions. */
CX_API size_t cx_infonce_mat_bwd_workspace_bytes(int n, int m, int k_max, int n_dims);
CX_API int cx_infonce_mat_bwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int n_dims, const int32_t* dims,
                       const float* wrel, float scale, const float* scale_dev, const float* rq, const float* rd, int label_offset,
                       int label_stride, const float* lse, float coef, const float* coef_gamma_dev, const float* inv_gamma_dev,
                       float* dq_raw, int64_t lddq, float* dd_raw, int64_t lddd, float* alpha, float* beta, void* workspace,
                       cx_stream_t stream);
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

/* ======================================================================================================== encoder ops
 * Memory-bound kernels of the encoder fwd/bwd.  Activations are bf16 [rows, d] row-major (rows = packed non-pad tokens),
 * LayerNorm parameters / statistics / parameter gradients are fp32.  d % 8 == 0, d <= 1024. */

/* ---- fused (dropout-)add-LayerNorm (replaces flash-attn dropout_add_layer_norm: layers/block.py:422-431,453-462,
 *      models/encoder/modeling_nomic_bert.py:531-535).  y = LN(a + b) * gamma + beta; b may be NULL;
 *      stats[rows][2] = (mean, rstd) for the backward; z_out (bf16, may be NULL) receives z = a + b, the residual
 *      stream of the pre-norm blocks (layers/block.py:293-388, ViT).  p_drop > 0: z = dropout(a)/(1-p) + b with a
 *      counter-based keep mask that is a pure function of (seed, row, column) -- the backward regenerates it. */
CX_API int cx_add_layernorm_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* y, float* stats,
                         int rows, int d, float eps, void* z_out, float p_drop, unsigned long long seed, cx_stream_t stream);
/* backward: z = a + b is recomputed, upstream gradient g = g1 + g2 (g2 may be NULL); writes dz (bf16, the gradient of
 * both a and b) and ADDS the parameter gradients into dgamma/dbeta (pass both NULL to skip them).
 * workspace: cx_layernorm_bwd_workspace_bytes(d).  gres (bf16, may be NULL): gradient arriving on the residual stream
 * z itself (pre-norm blocks), added to dz.  With dropout (same p_drop / seed as the forward) dz is the gradient of b and of
 * the residual stream, da_out (bf16, may be NULL) receives the gradient of the dropped branch a = dz * keep / (1 - p). */
CX_API size_t cx_layernorm_bwd_workspace_bytes(int d);
CX_API int cx_add_layernorm_bwd(const void* a, const void* b, const void* g1, const void* g2, const float* gamma,
                         const float* stats, void* dz, float* dgamma, float* dbeta, void* workspace, int rows, int d,
                         const void* gres, float p_drop, unsigned long long seed, void* da_out, cx_stream_t stream);
/* ---- embeddings + emb_ln (layers/embedding.py:594-615 + modeling_nomic_bert.py:531-534): y = LN(word[ids] + type[type_ids]).
 *      type_ids may be NULL (all zeros); p_drop > 0 applies emb_drop AFTER the LayerNorm (:535).  Backward scatters into the fp32 table gradients (atomic adds), ADDS dgamma/dbeta. */
CX_API int cx_embed_layernorm_fwd(const int64_t* ids, const int64_t* type_ids, const void* word_emb, const void* type_emb,
                           const float* gamma, const float* beta, void* y, float* stats, int rows, int d, float eps,
                           float p_drop, unsigned long long seed, cx_stream_t stream);
CX_API int cx_embed_layernorm_bwd(const int64_t* ids, const int64_t* type_ids, const void* word_emb, const void* type_emb,
                           const void* g1, const void* g2, const float* gamma, const float* stats, float* dword,
                           float* dtype_emb, float* dgamma, float* dbeta, void* workspace, int rows, int d,
                           int64_t padding_idx, float p_drop, unsigned long long seed, cx_stream_t stream);
/* ---- unpadded-token bookkeeping (flash-attn bert_padding: modeling_nomic_bert.py:333): pos[t] = t - cu_seqlens[seq(t)] */
CX_API int cx_token_positions(const int32_t* cu_seqlens, int nseq, int32_t* pos, int32_t* seq_id, cx_stream_t stream);
/* ---- rotary embedding, NeoX halves, in place on q and k of qkv [T,3,H,Dh] (layers/embedding.py:653-745);
 *      cos/sin fp32 [max_pos, Dh/2]; backward != 0 applies the transpose rotation; slots [first_slot, first_slot +
 *      num_slots) of {0 = q, 1 = k} are rotated. */
CX_API int cx_rope_inplace(void* qkv, const int32_t* pos, const float* cos_t, const float* sin_t, int T, int H, int Dh,
                    int backward, int first_slot, int num_slots, cx_stream_t stream);
/* fp32 dQ accumulator [T,H*Dh] -> bf16 q-slot of dqkv [T,3,H,Dh] with the transpose rotation fused */
CX_API int cx_dq_finalize_rope(const float* dq_acc, void* dqkv, const int32_t* pos, const float* cos_t, const float* sin_t,
                        int T, int H, int Dh, cx_stream_t stream);
/* ---- SwiGLU (layers/mlp.py:68-75): yg [T,2I] = [fc11(x) | fc12(x)];  out = y * silu(gate) */
CX_API int cx_swiglu_fwd(const void* yg, void* out, int64_t T, int I, cx_stream_t stream);
CX_API int cx_swiglu_bwd(const void* dout, const void* yg, void* dyg, int64_t T, int I, cx_stream_t stream);
/* ---- MeanPooling over each packed sequence (modeling_biencoder.py:79-90) and its backward */
CX_API int cx_mean_pool_fwd(const void* h, const int32_t* cu_seqlens, float* pooled, int nseq, int d, cx_stream_t stream);
CX_API int cx_mean_pool_bwd(const float* dpooled, const int32_t* cu_seqlens, void* dh, int nseq, int d, cx_stream_t stream);
/* ---- BiEncoder tail (modeling_biencoder.py:307-317): optional affine-free LN ("hamming"), cast to bf16, optional
 *      F.normalize, fp32 out.  save[rows][3] = (mean, rstd, inv_norm). */
CX_API int cx_embed_head_fwd(const float* pooled, float* out, float* save, int rows, int d, int hamming, int normalize,
                      cx_stream_t stream);
CX_API int cx_embed_head_bwd(const float* pooled, const float* gout, const float* save, float* gpooled, int rows, int d,
                      int hamming, int normalize, cx_stream_t stream);
/* ---- optimizer tail (optimizer.py:7-47, trainers/base.py:372-385) on flat fp32 buffers.
 *      cx_grad_clip_coef: out2[0] = ||grad||_2, out2[1] = min(1, max_norm/(norm+1e-6)) (device scalars, no host sync).
 *      cx_adamw_step: torch.optim.AdamW update, refreshes the bf16 shadow weights, optionally zeroes the gradient. */
CX_API size_t cx_grad_clip_workspace_bytes(void);
CX_API int cx_grad_clip_coef(const float* grad, int64_t n, float max_norm, float* out2, void* workspace, cx_stream_t stream);
CX_API int cx_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n, float lr,
                  float beta1, float beta2, float eps, float weight_decay, int step, const float* grad_scale_dev,
                  float grad_scale, int zero_grad, cx_stream_t stream);
CX_API int cx_cast_f32_bf16(const float* x, void* y, int64_t n, cx_stream_t stream);

/* ---- ViT tower pieces (models/vit/vit.py:176-276, layers/embedding.py:465-516, layers/mlp.py:8-34)
 * cx_linear_bias_bf16: y = x w^T + bias (FusedDense); cx_colsum_bf16: out[N] += column sums (bias gradients);
 * cx_act_*: kind 0 = gelu(erf), 1 = quick_gelu; cx_patchify: pixels [B,C,H,W] fp32 -> patch rows [B*gh*gw, C*p*p] bf16;
 * cx_vit_assemble_*: cls token + learned position embedding; cx_cls_select_*: ClsSelector (modeling_biencoder.py:44-49). */
CX_API int cx_linear_bias_bf16(const void* x, const void* w, const float* bias, void* y, int M, int N, int K, int64_t ldx,
                        int64_t ldw, int64_t ldy, cx_stream_t stream);
CX_API int cx_colsum_bf16(const void* x, int64_t T, int N, float* out, cx_stream_t stream);
CX_API int cx_act_fwd(const void* x, void* y, int64_t n, int kind, cx_stream_t stream);
CX_API int cx_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int kind, cx_stream_t stream);
CX_API int cx_patchify(const float* pixels, void* out, int B, int C, int Himg, int Wimg, int patch, cx_stream_t stream);
CX_API int cx_vit_assemble_fwd(const void* proj, const float* cls, const float* pos, void* z, int B, int nP, int d, cx_stream_t stream);
CX_API int cx_vit_assemble_bwd(const void* dz, void* dproj, float* dcls, float* dpos, int B, int nP, int d, cx_stream_t stream);
CX_API int cx_cls_select_fwd(const void* h, float* out, int B, int S, int d, cx_stream_t stream);
CX_API int cx_cls_select_bwd(const float* g, void* dh, int B, int S, int d, cx_stream_t stream);

/* ---- attention pooling: ONE learned query per head against every sequence's keys / values (FlashAttentionPooling inside
 * MultiHeadAttentionPooling, /root/reference/src/contrastors/layers/attention.py:313-440, models/biencoder/
 * modeling_biencoder.py:93-152).  q [H, 64] fp32 (Wq(latent)); kv [T, 2, H, 64] bf16 over packed tokens; out [nseq, H, 64] fp32;
 * lse [nseq, H].  Backward: dkv written (bf16), dq ACCUMULATED (the caller zeroes it). */
CX_API int cx_attn_pool_fwd(const float* q, const void* kv, const int32_t* cu_seqlens, float* out, float* lse, int nseq, int max_seqlen,
                     int H, int Dh, float softmax_scale, cx_stream_t stream);
CX_API int cx_attn_pool_bwd(const float* q, const void* kv, const int32_t* cu_seqlens, const float* dout, const float* lse, float* dq,
                     void* dkv, int nseq, int max_seqlen, int H, int Dh, float softmax_scale, cx_stream_t stream);

/* ---- varlen non-causal attention on tcgen05 (replaces flash_attn_varlen_qkvpacked_func: layers/attention.py:158-181)
 * qkv [T,3,H,Dh] bf16 (RoPE already applied), cu_seqlens int32[nseq+1]; out [T,H,Dh] bf16; lse [H,T] fp32 (natural log).
 * Dh == 64.  Backward: dqkv [T,3,H,Dh] bf16 (dk, dv written directly; dq through the fp32 accumulator dq_acc [T,H*Dh],
 * which cx_attn_bwd zeroes itself and the caller finalises with cx_dq_finalize_rope or cx_dq_finalize). delta [H,T] fp32 scratch.
 * dk_rope_inv_freq (32 fp32 values, or NULL): when given, dk leaves the kernel already rotated back (the transpose of the
 * rotary embedding applied to k before the forward; position = the key's index inside its sequence), so the caller runs no
 * rotary pass over dqkv. */
CX_API int cx_attn_fwd(const void* qkv, const int32_t* cu_seqlens, void* out, float* lse, int total_tokens, int nseq,
                int max_seqlen, int H, int Dh, float softmax_scale, cx_stream_t stream);
CX_API int cx_dq_finalize(const float* dq_acc, void* dqkv, int T, int H, int Dh, cx_stream_t stream);
CX_API int cx_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const int32_t* cu_seqlens,
                void* dqkv, float* dq_acc, float* delta, int total_tokens, int nseq, int max_seqlen, int H, int Dh,
                float softmax_scale, const float* dk_rope_inv_freq, cx_stream_t stream);

#ifdef __cplusplus
}
#endif
#endif /* CONTRASTORS_B200_H */
