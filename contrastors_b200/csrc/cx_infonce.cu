// Fused InfoNCE forward / backward on the tcgen05 GEMM core (C ABI: cx_infonce_*).
//
// Replaces the reference's unfused chain  matmul -> LogitScale -> F.cross_entropy -> argmax  and its autograd
// backward (/root/reference/src/contrastors/loss.py:105-130, modeling_biencoder.py:37-38).  The [n x m] logits never
// exist in HBM in fp32: the forward keeps them in TMEM and reduces them to per-row statistics in the GEMM epilogue;
// the backward recomputes them, emits dS = softmax - onehot (in [-1, 1], UNSCALED) as fp16 into a workspace (n*m*2 bytes,
// L2-resident at the 8-GPU per-rank shape), and contracts it twice (dQ = dS D^, dD = dS^T Q^) with the same GEMM core
// (MN-major operands, so no transposes are materialised) against fp16 copies of the operands that carry the per-row inverse
// norms of the normalised-prefix losses (Matryoshka / CLIP); coef and the logit scale ride in the contractions' alpha.
// One dS and one code path serve every loss variant at fp16's 11-bit mantissa (1e-3 parity needs more than bf16's 8).
#include <math.h>

#include "cx_gemm.cuh"

namespace cx {

static inline size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

struct NceWorkspace {
  float* part_max;
  float* part_sum;
  int* part_arg;
  float* dlogit_part;
  float* block_part;     // [2 * kMaxCombineBlocks]
  unsigned int* ticket;  // [1]
  __half* ds;
  int64_t ld_ds;
  __half* qh;  // fp16 copies of q * rq / d * rd for the backward contractions (exact for 2^-14 <= |x| <= 65504; saturating)
  __half* dh;
  int64_t ld_h;
  size_t bytes;
};
constexpr int kMaxCombineBlocks = 8192;
constexpr int kMaxGrid = 1024;

static NceWorkspace carve(void* base, int n, int m, int k) {
  NceWorkspace w{};
  const size_t ct = 2 * ((size_t)(m + 127) / 128);  // one partial per 64/128-column half tile
  size_t off = 0;
  auto take = [&](size_t bytes) {
    size_t o = off;
    off = align_up(off + bytes, 256);
    return reinterpret_cast<uint8_t*>(base) + o;
  };
  w.part_max = reinterpret_cast<float*>(take(ct * n * 4));
  w.part_sum = reinterpret_cast<float*>(take(ct * n * 4));
  w.part_arg = reinterpret_cast<int*>(take(ct * n * 4));
  w.dlogit_part = reinterpret_cast<float*>(take((size_t)kMaxGrid * 4));
  w.block_part = reinterpret_cast<float*>(take((size_t)2 * kMaxCombineBlocks * 4));
  w.ticket = reinterpret_cast<unsigned int*>(take(256));
  w.ld_ds = (int64_t)align_up((size_t)m, 8);
  w.ds = reinterpret_cast<__half*>(take((size_t)n * w.ld_ds * 2));
  w.ld_h = (int64_t)align_up((size_t)k, 8);
  w.qh = reinterpret_cast<__half*>(take((size_t)n * w.ld_h * 2));
  w.dh = reinterpret_cast<__half*>(take((size_t)m * w.ld_h * 2));
  w.bytes = off;
  return w;
}

// One warp per row: merge the per-column-tile partials (log2 domain).  Ties between tiles resolve to the lowest column
// index, so the result is exactly ATen's first-maximum argmax.  The loss sum / hit count are reduced in a fixed order
// (block partials summed by the last block to arrive), so results are run-to-run deterministic.
__global__ void nce_combine_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                   const int* __restrict__ part_arg, const float* __restrict__ label_logit, int n,
                                   int n_col_tiles, int label_offset, int label_stride, float* __restrict__ lse,
                                   int* __restrict__ argmax, float* __restrict__ stats, float* __restrict__ block_part,
                                   unsigned int* __restrict__ ticket) {
  constexpr float kLn2 = 0.6931471805599453f;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  float loss_term = 0.f, hit = 0.f;
  if (row < n) {
    float gmax = -INFINITY;
    int garg = 0x7fffffff;
    for (int t = lane; t < n_col_tiles; t += 32) {
      const float mx = part_max[(size_t)t * n + row];
      const int ag = part_arg[(size_t)t * n + row];
      if (mx > gmax || (mx == gmax && ag < garg)) {
        gmax = mx;
        garg = ag;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, gmax, o);
      const int oa = __shfl_xor_sync(0xffffffffu, garg, o);
      if (om > gmax || (om == gmax && oa < garg)) {
        gmax = om;
        garg = oa;
      }
    }
    float sum = 0.f;
    for (int t = lane; t < n_col_tiles; t += 32) {
      const float ps = part_sum[(size_t)t * n + row];
      if (ps > 0.f) sum += ps * exp2f(part_max[(size_t)t * n + row] - gmax);  // masked-out halves carry (-inf, 0)
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float l = (gmax + log2f(sum)) * kLn2;
    if (lane == 0) {
      lse[row] = l;
      argmax[row] = garg;
      loss_term = l - label_logit[row];
      hit = (garg == (row + label_offset) * label_stride) ? 1.f : 0.f;
    }
  }
  __shared__ float s_loss[32], s_hit[32];
  __shared__ bool is_last;
  if (lane == 0) {
    s_loss[warp] = loss_term;
    s_hit[warp] = hit;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += s_loss[w];
      b += s_hit[w];
    }
    block_part[2 * blockIdx.x] = a;
    block_part[2 * blockIdx.x + 1] = b;
    __threadfence();
    const unsigned int t = atomicAdd(ticket, 1u);
    is_last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (is_last && warp == 0) {
    __threadfence();
    float a = 0.f, b = 0.f;
    for (int i = lane; i < (int)gridDim.x; i += 32) {
      a += reinterpret_cast<volatile float*>(block_part)[2 * i];
      b += reinterpret_cast<volatile float*>(block_part)[2 * i + 1];
    }
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
    }
    if (lane == 0) {
      stats[0] = a;
      stats[1] = b;
      *ticket = 0;  // ready for the next call on this workspace
    }
  }
}

// sums the per-CTA partials (fixed order) into stats[2]
__global__ void nce_dlogit_kernel(const float* __restrict__ part, int count, float* __restrict__ stats) {
  const int lane = threadIdx.x;
  float a = 0.f;
  for (int i = lane; i < count; i += 32) a += part[i];
  for (int o = 16; o > 0; o >>= 1) a += __shfl_xor_sync(0xffffffffu, a, o);
  if (lane == 0) stats[2] = a;
}

// ---------------------------------------------------------------- row utilities
// One warp per row.
__global__ void rows_to_bf16_kernel(const float* __restrict__ x, int64_t ldx, __nv_bfloat16* __restrict__ y, int64_t ldy,
                                    float* __restrict__ inv_norm, int rows, int k, int normalize) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* xr = x + (size_t)row * ldx;
  float inv = 1.f;
  if (normalize) {
    float ss = 0.f;
    for (int j = lane; j < k; j += 32) ss = fmaf(xr[j], xr[j], ss);
    for (int o = 16; o > 0; o >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, o);
    inv = 1.f / fmaxf(sqrtf(ss), 1e-12f);  // F.normalize: x / max(||x||, eps)
  }
  if (inv_norm != nullptr && lane == 0) inv_norm[row] = inv;
  if (y != nullptr) {
    __nv_bfloat16* yr = y + (size_t)row * ldy;
    for (int j = lane; j < k; j += 32) yr[j] = __float2bfloat16_rn(xr[j] * inv);
  }
}

// y (fp16) = sat(x (bf16) * inv_norm[row]), 8 elements per thread (rows are 16-byte aligned with k % 8 == 0 on the fast
// path).  Saturating: an embedding entry beyond +-65504 (a logit beyond 1e6 at any useful scale) stays finite.
__global__ void bf16_to_f16_rows_kernel(const __nv_bfloat16* __restrict__ x, int64_t ldx, __half* __restrict__ y, int64_t ldy,
                                        const float* __restrict__ inv_norm, int rows, int k, int vec_ok) {
  const int per_row = (k + 7) / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)rows * per_row) return;
  const int r = (int)(i / per_row), c = (int)(i % per_row) * 8;
  const float inv = inv_norm != nullptr ? inv_norm[r] : 1.f;
  const __nv_bfloat16* xr = x + (size_t)r * ldx + c;
  __half* yr = y + (size_t)r * ldy + c;
  auto sat = [](float v) { return fminf(fmaxf(v, -65504.f), 65504.f); };
  if (vec_ok && c + 8 <= k) {
    const uint4 in = *reinterpret_cast<const uint4*>(xr);
    const __nv_bfloat162* b2 = reinterpret_cast<const __nv_bfloat162*>(&in);
    uint4 out;
    __half2* h2 = reinterpret_cast<__half2*>(&out);
#pragma unroll
    for (int j = 0; j < 4; ++j) h2[j] = __floats2half2_rn(sat(__low2float(b2[j]) * inv), sat(__high2float(b2[j]) * inv));
    *reinterpret_cast<uint4*>(yr) = out;
  } else {
    for (int j = 0; j < 8 && c + j < k; ++j) yr[j] = __float2half_rn(sat(__bfloat162float(xr[j]) * inv));
    for (int j = 0; j < 8 && c + j >= k && c + j < (int)ldy; ++j) yr[j] = __float2half_rn(0.f);  // zero the row padding
  }
}

// gx = g' - y (g'.y),  y = x*inv, g' = g * (g_prescaled ? 1 : inv)   (backward of F.normalize)
__global__ void l2norm_bwd_kernel(const float* __restrict__ x, int64_t ldx, const float* __restrict__ g, int64_t ldg,
                                  const float* __restrict__ inv_norm, float* __restrict__ gx, int64_t ldgx, int rows,
                                  int k, int g_prescaled, int accumulate) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float inv = inv_norm[row];
  const float gs = g_prescaled ? 1.f : inv;
  const float* xr = x + (size_t)row * ldx;
  const float* gr = g + (size_t)row * ldg;
  float dot = 0.f;
  for (int j = lane; j < k; j += 32) dot = fmaf(gr[j] * gs, xr[j] * inv, dot);
  for (int o = 16; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
  float* o_ = gx + (size_t)row * ldgx;
  for (int j = lane; j < k; j += 32) {
    const float v = gr[j] * gs - xr[j] * inv * dot;
    o_[j] = accumulate ? o_[j] + v : v;
  }
}

// host helper shared with cx_infonce_mat.cu: y (fp16) = sat(x (bf16) * inv_norm[row]) (inv_norm may be null)
int nce_rows_to_f16(const void* x, int64_t ldx, void* y, int64_t ldy, const float* inv_norm, int rows, int k, cudaStream_t stream) {
  const int threads = 256;
  const int64_t per_row = (k + 7) / 8;
  const int vec = (ldx % 8 == 0 && (reinterpret_cast<uintptr_t>(x) & 15) == 0) ? 1 : 0;
  bf16_to_f16_rows_kernel<<<(unsigned)(((int64_t)rows * per_row + threads - 1) / threads), threads, 0, stream>>>(
      reinterpret_cast<const __nv_bfloat16*>(x), ldx, reinterpret_cast<__half*>(y), ldy, inv_norm, rows, k, vec);
  CX_LAUNCH_CHECK();
  return 0;
}

}  // namespace cx

using namespace cx;

extern "C" size_t cx_infonce_workspace_bytes(int n, int m, int k_dim) {
  if (n <= 0 || m <= 0 || k_dim <= 0) return 0;
  return carve(nullptr, n, m, k_dim).bytes + 256;
}

static void* align256(void* p) { return reinterpret_cast<void*>((reinterpret_cast<uintptr_t>(p) + 255) & ~uintptr_t(255)); }

extern "C" int cx_infonce_fwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int k_dim,
                              float scale, const float* scale_dev, const float* rq, const float* rd, int label_offset, int label_stride,
                              float* lse, int32_t* argmax, float* label_logit, float* stats, void* workspace,
                              cx_stream_t stream_) {
  CX_REQUIRE(q && d && lse && argmax && label_logit && stats && workspace, "cx_infonce_fwd: null pointer");
  CX_REQUIRE(n > 0 && m > 0 && k_dim > 0, "cx_infonce_fwd: empty problem");
  CX_REQUIRE((long long)(n - 1 + label_offset) * label_stride < m && label_offset >= 0 && label_stride >= 1,
             "cx_infonce_fwd: labels out of range");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NceWorkspace w = carve(align256(workspace), n, m, k_dim);
  GemmArgs g{};
  g.A = q; g.B = d; g.C = nullptr;
  g.M = n; g.N = m; g.K = k_dim;
  g.a_mn = false; g.b_mn = false;
  g.lda = ldq; g.ldb = ldd; g.ldc = 0;
  g.out_f32 = false; g.accumulate = false; g.splits = 1;
  g.mode = EPI_NCE_STATS;
  g.ep.scale = scale; g.ep.scale_dev = scale_dev; g.ep.rq = rq; g.ep.rd = rd;
  g.ep.label_offset = label_offset; g.ep.label_stride = label_stride;
  g.ep.part_max = w.part_max; g.ep.part_sum = w.part_sum; g.ep.part_arg = w.part_arg;
  g.ep.label_logit = label_logit;
  g.stream = stream;
  int rc = launch_gemm(g);
  if (rc) return rc;
  const int bn = gemm_block_n(m);
  const int n_col_tiles = 2 * ((m + bn - 1) / bn);  // the epilogue emits one partial per column half
  const int threads = 256;  // 8 rows per block, one warp each
  const int blocks = (n + 7) / 8;
  CX_REQUIRE(blocks <= kMaxCombineBlocks, "cx_infonce_fwd: n too large");
  CX_CUDA_CHECK(cudaMemsetAsync(w.ticket, 0, 4, stream));
  nce_combine_kernel<<<blocks, threads, 0, stream>>>(w.part_max, w.part_sum, w.part_arg, label_logit, n, n_col_tiles,
                                                     label_offset, label_stride, lse, argmax, stats, w.block_part, w.ticket);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_infonce_bwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int k_dim,
                              float scale, const float* scale_dev, const float* rq, const float* rd, int label_offset, int label_stride,
                              const float* lse, float coef, const float* coef_dev, float* dq, int64_t lddq, float* dd, int64_t lddd,
                              int accumulate_dd, float* stats, void* workspace, cx_stream_t stream_) {
  CX_REQUIRE(q && d && lse && dq && dd && stats && workspace, "cx_infonce_bwd: null pointer");
  CX_REQUIRE(n > 0 && m > 0 && k_dim > 0, "cx_infonce_bwd: empty problem");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  NceWorkspace w = carve(align256(workspace), n, m, k_dim);
  // stage 1: dS (fp16) = softmax - onehot, unscaled; the logit-scale gradient partials carry coef
  GemmArgs g{};
  g.A = q; g.B = d; g.C = w.ds;
  g.M = n; g.N = m; g.K = k_dim;
  g.a_mn = false; g.b_mn = false;
  g.lda = ldq; g.ldb = ldd; g.ldc = w.ld_ds;
  g.out_f32 = false; g.accumulate = false; g.splits = 1;
  g.mode = EPI_NCE_DS;
  g.ep.scale = scale; g.ep.scale_dev = scale_dev; g.ep.rq = rq; g.ep.rd = rd;
  g.ep.label_offset = label_offset; g.ep.label_stride = label_stride;
  g.ep.lse = lse; g.ep.coef = coef; g.ep.coef_dev = coef_dev;
  g.ep.dlogit_part = w.dlogit_part;
  g.stream = stream;
  // every CTA of the stage-1 launch writes one partial; the array is zeroed first so the (launch-mode dependent) CTA
  // count never matters
  CX_CUDA_CHECK(cudaMemsetAsync(w.dlogit_part, 0, (size_t)kMaxGrid * sizeof(float), stream));
  int rc = launch_gemm(g);
  if (rc) return rc;
  nce_dlogit_kernel<<<1, 32, 0, stream>>>(w.dlogit_part, kMaxGrid, stats);
  CX_LAUNCH_CHECK();
  // fp16 B operands of the two contractions: q^ = q * rq, d^ = d * rd (rq / rd null = 1)
  {
    const int threads = 256;
    const int64_t per_row = (k_dim + 7) / 8;
    const int vq = (ldq % 8 == 0 && (reinterpret_cast<uintptr_t>(q) & 15) == 0) ? 1 : 0;
    const int vd = (ldd % 8 == 0 && (reinterpret_cast<uintptr_t>(d) & 15) == 0) ? 1 : 0;
    bf16_to_f16_rows_kernel<<<(unsigned)(((int64_t)n * per_row + threads - 1) / threads), threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(q), ldq, w.qh, w.ld_h, rq, n, k_dim, vq);
    CX_LAUNCH_CHECK();
    bf16_to_f16_rows_kernel<<<(unsigned)(((int64_t)m * per_row + threads - 1) / threads), threads, 0, stream>>>(
        reinterpret_cast<const __nv_bfloat16*>(d), ldd, w.dh, w.ld_h, rd, m, k_dim, vd);
    CX_LAUNCH_CHECK();
  }
  // stage 2a: dQ^[n,k] = scale * coef * dS[n,m] (K-major A) x D^[m,k] (MN-major B), split-K over m
  GemmArgs a{};
  a.A = w.ds; a.B = w.dh; a.C = dq;
  a.M = n; a.N = k_dim; a.K = m;
  a.a_mn = false; a.b_mn = true;
  a.lda = w.ld_ds; a.ldb = w.ld_h; a.ldc = lddq;
  a.out_f32 = true; a.accumulate = false; a.splits = 0;
  a.mode = EPI_STORE; a.ep.alpha = scale * coef; a.ep.alpha_dev = scale_dev; a.ep.alpha_dev2 = coef_dev;
  a.ep.ab_f16 = 1;
  a.stream = stream;
  rc = launch_gemm(a);
  if (rc) return rc;
  // stage 2b: dD^[m,k] = scale * coef * dS^T (MN-major A: stored [n,m]) x Q^[n,k] (MN-major B)
  GemmArgs b{};
  b.A = w.ds; b.B = w.qh; b.C = dd;
  b.M = m; b.N = k_dim; b.K = n;
  b.a_mn = true; b.b_mn = true;
  b.lda = w.ld_ds; b.ldb = w.ld_h; b.ldc = lddd;
  b.out_f32 = true; b.accumulate = accumulate_dd != 0; b.splits = 0;
  b.mode = EPI_STORE; b.ep.alpha = scale * coef; b.ep.alpha_dev = scale_dev; b.ep.alpha_dev2 = coef_dev;
  b.ep.ab_f16 = 1;
  b.stream = stream;
  return launch_gemm(b);
}

extern "C" int cx_rows_to_bf16(const float* x, int64_t ldx, void* y_bf16, int64_t ldy, float* inv_norm, int rows, int k,
                               int normalize, cx_stream_t stream) {
  CX_REQUIRE(x && (y_bf16 || inv_norm), "cx_rows_to_bf16: null pointer");
  if (rows <= 0 || k <= 0) return 0;
  const int warps = 8;
  rows_to_bf16_kernel<<<(rows + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, reinterpret_cast<__nv_bfloat16*>(y_bf16), ldy, inv_norm, rows, k, normalize);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_l2norm_bwd(const float* x, int64_t ldx, const float* g, int64_t ldg, const float* inv_norm, float* gx,
                             int64_t ldgx, int rows, int k, int g_prescaled, int accumulate, cx_stream_t stream) {
  CX_REQUIRE(x && g && inv_norm && gx, "cx_l2norm_bwd: null pointer");
  if (rows <= 0 || k <= 0) return 0;
  const int warps = 8;
  l2norm_bwd_kernel<<<(rows + warps - 1) / warps, warps * 32, 0, static_cast<cudaStream_t>(stream)>>>(
      x, ldx, g, ldg, inv_norm, gx, ldgx, rows, k, g_prescaled, accumulate);
  CX_LAUNCH_CHECK();
  return 0;
}
