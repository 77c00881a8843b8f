// Varlen non-causal multi-head attention forward/backward on tcgen05 / TMEM / TMA (Dh = 64), C ABI cx_attn_{fwd,bwd}.
//
// Replaces flash_attn_varlen_qkvpacked_func / flash_attn_qkvpacked_func (FA2, mma.sync) at
// /root/reference/src/contrastors/layers/attention.py:158-181,220-226.  Layouts follow the reference's packed format:
// qkv [T, 3, H, Dh] bf16 over unpadded tokens, cu_seqlens int32 [nseq+1].
//
// Forward (attn_fwd4_kernel, cx_attn_fwd.cuh): one CTA = (sequence, head, 128 query rows), two CTAs per SM, two softmax threads
//   per query row; S in TMEM, P (bf16) in its own TMEM columns feeding O += P V straight from tensor memory.
// Backward (attn_bwd4_kernel, cx_attn_bwd.cuh): one CTA = (sequence, head, 128 keys), loops over query tiles with TRANSPOSED
//   scores so that P^T / dS^T feed dV / dK from tensor memory; the per-query statistics enter through a fifth MMA k-step; dQ
//   partials leave through TMA reduce-add into an fp32 accumulator.
//   With rope_inv_freq the transposed rotary embedding is applied to dK in the epilogue (the key's position is its row index
//   inside the sequence), so no rotary pass over dqkv remains in the backward.
// Round 1 kept five forward and three backward generations selectable by environment variable; they were timed on hardware in
// round 2 (profiles/r02a_bench_attn_fwd4_bwd3_vs_flash_attn2.json) and the losers deleted; the backward was then re-scheduled
// from a phase trace (profiles/r02l_*, r02n_*: 325 -> 312 us), its predecessor deleted as well.
// What binds (tools/ubench, profiles/r01_ubench_tmem_mma.txt): forward, nearest hard bound the SFU (16384 exponentials per
// 128 x 128 tile at 16 / clk / SM = 1024 clk against 512 clk of tensor time); backward, the shared-memory port.
// The thread that ISSUES the MMAs must stay tight: the TMA / MMA warps run converged with elect.sync around the
// asynchronous instructions only (under `if (lane == 0)` every UTCHMMA sits in an elect-and-branch loop, ~80 clk each).
#include <math.h>
#include <stdlib.h>

#include "cx_host.h"
#include "cx_ptx.cuh"

namespace cx {

constexpr int kDh = 64;

constexpr int kFwdPolyDefault = 4;  // every other column quad: 99.1 vs 101.2 us at 64 x 512 x 12, 129.1 vs 131.1 at 256 x 197 x 12 (all on the SFU: 0; half: 109.3)
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

#include "cx_attn_fwd.cuh"
#include "cx_attn_bwd.cuh"

}  // namespace cx

using namespace cx;

extern "C" int cx_attn_fwd(const void* qkv, const int32_t* cu_seqlens, void* out, float* lse, int total_tokens, int nseq,
                           int max_seqlen, int H, int Dh, float softmax_scale, cx_stream_t stream_) {
  CX_REQUIRE(qkv && cu_seqlens && out && lse, "cx_attn_fwd: null pointer");
  CX_REQUIRE(Dh == kDh, "cx_attn_fwd: only head_dim 64 is implemented");
  CX_REQUIRE(total_tokens > 0 && nseq > 0 && max_seqlen > 0 && H > 0, "cx_attn_fwd: empty problem");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  CUtensorMap tm;
  int rc = make_tmap_2d(&tm, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H * Dh, (uint64_t)total_tokens,
                        (uint64_t)3 * H * Dh * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  // share of the exponentials evaluated on the FMA pipe instead of the SFU (read once; A/B knob of the round-2 measurement)
  static const int poly = [] {
    const char* e = getenv("CX_ATTN_POLY");
    const int v = (e && *e) ? atoi(e) : kFwdPolyDefault;
    return (v == 0 || v == 4 || v == 8) ? v : kFwdPolyDefault;
  }();
#define CX_FWD4(P)                                                                                                      \
  do {                                                                                                                  \
    CX_SET_SMEM_ONCE(attn_fwd4_kernel<P>, Fwd4Smem::kTotal);                                                            \
    attn_fwd4_kernel<P><<<grid, kFwd4Threads, Fwd4Smem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse,     \
                                                                         total_tokens, H, softmax_scale * kLog2e);    \
  } while (0)
  if (poly == 8) CX_FWD4(8);
  else if (poly == 4) CX_FWD4(4);
  else CX_FWD4(0);
#undef CX_FWD4
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const int32_t* cu_seqlens,
                           void* dqkv, float* dq_acc, float* delta, int total_tokens, int nseq, int max_seqlen, int H, int Dh,
                           float softmax_scale, const float* dk_rope_inv_freq, cx_stream_t stream_) {
  CX_REQUIRE(qkv && out && dout && lse && cu_seqlens && dqkv && dq_acc && delta, "cx_attn_bwd: null pointer");
  CX_REQUIRE(Dh == kDh, "cx_attn_bwd: only head_dim 64 is implemented");
  CX_REQUIRE(total_tokens > 0 && nseq > 0 && max_seqlen > 0 && H > 0, "cx_attn_bwd: empty problem");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int T = total_tokens;
  {
    const int64_t threads = (int64_t)T * H * 8;
    attn_delta_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, dq_acc, T, H);
    CX_LAUNCH_CHECK();
  }
  CUtensorMap tmQKV, tmDO, tmDQ;
  int rc = make_tmap_2d(&tmQKV, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, qkv, (uint64_t)3 * H * Dh, (uint64_t)T,
                        (uint64_t)3 * H * Dh * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmDO, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, dout, (uint64_t)H * Dh, (uint64_t)T, (uint64_t)H * Dh * 2, 64,
                    128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmDQ, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, dq_acc, (uint64_t)H * Dh, (uint64_t)T, (uint64_t)H * Dh * 4, 32,
                    128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  CX_SET_SMEM_ONCE(attn_bwd4_kernel, Bwd4Smem::kTotal);
  attn_bwd4_kernel<<<grid, kBwd4Threads, Bwd4Smem::kTotal, stream>>>(tmQKV, tmDO, tmDQ, cu_seqlens, lse, delta,
                                                                    (__nv_bfloat16*)dqkv, T, H, softmax_scale, dk_rope_inv_freq);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_dq_finalize(const float* dq_acc, void* dqkv, int T, int H, int Dh, cx_stream_t stream) {
  CX_REQUIRE(dq_acc && dqkv, "cx_dq_finalize: null pointer");
  CX_REQUIRE((H * Dh) % 8 == 0, "cx_dq_finalize: H*Dh must be a multiple of 8");
  if (T <= 0) return 0;
  const int64_t n = (int64_t)T * (H * Dh / 8);
  dq_finalize_kernel<<<(unsigned)((n + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream)>>>(dq_acc, (__nv_bfloat16*)dqkv, T, H * Dh);
  CX_LAUNCH_CHECK();
  return 0;
}

#ifdef CX_ATTN_TRACE
// trace build only (tools/trace_attn_bwd.py): point the backward kernel's phase trace at a device buffer (nullptr: off)
extern "C" __attribute__((visibility("default"))) int cx_attn_trace_set(long long* buf) {
  return cudaMemcpyToSymbol(cx::g_bwd_trace, &buf, sizeof(buf)) == cudaSuccess ? 0 : 1;
}
#endif
