// Varlen non-causal multi-head attention forward/backward on tcgen05 / TMEM / TMA (Dh = 64), C ABI cx_attn_{fwd,bwd}.
//
// Replaces flash_attn_varlen_qkvpacked_func / flash_attn_qkvpacked_func (FA2, mma.sync) at
// /root/reference/src/contrastors/layers/attention.py:158-181,220-226.  Layouts follow the reference's packed format:
// qkv [T, 3, H, Dh] bf16 over unpadded tokens, cu_seqlens int32 [nseq+1].
//
// Forward (attn_fwd3_kernel, default): one CTA = (sequence, head, 128 query rows), two CTAs per SM.
//   warp 0: TMA producer (Q once, K/V tiles double-buffered)     warp 1: MMA issuer (converged warp, one elected lane)
//   warp 2: TMEM allocator                                        warps 4-7: softmax, one thread per query row
//   per 128-key tile j:  S = Q K_j^T (one N = 128 chain, TMEM fp32)  ->  online softmax in two 64-column halves (exp2,
//   lazily raised maximum, packed fp32x2 arithmetic)  ->  P (bf16, its own TMEM columns)  ->  O += P V_j (A from TMEM).
//   S(j+1) is issued as soon as every thread has loaded its S(j) row, so it runs under the second half's exponentials.
// Backward (attn_bwd2_kernel, default): one CTA = (sequence, head, 128 keys); loops over query tiles; S and dP are
//   recomputed into TMEM, P / dS go to smem once and feed three contractions (dV += P^T dO, dK += dS^T Q, dQ += dS K);
//   dK/dV accumulate in TMEM, dQ partials leave through TMA reduce-add into an fp32 accumulator (own 4-warp group);
//   S(i+1) / dP(i+1) are issued as soon as P(i) / dS(i) have left the registers.
// Older generations (attn_fwd_kernel, attn_fwd2_kernel, attn_bwd_kernel) stay selectable for A/B timing (see below).
// What binds (tools/ubench, tools/trace_attn.py, profiles/r01_ubench_tmem_mma.txt): forward, nearest hard bound the SFU
// (16384 exponentials per 128 x 128 tile at 16 / clk / SM = 1024 clk against 512 clk of tensor time: S as N = 128 is 64 clk
// per MMA, PV from TMEM 32 clk), in practice the latency chain of the one softmax warp per SM sub-partition per CTA (the
// loop runs at 1200-1450 clk per tile) -- plus ~30 % of each CTA's lifetime in prologue (TMA round trip) and epilogue.  Backward, the shared-
// memory port: 240 KB of MMA operands + 128 KB of P / dS / dQ-staging traffic per tile at 128 B/clk = 2900 clk against
// 1664 clk of tensor time (an SS MMA with N = 64 takes 48 clk, not 32: its 6 KB of operands come through that port).
// The thread that ISSUES the MMAs must stay tight: the TMA / MMA warps run converged with elect.sync around the
// asynchronous instructions only (under `if (lane == 0)` every UTCHMMA sits in an elect-and-branch loop, ~80 clk each).
#include <math.h>
#include <stdlib.h>

#include <mutex>

#include "cx_host.h"
#include "cx_ptx.cuh"

namespace cx {

constexpr int kDh = 64;

// Profiling hook (cx_debug_attn_trace): when set, the pipelined kernels record clock64() stamps of their pipeline events,
// 64 slots per CTA, so a session can see where a CTA's lifetime goes.  Null in production (one predictable branch).
__device__ long long* g_attn_trace = nullptr;

__device__ __forceinline__ void trace_put(long long* tr, int slot) {
  if (tr != nullptr) tr[slot] = clock64();
}
constexpr float kLog2e = 1.4426950408889634f;
constexpr float kLn2 = 0.6931471805599453f;

#include "cx_attn_fwd.cuh"
#include "cx_attn_bwd.cuh"

}  // namespace cx

using namespace cx;

// Kernel generation selectable at run time for A/B timing and profiling sessions.  CX_ATTN_FWD = 1: serial 128-key tiles,
// 3: pipelined 64-key sub-tiles (P in tensor memory; carries the trace / ablation hooks), 6: wide-S = default, 7: wide-S
// with 3/8 of the exponentials on the FMA pipe, 8 / 9: EXPERIMENTAL two-threads-per-row wide-S kernel, non-persistent / persistent (not yet run on hardware);
// CX_ATTN_BWD = 1: serial, 2: pipelined = default, 3: EXPERIMENTAL transposed-score kernel (dV / dK fed from TMEM; not yet run
// on hardware).  Measured on B200,
// 64 x 512 tokens x 12 heads, L2 flushed: forward 138 / 127 / 110.6 / 113.6 us; backward (incl. delta, zero fill, dQ
// finalize) 415 / 376 us.  (Also tried and dropped: P through shared memory in the sub-tile kernel 131 us; 3/8 and 4/8
// polynomial exponentials there 145 / 150 us; holding back the second CTA of each SM to de-phase the pair: no gain.)
constexpr int kFwdDefaultMode = 8;  // two threads per row (round 2: 103 us vs 112 at 64 x 512 x 12, 131 vs 139 at 256 x 197 x 12)
constexpr int kBwdDefaultMode = 3;  // transposed scores (round 2: 351 us vs 376, 457 vs 488)
constexpr uint32_t kPoly38 = 0x52;  // column-pair pattern (period 8) routed to the polynomial
static int parse_mode(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  const int v = atoi(e);
  return (v == 1 || v == 2 || v == 3 || (v >= 6 && v <= 9)) ? v : dflt;
}
// the environment is read once per process (thread-safe static initialisation), not on every call; A/B sessions and the
// parity tests of the non-default generations switch at run time through cx_attn_select_kernels
static std::atomic<int> g_fwd_override{0}, g_bwd_override{0};
static int fwd_mode() {
  static const int m = parse_mode("CX_ATTN_FWD", kFwdDefaultMode);
  const int o = g_fwd_override.load(std::memory_order_relaxed);
  return o ? o : m;
}
static int bwd_mode() {
  static const int m = parse_mode("CX_ATTN_BWD", kBwdDefaultMode);
  const int o = g_bwd_override.load(std::memory_order_relaxed);
  return o ? o : m;
}

extern "C" int cx_attn_select_kernels(int fwd_generation, int bwd_generation) {
  auto ok = [](int v) { return v == 0 || v == 1 || v == 2 || v == 3 || (v >= 6 && v <= 9); };
  CX_REQUIRE(ok(fwd_generation) && ok(bwd_generation), "cx_attn_select_kernels: unknown kernel generation (0 = default)");
  g_fwd_override.store(fwd_generation, std::memory_order_relaxed);
  g_bwd_override.store(bwd_generation, std::memory_order_relaxed);
  return 0;
}

extern "C" int cx_debug_attn_trace(void* buf) {
#ifdef CX_DEBUG_HOOKS
  CX_CUDA_CHECK(cudaMemcpyToSymbol(g_attn_trace, &buf, sizeof(buf)));
  return 0;
#else
  (void)buf;
  CX_REQUIRE(false, "cx_debug_attn_trace: this library was built without CX_DEBUG_HOOKS (python -m contrastors_b200.build --debug-hooks)");
#endif
}

// timing ablations for profiling sessions (results are WRONG when set): CX_ATTN_ABLATE bit mask, see the kernels
static int attn_ablate() {
#ifdef CX_DEBUG_HOOKS
  static const int a = [] {
    const char* e = getenv("CX_ATTN_ABLATE");
    return (e && *e) ? atoi(e) : 0;
  }();
  return a;
#else
  return 0;  // release builds never ablate (the hook made the library return wrong results by environment variable)
#endif
}

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
  // once per process, thread-safe (forward runs on the Python thread, backward on autograd's worker thread)
  static std::once_flag configured;
  static cudaError_t cfg_err = cudaSuccess;
  std::call_once(configured, [] {
    auto set = [](const void* f, int bytes) {
      const cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (cfg_err == cudaSuccess) cfg_err = e;
    };
    set((const void*)attn_fwd_kernel, FwdSmem::kTotal);
    set((const void*)attn_fwd2_kernel<true, 0>, Fwd2Smem::kTotal);
    set((const void*)attn_fwd3_kernel<0>, Fwd2Smem::kTotal);
    set((const void*)attn_fwd4_kernel, Fwd4Smem::kTotal);
    set((const void*)attn_fwd5_kernel, Fwd5Smem::kTotal);
    set((const void*)attn_fwd3_kernel<kPoly38>, Fwd2Smem::kTotal);
  });
  CX_CUDA_CHECK(cfg_err);
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  const int mode = fwd_mode();
  const int ablate = attn_ablate();
  if (mode == 1)
    attn_fwd_kernel<<<grid, kFwdThreads, FwdSmem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse, total_tokens, H,
                                                                   softmax_scale * kLog2e);
  else if (mode == 3)
    attn_fwd2_kernel<true, 0><<<grid, kFwd2Threads, Fwd2Smem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse,
                                                                               total_tokens, H, softmax_scale * kLog2e, ablate);
  else if (mode == 9) {
    const int nqt = (max_seqlen + 127) / 128, n_items = nqt * H * nseq;
    int ctas = 2 * sm_count();
    if (ctas > n_items) ctas = n_items;
    attn_fwd5_kernel<<<ctas, kFwd4Threads, Fwd5Smem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse, total_tokens, H,
                                                                      softmax_scale * kLog2e, nqt, n_items);
  } else if (mode == 8)
    attn_fwd4_kernel<<<grid, kFwd4Threads, Fwd4Smem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse, total_tokens, H,
                                                                      softmax_scale * kLog2e);
  else if (mode == 7)
    attn_fwd3_kernel<kPoly38><<<grid, kFwd2Threads, Fwd2Smem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse, total_tokens,
                                                                               H, softmax_scale * kLog2e);
  else
    attn_fwd3_kernel<0><<<grid, kFwd2Threads, Fwd2Smem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse, total_tokens, H,
                                                                         softmax_scale * kLog2e);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_attn_bwd(const void* qkv, const void* out, const void* dout, const float* lse, const int32_t* cu_seqlens,
                           void* dqkv, float* dq_acc, float* delta, int total_tokens, int nseq, int max_seqlen, int H, int Dh,
                           float softmax_scale, cx_stream_t stream_) {
  CX_REQUIRE(qkv && out && dout && lse && cu_seqlens && dqkv && dq_acc && delta, "cx_attn_bwd: null pointer");
  CX_REQUIRE(Dh == kDh, "cx_attn_bwd: only head_dim 64 is implemented");
  CX_REQUIRE(total_tokens > 0 && nseq > 0 && max_seqlen > 0 && H > 0, "cx_attn_bwd: empty problem");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int T = total_tokens;
  {
    const int64_t threads = (int64_t)T * H * 8;
    attn_delta_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, stream>>>((const __nv_bfloat16*)out, (const __nv_bfloat16*)dout, delta, T, H);
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
  static std::once_flag configured;
  static cudaError_t cfg_err = cudaSuccess;
  std::call_once(configured, [] {
    auto set = [](const void* f, int bytes) {
      const cudaError_t e = cudaFuncSetAttribute(f, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
      if (cfg_err == cudaSuccess) cfg_err = e;
    };
    set((const void*)attn_bwd_kernel, BwdSmem::kTotal);
    set((const void*)attn_bwd2_kernel, BwdSmem::kTotal);
    set((const void*)attn_bwd3_kernel, Bwd3Smem::kTotal);
  });
  CX_CUDA_CHECK(cfg_err);
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  const int bmode = bwd_mode();
  if (bmode == 1)
    attn_bwd_kernel<<<grid, kBwdThreads, BwdSmem::kTotal, stream>>>(tmQKV, tmDO, tmDQ, cu_seqlens, lse, delta,
                                                                   (__nv_bfloat16*)dqkv, T, H, softmax_scale);
  else if (bmode == 3)
    attn_bwd3_kernel<<<grid, kBwd2Threads, Bwd3Smem::kTotal, stream>>>(tmQKV, tmDO, tmDQ, cu_seqlens, lse, delta,
                                                                      (__nv_bfloat16*)dqkv, T, H, softmax_scale);
  else
    attn_bwd2_kernel<<<grid, kBwd2Threads, BwdSmem::kTotal, stream>>>(tmQKV, tmDO, tmDQ, cu_seqlens, lse, delta,
                                                                     (__nv_bfloat16*)dqkv, T, H, softmax_scale, attn_ablate());
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
