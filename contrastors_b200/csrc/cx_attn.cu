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
// What binds (tools/ubench, tools/trace_attn.py, profiles/r01_ubench_tmem_mma.txt): forward, the SFU: 16384 exponentials
// per 128 x 128 tile at 16 / clk / SM = 1024 clk against 512 clk of tensor time (S as N = 128: 64 clk per MMA; PV from TMEM:
// 32 clk per MMA) -- plus ~30 % of each CTA's lifetime in prologue (TMA round trip) and epilogue.  Backward, the shared-
// memory port: 240 KB of MMA operands + 128 KB of P / dS / dQ-staging traffic per tile at 128 B/clk = 2900 clk against
// 1664 clk of tensor time (an SS MMA with N = 64 takes 48 clk, not 32: its 6 KB of operands come through that port).
// The thread that ISSUES the MMAs must stay tight: the TMA / MMA warps run converged with elect.sync around the
// asynchronous instructions only (under `if (lane == 0)` every UTCHMMA sits in an elect-and-branch loop, ~80 clk each).
#include <math.h>
#include <stdlib.h>

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

// ============================================================================================== forward
// One CTA = (sequence, head, 128 query rows); two CTAs are co-resident per SM (112.6 KB smem, 256 TMEM columns each), so
// one CTA's softmax overlaps the other's MMAs and prologue.  Softmax: 8 warps, two threads per query row (64 key columns
// each); the row maximum is agreed through a 512-byte bf16 exchange (rounded up, so it is a valid stabiliser for both).
constexpr int kFwdThreads = 384;
struct FwdSmem {
  static constexpr int kTile = 128 * kDh * 2;      // 16 KB: 128 rows x 128 B
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kTile;            // 2 stages
  static constexpr int kV = kK + 2 * kTile;        // 2 stages
  static constexpr int kP = kV + 2 * kTile;        // 32 KB
  static constexpr int kSmax = kP + 32768;         // [2 groups][128 rows] bf16
  static constexpr int kBars = kSmax + 512;
  static constexpr int kTotal = kBars + 112;       // 115,312 B <= 115,712: two CTAs fit in one SM's 228 KB
};

__global__ void __launch_bounds__(kFwdThreads, 2)
attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ cu_seqlens,
                __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + FwdSmem::kBars);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [1]
  uint64_t* p_full = bars + 10;     // [1]
  uint64_t* o_full = bars + 11;     // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int q0 = blockIdx.x * 128;
  if (q0 >= len) return;  // uniform per CTA, before any barrier/TMEM use
  const int nk = (len + 127) / 128;
  if ((smem_u32(smem) & 1023u) != 0) __trap();  // the swizzled tiles need a 1024-byte aligned base

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(p_full, 256);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;  // S: columns [0,128), O: [128,192)

  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, FwdSmem::kTile);
      tma_load_2d(smem + FwdSmem::kQ, &tmQKV, q_full, col_q, seq_begin + q0);
      for (int j = 0; j < nk; ++j) {
        const int st = j & 1;
        const uint32_t ph = (j >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&k_full[st], FwdSmem::kTile);
        tma_load_2d(smem + FwdSmem::kK + st * FwdSmem::kTile, &tmQKV, &k_full[st], col_k, seq_begin + j * 128);
        mbar_wait(&v_empty[st], ph ^ 1);
        mbar_arrive_expect_tx(&v_full[st], FwdSmem::kTile);
        tma_load_2d(smem + FwdSmem::kV + st * FwdSmem::kTile, &tmQKV, &v_full[st], col_v, seq_begin + j * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S = Q K^T: both K-major
      constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O = P V  : A K-major, B (V) MN-major
      const uint32_t q_addr = smem_u32(smem + FwdSmem::kQ), k_addr = smem_u32(smem + FwdSmem::kK);
      const uint32_t v_addr = smem_u32(smem + FwdSmem::kV), p_addr = smem_u32(smem + FwdSmem::kP);
      auto issue_s = [&](int st) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem_base, make_smem_desc_sw128(q_addr + kk * 32, 0, 1024),
                      make_smem_desc_sw128(k_addr + st * FwdSmem::kTile + kk * 32, 0, 1024), idesc_s, kk > 0 ? 1u : 0u);
      };
      auto issue_pv = [&](int st, bool accumulate) {
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ss(tmem_base + 128, make_smem_desc_sw128(p_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                      make_smem_desc_sw128(v_addr + st * FwdSmem::kTile + kk * 2048, 8192, 1024), idesc_o,
                      (accumulate || kk > 0) ? 1u : 0u);
      };
      mbar_wait(q_full, 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0);
      umma_commit(s_full);
      umma_commit(&k_empty[0]);
      for (int j = 0; j < nk; ++j) {
        const int st = j & 1;
        mbar_wait(&v_full[st], (j >> 1) & 1);
        mbar_wait(p_full, j & 1);
        tc_fence_after();
        issue_pv(st, j > 0);
        umma_commit(&v_empty[st]);
        if (j + 1 < nk) {
          const int ns = (j + 1) & 1;
          mbar_wait(&k_full[ns], ((j + 1) >> 1) & 1);
          tc_fence_after();
          issue_s(ns);
          umma_commit(s_full);
          umma_commit(&k_empty[ns]);
        } else {
          umma_commit(o_full);
        }
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax warps (two threads per query row)
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;             // key columns [grp*64, grp*64+64) of each tile; O columns [grp*32, +32)
    const int r = ew * 32 + lane;                // row within the query tile
    const int q_row = q0 + r;                    // row within the sequence
    const uint32_t t_s = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + grp * 64;
    const uint32_t t_o = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + 128 + grp * 32;
    uint8_t* p_smem = smem + FwdSmem::kP + grp * 16384;
    __nv_bfloat16* smax = reinterpret_cast<__nv_bfloat16*>(smem + FwdSmem::kSmax);
    float m_run = -INFINITY, l_run = 0.f;        // l_run: this thread's 64-column share of the row sum
    for (int j = 0; j < nk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      const int kv_valid = min(128, len - j * 128) - grp * 64;  // valid columns among this thread's 64
      const bool full = kv_valid >= 64;
      // one TMEM read of this thread's 64 scores, kept in registers for both the maximum and the exponentials
      uint32_t va[32], vb[32];
      tmem_ld_32x32(t_s, va);
      tmem_ld_32x32(t_s + 32, vb);
      tmem_ld_wait();
      if (!full) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) va[i] = 0xff800000u;       // -inf: never the maximum, exp2 -> 0
          if (32 + i >= kv_valid) vb[i] = 0xff800000u;
        }
      }
      float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        a0 = fmax3(a0, __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
        a1 = fmax3(a1, __uint_as_float(va[i + 2]), __uint_as_float(va[i + 3]));
        a2 = fmax3(a2, __uint_as_float(va[i + 4]), __uint_as_float(va[i + 5]));
        a3 = fmax3(a3, __uint_as_float(va[i + 6]), __uint_as_float(va[i + 7]));
        a0 = fmax3(a0, __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
        a1 = fmax3(a1, __uint_as_float(vb[i + 2]), __uint_as_float(vb[i + 3]));
        a2 = fmax3(a2, __uint_as_float(vb[i + 4]), __uint_as_float(vb[i + 5]));
        a3 = fmax3(a3, __uint_as_float(vb[i + 6]), __uint_as_float(vb[i + 7]));
      }
      const float mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
      // agree on the row maximum with the thread that owns the other 64 columns (values rounded UP to bf16, so the
      // agreed stabiliser is >= the true maximum and identical in both threads)
      const __nv_bfloat16 mine = __float2bfloat16_ru(mx * scale2);
      smax[grp * 128 + r] = mine;
      named_bar_sync(2, 256);
      const float m_new = fmax3(m_run, __bfloat162float(mine), __bfloat162float(smax[(grp ^ 1) * 128 + r]));
      const float alpha = fast_exp2(m_run - m_new);  // 0 on the first tile (m_run = -inf)
      // P = exp2(s*scale2 - m_new) -> bf16 smem (this group's 64-column swizzled block), row-sum share
      float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
      uint8_t* dst = p_smem + r * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        float p[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const uint32_t raw = (q < 4) ? va[8 * q + i] : vb[8 * (q - 4) + i];
          p[i] = fast_exp2(fmaf(__uint_as_float(raw), scale2, -m_new));
        }
        rs0 += p[0] + p[4];
        rs1 += p[1] + p[5];
        rs2 += p[2] + p[6];
        rs3 += p[3] + p[7];
        uint4 w;
        w.x = pack_bf16x2(p[0], p[1]);
        w.y = pack_bf16x2(p[2], p[3]);
        w.z = pack_bf16x2(p[4], p[5]);
        w.w = pack_bf16x2(p[6], p[7]);
        *reinterpret_cast<uint4*>(dst + ((q ^ (r & 7)) << 4)) = w;
      }
      l_run = l_run * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m_run = m_new;
      // rescale this thread's 32 output columns only if some row of the warp moved its maximum
      if (j > 0 && __any_sync(0xffffffffu, alpha != 1.f)) {
        uint32_t v[32];
        tmem_ld_32x32(t_o, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st_32x32(t_o, v);
        tmem_st_wait();
      }
      fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
      tc_fence_before();
      mbar_arrive(p_full);
    }
    // epilogue: combine the two row-sum shares, O / l -> bf16 (this thread's 32 columns), lse
    mbar_wait(o_full, 0);
    tc_fence_after();
    float* lsum = reinterpret_cast<float*>(smem + FwdSmem::kP);  // P is dead once o_full fired
    lsum[grp * 128 + r] = l_run;
    named_bar_sync(2, 256);
    const float l_tot = l_run + lsum[(grp ^ 1) * 128 + r];
    const float inv_l = 1.f / l_tot;
    const bool row_ok = q_row < len;
    __nv_bfloat16* orow = out + ((size_t)(seq_begin + q_row) * H + head) * kDh + grp * 32;
    {
      uint32_t v[32];
      tmem_ld_32x32(t_o, v);
      tmem_ld_wait();
      if (row_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
          w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
          w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
          w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
          *reinterpret_cast<uint4*>(orow + q * 8) = w;
        }
      }
    }
    if (row_ok && grp == 0) lse[(size_t)head * T + seq_begin + q_row] = (m_run + log2f(l_tot)) * kLn2;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- forward, pipelined
// One CTA = (sequence, head, 128 query rows), two CTAs per SM.  Keys are consumed in 64-wide sub-tiles so that ONE thread
// owns a whole query row (no cross-thread maximum exchange) and the score accumulator is double-buffered in TMEM: the MMA
// thread runs S(u+1) while the softmax warps are still busy with S(u).  The row maximum is only raised when it grows by
// more than 2^8 (the final normalisation uses the same stabiliser, so the result is exact); the O accumulator is then
// rescaled in TMEM, which almost never happens after the first sub-tile.  P (bf16) either overwrites the first 32 columns
// of its own score buffer and feeds the PV contraction straight from tensor memory (kPTmem), or goes through swizzled
// shared memory.  Warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-7 softmax (thread = query row).
constexpr int kFwd2Threads = 256;
struct Fwd2Smem {
  static constexpr int kTile = 128 * kDh * 2;      // 16 KB: 128 rows x 128 B
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kTile;            // 2 stages of 128 keys
  static constexpr int kV = kK + 2 * kTile;        // 2 stages
  static constexpr int kP = kV + 2 * kTile;        // 2 buffers [128 q x 64 keys] bf16 (smem-P mode only)
  static constexpr int kBars = kP + 2 * kTile;
  static constexpr int kTotal = kBars + 256;       // 114,944 B: two CTAs per SM (17 barriers + the TMEM base word)
};

// kPolyMask: bit (t & 7) set => the t-th column pair of a row takes its exponentials from the FMA-pipe polynomial
// instead of MUFU.EX2 (the forward pass at head dim 64 is bound by the 16 exponentials / clk / SM of the SFU).
template <bool kPTmem, uint32_t kPolyMask>
__global__ void __launch_bounds__(kFwd2Threads, 2)
attn_fwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ cu_seqlens,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale2, int ablate) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Fwd2Smem::kBars);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [2]  S(u) in score buffer u & 1
  uint64_t* p_ready = bars + 11;    // [2]  P(u) written, score registers loaded (128 arrivals)
  uint64_t* pv_done = bars + 13;    // [1]  PV(u) complete
  uint64_t* o_full = bars + 14;     // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int q0 = blockIdx.x * 128;
  if (q0 >= len) return;  // uniform per CTA, before any barrier/TMEM use
  const int nk = (len + 127) / 128;  // 128-key TMA tiles
  const int nu = (len + 63) / 64;    // 64-key sub-tiles
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  long long* tr = g_attn_trace;
  if (tr != nullptr) {
    tr += ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64;
    if (threadIdx.x == 0) {
      uint32_t smid;
      unsigned long long gt;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      tr[0] = clock64();
      tr[1] = smid;
      tr[2] = (long long)gt;
    }
  }

  if (warp == 0 && lane == 0) tma_prefetch_desc(&tmQKV);
  if (warp == 1 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&s_full[i], 1);
      mbar_init(&p_ready[i], 128);
    }
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;  // S0 / P0: [0,64)   S1 / P1: [64,128)   O: [128,192)
  if (threadIdx.x == 0) trace_put(tr, 3);

  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;

  // The TMA and MMA warps run CONVERGED (all 32 lanes wait on the barriers) and elect one lane around the asynchronous
  // instructions only: code under `if (lane == 0)` makes ptxas wrap every UTCHMMA / UTCBAR in an elect-and-branch loop
  // with its descriptor arithmetic in between (~80 clk per MMA), which made the issuing thread the critical path.
  if (warp == 0) {
    if (elect_one()) {
      mbar_arrive_expect_tx(q_full, Fwd2Smem::kTile);
      tma_load_2d(smem + Fwd2Smem::kQ, &tmQKV, q_full, col_q, seq_begin + q0);
    }
    __syncwarp();
    for (int j = 0; j < nk; ++j) {
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[st], Fwd2Smem::kTile);
        tma_load_2d(smem + Fwd2Smem::kK + st * Fwd2Smem::kTile, &tmQKV, &k_full[st], col_k, seq_begin + j * 128);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[st], Fwd2Smem::kTile);
        tma_load_2d(smem + Fwd2Smem::kV + st * Fwd2Smem::kTile, &tmQKV, &v_full[st], col_v, seq_begin + j * 128);
      }
      __syncwarp();
      if (lane == 0 && j < 4) trace_put(tr, 26 + j);
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 64, 0, 0);   // S = Q K^T (64 keys): both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O += P V   : A K-major, B (V) MN-major
    // descriptor of (base + off) = descriptor of base + (off >> 4): the address field never carries out of its 14 bits
    const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + Fwd2Smem::kQ), 0, 1024);
    const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + Fwd2Smem::kK), 0, 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + Fwd2Smem::kV), 8192, 1024);
    const uint64_t pd = make_smem_desc_sw128(smem_u32(smem + Fwd2Smem::kP), 0, 1024);
    // sub-tile u (uu = u & 3 is a compile-time constant after unrolling): key tile u >> 1 in stage (u >> 1) & 1, half
    // u & 1 (+8 KB), score buffer u & 1
    auto issue_s = [&](const int uu, const int u) {
      if (elect_one()) {
        const uint32_t off = ((uu >> 1) & 1) * Fwd2Smem::kTile + (uu & 1) * 8192;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem_base + (uu & 1) * 64, qd + ((kk * 32) >> 4), kd + ((off + kk * 32) >> 4), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(&s_full[uu & 1]);
        if ((uu & 1) || u == nu - 1) umma_commit(&k_empty[(uu >> 1) & 1]);  // last reader of this key tile
      }
      __syncwarp();
    };
    auto issue_pv = [&](const int uu, const int u) {
      if (elect_one()) {
        const uint32_t off = ((uu >> 1) & 1) * Fwd2Smem::kTile + (uu & 1) * 8192;
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint64_t bdesc = vd + ((off + kk * 2048) >> 4);
          if (kPTmem)
            umma_f16_ts(tmem_base + 128, tmem_base + (uu & 1) * 64 + kk * 8, bdesc, idesc_o, (u > 0 || kk > 0) ? 1u : 0u);
          else
            umma_f16_ss(tmem_base + 128, pd + (((uu & 1) * Fwd2Smem::kTile + kk * 32) >> 4), bdesc, idesc_o,
                        (u > 0 || kk > 0) ? 1u : 0u);
        }
        if ((uu & 1) || u == nu - 1) umma_commit(&v_empty[(uu >> 1) & 1]);
        umma_commit(pv_done);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    if (lane == 0) trace_put(tr, 4);
    issue_s(0, 0);
    if (nu > 1) issue_s(1, 1);
    for (int u0 = 0; u0 < nu; u0 += 4) {
#pragma unroll
      for (int uu = 0; uu < 4; ++uu) {
        const int u = u0 + uu;
        if (u < nu) {  // warp-uniform
          if ((uu & 1) == 0) mbar_wait(&v_full[(uu >> 1) & 1], (u >> 2) & 1);
          mbar_wait(&p_ready[uu & 1], (u >> 1) & 1);
          tc_fence_after();
          if (lane == 0 && u < 8) trace_put(tr, 32 + u);
          if ((ablate & 32) && u + 2 < nu) {  // interleave the k-steps of PV(u) and S(u+2) (independent accumulators)
            if ((uu & 1) == 0) {
              mbar_wait(&k_full[((uu + 2) >> 1) & 1], ((u + 2) >> 2) & 1);
              tc_fence_after();
            }
            if (elect_one()) {
              const int u2 = (uu + 2) & 3;
              const uint32_t offv = ((uu >> 1) & 1) * Fwd2Smem::kTile + (uu & 1) * 8192;
              const uint32_t offk = ((u2 >> 1) & 1) * Fwd2Smem::kTile + (u2 & 1) * 8192;
#pragma unroll
              for (int kk = 0; kk < 4; ++kk) {
                const uint64_t bdesc = vd + ((offv + kk * 2048) >> 4);
                if (kPTmem)
                  umma_f16_ts(tmem_base + 128, tmem_base + (uu & 1) * 64 + kk * 8, bdesc, idesc_o, (u > 0 || kk > 0) ? 1u : 0u);
                else
                  umma_f16_ss(tmem_base + 128, pd + (((uu & 1) * Fwd2Smem::kTile + kk * 32) >> 4), bdesc, idesc_o, (u > 0 || kk > 0) ? 1u : 0u);
                if (kk == 3) {  // PV(u) complete in issue order before S(u+2) finishes overwriting the P columns
                  if ((uu & 1) || u == nu - 1) umma_commit(&v_empty[(uu >> 1) & 1]);
                  umma_commit(pv_done);
                }
              }
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ss(tmem_base + (u2 & 1) * 64, qd + ((kk * 32) >> 4), kd + ((offk + kk * 32) >> 4), idesc_s, kk > 0 ? 1u : 0u);
              umma_commit(&s_full[u2 & 1]);
              if ((u2 & 1) || u + 2 == nu - 1) umma_commit(&k_empty[(u2 >> 1) & 1]);
            }
            __syncwarp();
          } else {
          issue_pv(uu, u);
          if (u + 2 < nu) {
            if ((uu & 1) == 0) {
              mbar_wait(&k_full[((uu + 2) >> 1) & 1], ((u + 2) >> 2) & 1);
              tc_fence_after();
            }
            issue_s((uu + 2) & 3, u + 2);
          }
          }
        }
      }
    }
    if (elect_one()) umma_commit(o_full);
    __syncwarp();
    if (lane == 0) trace_put(tr, 5);
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax: one thread per query row
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const int q_row = q0 + r;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t t_o = tmem_base + lane_base + 128;
    const float2 sc2 = make_float2(scale2, scale2);
    float m_run = -INFINITY, l_run = 0.f;
    for (int u = 0; u < nu; ++u) {
      const int b = u & 1;
      const uint32_t t_s = tmem_base + lane_base + b * 64;
      mbar_wait(&s_full[b], (u >> 1) & 1);
      tc_fence_after();
      if (threadIdx.x == 128 && u < 8) trace_put(tr, 6 + u);
      uint32_t va[32], vb[32];
      if (!(ablate & 8)) {
        tmem_ld_32x32(t_s, va);
        tmem_ld_32x32(t_s + 32, vb);
        tmem_ld_wait();
      } else {  // timing ablation: no score read (results are wrong)
#pragma unroll
        for (int i = 0; i < 32; ++i) va[i] = vb[i] = __float_as_uint(0.01f * (float)(i + lane));
      }
      const int kv_valid = len - u * 64;
      if (kv_valid < 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) va[i] = 0xff800000u;       // -inf: never the maximum, exp2 -> 0
          if (32 + i >= kv_valid) vb[i] = 0xff800000u;
        }
      }
      float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        a0 = fmax3(a0, __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
        a1 = fmax3(a1, __uint_as_float(va[i + 2]), __uint_as_float(va[i + 3]));
        a2 = fmax3(a2, __uint_as_float(va[i + 4]), __uint_as_float(va[i + 5]));
        a3 = fmax3(a3, __uint_as_float(va[i + 6]), __uint_as_float(va[i + 7]));
        a0 = fmax3(a0, __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
        a1 = fmax3(a1, __uint_as_float(vb[i + 2]), __uint_as_float(vb[i + 3]));
        a2 = fmax3(a2, __uint_as_float(vb[i + 4]), __uint_as_float(vb[i + 5]));
        a3 = fmax3(a3, __uint_as_float(vb[i + 6]), __uint_as_float(vb[i + 7]));
      }
      const float m_c = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * scale2;
      const bool raise = m_c > m_run + 8.f;              // always true on the first sub-tile (m_run = -inf)
      float alpha = 1.f;
      if (raise) {
        alpha = fast_exp2(m_run - m_c);                  // 0 on the first sub-tile
        m_run = m_c;
      }
      const float2 nm2 = make_float2(-m_run, -m_run);
      // P = exp2(s * scale2 - m_run) -> bf16 pairs, row-sum
      uint32_t pp[32];
      float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        const uint32_t* v0 = (t < 16) ? &va[2 * t] : &vb[2 * (t - 16)];
        float2 x0 = ffma2(make_float2(__uint_as_float(v0[0]), __uint_as_float(v0[1])), sc2, nm2);
        float2 x1 = ffma2(make_float2(__uint_as_float(v0[2]), __uint_as_float(v0[3])), sc2, nm2);
        if (!(ablate & 2)) {  // (ablation bit 2: no exponentials)
          x0 = ((kPolyMask >> (t & 7)) & 1u) ? exp2_poly2(x0) : make_float2(fast_exp2(x0.x), fast_exp2(x0.y));
          x1 = ((kPolyMask >> ((t + 1) & 7)) & 1u) ? exp2_poly2(x1) : make_float2(fast_exp2(x1.x), fast_exp2(x1.y));
        }
        rs0 = fadd2(rs0, x0);
        rs1 = fadd2(rs1, x1);
        pp[t] = pack_bf16x2(x0.x, x0.y);
        pp[t + 1] = pack_bf16x2(x1.x, x1.y);
      }
      l_run = l_run * alpha + ((rs0.x + rs0.y) + (rs1.x + rs1.y));
      if (ablate & 4) {          // timing ablation: P is not written
      } else if (kPTmem) {
        tmem_st_32x32(t_s, pp);  // P overwrites the first 32 columns of its own (already loaded) score buffer
      } else {
        uint8_t* dst = smem + Fwd2Smem::kP + b * Fwd2Smem::kTile + r * 128;  // free: S(u) was committed after PV(u-2)
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<uint4*>(dst + ((ch ^ (r & 7)) << 4)) = make_uint4(pp[4 * ch], pp[4 * ch + 1], pp[4 * ch + 2], pp[4 * ch + 3]);
      }
      // rescale O only if some row of the warp raised its maximum; PV(u-1) must have completed, PV(u) is not issued yet
      if (u > 0 && __any_sync(0xffffffffu, raise)) {
        mbar_wait(pv_done, (u - 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_o + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(t_o + c * 32, v);
        }
      }
      if (kPTmem) {
        tmem_st_wait();
      } else {
        tmem_st_wait();
        fence_proxy_async_smem();  // P (generic-proxy stores) -> visible to the tensor core's async proxy
      }
      tc_fence_before();
      mbar_arrive(&p_ready[b]);
      if (threadIdx.x == 128 && u < 8) trace_put(tr, 14 + u);
    }
    // epilogue: O / l -> bf16, lse
    mbar_wait(o_full, 0);
    tc_fence_after();
    if (threadIdx.x == 128) trace_put(tr, 22);
    const float inv_l = 1.f / l_run;
    const bool row_ok = q_row < len;
    // O / l -> bf16 -> the (dead) Q tile in smem, one swizzled 128-byte row per thread; then the 128 softmax threads copy
    // the tile out with 16-byte stores that are contiguous along each row (a warp store covers 4 full 128-byte rows
    // instead of 32 partial ones: the per-thread row stores cost ~1800 clk per CTA)
    uint8_t* stg = smem + Fwd2Smem::kQ;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_o + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
        w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
        w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
        w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stg + r * 128 + (((c * 4 + q) ^ (r & 7)) << 4)) = w;
      }
    }
    named_bar_sync(1, 128);
    {
      const int tid = threadIdx.x - 128;
      const int rows_ok = min(128, len - q0);
      uint8_t* obase = reinterpret_cast<uint8_t*>(out + ((size_t)(seq_begin + q0) * H + head) * kDh);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + tid, row = idx >> 3, ch = idx & 7;
        if (row < rows_ok)
          *reinterpret_cast<uint4*>(obase + (size_t)row * H * kDh * 2 + ch * 16) =
              *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
    if (row_ok) lse[(size_t)head * T + seq_begin + q_row] = (m_run + log2f(l_run)) * kLn2;
    if (threadIdx.x == 128) trace_put(tr, 23);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
  if (threadIdx.x == 0) trace_put(tr, 24);
}

// ---------------------------------------------------------------------------------------------- forward, wide-S variant
// As attn_fwd2_kernel, but the scores of a whole 128-key tile come from ONE chain of four N = 128 MMAs (half the
// instructions to issue, and Q is read from shared memory once per key tile instead of twice), while the softmax still
// works in 64-column halves with one thread per row.  P (bf16) gets its own 2 x 32 TMEM columns, so
// the score columns are free as soon as every thread has LOADED the second half into registers (s_free): S(j+1) then
// runs under the exponentials of the second half of tile j.  TMEM: S [0,128)  O [128,192)  P0 [192,224)  P1 [224,256).
template <uint32_t kPolyMask>
__global__ void __launch_bounds__(kFwd2Threads, 2)
attn_fwd3_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ cu_seqlens,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Fwd2Smem::kBars);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [1]  S(j) (128 keys) in TMEM
  uint64_t* s_free = bars + 10;     // [1]  every softmax thread holds its S(j) row in registers (128 arrivals per tile)
  uint64_t* p_ready = bars + 11;    // [2]  P of half h written (128 arrivals per tile)
  uint64_t* pv_done = bars + 13;    // [2]  PV of half h complete (one completion per tile)
  uint64_t* o_full = bars + 15;     // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 16);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int q0 = blockIdx.x * 128;
  if (q0 >= len) return;  // uniform per CTA, before any barrier/TMEM use
  const int nk = (len + 127) / 128;  // 128-key tiles
  const int nu = (len + 63) / 64;    // 64-key halves
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;
  if (warp == 0 && lane == 0) {
    // the producer lane initialises the barriers itself and starts Q, K_0, V_0 before the CTA-wide sync, so the TMA round
    // trip overlaps the TMEM allocation and the rest of the set-up
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
      mbar_init(&p_ready[i], 128);
      mbar_init(&pv_done[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 128);
    mbar_init(o_full, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(q_full, Fwd2Smem::kTile);
    tma_load_2d(smem + Fwd2Smem::kQ, &tmQKV, q_full, col_q, seq_begin + q0);
    mbar_arrive_expect_tx(&k_full[0], Fwd2Smem::kTile);
    tma_load_2d(smem + Fwd2Smem::kK, &tmQKV, &k_full[0], col_k, seq_begin);
    mbar_arrive_expect_tx(&v_full[0], Fwd2Smem::kTile);
    tma_load_2d(smem + Fwd2Smem::kV, &tmQKV, &v_full[0], col_v, seq_begin);
  }
  if (warp == 2) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;


  if (warp == 0) {
    for (int j = 1; j < nk; ++j) {  // Q and key tile 0 were issued during set-up
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[st], Fwd2Smem::kTile);
        tma_load_2d(smem + Fwd2Smem::kK + st * Fwd2Smem::kTile, &tmQKV, &k_full[st], col_k, seq_begin + j * 128);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[st], Fwd2Smem::kTile);
        tma_load_2d(smem + Fwd2Smem::kV + st * Fwd2Smem::kTile, &tmQKV, &v_full[st], col_v, seq_begin + j * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S = Q K^T (128 keys): both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O += P V: A (P) from TMEM, B (V) MN-major
    const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + Fwd2Smem::kQ), 0, 1024);
    const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + Fwd2Smem::kK), 0, 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + Fwd2Smem::kV), 8192, 1024);
    auto issue_s = [&](const int st) {  // st = stage of the key tile (compile-time after unrolling)
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem_base, qd + ((kk * 32) >> 4), kd + ((st * Fwd2Smem::kTile + kk * 32) >> 4), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j0 = 0; j0 < nk; j0 += 2) {
#pragma unroll
      for (int st = 0; st < 2; ++st) {  // st = j & 1
        const int j = j0 + st;
        if (j < nk) {  // warp-uniform
          const int halves = min(2, nu - 2 * j);
          mbar_wait(&v_full[st], (j >> 1) & 1);
#pragma unroll
          for (int h = 0; h < 2; ++h) {
            if (h < halves) {
              mbar_wait(&p_ready[h], j & 1);
              tc_fence_after();
              if (elect_one()) {
#pragma unroll
                for (int kk = 0; kk < 4; ++kk)
                  umma_f16_ts(tmem_base + 128, tmem_base + 192 + h * 32 + kk * 8,
                              vd + ((st * Fwd2Smem::kTile + h * 8192 + kk * 2048) >> 4), idesc_o, (j > 0 || h > 0 || kk > 0) ? 1u : 0u);
                if (h == halves - 1) umma_commit(&v_empty[st]);
                umma_commit(&pv_done[h]);
              }
              __syncwarp();
              if (h == 0 && j + 1 < nk) {  // S(j+1) as soon as the score columns have been read out
                mbar_wait(s_free, j & 1);
                mbar_wait(&k_full[st ^ 1], ((j + 1) >> 1) & 1);
                tc_fence_after();
                issue_s(st ^ 1);
              }
            }
          }
        }
      }
    }
    if (elect_one()) umma_commit(o_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax: one thread per query row
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const int q_row = q0 + r;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t t_o = tmem_base + lane_base + 128;
    const float2 sc2 = make_float2(scale2, scale2);
    float m_run = -INFINITY, l_run = 0.f;
    for (int u = 0; u < nu; ++u) {
      const int h = u & 1, j = u >> 1;
      if (h == 0) {
        mbar_wait(s_full, j & 1);
        tc_fence_after();
      }
      uint32_t va[32], vb[32];
      tmem_ld_32x32(tmem_base + lane_base + h * 64, va);
      tmem_ld_32x32(tmem_base + lane_base + h * 64 + 32, vb);
      tmem_ld_wait();
      if (h == 1 || u == nu - 1) {  // this thread's S(j) row is in registers: exactly one arrival per tile
        tc_fence_before();
        mbar_arrive(s_free);
      }
      const int kv_valid = len - u * 64;
      if (kv_valid < 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) va[i] = 0xff800000u;       // -inf: never the maximum, exp2 -> 0
          if (32 + i >= kv_valid) vb[i] = 0xff800000u;
        }
      }
      float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        a0 = fmax3(a0, __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
        a1 = fmax3(a1, __uint_as_float(va[i + 2]), __uint_as_float(va[i + 3]));
        a2 = fmax3(a2, __uint_as_float(va[i + 4]), __uint_as_float(va[i + 5]));
        a3 = fmax3(a3, __uint_as_float(va[i + 6]), __uint_as_float(va[i + 7]));
        a0 = fmax3(a0, __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
        a1 = fmax3(a1, __uint_as_float(vb[i + 2]), __uint_as_float(vb[i + 3]));
        a2 = fmax3(a2, __uint_as_float(vb[i + 4]), __uint_as_float(vb[i + 5]));
        a3 = fmax3(a3, __uint_as_float(vb[i + 6]), __uint_as_float(vb[i + 7]));
      }
      const float m_c = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3)) * scale2;
      const bool raise = m_c > m_run + 8.f;              // always true on the first half (m_run = -inf)
      float alpha = 1.f;
      if (raise) {
        alpha = fast_exp2(m_run - m_c);                  // 0 on the first half
        m_run = m_c;
      }
      const float2 nm2 = make_float2(-m_run, -m_run);
      uint32_t pp[32];
      float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int t = 0; t < 32; t += 2) {
        const uint32_t* v0 = (t < 16) ? &va[2 * t] : &vb[2 * (t - 16)];
        float2 x0 = ffma2(make_float2(__uint_as_float(v0[0]), __uint_as_float(v0[1])), sc2, nm2);
        float2 x1 = ffma2(make_float2(__uint_as_float(v0[2]), __uint_as_float(v0[3])), sc2, nm2);
        x0 = ((kPolyMask >> (t & 7)) & 1u) ? exp2_poly2(x0) : make_float2(fast_exp2(x0.x), fast_exp2(x0.y));
        x1 = ((kPolyMask >> ((t + 1) & 7)) & 1u) ? exp2_poly2(x1) : make_float2(fast_exp2(x1.x), fast_exp2(x1.y));
        rs0 = fadd2(rs0, x0);
        rs1 = fadd2(rs1, x1);
        pp[t] = pack_bf16x2(x0.x, x0.y);
        pp[t + 1] = pack_bf16x2(x1.x, x1.y);
      }
      l_run = l_run * alpha + ((rs0.x + rs0.y) + (rs1.x + rs1.y));
      if (j > 0) {  // the P columns of this half were last read by PV(u-2)
        mbar_wait(&pv_done[h], (j - 1) & 1);
        tc_fence_after();
      }
      tmem_st_32x32(tmem_base + lane_base + 192 + h * 32, pp);
      // rescale O only if some row of the warp raised its maximum; PV(u-1) must have completed, PV(u) is not issued yet
      if (u > 0 && __any_sync(0xffffffffu, raise)) {
        mbar_wait(&pv_done[h ^ 1], ((u - 1) >> 1) & 1);
        tc_fence_after();
#pragma unroll
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_o + c * 32, v);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
          tmem_st_32x32(t_o + c * 32, v);
        }
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_ready[h]);
    }
    // epilogue: O / l -> bf16 (staged through the dead Q tile for row-contiguous stores), lse
    mbar_wait(o_full, 0);
    tc_fence_after();
    const float inv_l = 1.f / l_run;
    const bool row_ok = q_row < len;
    uint8_t* stg = smem + Fwd2Smem::kQ;
#pragma unroll
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(t_o + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
        w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
        w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
        w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stg + r * 128 + (((c * 4 + q) ^ (r & 7)) << 4)) = w;
      }
    }
    named_bar_sync(1, 128);
    {
      const int tid = threadIdx.x - 128;
      const int rows_ok = min(128, len - q0);
      uint8_t* obase = reinterpret_cast<uint8_t*>(out + ((size_t)(seq_begin + q0) * H + head) * kDh);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + tid, row = idx >> 3, ch = idx & 7;
        if (row < rows_ok)
          *reinterpret_cast<uint4*>(obase + (size_t)row * H * kDh * 2 + ch * 16) =
              *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
    if (row_ok) lse[(size_t)head * T + seq_begin + q_row] = (m_run + log2f(l_run)) * kLn2;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- forward, two threads per row
// EXPERIMENTAL (CX_ATTN_FWD=8; written after the last GPU minute of round 1: it compiles, it has NOT run on hardware yet).
// attn_fwd3_kernel's pipeline with the serial kernel's softmax layout: 8 softmax warps per CTA, two threads per query row,
// each owning 64 of the 128 key columns of a tile, so a whole key tile is one softmax step and every SM sub-partition has
// four softmax warps to overlap (tools/trace_attn.py: with one warp per sub-partition per CTA the loop is bound by that
// warp's own latency chain, ~1100 clk per 64 exponentials per thread, not by the SFU).  The two threads of a row agree on
// the (lazily raised) maximum through a 2-byte exchange in shared memory.
// TMEM: S [0,128)  O [128,192)  P [192,256) (128 keys as bf16 pairs).  Warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 softmax.
constexpr int kFwd4Threads = 384;
struct Fwd4Smem {
  static constexpr int kTile = 128 * kDh * 2;      // 16 KB: 128 rows x 128 B
  static constexpr int kQ = 0;
  static constexpr int kK = kQ + kTile;            // 2 stages of 128 keys
  static constexpr int kV = kK + 2 * kTile;        // 2 stages
  static constexpr int kX = kV + 2 * kTile;        // exchange: 2 x [2 groups][128 rows] bf16 maxima, then [2][128] fp32 row sums
  static constexpr int kBars = kX + 1024;
  static constexpr int kTotal = kBars + 256;       // 83,200 B: two CTAs per SM
};

__global__ void __launch_bounds__(kFwd4Threads, 2)
attn_fwd4_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ cu_seqlens,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale2) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Fwd4Smem::kBars);
  uint64_t* q_full = bars;          // [1]
  uint64_t* k_full = bars + 1;      // [2]
  uint64_t* k_empty = bars + 3;     // [2]
  uint64_t* v_full = bars + 5;      // [2]
  uint64_t* v_empty = bars + 7;     // [2]
  uint64_t* s_full = bars + 9;      // [1]  S(j) in TMEM
  uint64_t* s_free = bars + 10;     // [1]  every softmax thread holds its 64 scores of S(j) in registers (256 arrivals)
  uint64_t* p_ready = bars + 11;    // [1]  P(j) written (256 arrivals)
  uint64_t* pv_done = bars + 12;    // [1]  PV(j) complete
  uint64_t* o_full = bars + 13;     // [1]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 14);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int q0 = blockIdx.x * 128;
  if (q0 >= len) return;  // uniform per CTA, before any barrier/TMEM use
  const int nk = (len + 127) / 128;
  if ((smem_u32(smem) & 1023u) != 0) __trap();

  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;
  if (warp == 0 && lane == 0) {
    mbar_init(q_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&k_full[i], 1);
      mbar_init(&k_empty[i], 1);
      mbar_init(&v_full[i], 1);
      mbar_init(&v_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(s_free, 256);
    mbar_init(p_ready, 256);
    mbar_init(pv_done, 1);
    mbar_init(o_full, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(q_full, Fwd4Smem::kTile);
    tma_load_2d(smem + Fwd4Smem::kQ, &tmQKV, q_full, col_q, seq_begin + q0);
    mbar_arrive_expect_tx(&k_full[0], Fwd4Smem::kTile);
    tma_load_2d(smem + Fwd4Smem::kK, &tmQKV, &k_full[0], col_k, seq_begin);
    mbar_arrive_expect_tx(&v_full[0], Fwd4Smem::kTile);
    tma_load_2d(smem + Fwd4Smem::kV, &tmQKV, &v_full[0], col_v, seq_begin);
  }
  if (warp == 2) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    for (int j = 1; j < nk; ++j) {  // Q and key tile 0 were issued during set-up
      const int st = j & 1;
      const uint32_t ph = (j >> 1) & 1;
      mbar_wait(&k_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&k_full[st], Fwd4Smem::kTile);
        tma_load_2d(smem + Fwd4Smem::kK + st * Fwd4Smem::kTile, &tmQKV, &k_full[st], col_k, seq_begin + j * 128);
      }
      __syncwarp();
      mbar_wait(&v_empty[st], ph ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&v_full[st], Fwd4Smem::kTile);
        tma_load_2d(smem + Fwd4Smem::kV + st * Fwd4Smem::kTile, &tmQKV, &v_full[st], col_v, seq_begin + j * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S = Q K^T (128 keys): both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O += P V: A (P) from TMEM, B (V) MN-major
    const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + Fwd4Smem::kQ), 0, 1024);
    const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + Fwd4Smem::kK), 0, 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + Fwd4Smem::kV), 8192, 1024);
    auto issue_s = [&](const int st) {
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem_base, qd + ((kk * 32) >> 4), kd + ((st * Fwd4Smem::kTile + kk * 32) >> 4), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    mbar_wait(q_full, 0);
    mbar_wait(&k_full[0], 0);
    tc_fence_after();
    issue_s(0);
    for (int j0 = 0; j0 < nk; j0 += 2) {
#pragma unroll
      for (int st = 0; st < 2; ++st) {  // st = j & 1
        const int j = j0 + st;
        if (j < nk) {  // warp-uniform
          if (j + 1 < nk) {  // S(j+1) as soon as the score columns have been read out: it runs under softmax(j)
            mbar_wait(s_free, j & 1);
            mbar_wait(&k_full[st ^ 1], ((j + 1) >> 1) & 1);
            tc_fence_after();
            issue_s(st ^ 1);
          }
          mbar_wait(&v_full[st], (j >> 1) & 1);
          mbar_wait(p_ready, j & 1);
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16_ts(tmem_base + 128, tmem_base + 192 + kk * 8, vd + ((st * Fwd4Smem::kTile + kk * 2048) >> 4), idesc_o,
                          (j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&v_empty[st]);
            umma_commit(pv_done);
          }
          __syncwarp();
        }
      }
    }
    if (elect_one()) umma_commit(o_full);
    __syncwarp();
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax: two threads per query row
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;             // key columns [grp*64, +64) of each tile; O columns [grp*32, +32)
    const int r = ew * 32 + lane;
    const int q_row = q0 + r;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + grp * 64;
    const uint32_t t_o = tmem_base + lane_base + 128 + grp * 32;
    const uint32_t t_p = tmem_base + lane_base + 192 + grp * 32;
    __nv_bfloat16* smax = reinterpret_cast<__nv_bfloat16*>(smem + Fwd4Smem::kX);  // [2 parities][2 groups][128 rows]
    const float2 sc2 = make_float2(scale2, scale2);
    float m_run = -INFINITY, l_run = 0.f;        // l_run: this thread's 64-column share of the row sum
    for (int j = 0; j < nk; ++j) {
      mbar_wait(s_full, j & 1);
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld_32x32(t_s, va);
      tmem_ld_32x32(t_s + 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(s_free);                       // this thread's scores are in registers
      const int kv_valid = min(128, len - j * 128) - grp * 64;
      if (kv_valid < 64) {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
          if (i >= kv_valid) va[i] = 0xff800000u;       // -inf: never the maximum, exp2 -> 0
          if (32 + i >= kv_valid) vb[i] = 0xff800000u;
        }
      }
      float a0 = -INFINITY, a1 = -INFINITY, a2 = -INFINITY, a3 = -INFINITY;
#pragma unroll
      for (int i = 0; i < 32; i += 8) {
        a0 = fmax3(a0, __uint_as_float(va[i]), __uint_as_float(va[i + 1]));
        a1 = fmax3(a1, __uint_as_float(va[i + 2]), __uint_as_float(va[i + 3]));
        a2 = fmax3(a2, __uint_as_float(va[i + 4]), __uint_as_float(va[i + 5]));
        a3 = fmax3(a3, __uint_as_float(va[i + 6]), __uint_as_float(va[i + 7]));
        a0 = fmax3(a0, __uint_as_float(vb[i]), __uint_as_float(vb[i + 1]));
        a1 = fmax3(a1, __uint_as_float(vb[i + 2]), __uint_as_float(vb[i + 3]));
        a2 = fmax3(a2, __uint_as_float(vb[i + 4]), __uint_as_float(vb[i + 5]));
        a3 = fmax3(a3, __uint_as_float(vb[i + 6]), __uint_as_float(vb[i + 7]));
      }
      const float mx = fmaxf(fmaxf(a0, a1), fmaxf(a2, a3));
      // agree on the tile maximum with the thread that owns the other 64 columns (rounded UP to bf16: a valid stabiliser,
      // identical in both threads; a fully masked half contributes -inf)
      const __nv_bfloat16 mine = __float2bfloat16_ru(mx * scale2);
      __nv_bfloat16* xm = smax + (j & 1) * 256;
      xm[grp * 128 + r] = mine;
      named_bar_sync(2, 256);
      const float m_c = fmaxf(__bfloat162float(mine), __bfloat162float(xm[(grp ^ 1) * 128 + r]));
      const bool raise = m_c > m_run + 8.f;      // always true on the first tile (m_run = -inf); same in both threads
      float alpha = 1.f;
      if (raise) {
        alpha = fast_exp2(m_run - m_c);          // 0 on the first tile
        m_run = m_c;
      }
      const float2 nm2 = make_float2(-m_run, -m_run);
      // P = exp2(s * scale2 - m_run) in two 32-column batches, each stored to its 16 TMEM columns as soon as it is packed
      // (keeps the live registers under the 80 the 2-CTA/SM occupancy allows)
      float2 rs0 = make_float2(0.f, 0.f), rs1 = make_float2(0.f, 0.f);
#pragma unroll
      for (int b = 0; b < 2; ++b) {
        uint32_t pp[16];
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          const uint32_t* v0 = (b == 0) ? &va[2 * t] : &vb[2 * t];
          float2 x0 = ffma2(make_float2(__uint_as_float(v0[0]), __uint_as_float(v0[1])), sc2, nm2);
          float2 x1 = ffma2(make_float2(__uint_as_float(v0[2]), __uint_as_float(v0[3])), sc2, nm2);
          x0 = make_float2(fast_exp2(x0.x), fast_exp2(x0.y));
          x1 = make_float2(fast_exp2(x1.x), fast_exp2(x1.y));
          rs0 = fadd2(rs0, x0);
          rs1 = fadd2(rs1, x1);
          pp[t] = pack_bf16x2(x0.x, x0.y);
          pp[t + 1] = pack_bf16x2(x1.x, x1.y);
        }
        if (b == 0 && j > 0) {  // PV(j-1) has finished reading the P columns and accumulating into O
          mbar_wait_quiet(pv_done, (j - 1) & 1);
          tc_fence_after();
        }
        tmem_st_32x16(t_p + b * 16, pp);
      }
      l_run = l_run * alpha + ((rs0.x + rs0.y) + (rs1.x + rs1.y));
      // rescale this thread's 32 output columns only if some row of the warp raised its maximum
      if (j > 0 && __any_sync(0xffffffffu, raise)) {
        uint32_t v[32];
        tmem_ld_32x32(t_o, v);
        tmem_ld_wait();
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = __float_as_uint(__uint_as_float(v[i]) * alpha);
        tmem_st_32x32(t_o, v);
      }
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
    }
    // epilogue: combine the two row-sum shares, O / l -> bf16 (staged through the dead Q tile), lse
    mbar_wait(o_full, 0);
    tc_fence_after();
    float* lsum = reinterpret_cast<float*>(smem + Fwd4Smem::kX);  // the maxima are dead once o_full fired
    named_bar_sync(2, 256);                                        // ... in EVERY thread (last exchange read is behind us)
    lsum[grp * 128 + r] = l_run;
    named_bar_sync(2, 256);
    const float l_tot = l_run + lsum[(grp ^ 1) * 128 + r];
    const float inv_l = 1.f / l_tot;
    const bool row_ok = q_row < len;
    uint8_t* stg = smem + Fwd4Smem::kQ;
    {
      uint32_t v[32];
      tmem_ld_32x32(t_o, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
        w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
        w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
        w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stg + r * 128 + (((grp * 4 + q) ^ (r & 7)) << 4)) = w;
      }
    }
    named_bar_sync(2, 256);
    {
      const int tid = threadIdx.x - 128;
      const int rows_ok = min(128, len - q0);
      uint8_t* obase = reinterpret_cast<uint8_t*>(out + ((size_t)(seq_begin + q0) * H + head) * kDh);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * 256 + tid, row = idx >> 3, ch = idx & 7;
        if (row < rows_ok)
          *reinterpret_cast<uint4*>(obase + (size_t)row * H * kDh * 2 + ch * 16) =
              *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
    if (row_ok && grp == 0) lse[(size_t)head * T + seq_begin + q_row] = (m_run + log2f(l_tot)) * kLn2;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

// ============================================================================================== backward
// delta[h, t] = sum_d dO[t,h,d] * O[t,h,d]; 8 threads per (t, h) row of 64, 16-byte loads
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, int T, int H) {
  const int64_t gid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int64_t row = gid >> 3;  // (t, h) flattened: element offset row * 64
  const int sub = (int)(gid & 7);
  float s = 0.f;
  if (row < (int64_t)T * H) {
    const uint4 a = *reinterpret_cast<const uint4*>(o + row * kDh + sub * 8);
    const uint4 b = *reinterpret_cast<const uint4*>(dout + row * kDh + sub * 8);
    const __nv_bfloat162* pa = reinterpret_cast<const __nv_bfloat162*>(&a);
    const __nv_bfloat162* pb = reinterpret_cast<const __nv_bfloat162*>(&b);
#pragma unroll
    for (int i = 0; i < 4; ++i) s += __low2float(pa[i]) * __low2float(pb[i]) + __high2float(pa[i]) * __high2float(pb[i]);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (sub == 0 && row < (int64_t)T * H) {
    const int t = (int)(row / H), h = (int)(row % H);
    delta[(size_t)h * T + t] = s;
  }
}

constexpr int kBwdThreads = 384;  // 4 control warps + 2 x 4 worker warps (each group owns 64 of the 128 key columns)
struct BwdSmem {
  static constexpr int kTile = 128 * kDh * 2;   // 16 KB
  static constexpr int kK = 0;                  // K_j  (B of S, B of dQ as MN-major)
  static constexpr int kV = kK + kTile;         // V_j  (B of dP)
  static constexpr int kQ = kV + kTile;         // 2 stages: Q_i (A of S, B of dK as MN-major)
  static constexpr int kDO = kQ + 2 * kTile;    // 2 stages: dO_i (A of dP, B of dV as MN-major)
  static constexpr int kP = kDO + 2 * kTile;    // P  [128 q x 128 keys] bf16 (A of dV, MN-major)
  static constexpr int kDS = kP + 32768;        // dS [128 q x 128 keys] bf16 (A of dK MN-major, A of dQ K-major)
  static constexpr int kDQ = kDS + 32768;       // fp32 staging for the dQ reduce-add: 2 x [128 x 32] (128 B rows)
  static constexpr int kBars = kDQ + 2 * 16384;
  static constexpr int kTotal = kBars + 256 + 1024;
};

// TMEM columns: S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ [384,448)
__global__ void __launch_bounds__(kBwdThreads, 1)
attn_bwd_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                const __grid_constant__ CUtensorMap tmDQ, const int* __restrict__ cu_seqlens,
                const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T,
                int H, float softmax_scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBars);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [2]  (Q_i and dO_i of a stage)
  uint64_t* q_empty = bars + 3;    // [2]
  uint64_t* sdp_full = bars + 5;   // [1]  S and dP of the current tile are in TMEM
  uint64_t* pds_full = bars + 6;   // [1]  P and dS are in smem (128 arrivals)
  uint64_t* dq_full = bars + 7;    // [1]  dQ partial of the current tile is in TMEM
  uint64_t* dq_free = bars + 8;    // [1]  dQ TMEM drained by the epilogue warps (128 arrivals)
  uint64_t* acc_full = bars + 9;   // [1]  dK / dV complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 10);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int k0 = blockIdx.x * 128;
  if (k0 >= len) return;
  const int nq = (len + 127) / 128;
  const float scale2 = softmax_scale * kLog2e;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmQKV);
    tma_prefetch_desc(&tmDO);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 1 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(sdp_full, 1);
    mbar_init(pds_full, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 256);
    mbar_init(acc_full, 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;
  const int col_o = head * kDh;

  if (warp == 0) {
    if (lane == 0) {
      mbar_arrive_expect_tx(kv_full, 2 * BwdSmem::kTile);
      tma_load_2d(smem + BwdSmem::kK, &tmQKV, kv_full, col_k, seq_begin + k0);
      tma_load_2d(smem + BwdSmem::kV, &tmQKV, kv_full, col_v, seq_begin + k0);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
        mbar_arrive_expect_tx(&q_full[st], 2 * BwdSmem::kTile);
        tma_load_2d(smem + BwdSmem::kQ + st * BwdSmem::kTile, &tmQKV, &q_full[st], col_q, seq_begin + i * 128);
        tma_load_2d(smem + BwdSmem::kDO + st * BwdSmem::kTile, &tmDO, &q_full[st], col_o, seq_begin + i * 128);
      }
    }
  } else if (warp == 1) {
    if (lane == 0) {
      constexpr uint32_t id_kk = make_idesc_bf16(128, 128, 0, 0);  // S, dP: A K-major, B K-major, N = 128
      constexpr uint32_t id_mm = make_idesc_bf16(128, 64, 1, 1);   // dV, dK: A MN-major (P^T / dS^T), B MN-major, N = 64
      constexpr uint32_t id_km = make_idesc_bf16(128, 64, 0, 1);   // dQ: A K-major (dS), B MN-major (K_j), N = 64
      const uint32_t k_addr = smem_u32(smem + BwdSmem::kK), v_addr = smem_u32(smem + BwdSmem::kV);
      const uint32_t q_addr = smem_u32(smem + BwdSmem::kQ), do_addr = smem_u32(smem + BwdSmem::kDO);
      const uint32_t p_addr = smem_u32(smem + BwdSmem::kP), ds_addr = smem_u32(smem + BwdSmem::kDS);
      mbar_wait(kv_full, 0);
      for (int i = 0; i < nq; ++i) {
        const int st = i & 1;
        mbar_wait(&q_full[st], (i >> 1) & 1);
        tc_fence_after();
        // S = Q_i K_j^T ; dP = dO_i V_j^T      (K = Dh = 64: 4 k-steps each)
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem_base + 0, make_smem_desc_sw128(q_addr + st * BwdSmem::kTile + kk * 32, 0, 1024),
                      make_smem_desc_sw128(k_addr + kk * 32, 0, 1024), id_kk, kk > 0 ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem_base + 128, make_smem_desc_sw128(do_addr + st * BwdSmem::kTile + kk * 32, 0, 1024),
                      make_smem_desc_sw128(v_addr + kk * 32, 0, 1024), id_kk, kk > 0 ? 1u : 0u);
        umma_commit(sdp_full);
        // wait for P / dS in smem
        mbar_wait(pds_full, i & 1);
        tc_fence_after();
        // dV += P^T dO_i ; dK += dS^T Q_i   (contraction over the 128 query rows: 8 k-steps of 16 rows = +2048 B;
        //                                      A atoms (64 keys each) are 16 KB apart => LBO = 16384)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ss(tmem_base + 256, make_smem_desc_sw128(p_addr + kk * 2048, 16384, 1024),
                      make_smem_desc_sw128(do_addr + st * BwdSmem::kTile + kk * 2048, 8192, 1024), id_mm,
                      (i > 0 || kk > 0) ? 1u : 0u);
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ss(tmem_base + 320, make_smem_desc_sw128(ds_addr + kk * 2048, 16384, 1024),
                      make_smem_desc_sw128(q_addr + st * BwdSmem::kTile + kk * 2048, 8192, 1024), id_mm,
                      (i > 0 || kk > 0) ? 1u : 0u);
        // dQ_i(partial) = dS K_j   (contraction over 128 keys: A K-major two 64-key blocks, B = K_j MN-major)
        mbar_wait(dq_free, (i & 1) ^ 1);
        tc_fence_after();
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ss(tmem_base + 384, make_smem_desc_sw128(ds_addr + (kk >> 2) * 16384 + (kk & 3) * 32, 0, 1024),
                      make_smem_desc_sw128(k_addr + kk * 2048, 8192, 1024), id_km, kk > 0 ? 1u : 0u);
        umma_commit(dq_full);
        umma_commit(&q_empty[st]);
      }
      umma_commit(acc_full);
    }
  } else if (warp >= 4) {
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;  // column group: keys [grp*64, grp*64+64) of the tile; dQ columns [grp*32, +32)
    const int r = ew * 32 + lane;     // query row within the tile (S/dP/dQ) or key row within the tile (dK/dV)
    const int etid = (threadIdx.x - 128) & 127;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const int kv_valid = min(128, len - k0);
    const bool full = kv_valid == 128;
    uint8_t* p_smem = smem + BwdSmem::kP + grp * 16384;    // this group's 64-key block of P / dS
    uint8_t* ds_smem = smem + BwdSmem::kDS + grp * 16384;
    uint8_t* dq_smem = smem + BwdSmem::kDQ + grp * 16384;
    for (int i = 0; i < nq; ++i) {
      const int q_row = i * 128 + r;
      const bool row_ok = q_row < len;
      const float lse2 = row_ok ? lse[(size_t)head * T + seq_begin + q_row] * kLog2e : INFINITY;  // +inf => P = 0
      const float dl = row_ok ? delta[(size_t)head * T + seq_begin + q_row] * softmax_scale : 0.f;  // pre-scaled
      mbar_wait(sdp_full, i & 1);
      tc_fence_after();
#pragma unroll 1
      for (int c = 0; c < 2; ++c) {
        uint32_t vs[32], vd[32];
        tmem_ld_32x32(tmem_base + lane_base + grp * 64 + c * 32, vs);
        tmem_ld_32x32(tmem_base + lane_base + 128 + grp * 64 + c * 32, vd);
        tmem_ld_wait();
        if (!full) {
#pragma unroll
          for (int t = 0; t < 32; ++t)
            if (grp * 64 + c * 32 + t >= kv_valid) vs[t] = 0xff800000u;  // -inf => P = 0
        }
        uint8_t* dp = p_smem + r * 128;
        uint8_t* dd = ds_smem + r * 128;
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          float p[8], ds[8];
#pragma unroll
          for (int t = 0; t < 8; ++t) {
            p[t] = fast_exp2(fmaf(__uint_as_float(vs[8 * q + t]), scale2, -lse2));
            ds[t] = p[t] * fmaf(__uint_as_float(vd[8 * q + t]), softmax_scale, -dl);
          }
          uint4 w, x;
          w.x = pack_bf16x2(p[0], p[1]);
          w.y = pack_bf16x2(p[2], p[3]);
          w.z = pack_bf16x2(p[4], p[5]);
          w.w = pack_bf16x2(p[6], p[7]);
          x.x = pack_bf16x2(ds[0], ds[1]);
          x.y = pack_bf16x2(ds[2], ds[3]);
          x.z = pack_bf16x2(ds[4], ds[5]);
          x.w = pack_bf16x2(ds[6], ds[7]);
          const int chunk = c * 4 + q;
          *reinterpret_cast<uint4*>(dp + ((chunk ^ (r & 7)) << 4)) = w;
          *reinterpret_cast<uint4*>(dd + ((chunk ^ (r & 7)) << 4)) = x;
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(pds_full);
      // drain this tile's dQ partial (this group's 32 columns): TMEM -> fp32 smem stage -> TMA reduce-add into dq_acc
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + lane_base + 384 + grp * 32, v);
        tmem_ld_wait();
        if (etid == 0) tma_store_wait_read<0>();
        named_bar_sync(1 + grp, 128);
        uint8_t* dst = dq_smem + r * 128;
#pragma unroll
        for (int q = 0; q < 8; ++q) {
          uint4 w = make_uint4(row_ok ? v[4 * q] : 0u, row_ok ? v[4 * q + 1] : 0u, row_ok ? v[4 * q + 2] : 0u,
                               row_ok ? v[4 * q + 3] : 0u);
          *reinterpret_cast<uint4*>(dst + ((q ^ (r & 7)) << 4)) = w;
        }
        fence_proxy_async_smem();
        named_bar_sync(1 + grp, 128);
        if (etid == 0) {
          tma_reduce_add_2d(&tmDQ, dq_smem, col_o + grp * 32, seq_begin + i * 128);
          tma_store_commit();
        }
      }
      tc_fence_before();
      mbar_arrive(dq_free);
    }
    // dV (group 0) / dK (group 1): TMEM -> bf16 -> dqkv rows of this key tile
    mbar_wait(acc_full, 0);
    tc_fence_after();
    const int k_row = k0 + r;
    const bool krow_ok = k_row < len;
    __nv_bfloat16* dst = dqkv + ((size_t)(seq_begin + k_row) * 3 + (grp == 0 ? 2 : 1)) * H * kDh + (size_t)head * kDh;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + lane_base + 256 + grp * 64 + c * 32, v);
      tmem_ld_wait();
      if (krow_ok) {
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint4 w;
          w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
          w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
          w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
          w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
          *reinterpret_cast<uint4*>(dst + c * 32 + q * 8) = w;
        }
      }
    }
    if (etid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// ---------------------------------------------------------------------------------------------- backward, pipelined
// Same tiling and smem/TMEM layout as attn_bwd_kernel, but the five contractions of consecutive query tiles overlap with
// the exponentials:  the MMA thread issues S(i+1) as soon as P(i) has left the registers (S columns are free again) and
// dP(i+1) as soon as dS(i) is in smem, so the tensor core runs dV(i) / dK(i) / dQ(i) while the 8 worker warps compute
// the next tile's exponentials; a separate 4-warp group drains the dQ partials (TMEM -> smem -> TMA reduce-add), so the
// workers never wait for it.  16 warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 workers, 12-15 dQ drain.
constexpr int kBwd2Threads = 512;

__global__ void __launch_bounds__(kBwd2Threads, 1)
attn_bwd2_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const int* __restrict__ cu_seqlens,
                 const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T,
                 int H, float softmax_scale, int ablate) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + BwdSmem::kBars);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [2]  Q_i and dO_i of a stage landed
  uint64_t* q_empty = bars + 3;    // [2]  every MMA reading the stage has completed
  uint64_t* s_full = bars + 5;     // S(i) in TMEM
  uint64_t* dp_full = bars + 6;    // dP(i) in TMEM
  uint64_t* p_ready = bars + 7;    // P(i) in smem, S columns free (256 arrivals)
  uint64_t* ds_ready = bars + 8;   // dS(i) in smem, dP columns free (256 arrivals)
  uint64_t* p_free = bars + 9;     // dV(i) has finished reading P(i)
  uint64_t* dq_full = bars + 10;   // dQ(i) partial in TMEM; also: dK(i), dQ(i) have finished reading dS(i)
  uint64_t* dq_free = bars + 11;   // dQ TMEM columns drained (128 arrivals)
  uint64_t* acc_full = bars + 12;  // dK / dV complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 13);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int k0 = blockIdx.x * 128;
  if (k0 >= len) return;
  const int nq = (len + 127) / 128;
  const float scale2 = softmax_scale * kLog2e;
  long long* tr = g_attn_trace;
  if (tr != nullptr) {
    tr += ((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 64;
    if (threadIdx.x == 0) {
      uint32_t smid;
      unsigned long long gt;
      asm volatile("mov.u32 %0, %%smid;" : "=r"(smid));
      asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(gt));
      tr[0] = clock64();
      tr[1] = smid;
      tr[2] = (long long)gt;
    }
  }

  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;
  const int col_o = head * kDh;
  if (warp == 0 && lane == 0) {
    // the producer lane initialises the barriers itself and starts the first loads before the CTA-wide sync, so the TMA
    // round trip (~2000 clk) overlaps the TMEM allocation and the rest of the set-up
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(ds_ready, 256);
    mbar_init(p_free, 1);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 128);
    mbar_init(acc_full, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(kv_full, 2 * BwdSmem::kTile);
    tma_load_2d(smem + BwdSmem::kK, &tmQKV, kv_full, col_k, seq_begin + k0);
    tma_load_2d(smem + BwdSmem::kV, &tmQKV, kv_full, col_v, seq_begin + k0);
    mbar_arrive_expect_tx(&q_full[0], 2 * BwdSmem::kTile);
    tma_load_2d(smem + BwdSmem::kQ, &tmQKV, &q_full[0], col_q, seq_begin);
    tma_load_2d(smem + BwdSmem::kDO, &tmDO, &q_full[0], col_o, seq_begin);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;  // S [0,128)  dP [128,256)  dV [256,320)  dK [320,384)  dQ [384,448)
  if (threadIdx.x == 0) trace_put(tr, 3);

  // TMA and MMA warps run converged and elect one lane around the asynchronous instructions (see attn_fwd2_kernel)
  if (warp == 0) {
    for (int i = 1; i < nq; ++i) {  // K, V and the first query tile were issued during set-up
      const int st = i & 1;
      mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[st], 2 * BwdSmem::kTile);
        tma_load_2d(smem + BwdSmem::kQ + st * BwdSmem::kTile, &tmQKV, &q_full[st], col_q, seq_begin + i * 128);
        tma_load_2d(smem + BwdSmem::kDO + st * BwdSmem::kTile, &tmDO, &q_full[st], col_o, seq_begin + i * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t id_kk = make_idesc_bf16(128, 128, 0, 0);  // S, dP: A K-major, B K-major, N = 128
    constexpr uint32_t id_mm = make_idesc_bf16(128, 64, 1, 1);   // dV, dK: A MN-major (P^T / dS^T), B MN-major, N = 64
    constexpr uint32_t id_km = make_idesc_bf16(128, 64, 0, 1);   // dQ: A K-major (dS), B MN-major (K_j), N = 64
    // base descriptors; descriptor of (base + off) = base descriptor + (off >> 4)
    const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kK), 0, 1024);          // K_j  K-major (B of S)
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kV), 0, 1024);          // V_j  K-major (B of dP)
    const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kQ), 0, 1024);          // Q_i  K-major (A of S)
    const uint64_t dod = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kDO), 0, 1024);        // dO_i K-major (A of dP)
    const uint64_t km = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kK), 8192, 1024);       // K_j  MN-major (B of dQ)
    const uint64_t qm = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kQ), 8192, 1024);       // Q_i  MN-major (B of dK)
    const uint64_t dom = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kDO), 8192, 1024);     // dO_i MN-major (B of dV)
    const uint64_t pm = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kP), 16384, 1024);      // P    MN-major (A of dV)
    const uint64_t dsm = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kDS), 16384, 1024);    // dS   MN-major (A of dK)
    const uint64_t dsk = make_smem_desc_sw128(smem_u32(smem + BwdSmem::kDS), 0, 1024);        // dS   K-major  (A of dQ)
    const int nmma = (ablate & 16) ? 1 : 8;
    mbar_wait(kv_full, 0);
    mbar_wait(&q_full[0], 0);
    tc_fence_after();
    if (lane == 0) trace_put(tr, 4);
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_base + 0, qd + ((kk * 32) >> 4), kd + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
      umma_commit(s_full);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_base + 128, dod + ((kk * 32) >> 4), vd + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
      umma_commit(dp_full);
    }
    __syncwarp();
    for (int i0 = 0; i0 < nq; i0 += 2) {
#pragma unroll
      for (int st = 0; st < 2; ++st) {  // st = i & 1 is a compile-time constant after unrolling
        const int i = i0 + st;
        if (i < nq) {                   // warp-uniform
          constexpr int kT = BwdSmem::kTile;
          const int ns = st ^ 1;
          const bool more = i + 1 < nq;
          mbar_wait(p_ready, i & 1);
          if (more) mbar_wait(&q_full[ns], ((i + 1) >> 1) & 1);
          tc_fence_after();
          if (lane == 0 && i < 4) trace_put(tr, 40 + 2 * i);
          if (elect_one()) {
            if (ablate & 32) {  // interleave the k-steps of the two independent accumulation chains
#pragma unroll
              for (int kk = 0; kk < 8; ++kk) {
                if (more && (kk & 1) == 0)
                  umma_f16_ss(tmem_base + 0, qd + ((ns * kT + (kk >> 1) * 32) >> 4), kd + (((kk >> 1) * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
                umma_f16_ss(tmem_base + 256, pm + ((kk * 2048) >> 4), dom + ((st * kT + kk * 2048) >> 4), id_mm,
                            (i > 0 || kk > 0) ? 1u : 0u);
                if (more && kk == 6) umma_commit(s_full);
              }
              umma_commit(p_free);
            } else {
            if (more) {  // S(i+1) = Q_{i+1} K_j^T: the score columns are free again
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ss(tmem_base + 0, qd + ((ns * kT + kk * 32) >> 4), kd + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
              umma_commit(s_full);
            }
            // dV += P^T dO_i  (contraction over the 128 query rows: 8 k-steps of 16 rows = +2048 B; A atoms 16 KB apart)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (kk < nmma)
                umma_f16_ss(tmem_base + 256, pm + ((kk * 2048) >> 4), dom + ((st * kT + kk * 2048) >> 4), id_mm,
                            (i > 0 || kk > 0) ? 1u : 0u);
            umma_commit(p_free);
            }
          }
          __syncwarp();
          mbar_wait(ds_ready, i & 1);
          if (i > 0) mbar_wait(dq_free, (i - 1) & 1);
          tc_fence_after();
          if (lane == 0 && i < 4) trace_put(tr, 41 + 2 * i);
          if (elect_one()) {
          if (ablate & 32) {  // interleaved: dP(i+1) k, dK(i) 2k, dQ(i) 2k, dK(i) 2k+1, dQ(i) 2k+1
#pragma unroll
            for (int kk = 0; kk < 8; ++kk) {
              if (more && (kk & 1) == 0)
                umma_f16_ss(tmem_base + 128, dod + ((ns * kT + (kk >> 1) * 32) >> 4), vd + (((kk >> 1) * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
              umma_f16_ss(tmem_base + 320, dsm + ((kk * 2048) >> 4), qm + ((st * kT + kk * 2048) >> 4), id_mm, (i > 0 || kk > 0) ? 1u : 0u);
              umma_f16_ss(tmem_base + 384, dsk + (((kk >> 2) * 16384 + (kk & 3) * 32) >> 4), km + ((kk * 2048) >> 4), id_km, kk > 0 ? 1u : 0u);
              if (more && kk == 6) umma_commit(dp_full);
            }
            umma_commit(dq_full);
            umma_commit(&q_empty[st]);
          } else {
            if (more) {  // dP(i+1) = dO_{i+1} V_j^T
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ss(tmem_base + 128, dod + ((ns * kT + kk * 32) >> 4), vd + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
              umma_commit(dp_full);
            }
            // dK += dS^T Q_i
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (kk < nmma)
                umma_f16_ss(tmem_base + 320, dsm + ((kk * 2048) >> 4), qm + ((st * kT + kk * 2048) >> 4), id_mm,
                            (i > 0 || kk > 0) ? 1u : 0u);
            // dQ_i(partial) = dS K_j   (contraction over 128 keys: A K-major two 64-key blocks, B = K_j MN-major)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              if (kk < nmma)
                umma_f16_ss(tmem_base + 384, dsk + (((kk >> 2) * 16384 + (kk & 3) * 32) >> 4), km + ((kk * 2048) >> 4), id_km,
                            kk > 0 ? 1u : 0u);
            umma_commit(dq_full);
            umma_commit(&q_empty[st]);
          }
          }
          __syncwarp();
        }
      }
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
    if (lane == 0) trace_put(tr, 5);
  } else if (warp >= 4 && warp < 12) {
    // ---------------------------------------------------------------- workers: two threads per query row
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;  // keys [grp*64, grp*64+64) of the tile
    const int r = ew * 32 + lane;     // query row within the tile (S/dP) or key row within the tile (dK/dV epilogue)
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const int kv_valid = min(128, len - k0) - grp * 64;  // valid keys among this thread's 64
    const bool full = kv_valid >= 64;
    uint8_t* dp = smem + BwdSmem::kP + grp * 16384 + r * 128;    // this thread's 128-byte row of P
    uint8_t* dd = smem + BwdSmem::kDS + grp * 16384 + r * 128;   // ... and of dS
    const float2 sc2 = make_float2(scale2, scale2), ss2 = make_float2(softmax_scale, softmax_scale);
    const float* lse_h = lse + (size_t)head * T + seq_begin;
    const float* dl_h = delta + (size_t)head * T + seq_begin;
    float lse_n = (r < len) ? lse_h[r] : INFINITY;
    float dl_n = (r < len) ? dl_h[r] : 0.f;
    for (int i = 0; i < nq; ++i) {
      const float nlse2 = -lse_n * kLog2e;           // -inf for rows past the sequence end => P = 0
      const float ndl = -dl_n * softmax_scale;
      {
        const int nr = (i + 1) * 128 + r;            // prefetch the next tile's row statistics
        lse_n = (nr < len) ? lse_h[nr] : INFINITY;
        dl_n = (nr < len) ? dl_h[nr] : 0.f;
      }
      const float2 nl2 = make_float2(nlse2, nlse2), nd2 = make_float2(ndl, ndl);
      // ---- X: P = exp2(S * scale2 - lse2) -> bf16 registers -> smem
      uint32_t pp[32];
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      if (threadIdx.x == 128 && i < 4) trace_put(tr, 8 + 4 * i);
      {
        uint32_t va[32], vb[32];
        if (!(ablate & 8)) {
          tmem_ld_32x32(tmem_base + lane_base + grp * 64, va);
          tmem_ld_32x32(tmem_base + lane_base + grp * 64 + 32, vb);
          tmem_ld_wait();
        } else {  // timing ablation: no score read (results are wrong)
#pragma unroll
          for (int t = 0; t < 32; ++t) va[t] = vb[t] = __float_as_uint(0.01f * (float)(t + lane));
        }
        if (!full) {
#pragma unroll
          for (int t = 0; t < 32; ++t) {
            if (t >= kv_valid) va[t] = 0xff800000u;  // -inf => P = 0
            if (32 + t >= kv_valid) vb[t] = 0xff800000u;
          }
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float2 xa = ffma2(make_float2(__uint_as_float(va[2 * t]), __uint_as_float(va[2 * t + 1])), sc2, nl2);
          pp[t] = (ablate & 2) ? pack_bf16x2(xa.x, xa.y) : pack_bf16x2(fast_exp2(xa.x), fast_exp2(xa.y));
        }
#pragma unroll
        for (int t = 0; t < 16; ++t) {
          const float2 xb = ffma2(make_float2(__uint_as_float(vb[2 * t]), __uint_as_float(vb[2 * t + 1])), sc2, nl2);
          pp[16 + t] = (ablate & 2) ? pack_bf16x2(xb.x, xb.y) : pack_bf16x2(fast_exp2(xb.x), fast_exp2(xb.y));
        }
      }
      if (i > 0) mbar_wait(p_free, (i - 1) & 1);  // dV(i-1) has finished reading the P buffer
      if (!(ablate & 4)) {
#pragma unroll
        for (int ch = 0; ch < 8; ++ch)
          *reinterpret_cast<uint4*>(dp + ((ch ^ (r & 7)) << 4)) = make_uint4(pp[4 * ch], pp[4 * ch + 1], pp[4 * ch + 2], pp[4 * ch + 3]);
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(p_ready);
      if (threadIdx.x == 128 && i < 4) trace_put(tr, 9 + 4 * i);
      // ---- Y: dS = P * (dP * scale - delta * scale) -> smem
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
      if (i > 0) mbar_wait(dq_full, (i - 1) & 1);  // dK(i-1), dQ(i-1) have finished reading the dS buffer
      if (threadIdx.x == 128 && i < 4) trace_put(tr, 10 + 4 * i);
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t vd[32];
        if (!(ablate & 8)) {
          tmem_ld_32x32(tmem_base + lane_base + 128 + grp * 64 + c * 32, vd);
          tmem_ld_wait();
        } else {
#pragma unroll
          for (int t = 0; t < 32; ++t) vd[t] = __float_as_uint(0.01f * (float)(t + lane));
        }
#pragma unroll
        for (int q = 0; q < 4; ++q) {
          uint32_t w[4];
#pragma unroll
          for (int t = 0; t < 4; ++t) {
            const float2 pv = unpack_bf16x2(pp[c * 16 + q * 4 + t]);
            const float2 g = ffma2(make_float2(__uint_as_float(vd[8 * q + 2 * t]), __uint_as_float(vd[8 * q + 2 * t + 1])), ss2, nd2);
            const float2 ds = fmul2(pv, g);
            w[t] = pack_bf16x2(ds.x, ds.y);
          }
          const int ch = c * 4 + q;
          if (!(ablate & 4)) *reinterpret_cast<uint4*>(dd + ((ch ^ (r & 7)) << 4)) = make_uint4(w[0], w[1], w[2], w[3]);
        }
      }
      fence_proxy_async_smem();
      tc_fence_before();
      mbar_arrive(ds_ready);
      if (threadIdx.x == 128 && i < 4) trace_put(tr, 11 + 4 * i);
    }
    // dV (group 0) / dK (group 1): TMEM -> bf16 -> dqkv rows of this key tile
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (threadIdx.x == 128) trace_put(tr, 24);
    // dV (group 0) / dK (group 1) -> bf16 -> the (dead) P buffer, one swizzled 128-byte row per thread; then each group
    // copies its tile out with row-contiguous 16-byte stores (see attn_fwd2_kernel)
    uint8_t* stg = smem + BwdSmem::kP + grp * 16384;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + lane_base + 256 + grp * 64 + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
        w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
        w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
        w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
        *reinterpret_cast<uint4*>(stg + r * 128 + (((c * 4 + q) ^ (r & 7)) << 4)) = w;
      }
    }
    named_bar_sync(2 + grp, 128);
    {
      const int tid = (threadIdx.x - 128) & 127;
      const int rows_ok = min(128, len - k0);
      uint8_t* obase = reinterpret_cast<uint8_t*>(dqkv + ((size_t)(seq_begin + k0) * 3 + (grp == 0 ? 2 : 1)) * H * kDh + (size_t)head * kDh);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + tid, row = idx >> 3, ch = idx & 7;
        if (row < rows_ok)
          *reinterpret_cast<uint4*>(obase + (size_t)row * 3 * H * kDh * 2 + ch * 16) =
              *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
  } else if (warp >= 12) {
    // ---------------------------------------------------------------- dQ drain: one thread per query row, 64 columns
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const int etid = threadIdx.x - 384;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    uint8_t* stage = smem + BwdSmem::kDQ;  // two [128 x 32] fp32 boxes (128-byte rows, 128B swizzle)
    for (int i = 0; i < nq; ++i) {
      const bool row_ok = i * 128 + r < len;
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      if (etid == 0 && i < 4) trace_put(tr, 28 + 3 * i);
      uint32_t va[32], vb[32];
      tmem_ld_32x32(tmem_base + lane_base + 384, va);
      tmem_ld_32x32(tmem_base + lane_base + 384 + 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);  // the next tile's dQ MMA may overwrite the columns
      if (etid == 0) tma_store_wait_read<0>();  // the previous reduce-add has finished reading the stage
      named_bar_sync(1, 128);
      if (etid == 0 && i < 4) trace_put(tr, 29 + 3 * i);
      uint8_t* d0 = stage + r * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        *reinterpret_cast<uint4*>(d0 + ((q ^ (r & 7)) << 4)) =
            make_uint4(row_ok ? va[4 * q] : 0u, row_ok ? va[4 * q + 1] : 0u, row_ok ? va[4 * q + 2] : 0u, row_ok ? va[4 * q + 3] : 0u);
        *reinterpret_cast<uint4*>(d0 + 16384 + ((q ^ (r & 7)) << 4)) =
            make_uint4(row_ok ? vb[4 * q] : 0u, row_ok ? vb[4 * q + 1] : 0u, row_ok ? vb[4 * q + 2] : 0u, row_ok ? vb[4 * q + 3] : 0u);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (etid == 0 && !(ablate & 1)) {  // (ablation bit 1: the dQ partial is dropped)
        tma_reduce_add_2d(&tmDQ, stage, col_o, seq_begin + i * 128);
        tma_reduce_add_2d(&tmDQ, stage + 16384, col_o + 32, seq_begin + i * 128);
        tma_store_commit();
      }
      if (etid == 0 && i < 4) trace_put(tr, 30 + 3 * i);
    }
    if (etid == 0) tma_store_wait_read<0>();  // the stage must outlive the TMA reads; the adds complete by kernel end
    if (etid == 0) trace_put(tr, 48);
  }
  if (threadIdx.x == 128) trace_put(tr, 25);

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
  if (threadIdx.x == 0) trace_put(tr, 50);
}

// ---------------------------------------------------------------------------------------------- backward, transposed scores
// EXPERIMENTAL (CX_ATTN_BWD=3; written after the last GPU minute of round 1: it compiles, it has NOT run on hardware yet).
// attn_bwd2_kernel is bound by the shared-memory port (368 KB per 128 x 128 tile: MMA operands + P / dS stores + dQ staging).
// Here the scores are computed TRANSPOSED, S^T = K_j Q_i^T and dP^T = V_j dO_i^T (lane = key, column = query), so that P^T and
// dS^T -- the A operands of dV += P^T dO and dK += dS^T Q -- are written back into tensor memory (bf16 pairs, over the
// thread's own first 32 score columns) and those two contractions read only their B operand from shared memory (a TS MMA
// with N = 64 runs at 32 clk instead of 48, profiles/r01_ubench_tmem_mma.txt).  dS still goes to shared memory once, as the
// (MN-major) A operand of dQ = dS K.  Shared-memory traffic per tile: 240 KB.  The row statistics are per COLUMN now: the
// workers publish -lse*log2(e) and -delta*scale of the query tile in shared memory and read them with broadcast loads.
// TMEM: S^T [0,128)  dP^T [128,256)  dV [256,320)  dK [320,384)  dQ [384,448);  P^T over S^T columns [0,32) + [64,96),
// dS^T over dP^T columns [128,160) + [192,224) (each thread overwrites only columns it has already loaded itself).
// 16 warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-11 workers (two threads per key row, 64 query columns each), 12-15 dQ drain.
struct Bwd3Smem {
  static constexpr int kTile = 128 * kDh * 2;   // 16 KB
  static constexpr int kK = 0;                  // K_j  (A of S^T, B of dQ as MN-major)
  static constexpr int kV = kK + kTile;         // V_j  (A of dP^T)
  static constexpr int kQ = kV + kTile;         // 2 stages: Q_i (B of S^T, B of dK as MN-major)
  static constexpr int kDO = kQ + 2 * kTile;    // 2 stages: dO_i (B of dP^T, B of dV as MN-major)
  static constexpr int kDS = kDO + 2 * kTile;   // dS^T [128 keys x 128 q] bf16: 2 blocks (q halves) x 128 rows x 128 B
  static constexpr int kDQ = kDS + 32768;       // fp32 staging for the dQ reduce-add: 2 x [128 x 32] (128 B rows)
  static constexpr int kStat = kDQ + 2 * 16384; // 2 buffers x { -lse*log2e [128], -delta*scale [128] } fp32
  static constexpr int kBars = kStat + 2048;
  static constexpr int kTotal = kBars + 256 + 1024;
};

__global__ void __launch_bounds__(kBwd2Threads, 1)
attn_bwd3_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const int* __restrict__ cu_seqlens,
                 const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T,
                 int H, float softmax_scale) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Bwd3Smem::kBars);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [2]  Q_i and dO_i of a stage landed
  uint64_t* q_empty = bars + 3;    // [2]  every MMA reading the stage has completed
  uint64_t* s_full = bars + 5;     // S^T(i) in TMEM
  uint64_t* dp_full = bars + 6;    // dP^T(i) in TMEM
  uint64_t* p_ready = bars + 7;    // P^T(i) in TMEM, every worker has loaded its S^T(i) columns (256 arrivals)
  uint64_t* ds_ready = bars + 8;   // dS^T(i) in TMEM and in smem, every worker has loaded its dP^T(i) columns (256 arrivals)
  uint64_t* dq_full = bars + 9;    // dQ(i) partial in TMEM; also: dQ(i) has finished reading dS(i) from smem
  uint64_t* dq_free = bars + 10;   // dQ TMEM columns drained (128 arrivals)
  uint64_t* acc_full = bars + 11;  // dK / dV complete
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 12);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int k0 = blockIdx.x * 128;
  if (k0 >= len) return;
  const int nq = (len + 127) / 128;
  const float scale2 = softmax_scale * kLog2e;

  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;
  const int col_o = head * kDh;
  if (warp == 0 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, 256);
    mbar_init(ds_ready, 256);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 128);
    mbar_init(acc_full, 1);
    fence_barrier_init();
    mbar_arrive_expect_tx(kv_full, 2 * Bwd3Smem::kTile);
    tma_load_2d(smem + Bwd3Smem::kK, &tmQKV, kv_full, col_k, seq_begin + k0);
    tma_load_2d(smem + Bwd3Smem::kV, &tmQKV, kv_full, col_v, seq_begin + k0);
    mbar_arrive_expect_tx(&q_full[0], 2 * Bwd3Smem::kTile);
    tma_load_2d(smem + Bwd3Smem::kQ, &tmQKV, &q_full[0], col_q, seq_begin);
    tma_load_2d(smem + Bwd3Smem::kDO, &tmDO, &q_full[0], col_o, seq_begin);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    for (int i = 1; i < nq; ++i) {  // K, V and the first query tile were issued during set-up
      const int st = i & 1;
      mbar_wait(&q_empty[st], ((i >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[st], 2 * Bwd3Smem::kTile);
        tma_load_2d(smem + Bwd3Smem::kQ + st * Bwd3Smem::kTile, &tmQKV, &q_full[st], col_q, seq_begin + i * 128);
        tma_load_2d(smem + Bwd3Smem::kDO + st * Bwd3Smem::kTile, &tmDO, &q_full[st], col_o, seq_begin + i * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t id_kk = make_idesc_bf16(128, 128, 0, 0);  // S^T, dP^T: A (K_j / V_j) K-major, B (Q_i / dO_i) K-major
    constexpr uint32_t id_tm = make_idesc_bf16(128, 64, 0, 1);   // dV, dK: A from TMEM, B (dO_i / Q_i) MN-major, N = 64
    constexpr uint32_t id_mm = make_idesc_bf16(128, 64, 1, 1);   // dQ: A (dS^T in smem) MN-major, B (K_j) MN-major, N = 64
    const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kK), 0, 1024);        // K_j  K-major (A of S^T)
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kV), 0, 1024);        // V_j  K-major (A of dP^T)
    const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kQ), 0, 1024);        // Q_i  K-major (B of S^T)
    const uint64_t dod = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kDO), 0, 1024);      // dO_i K-major (B of dP^T)
    const uint64_t km = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kK), 8192, 1024);     // K_j  MN-major (B of dQ)
    const uint64_t qm = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kQ), 8192, 1024);     // Q_i  MN-major (B of dK)
    const uint64_t dom = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kDO), 8192, 1024);   // dO_i MN-major (B of dV)
    const uint64_t dsm = make_smem_desc_sw128(smem_u32(smem + Bwd3Smem::kDS), 16384, 1024);  // dS^T MN-major (A of dQ)
    mbar_wait(kv_full, 0);
    mbar_wait(&q_full[0], 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_base + 0, kd + ((kk * 32) >> 4), qd + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
      umma_commit(s_full);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_base + 128, vd + ((kk * 32) >> 4), dod + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
      umma_commit(dp_full);
    }
    __syncwarp();
    for (int i0 = 0; i0 < nq; i0 += 2) {
#pragma unroll
      for (int st = 0; st < 2; ++st) {  // st = i & 1 is a compile-time constant after unrolling
        const int i = i0 + st;
        if (i < nq) {                   // warp-uniform
          constexpr int kT = Bwd3Smem::kTile;
          const int ns = st ^ 1;
          const bool more = i + 1 < nq;
          mbar_wait(p_ready, i & 1);
          if (more) mbar_wait(&q_full[ns], ((i + 1) >> 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            // dV += P^T dO_i: A = P^T from TMEM (16 queries = 8 columns per k-step; queries 64.. live at column 64),
            // B = dO_i MN-major (16 query rows = +2048 B per k-step)
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16_ts(tmem_base + 256, tmem_base + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8), dom + ((st * kT + kk * 2048) >> 4),
                          id_tm, (i > 0 || kk > 0) ? 1u : 0u);
            if (more) {  // S^T(i+1) = K_j Q_{i+1}^T overwrites the score columns (and P^T) behind dV(i), in issue order
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ss(tmem_base + 0, kd + ((kk * 32) >> 4), qd + ((ns * kT + kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
              umma_commit(s_full);
            }
          }
          __syncwarp();
          mbar_wait(ds_ready, i & 1);
          if (i > 0) mbar_wait(dq_free, (i - 1) & 1);
          tc_fence_after();
          if (elect_one()) {
            // dK += dS^T Q_i: A = dS^T from TMEM (over the dP^T columns), B = Q_i MN-major
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16_ts(tmem_base + 320, tmem_base + 128 + (kk < 4 ? kk * 8 : 64 + (kk - 4) * 8),
                          qm + ((st * kT + kk * 2048) >> 4), id_tm, (i > 0 || kk > 0) ? 1u : 0u);
            if (more) {  // dP^T(i+1) = V_j dO_{i+1}^T overwrites dS^T behind dK(i)
#pragma unroll
              for (int kk = 0; kk < 4; ++kk)
                umma_f16_ss(tmem_base + 128, vd + ((kk * 32) >> 4), dod + ((ns * kT + kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
              umma_commit(dp_full);
            }
            // dQ_i(partial) = dS K_j: A = dS^T in smem read MN-major (16 keys = +2048 B per k-step, the two 64-query atoms
            // 16 KB apart), B = K_j MN-major
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16_ss(tmem_base + 384, dsm + ((kk * 2048) >> 4), km + ((kk * 2048) >> 4), id_mm, kk > 0 ? 1u : 0u);
            umma_commit(dq_full);
            umma_commit(&q_empty[st]);
          }
          __syncwarp();
        }
      }
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4 && warp < 12) {
    // ---------------------------------------------------------------- workers: two threads per KEY row, 64 queries each
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;  // queries [grp*64, grp*64+64) of the tile
    const int r = ew * 32 + lane;     // key row within the tile
    const int wt = grp * 128 + r;     // worker thread index 0..255
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const bool key_ok = k0 + r < len;
    uint8_t* dd = smem + Bwd3Smem::kDS + grp * 16384 + r * 128;   // this thread's 128-byte row of dS^T (its query half)
    float* stat = reinterpret_cast<float*>(smem + Bwd3Smem::kStat);
    const float2 sc2 = make_float2(scale2, scale2), ss2 = make_float2(softmax_scale, softmax_scale);
    const float* lse_h = lse + (size_t)head * T + seq_begin;
    const float* dl_h = delta + (size_t)head * T + seq_begin;
    // thread wt publishes one column statistic per tile: wt < 128: -lse * log2(e) of query wt; else -delta * scale of query wt - 128
    const int sq = wt & 127;
    float st_n = (sq < len) ? ((wt < 128) ? -lse_h[sq] * kLog2e : -dl_h[sq] * softmax_scale) : ((wt < 128) ? -INFINITY : 0.f);
    for (int i = 0; i < nq; ++i) {
      float* sb = stat + (i & 1) * 256;
      sb[wt] = st_n;  // [0,128): -lse2 per query (-inf past the sequence end => P = 0), [128,256): -delta*scale
      {
        const int nqr = (i + 1) * 128 + sq;  // prefetch the next tile's statistic
        st_n = (nqr < len) ? ((wt < 128) ? -lse_h[nqr] * kLog2e : -dl_h[nqr] * softmax_scale) : ((wt < 128) ? -INFINITY : 0.f);
      }
      named_bar_sync(4, 256);
      const float* nl = sb + grp * 64;        // -lse2 of this thread's 64 queries
      const float* nd = sb + 128 + grp * 64;  // -delta*scale
      // ---- X: P^T = exp2(S^T * scale2 - lse2[q]) -> bf16 pairs -> this thread's first 32 score columns
      uint32_t pp[32];
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      {
        uint32_t va[32], vb[32];
        tmem_ld_32x32(tmem_base + lane_base + grp * 64, va);
        tmem_ld_32x32(tmem_base + lane_base + grp * 64 + 32, vb);
        tmem_ld_wait();
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          const float4 c = *reinterpret_cast<const float4*>(nl + 2 * t);  // broadcast: every lane reads the same address
          const float2 xa = ffma2(make_float2(__uint_as_float(va[2 * t]), __uint_as_float(va[2 * t + 1])), sc2, make_float2(c.x, c.y));
          const float2 xb = ffma2(make_float2(__uint_as_float(va[2 * t + 2]), __uint_as_float(va[2 * t + 3])), sc2, make_float2(c.z, c.w));
          pp[t] = pack_bf16x2(fast_exp2(xa.x), fast_exp2(xa.y));
          pp[t + 1] = pack_bf16x2(fast_exp2(xb.x), fast_exp2(xb.y));
        }
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          const float4 c = *reinterpret_cast<const float4*>(nl + 32 + 2 * t);
          const float2 xa = ffma2(make_float2(__uint_as_float(vb[2 * t]), __uint_as_float(vb[2 * t + 1])), sc2, make_float2(c.x, c.y));
          const float2 xb = ffma2(make_float2(__uint_as_float(vb[2 * t + 2]), __uint_as_float(vb[2 * t + 3])), sc2, make_float2(c.z, c.w));
          pp[16 + t] = pack_bf16x2(fast_exp2(xa.x), fast_exp2(xa.y));
          pp[16 + t + 1] = pack_bf16x2(fast_exp2(xb.x), fast_exp2(xb.y));
        }
      }
      if (!key_ok) {  // keys past the sequence end contribute nothing
#pragma unroll
        for (int t = 0; t < 32; ++t) pp[t] = 0u;
      }
      tmem_st_32x32(tmem_base + lane_base + grp * 64, pp);  // over this thread's own (already loaded) score columns
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      // ---- Y: dS^T = P^T * (dP^T * scale - delta[q] * scale) -> TMEM (A of dK) and smem (A of dQ)
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
      if (i > 0) mbar_wait_quiet(dq_full, (i - 1) & 1);  // dQ(i-1) has finished reading the dS buffer
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t vd[32];
        tmem_ld_32x32(tmem_base + lane_base + 128 + grp * 64 + c * 32, vd);
        tmem_ld_wait();
        uint32_t w[16];
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          const float4 dc = *reinterpret_cast<const float4*>(nd + c * 32 + 2 * t);
          const float2 ga = ffma2(make_float2(__uint_as_float(vd[2 * t]), __uint_as_float(vd[2 * t + 1])), ss2, make_float2(dc.x, dc.y));
          const float2 gb = ffma2(make_float2(__uint_as_float(vd[2 * t + 2]), __uint_as_float(vd[2 * t + 3])), ss2, make_float2(dc.z, dc.w));
          const float2 da = fmul2(unpack_bf16x2(pp[c * 16 + t]), ga);
          const float2 db = fmul2(unpack_bf16x2(pp[c * 16 + t + 1]), gb);
          w[t] = pack_bf16x2(da.x, da.y);
          w[t + 1] = pack_bf16x2(db.x, db.y);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dd + (((c * 4 + q) ^ (r & 7)) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        // over this thread's own dP^T columns [0,16) / [16,32): both lie inside chunk 0, which is in registers by now
        tmem_st_32x16(tmem_base + lane_base + 128 + grp * 64 + c * 16, w);
      }
      fence_proxy_async_smem();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_ready);
    }
    // dV (group 0) / dK (group 1) -> bf16 -> the (dead) dS buffer, one swizzled 128-byte row per thread; then each group
    // copies its tile out with row-contiguous 16-byte stores
    mbar_wait(acc_full, 0);
    tc_fence_after();
    uint8_t* stg = smem + Bwd3Smem::kDS + grp * 16384;
#pragma unroll 1
    for (int c = 0; c < 2; ++c) {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + lane_base + 256 + grp * 64 + c * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
        w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
        w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
        w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
        *reinterpret_cast<uint4*>(stg + r * 128 + (((c * 4 + q) ^ (r & 7)) << 4)) = w;
      }
    }
    named_bar_sync(2 + grp, 128);
    {
      const int tid = (threadIdx.x - 128) & 127;
      const int rows_ok = min(128, len - k0);
      uint8_t* obase = reinterpret_cast<uint8_t*>(dqkv + ((size_t)(seq_begin + k0) * 3 + (grp == 0 ? 2 : 1)) * H * kDh + (size_t)head * kDh);
#pragma unroll
      for (int it = 0; it < 8; ++it) {
        const int idx = it * 128 + tid, row = idx >> 3, ch = idx & 7;
        if (row < rows_ok)
          *reinterpret_cast<uint4*>(obase + (size_t)row * 3 * H * kDh * 2 + ch * 16) =
              *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
  } else if (warp >= 12) {
    // ---------------------------------------------------------------- dQ drain: one thread per query row, 64 columns
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const int etid = threadIdx.x - 384;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    uint8_t* stage = smem + Bwd3Smem::kDQ;  // two [128 x 32] fp32 boxes (128-byte rows, 128B swizzle)
    for (int i = 0; i < nq; ++i) {
      const bool row_ok = i * 128 + r < len;
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      uint32_t va[32], vb[32];
      tmem_ld_32x32(tmem_base + lane_base + 384, va);
      tmem_ld_32x32(tmem_base + lane_base + 384 + 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);  // the next tile's dQ MMA may overwrite the columns
      if (etid == 0) tma_store_wait_read<0>();  // the previous reduce-add has finished reading the stage
      named_bar_sync(1, 128);
      uint8_t* d0 = stage + r * 128;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        *reinterpret_cast<uint4*>(d0 + ((q ^ (r & 7)) << 4)) =
            make_uint4(row_ok ? va[4 * q] : 0u, row_ok ? va[4 * q + 1] : 0u, row_ok ? va[4 * q + 2] : 0u, row_ok ? va[4 * q + 3] : 0u);
        *reinterpret_cast<uint4*>(d0 + 16384 + ((q ^ (r & 7)) << 4)) =
            make_uint4(row_ok ? vb[4 * q] : 0u, row_ok ? vb[4 * q + 1] : 0u, row_ok ? vb[4 * q + 2] : 0u, row_ok ? vb[4 * q + 3] : 0u);
      }
      fence_proxy_async_smem();
      named_bar_sync(1, 128);
      if (etid == 0) {
        tma_reduce_add_2d(&tmDQ, stage, col_o, seq_begin + i * 128);
        tma_reduce_add_2d(&tmDQ, stage + 16384, col_o + 32, seq_begin + i * 128);
        tma_store_commit();
      }
    }
    if (etid == 0) tma_store_wait_read<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// plain fp32 -> bf16 conversion of the dQ accumulator into the q slot of dqkv (no rotary: ViT path)
__global__ void dq_finalize_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv, int T, int HD) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per = HD / 8;
  if (i >= (int64_t)T * per) return;
  const int t = (int)(i / per), c = (int)(i % per) * 8;
  const float4 a = *reinterpret_cast<const float4*>(dq_acc + (size_t)t * HD + c);
  const float4 b = *reinterpret_cast<const float4*>(dq_acc + (size_t)t * HD + c + 4);
  uint4 w;
  w.x = pack_bf16x2(a.x, a.y);
  w.y = pack_bf16x2(a.z, a.w);
  w.z = pack_bf16x2(b.x, b.y);
  w.w = pack_bf16x2(b.z, b.w);
  *reinterpret_cast<uint4*>(dqkv + (size_t)t * 3 * HD + c) = w;
}

}  // namespace cx

using namespace cx;

// Kernel generation selectable at run time for A/B timing and profiling sessions.  CX_ATTN_FWD = 1: serial 128-key tiles,
// 3: pipelined 64-key sub-tiles (P in tensor memory; carries the trace / ablation hooks), 6: wide-S = default, 7: wide-S
// with 3/8 of the exponentials on the FMA pipe, 8: EXPERIMENTAL two-threads-per-row wide-S kernel (not yet run on hardware);
// CX_ATTN_BWD = 1: serial, 2: pipelined = default, 3: EXPERIMENTAL transposed-score kernel (dV / dK fed from TMEM; not yet run
// on hardware).  Measured on B200,
// 64 x 512 tokens x 12 heads, L2 flushed: forward 138 / 127 / 110.6 / 113.6 us; backward (incl. delta, zero fill, dQ
// finalize) 415 / 376 us.  (Also tried and dropped: P through shared memory in the sub-tile kernel 131 us; 3/8 and 4/8
// polynomial exponentials there 145 / 150 us; holding back the second CTA of each SM to de-phase the pair: no gain.)
constexpr int kFwdDefaultMode = 6;
constexpr int kBwdDefaultMode = 2;
constexpr uint32_t kPoly38 = 0x52;  // column-pair pattern (period 8) routed to the polynomial
static int attn_mode(const char* name, int dflt) {
  const char* e = getenv(name);
  if (!e || !*e) return dflt;
  const int v = atoi(e);
  return (v == 1 || v == 2 || v == 3 || v == 6 || v == 7 || v == 8) ? v : dflt;
}

extern "C" int cx_debug_attn_trace(void* buf) {
  CX_CUDA_CHECK(cudaMemcpyToSymbol(g_attn_trace, &buf, sizeof(buf)));
  return 0;
}

// timing ablations for profiling sessions (results are WRONG when set): CX_ATTN_ABLATE bit mask, see the kernels
static int attn_ablate() {
  const char* e = getenv("CX_ATTN_ABLATE");
  return (e && *e) ? atoi(e) : 0;
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
  static bool configured = false;
  if (!configured) {
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, FwdSmem::kTotal));
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd2_kernel<true, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem::kTotal));
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd3_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem::kTotal));
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd4_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd4Smem::kTotal));
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_fwd3_kernel<kPoly38>, cudaFuncAttributeMaxDynamicSharedMemorySize, Fwd2Smem::kTotal));
    configured = true;
  }
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  const int mode = attn_mode("CX_ATTN_FWD", kFwdDefaultMode);
  const int ablate = attn_ablate();
  if (mode == 1)
    attn_fwd_kernel<<<grid, kFwdThreads, FwdSmem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse, total_tokens, H,
                                                                   softmax_scale * kLog2e);
  else if (mode == 3)
    attn_fwd2_kernel<true, 0><<<grid, kFwd2Threads, Fwd2Smem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse,
                                                                               total_tokens, H, softmax_scale * kLog2e, ablate);
  else if (mode == 8)
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
  static bool configured = false;
  if (!configured) {
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal));
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd2_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, BwdSmem::kTotal));
    CX_CUDA_CHECK(cudaFuncSetAttribute(attn_bwd3_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, Bwd3Smem::kTotal));
    configured = true;
  }
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  const int bmode = attn_mode("CX_ATTN_BWD", kBwdDefaultMode);
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
