// Varlen non-causal multi-head attention forward/backward on tcgen05 / TMEM / TMA (Dh = 64), C ABI cx_attn_{fwd,bwd}.
//
// Replaces flash_attn_varlen_qkvpacked_func / flash_attn_qkvpacked_func (FA2, mma.sync) at
// /root/reference/src/contrastors/layers/attention.py:158-181,220-226.  Layouts follow the reference's packed format:
// qkv [T, 3, H, Dh] bf16 over unpadded tokens, cu_seqlens int32 [nseq+1].
//
// Forward: one CTA = (sequence, head, 256 query rows) = two 128-row query tiles that ping-pong on the tensor core
//   warp 0: TMA producer (Q once, K/V tiles double-buffered)     warp 1: MMA issuer (one thread)
//   warp 2: TMEM allocator                                        warps 4-7 / 8-11: softmax group of query tile 0 / 1
//   per key tile j:  S = Q K_j^T (TMEM, fp32)  ->  online softmax in registers (exp2, running max/sum)  ->  P (bf16,
//   swizzled smem)  ->  O += P V_j (TMEM).  O is rescaled in TMEM only when a row maximum moves.
// Backward: one CTA = (sequence, head, 128 keys); loops over query tiles; S and dP are recomputed into TMEM,
//   P / dS go to smem once and feed three contractions (dV += P^T dO, dK += dS^T Q, dQ += dS K); dK/dV accumulate in
//   TMEM, dQ partials leave through TMA reduce-add into an fp32 accumulator.
#include <math.h>

#include "cx_host.h"
#include "cx_ptx.cuh"

namespace cx {

constexpr int kDh = 64;
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
    configured = true;
  }
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  attn_fwd_kernel<<<grid, kFwdThreads, FwdSmem::kTotal, stream>>>(tm, cu_seqlens, (__nv_bfloat16*)out, lse, total_tokens, H,
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
    configured = true;
  }
  dim3 grid((max_seqlen + 127) / 128, H, nseq);
  attn_bwd_kernel<<<grid, kBwdThreads, BwdSmem::kTotal, stream>>>(tmQKV, tmDO, tmDQ, cu_seqlens, lse, delta,
                                                                 (__nv_bfloat16*)dqkv, T, H, softmax_scale);
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
