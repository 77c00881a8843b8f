// Attention forward kernels (included by cx_attn.cu inside namespace cx; see the overview there).
#pragma once

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

// ---------------------------------------------------------------------------------------------- forward, persistent
// EXPERIMENTAL (CX_ATTN_FWD=9; written after the last GPU minute of round 1: it compiles and its mbarrier protocol runs
// clean on tools/sim_attn_protocol.py, it has NOT run on hardware yet).  attn_fwd4_kernel made persistent: 2 CTAs per SM
// loop over the work items (sequence, head, 128 query rows), so the next item's Q / K / V loads and its first score tile
// run under the current item's loop and epilogue (tools/trace_attn.py: ~30 % of a non-persistent CTA's lifetime is the TMA
// round trip of its first loads plus its epilogue).  All pipeline barriers run on counters that continue across items; Q is
// double-buffered and the finished item's Q buffer doubles as the staging tile of its epilogue (q_empty = the MMA commit +
// 256 softmax arrivals after the copy-out); the single O accumulator is handed back through o_free before the next item's
// first PV overwrites it.  TMEM as attn_fwd4_kernel: S [0,128)  O [128,192)  P [192,256).
struct Fwd5Smem {
  static constexpr int kTile = 128 * kDh * 2;      // 16 KB: 128 rows x 128 B
  static constexpr int kQ = 0;                     // 2 buffers (item parity)
  static constexpr int kK = kQ + 2 * kTile;        // 2 stages of 128 keys
  static constexpr int kV = kK + 2 * kTile;        // 2 stages
  static constexpr int kX = kV + 2 * kTile;        // exchange: 2 x [2 groups][128 rows] bf16 maxima, then [2][128] fp32 row sums
  static constexpr int kBars = kX + 1024;
  static constexpr int kTotal = kBars + 256;       // 99,584 B: two CTAs per SM
};

struct Fwd5Item {
  int seq_begin, len, q0, nk, head;
};
// work item w -> (sequence, head, query tile); false if the tile lies past the end of its sequence
__device__ __forceinline__ bool fwd5_item(int w, int nqt, int H, const int* __restrict__ cu_seqlens, Fwd5Item& it) {
  const int qt = w % nqt, rest = w / nqt;
  it.head = rest % H;
  const int seq = rest / H;
  it.seq_begin = cu_seqlens[seq];
  it.len = cu_seqlens[seq + 1] - it.seq_begin;
  it.q0 = qt * 128;
  it.nk = (it.len + 127) / 128;
  return it.q0 < it.len;
}

__global__ void __launch_bounds__(kFwd4Threads, 2)
attn_fwd5_kernel(const __grid_constant__ CUtensorMap tmQKV, const int* __restrict__ cu_seqlens,
                 __nv_bfloat16* __restrict__ out, float* __restrict__ lse, int T, int H, float scale2, int nqt, int n_items) {
  extern __shared__ __align__(1024) uint8_t smem[];
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Fwd5Smem::kBars);
  uint64_t* q_full = bars;          // [2]
  uint64_t* q_empty = bars + 2;     // [2]  257 arrivals: the MMA commit behind the item's last S + the 256 epilogue threads
  uint64_t* k_full = bars + 4;      // [2]
  uint64_t* k_empty = bars + 6;     // [2]
  uint64_t* v_full = bars + 8;      // [2]
  uint64_t* v_empty = bars + 10;    // [2]
  uint64_t* s_full = bars + 12;     // S(g) in TMEM
  uint64_t* s_free = bars + 13;     // every softmax thread holds its 64 scores of S(g) in registers (256 arrivals)
  uint64_t* p_ready = bars + 14;    // P(g) written (256 arrivals)
  uint64_t* pv_done = bars + 15;    // PV(g) complete
  uint64_t* o_full = bars + 16;     // item n: O complete
  uint64_t* o_free = bars + 17;     // item n: O read out by the epilogue (256 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 18);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if ((smem_u32(smem) & 1023u) != 0) __trap();
  if (warp == 0 && lane == 0) {
    for (int i = 0; i < 2; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 257);
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
    mbar_init(o_free, 256);
    fence_barrier_init();
    tma_prefetch_desc(&tmQKV);
  }
  if (warp == 2) tmem_alloc<256>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: runs ahead across items
    int g = 0, n = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      Fwd5Item it;
      if (!fwd5_item(w, nqt, H, cu_seqlens, it)) continue;
      const int col_q = (0 * H + it.head) * kDh, col_k = (1 * H + it.head) * kDh, col_v = (2 * H + it.head) * kDh;
      const int qb = n & 1;
      mbar_wait(&q_empty[qb], ((n >> 1) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[qb], Fwd5Smem::kTile);
        tma_load_2d(smem + Fwd5Smem::kQ + qb * Fwd5Smem::kTile, &tmQKV, &q_full[qb], col_q, it.seq_begin + it.q0);
      }
      __syncwarp();
      for (int j = 0; j < it.nk; ++j, ++g) {
        const int st = g & 1;
        const uint32_t ph = (g >> 1) & 1;
        mbar_wait(&k_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&k_full[st], Fwd5Smem::kTile);
          tma_load_2d(smem + Fwd5Smem::kK + st * Fwd5Smem::kTile, &tmQKV, &k_full[st], col_k, it.seq_begin + j * 128);
        }
        __syncwarp();
        mbar_wait(&v_empty[st], ph ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(&v_full[st], Fwd5Smem::kTile);
          tma_load_2d(smem + Fwd5Smem::kV + st * Fwd5Smem::kTile, &tmQKV, &v_full[st], col_v, it.seq_begin + j * 128);
        }
        __syncwarp();
      }
      ++n;
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: one flat sequence of key tiles
    constexpr uint32_t idesc_s = make_idesc_bf16(128, 128, 0, 0);  // S = Q K^T (128 keys): both K-major
    constexpr uint32_t idesc_o = make_idesc_bf16(128, 64, 0, 1);   // O += P V: A (P) from TMEM, B (V) MN-major
    const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + Fwd5Smem::kQ), 0, 1024);
    const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + Fwd5Smem::kK), 0, 1024);
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + Fwd5Smem::kV), 8192, 1024);
    auto issue_s = [&](const int qb, const int st) {  // S(next tile) from Q buffer qb and key stage st
      if (elect_one()) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk)
          umma_f16_ss(tmem_base, qd + (uint64_t)((qb * Fwd5Smem::kTile + kk * 32) >> 4),
                      kd + (uint64_t)((st * Fwd5Smem::kTile + kk * 32) >> 4), idesc_s, kk > 0 ? 1u : 0u);
        umma_commit(s_full);
        umma_commit(&k_empty[st]);
      }
      __syncwarp();
    };
    // cursor over the valid items of this CTA: (w, nk) of the current item and of the next one
    int w_cur = blockIdx.x, nk_cur = 0;
    {
      Fwd5Item it;
      while (w_cur < n_items && !fwd5_item(w_cur, nqt, H, cu_seqlens, it)) w_cur += gridDim.x;
      nk_cur = it.nk;
    }
    if (w_cur < n_items) {
      int g = 0, n = 0;
      mbar_wait(&q_full[0], 0);
      mbar_wait(&k_full[0], 0);
      tc_fence_after();
      issue_s(0, 0);
      while (w_cur < n_items) {
        int w_nxt = w_cur + gridDim.x, nk_nxt = 0;
        {
          Fwd5Item it;
          while (w_nxt < n_items && !fwd5_item(w_nxt, nqt, H, cu_seqlens, it)) w_nxt += gridDim.x;
          nk_nxt = it.nk;
        }
        for (int j = 0; j < nk_cur; ++j, ++g) {
          const bool last = j == nk_cur - 1;
          const bool has_next = !last || w_nxt < n_items;
          if (has_next) {  // S(g+1) as soon as the score columns have been read out: it runs under softmax(g)
            const int qb = last ? ((n + 1) & 1) : (n & 1);
            mbar_wait(s_free, g & 1);
            if (last) mbar_wait(&q_full[qb], ((n + 1) >> 1) & 1);
            mbar_wait(&k_full[(g + 1) & 1], ((g + 1) >> 1) & 1);
            tc_fence_after();
            issue_s(qb, (g + 1) & 1);
          }
          if (last) {  // every S of item n has been issued: its Q buffer is free once they complete (+ the epilogue's copy-out)
            if (elect_one()) umma_commit(&q_empty[n & 1]);
            __syncwarp();
          }
          mbar_wait(&v_full[g & 1], (g >> 1) & 1);
          mbar_wait(p_ready, g & 1);
          if (j == 0 && n > 0) mbar_wait(o_free, (n - 1) & 1);  // the previous item's O has been read out
          tc_fence_after();
          if (elect_one()) {
#pragma unroll
            for (int kk = 0; kk < 8; ++kk)
              umma_f16_ts(tmem_base + 128, tmem_base + 192 + kk * 8,
                          vd + (uint64_t)(((g & 1) * Fwd5Smem::kTile + kk * 2048) >> 4), idesc_o, (j > 0 || kk > 0) ? 1u : 0u);
            umma_commit(&v_empty[g & 1]);
            umma_commit(pv_done);
            if (last) umma_commit(o_full);
          }
          __syncwarp();
        }
        ++n;
        w_cur = w_nxt;
        nk_cur = nk_nxt;
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- softmax + epilogue: two threads per query row
    const int ew = warp & 3;
    const int grp = (warp - 4) >> 2;             // key columns [grp*64, +64) of each tile; O columns [grp*32, +32)
    const int r = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const uint32_t t_s = tmem_base + lane_base + grp * 64;
    const uint32_t t_o = tmem_base + lane_base + 128 + grp * 32;
    const uint32_t t_p = tmem_base + lane_base + 192 + grp * 32;
    __nv_bfloat16* smax = reinterpret_cast<__nv_bfloat16*>(smem + Fwd5Smem::kX);  // [2 parities][2 groups][128 rows]
    float* lsum = reinterpret_cast<float*>(smem + Fwd5Smem::kX);
    const float2 sc2 = make_float2(scale2, scale2);
    int g = 0, n = 0;
    for (int w = blockIdx.x; w < n_items; w += gridDim.x) {
      Fwd5Item it;
      if (!fwd5_item(w, nqt, H, cu_seqlens, it)) continue;
      const int len = it.len, q0 = it.q0, head = it.head, seq_begin = it.seq_begin;
      const int q_row = q0 + r;
      float m_run = -INFINITY, l_run = 0.f;        // l_run: this thread's 64-column share of the row sum
      for (int j = 0; j < it.nk; ++j, ++g) {
        mbar_wait(s_full, g & 1);
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
        const __nv_bfloat16 mine = __float2bfloat16_ru(mx * scale2);
        __nv_bfloat16* xm = smax + (g & 1) * 256;
        xm[grp * 128 + r] = mine;
        named_bar_sync(2, 256);
        const float m_c = fmaxf(__bfloat162float(mine), __bfloat162float(xm[(grp ^ 1) * 128 + r]));
        const bool raise = m_c > m_run + 8.f;      // always true on an item's first tile (m_run = -inf); same in both threads
        float alpha = 1.f;
        if (raise) {
          alpha = fast_exp2(m_run - m_c);          // 0 on the first tile
          m_run = m_c;
        }
        const float2 nm2 = make_float2(-m_run, -m_run);
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
          if (b == 0 && g > 0) {  // PV(g-1) (possibly the previous item's last) has finished reading the P columns
            mbar_wait_quiet(pv_done, (g - 1) & 1);
            tc_fence_after();
          }
          tmem_st_32x16(t_p + b * 16, pp);
        }
        l_run = l_run * alpha + ((rs0.x + rs0.y) + (rs1.x + rs1.y));
        // rescale this thread's 32 output columns only if some row of the warp raised its maximum (never on the first tile:
        // its PV starts a fresh accumulator)
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
      // ---- epilogue of item n: read O out (frees the accumulator), O / l -> bf16 -> the item's own (dead) Q buffer -> global
      mbar_wait(o_full, n & 1);
      tc_fence_after();
      uint32_t v[32];
      tmem_ld_32x32(t_o, v);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(o_free);
      named_bar_sync(2, 256);                      // every thread is past its last exchange read of this item
      lsum[grp * 128 + r] = l_run;
      named_bar_sync(2, 256);
      const float l_tot = l_run + lsum[(grp ^ 1) * 128 + r];
      const float inv_l = 1.f / l_tot;
      const bool row_ok = q_row < len;
      uint8_t* stg = smem + Fwd5Smem::kQ + (n & 1) * Fwd5Smem::kTile;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 wv;
        wv.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]) * inv_l, __uint_as_float(v[8 * q + 1]) * inv_l);
        wv.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]) * inv_l, __uint_as_float(v[8 * q + 3]) * inv_l);
        wv.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]) * inv_l, __uint_as_float(v[8 * q + 5]) * inv_l);
        wv.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]) * inv_l, __uint_as_float(v[8 * q + 7]) * inv_l);
        *reinterpret_cast<uint4*>(stg + r * 128 + (((grp * 4 + q) ^ (r & 7)) << 4)) = wv;
      }
      named_bar_sync(2, 256);                      // (also orders the next item's first exchange write behind the lsum reads)
      {
        const int tid = threadIdx.x - 128;
        const int rows_ok = min(128, len - q0);
        uint8_t* obase = reinterpret_cast<uint8_t*>(out + ((size_t)(seq_begin + q0) * H + head) * kDh);
#pragma unroll
        for (int i4 = 0; i4 < 4; ++i4) {
          const int idx = i4 * 256 + tid, row = idx >> 3, ch = idx & 7;
          if (row < rows_ok)
            *reinterpret_cast<uint4*>(obase + (size_t)row * H * kDh * 2 + ch * 16) =
                *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
        }
      }
      fence_proxy_async_smem();                    // the staging reads precede the TMA write that refills this Q buffer
      mbar_arrive(&q_empty[n & 1]);
      if (row_ok && grp == 0) lse[(size_t)head * T + seq_begin + q_row] = (m_run + log2f(l_tot)) * kLn2;
      ++n;
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<256>(tmem_base);
  }
}

