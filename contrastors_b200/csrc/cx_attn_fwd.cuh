// Attention forward kernel (included by cx_attn.cu inside namespace cx; see the overview there).
#pragma once

// ============================================================================================== forward
// One CTA = (sequence, head, 128 query rows); two CTAs are co-resident per SM (83 KB smem, 256 TMEM columns each).
//   warp 0: TMA producer (Q once, K/V tiles double-buffered)     warp 1: MMA issuer (converged warp, one elected lane)
//   warp 2: TMEM allocator                                        warps 4-11: softmax, TWO threads per query row
// Per 128-key tile j: S = Q K_j^T (one N = 128 chain, TMEM fp32) -> online softmax, each thread owning 64 of the tile's key
// columns (exp2, lazily raised maximum agreed between the two threads of a row through a 2-byte exchange in shared memory,
// packed fp32x2 arithmetic) -> P (bf16, its own TMEM columns) -> O += P V_j (A from TMEM).  S(j+1) is issued as soon as every
// thread holds its S(j) scores in registers, so it runs under softmax(j).  With 8 softmax warps per CTA every SM sub-partition
// has four of them to overlap: the one-thread-per-row predecessor was bound by the latency chain of its single warp per
// sub-partition (round 1: 110.6 us at 64 x 512 x 12; this kernel 103.3 us; FlashAttention-2 on the same box 174 us).
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

// kPoly: how many of the 8 column quads of each 32-column batch take the second pair of their exponentials from the FMA-pipe
// polynomial (exp2_poly2) instead of MUFU.EX2: 0 = none, 4 = every other quad (25 % of the exponentials), 8 = all (50 %).  The
// forward is bound by the SFU (16 exponentials / clk / SM = 1024 clk per 128 x 128 tile against 512 clk of tensor time).
template <int kPoly>
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
          if (kPoly == 8 || (kPoly == 4 && ((t >> 1) & 1))) x1 = exp2_poly2(x1);
          else x1 = make_float2(fast_exp2(x1.x), fast_exp2(x1.y));
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
