// Attention backward kernels (included by cx_attn.cu inside namespace cx; see the overview there).
#pragma once

// ============================================================================================== backward
// delta[h, t] = sum_d dO[t,h,d] * O[t,h,d]; 8 threads per (t, h) row of 64, 16-byte loads.  The same threads zero the fp32 dQ
// accumulator row the backward kernel reduce-adds into (32 bytes each): no separate 100 MB fill launch per layer.
__global__ void attn_delta_kernel(const __nv_bfloat16* __restrict__ o, const __nv_bfloat16* __restrict__ dout,
                                  float* __restrict__ delta, float* __restrict__ dq_acc, int T, int H) {
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
    float4* z = reinterpret_cast<float4*>(dq_acc + row * kDh + sub * 8);
    z[0] = make_float4(0.f, 0.f, 0.f, 0.f);
    z[1] = make_float4(0.f, 0.f, 0.f, 0.f);
  }
  s += __shfl_xor_sync(0xffffffffu, s, 4);
  s += __shfl_xor_sync(0xffffffffu, s, 2);
  s += __shfl_xor_sync(0xffffffffu, s, 1);
  if (sub == 0 && row < (int64_t)T * H) {
    const int t = (int)(row / H), h = (int)(row % H);
    delta[(size_t)h * T + t] = s;
  }
}

// Optional phase trace (tools/trace_attn_bwd.py builds a separate library with -DCX_ATTN_TRACE; the product build compiles none
// of it): one lane of one warp per role stamps clock64() at its phase boundaries, [block][role 5][tile 8][point 8].
#ifdef CX_ATTN_TRACE
__device__ long long* g_bwd_trace = nullptr;
// the buffer pointer is read ONCE per thread (CX_TR_INIT): a stamp is then a clock read and a fire-and-forget store
#define CX_TR_INIT() long long* const cx_tr_buf = g_bwd_trace
#define CX_TR(role, tile, pt)                                                                                          \
  do {                                                                                                                 \
    if (cx_tr_buf != nullptr && lane == 0 && (tile) < 8)                                                               \
      cx_tr_buf[((((size_t)(blockIdx.z * gridDim.y + blockIdx.y) * gridDim.x + blockIdx.x) * 5 + (role)) * 8 + (tile)) * 8 + (pt)] = clock64(); \
  } while (0)
#define CX_TRW(i, pt) do { if (tr_role < 5) CX_TR(tr_role, i, pt); } while (0)
#else
#define CX_TR_INIT() do { } while (0)
#define CX_TR(role, tile, pt) do { } while (0)
#define CX_TRW(i, pt) do { } while (0)
#endif

// ---------------------------------------------------------------------------------------------- backward, transposed scores
// One CTA = (sequence, head, 128 keys); loops over the query tiles of the sequence.  The straightforward formulation (S = Q K^T,
// P and dS through shared memory as the A operands of dV += P^T dO and dK += dS^T Q) is bound by the shared-memory port.  Here
// the scores are computed TRANSPOSED, S^T = K_j Q_i^T and dP^T = V_j dO_i^T (lane = key, column = query), so that P^T and dS^T --
// the A operands of dV += P^T dO and dK += dS^T Q -- are written back into tensor memory (bf16 pairs, over the thread's own
// score columns) and those two contractions read only their B operand from shared memory.  dS still goes to shared memory once,
// as the (MN-major) A operand of dQ = dS K; dQ partials leave through a TMA reduce-add into an fp32 accumulator.
//
// What the phase trace (tools/trace_attn_bwd.py, profiles/r02l_*, r02n_*) showed about the first version of this kernel (8 worker
// warps, two threads per key row, statistics read from shared memory; 325 us incl. delta + finalize at 64 x 512 x 12) and what
// this version does about it:
//  * the query-tile period (~3400 clk) is the SUM of the two element-wise phases of the worker warps (X: P^T = exp2(..), Y: dS^T),
//    the tensor pipe (~1550 clk of MMAs per tile) being hidden underneath; every consumer of the shared-memory port that is
//    removed shortens the period by about its wavefront count.  The per-COLUMN statistics were the largest avoidable one: a
//    broadcast LDS.128 still delivers 16 bytes to each lane = 4 wavefronts, 1024 per tile.  They now enter through the MMA: a
//    fifth k-step multiplies a ones operand with -lse/scale (resp. -delta) of the query tile, split into three bf16 terms (exact
//    to 2^-24), so S^T and dP^T arrive with the statistic already subtracted and the workers only scale;
//  * 16 worker warps, FOUR threads per key row with 32 query columns each (the dependent FFMA -> MUFU -> F2F chains of 64
//    elements per thread were latency-bound with two warps per scheduler), a quarter of the exponentials on the FMA pipe;
//  * a three-stage Q_i / dO_i ring (with two, the request for tile i+1 went out only when tile i-1's last MMA had completed);
//  * the epilogue (dV / dK conversion, rotary transpose, copy-out) is shared by twice as many threads.
// Measured and NOT adopted (each slower, the traces are in profiles/): X and Y on separate warp groups with P^T in its own
// columns (r02m: 343 us -- the two groups contend for the same port), dQ issued one slot late (r02o: 325 us), K_j / V_j as
// tensor-memory A operands of S^T / dP^T (320 us -- TS MMAs with N = 128 collide with the workers' tcgen05.ld / st).
// TMEM: S^T [0,128)  dP^T [128,256)  dV [256,320)  dK [320,384)  dQ [384,448);  P^T of query quarter qq over S^T columns
// [qq*32, qq*32+16), dS^T over dP^T columns [128+qq*32, +16) (each thread overwrites only columns it has already loaded itself).
// 24 warps: 0 TMA, 1 MMA, 2 TMEM alloc, 3 idle, 4-19 workers, 20-23 dQ drain.
constexpr int kBwd4Threads = 768;
constexpr int kBwd4Poly = 4;  // every other quad of the second half-pair: 25 % of the exponentials on the FMA pipe

struct Bwd4Smem {
  static constexpr int kTile = 128 * kDh * 2;   // 16 KB
  static constexpr int kK = 0;                  // K_j  (A of S^T, B of dQ as MN-major)
  static constexpr int kV = kK + kTile;         // V_j  (A of dP^T)
  static constexpr int kStages = 3;             // Q_i / dO_i ring: tile i+1 is requested when tile i-2's last MMA completes
  static constexpr int kQ = kV + kTile;         // kStages x Q_i (B of S^T, B of dK as MN-major)
  static constexpr int kDO = kQ + kStages * kTile;   // kStages x dO_i (B of dP^T, B of dV as MN-major)
  static constexpr int kDS = kDO + kStages * kTile;  // dS^T [128 keys x 128 q] bf16: 2 blocks (q halves) x 128 rows x 128 B
  static constexpr int kDQ = kDS + 32768;       // fp32 staging for the dQ reduce-add: 2 x [128 x 32] (128 B rows)
  // auxiliary K-major operand tile [128 rows x 64 bf16]: k-step 0 (columns 0..15) = ones (A), k-step 1 = -lse/scale of the query
  // tile split into three bf16 terms (B of S^T), k-step 2 = -delta split the same way (B of dP^T); k-step 3 unused
  static constexpr int kAux = kDQ + 2 * 16384;
  static constexpr int kBars = kAux + kTile;
  static constexpr int kTotal = kBars + 256 + 1024;
};

__global__ void __launch_bounds__(kBwd4Threads, 1)
attn_bwd4_kernel(const __grid_constant__ CUtensorMap tmQKV, const __grid_constant__ CUtensorMap tmDO,
                 const __grid_constant__ CUtensorMap tmDQ, const int* __restrict__ cu_seqlens,
                 const float* __restrict__ lse, const float* __restrict__ delta, __nv_bfloat16* __restrict__ dqkv, int T,
                 int H, float softmax_scale, const float* __restrict__ rope_inv_freq) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + Bwd4Smem::kBars);
  uint64_t* kv_full = bars;        // [1]
  uint64_t* q_full = bars + 1;     // [3]  Q_i and dO_i of a stage landed
  uint64_t* q_empty = bars + 4;    // [3]  every MMA reading the stage has completed
  uint64_t* s_full = bars + 7;     // S^T(i) in TMEM
  uint64_t* dp_full = bars + 8;    // dP^T(i) in TMEM
  uint64_t* p_ready = bars + 9;    // P^T(i) in TMEM, every worker has loaded its S^T(i) columns (512 arrivals)
  uint64_t* ds_ready = bars + 10;  // dS^T(i) in TMEM and in smem, every worker has loaded its dP^T(i) columns (512 arrivals)
  uint64_t* dq_full = bars + 11;   // dQ(i) partial in TMEM; also: dQ(i) has finished reading dS(i) from smem
  uint64_t* dq_free = bars + 12;   // dQ TMEM columns drained (128 arrivals)
  uint64_t* acc_full = bars + 13;  // dK / dV complete
  uint64_t* aux_init = bars + 14;  // the auxiliary operand tile holds ones and tile 0's statistics (512 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 15);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  CX_TR_INIT();
  const int seq = blockIdx.z, head = blockIdx.y;
  const int seq_begin = cu_seqlens[seq];
  const int len = cu_seqlens[seq + 1] - seq_begin;
  const int k0 = blockIdx.x * 128;
  if (k0 >= len) return;
  const int nq = (len + 127) / 128;
  const float scale2 = softmax_scale * kLog2e;
  if (warp == 0) CX_TR(4, 0, 0);

  const int col_q = (0 * H + head) * kDh, col_k = (1 * H + head) * kDh, col_v = (2 * H + head) * kDh;
  const int col_o = head * kDh;
  if (warp == 0 && lane == 0) {
    mbar_init(kv_full, 1);
    for (int i = 0; i < Bwd4Smem::kStages; ++i) {
      mbar_init(&q_full[i], 1);
      mbar_init(&q_empty[i], 1);
    }
    mbar_init(s_full, 1);
    mbar_init(dp_full, 1);
    mbar_init(p_ready, 512);
    mbar_init(ds_ready, 512);
    mbar_init(dq_full, 1);
    mbar_init(dq_free, 128);
    mbar_init(acc_full, 1);
    mbar_init(aux_init, 512);
    fence_barrier_init();
    mbar_arrive_expect_tx(kv_full, 2 * Bwd4Smem::kTile);
    tma_load_2d(smem + Bwd4Smem::kK, &tmQKV, kv_full, col_k, seq_begin + k0);
    tma_load_2d(smem + Bwd4Smem::kV, &tmQKV, kv_full, col_v, seq_begin + k0);
    mbar_arrive_expect_tx(&q_full[0], 2 * Bwd4Smem::kTile);
    tma_load_2d(smem + Bwd4Smem::kQ, &tmQKV, &q_full[0], col_q, seq_begin);
    tma_load_2d(smem + Bwd4Smem::kDO, &tmDO, &q_full[0], col_o, seq_begin);
    tma_prefetch_desc(&tmDQ);
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;
  if (warp == 0) CX_TR(4, 0, 1);

  if (warp == 0) {
    for (int i = 1; i < nq; ++i) {  // K, V and the first query tile were issued during set-up
      const int st = i % Bwd4Smem::kStages;
      mbar_wait(&q_empty[st], ((i / Bwd4Smem::kStages) & 1) ^ 1);
      if (elect_one()) {
        mbar_arrive_expect_tx(&q_full[st], 2 * Bwd4Smem::kTile);
        tma_load_2d(smem + Bwd4Smem::kQ + st * Bwd4Smem::kTile, &tmQKV, &q_full[st], col_q, seq_begin + i * 128);
        tma_load_2d(smem + Bwd4Smem::kDO + st * Bwd4Smem::kTile, &tmDO, &q_full[st], col_o, seq_begin + i * 128);
      }
      __syncwarp();
    }
  } else if (warp == 1) {
    constexpr uint32_t id_kk = make_idesc_bf16(128, 128, 0, 0);  // S^T, dP^T: A (K_j / V_j) K-major, B (Q_i / dO_i) K-major
    constexpr uint32_t id_tm = make_idesc_bf16(128, 64, 0, 1);   // dV, dK: A from TMEM, B (dO_i / Q_i) MN-major, N = 64
    constexpr uint32_t id_mm = make_idesc_bf16(128, 64, 1, 1);   // dQ: A (dS^T in smem) MN-major, B (K_j) MN-major, N = 64
    const uint64_t kd = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kK), 0, 1024);        // K_j  K-major (A of S^T)
    const uint64_t vd = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kV), 0, 1024);        // V_j  K-major (A of dP^T)
    const uint64_t qd = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kQ), 0, 1024);        // Q_i  K-major (B of S^T)
    const uint64_t dod = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kDO), 0, 1024);      // dO_i K-major (B of dP^T)
    const uint64_t km = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kK), 8192, 1024);     // K_j  MN-major (B of dQ)
    const uint64_t qm = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kQ), 8192, 1024);     // Q_i  MN-major (B of dK)
    const uint64_t dom = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kDO), 8192, 1024);   // dO_i MN-major (B of dV)
    const uint64_t dsm = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kDS), 16384, 1024);  // dS^T MN-major (A of dQ)
    // fifth k-step of S^T and dP^T: ones [128 keys x 16] times the statistics [128 queries x 16] adds -lse/scale resp. -delta
    // of each query column to the whole column, inside the accumulation (the workers then only scale)
    const uint64_t xd = make_smem_desc_sw128(smem_u32(smem + Bwd4Smem::kAux), 0, 1024);
    const uint64_t x_ones = xd, x_lse = xd + (32 >> 4), x_delta = xd + (64 >> 4);
    mbar_wait(kv_full, 0);
    mbar_wait(&q_full[0], 0);
    mbar_wait(aux_init, 0);
    tc_fence_after();
    if (elect_one()) {
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_base + 0, kd + ((kk * 32) >> 4), qd + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
      umma_f16_ss(tmem_base + 0, x_ones, x_lse, id_kk, 1u);
      umma_commit(s_full);
#pragma unroll
      for (int kk = 0; kk < 4; ++kk) umma_f16_ss(tmem_base + 128, vd + ((kk * 32) >> 4), dod + ((kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
      umma_f16_ss(tmem_base + 128, x_ones, x_delta, id_kk, 1u);
      umma_commit(dp_full);
    }
    __syncwarp();
    for (int i = 0; i < nq; ++i) {
      constexpr int kT = Bwd4Smem::kTile;
      const int st = i % Bwd4Smem::kStages, ns = (i + 1) % Bwd4Smem::kStages;
      const bool more = i + 1 < nq;
      CX_TR(2, i, 0);
      if (more) mbar_wait(&q_full[ns], ((i + 1) / Bwd4Smem::kStages) & 1);
      CX_TR(2, i, 6);
      mbar_wait(p_ready, i & 1);
      tc_fence_after();
      CX_TR(2, i, 1);
      if (elect_one()) {
        // dV += P^T dO_i: A = P^T from TMEM (16 queries = 8 packed columns per k-step; query quarter qq starts at column
        // qq * 32), B = dO_i MN-major (16 query rows = +2048 B per k-step)
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ts(tmem_base + 256, tmem_base + (kk >> 1) * 32 + (kk & 1) * 8, dom + ((st * kT + kk * 2048) >> 4), id_tm,
                      (i > 0 || kk > 0) ? 1u : 0u);
        if (more) {  // S^T(i+1) = K_j Q_{i+1}^T overwrites the score columns (and P^T) behind dV(i), in issue order
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_ss(tmem_base + 0, kd + ((kk * 32) >> 4), qd + ((ns * kT + kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
          umma_f16_ss(tmem_base + 0, x_ones, x_lse, id_kk, 1u);
          umma_commit(s_full);
        }
      }
      __syncwarp();
      CX_TR(2, i, 2);
      mbar_wait(ds_ready, i & 1);
      CX_TR(2, i, 3);
      if (i > 0) mbar_wait(dq_free, (i - 1) & 1);
      tc_fence_after();
      CX_TR(2, i, 4);
      if (elect_one()) {
        // dK += dS^T Q_i: A = dS^T from TMEM (over the dP^T columns), B = Q_i MN-major
#pragma unroll
        for (int kk = 0; kk < 8; ++kk)
          umma_f16_ts(tmem_base + 320, tmem_base + 128 + (kk >> 1) * 32 + (kk & 1) * 8, qm + ((st * kT + kk * 2048) >> 4), id_tm,
                      (i > 0 || kk > 0) ? 1u : 0u);
        if (more) {  // dP^T(i+1) = V_j dO_{i+1}^T overwrites dS^T behind dK(i)
#pragma unroll
          for (int kk = 0; kk < 4; ++kk)
            umma_f16_ss(tmem_base + 128, vd + ((kk * 32) >> 4), dod + ((ns * kT + kk * 32) >> 4), id_kk, kk > 0 ? 1u : 0u);
          umma_f16_ss(tmem_base + 128, x_ones, x_delta, id_kk, 1u);
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
      CX_TR(2, i, 5);
    }
    if (elect_one()) umma_commit(acc_full);
    __syncwarp();
  } else if (warp >= 4 && warp < 20) {
    // ---------------------------------------------------------------- workers: four threads per KEY row, 32 queries each
    const int ew = warp & 3;
    const int qq = (warp - 4) >> 2;   // queries [qq*32, qq*32+32) of the tile
    const int r = ew * 32 + lane;     // key row within the tile
    const int wt = qq * 128 + r;      // worker thread index 0..511
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const bool key_ok = k0 + r < len;
    // this thread's 64 bytes of its dS^T row: block = query half, chunks (qq & 1) * 4 .. + 3 of the 128-byte row
    uint8_t* dd = smem + Bwd4Smem::kDS + (qq >> 1) * 16384 + r * 128;
    const int ch0 = (qq & 1) * 4;
    const float2 sc2 = make_float2(scale2, scale2), ss2 = make_float2(softmax_scale, softmax_scale);
    // The per-query statistics enter through the MMA (see the auxiliary operand tile): row `sq` of the tile holds, as three bf16
    // terms each, -lse/scale (k-step 1) and -delta (k-step 2) of query sq of the CURRENT tile.  Quarter 1's threads rewrite the lse
    // terms for tile i+1 once S^T(i) has completed (behind the s_full(i) wait, before their p_ready(i) arrival, which S^T(i+1)
    // waits for); quarter 2's threads rewrite the delta terms once dP^T(i) has completed (behind dp_full(i), before ds_ready(i),
    // which dP^T(i+1) waits for).  The raw values are loaded from global memory a tile earlier still.
    const int sq = wt & 127;
    uint8_t* aux_row = smem + Bwd4Smem::kAux + sq * 128;
    const float inv_scale = 1.f / softmax_scale;
    const float* st_src = (qq == 1 ? lse : delta) + (size_t)head * T + seq_begin;
    auto stat_terms = [&](float raw, bool ok) -> uint4 {  // value = hi + mid + lo exactly to 2^-24 relative
      const float v = (qq == 1) ? (ok ? -raw * inv_scale : -1e30f) : (ok ? -raw : 0.f);  // -1e30 => P = 0 for queries past the end
      const __nv_bfloat16 h = __float2bfloat16_rn(v);
      const float r1 = v - __bfloat162float(h);
      const __nv_bfloat16 m = __float2bfloat16_rn(r1);
      const __nv_bfloat16 l = __float2bfloat16_rn(r1 - __bfloat162float(m));
      return make_uint4((uint32_t)__bfloat16_as_ushort(h) | ((uint32_t)__bfloat16_as_ushort(m) << 16), (uint32_t)__bfloat16_as_ushort(l), 0u, 0u);
    };
    float st_raw = 0.f;
    {
      // quarter qq initialises logical 16-byte chunks 2*qq and 2*qq+1 of row sq: k-step qq's 16 columns
      uint4 first = make_uint4(0u, 0u, 0u, 0u);
      if (qq == 0) first = make_uint4(0x3F803F80u, 0x00003F80u, 0u, 0u);  // ones in columns 0..2
      if (qq == 1 || qq == 2) {
        first = stat_terms((sq < len) ? st_src[sq] : 0.f, sq < len);
        st_raw = (128 + sq < len) ? st_src[128 + sq] : 0.f;
      }
      *reinterpret_cast<uint4*>(aux_row + (((2 * qq) ^ (sq & 7)) << 4)) = first;
      *reinterpret_cast<uint4*>(aux_row + (((2 * qq + 1) ^ (sq & 7)) << 4)) = make_uint4(0u, 0u, 0u, 0u);
      fence_proxy_async_smem();
      mbar_arrive(aux_init);
    }
    const int tr_role = (warp == 4) ? 0 : (warp == 12 ? 1 : 5);  // trace: the first warp of query quarters 0 and 2
    (void)tr_role;
    for (int i = 0; i < nq; ++i) {
      CX_TRW(i, 0);
      // ---- X: P^T = exp2((S^T - lse/scale) * scale2) -> bf16 pairs -> this thread's first 16 score columns
      uint32_t pp[16];
      mbar_wait(s_full, i & 1);
      tc_fence_after();
      CX_TRW(i, 1);
      if (qq == 1 && i + 1 < nq) {
        *reinterpret_cast<uint4*>(aux_row + ((2 ^ (sq & 7)) << 4)) = stat_terms(st_raw, (i + 1) * 128 + sq < len);
        const int nqr = (i + 2) * 128 + sq;
        st_raw = (nqr < len) ? st_src[nqr] : 0.f;
        fence_proxy_async_smem();
      }
      {
        uint32_t va[32];
        tmem_ld_32x32(tmem_base + lane_base + qq * 32, va);
        tmem_ld_wait();
        CX_TRW(i, 2);
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          float2 xa = fmul2(make_float2(__uint_as_float(va[2 * t]), __uint_as_float(va[2 * t + 1])), sc2);
          float2 xb = fmul2(make_float2(__uint_as_float(va[2 * t + 2]), __uint_as_float(va[2 * t + 3])), sc2);
          xa = make_float2(fast_exp2(xa.x), fast_exp2(xa.y));
          if (kBwd4Poly == 8 || (kBwd4Poly == 4 && ((t >> 1) & 1))) xb = exp2_poly2(xb);
          else xb = make_float2(fast_exp2(xb.x), fast_exp2(xb.y));
          pp[t] = pack_bf16x2(xa.x, xa.y);
          pp[t + 1] = pack_bf16x2(xb.x, xb.y);
        }
      }
      if (!key_ok) {  // keys past the sequence end contribute nothing
#pragma unroll
        for (int t = 0; t < 16; ++t) pp[t] = 0u;
      }
      tmem_st_32x16(tmem_base + lane_base + qq * 32, pp);  // over this thread's own (already loaded) score columns
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(p_ready);
      CX_TRW(i, 3);
      // ---- Y: dS^T = P^T * (dP^T - delta[q]) * scale -> TMEM (A of dK) and smem (A of dQ)
      mbar_wait(dp_full, i & 1);
      tc_fence_after();
      if (qq == 2 && i + 1 < nq) {  // made visible to the MMA by the proxy fence ahead of the ds_ready arrival below
        *reinterpret_cast<uint4*>(aux_row + ((4 ^ (sq & 7)) << 4)) = stat_terms(st_raw, (i + 1) * 128 + sq < len);
        const int nqr = (i + 2) * 128 + sq;
        st_raw = (nqr < len) ? st_src[nqr] : 0.f;
      }
      if (i > 0) mbar_wait_quiet(dq_full, (i - 1) & 1);  // dQ(i-1) has finished reading the dS buffer
      CX_TRW(i, 4);
      {
        uint32_t vd[32];
        tmem_ld_32x32(tmem_base + lane_base + 128 + qq * 32, vd);
        tmem_ld_wait();
        CX_TRW(i, 5);
        uint32_t w[16];
#pragma unroll
        for (int t = 0; t < 16; t += 2) {
          const float2 ga = fmul2(make_float2(__uint_as_float(vd[2 * t]), __uint_as_float(vd[2 * t + 1])), ss2);
          const float2 gb = fmul2(make_float2(__uint_as_float(vd[2 * t + 2]), __uint_as_float(vd[2 * t + 3])), ss2);
          const float2 da = fmul2(unpack_bf16x2(pp[t]), ga);
          const float2 db = fmul2(unpack_bf16x2(pp[t + 1]), gb);
          w[t] = pack_bf16x2(da.x, da.y);
          w[t + 1] = pack_bf16x2(db.x, db.y);
        }
#pragma unroll
        for (int q = 0; q < 4; ++q)
          *reinterpret_cast<uint4*>(dd + (((ch0 + q) ^ (r & 7)) << 4)) = make_uint4(w[4 * q], w[4 * q + 1], w[4 * q + 2], w[4 * q + 3]);
        tmem_st_32x16(tmem_base + lane_base + 128 + qq * 32, w);  // over this thread's own (already loaded) dP^T columns
      }
      CX_TRW(i, 6);
      fence_proxy_async_smem();
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(ds_ready);
      CX_TRW(i, 7);
    }
    // dV (quarters 0, 1) / dK (quarters 2, 3) -> bf16 -> the (dead) dS buffer, half a swizzled 128-byte row per thread; then each
    // group of 256 threads copies its tile out with row-contiguous 16-byte stores
    mbar_wait(acc_full, 0);
    tc_fence_after();
    if (warp == 4) CX_TR(4, 0, 2);
    const int grp = qq >> 1, hf = qq & 1;
    uint8_t* stg = smem + Bwd4Smem::kDS + grp * 16384;
    if (grp == 1 && rope_inv_freq != nullptr) {
      // dK with the transposed rotary embedding (the keys were rotated before the scores were formed): a thread owns 16 of the 32
      // (x1, x2) = (column j, column 32 + j) pairs of its key row; the key's position is its row index in the sequence
      //   d x1 = g1 cos + g2 sin,  d x2 = g2 cos - g1 sin      (forward: o1 = x1 cos - x2 sin, o2 = x2 cos + x1 sin)
      uint32_t v1[16], v2[16];
      tmem_ld_32x16(tmem_base + lane_base + 320 + hf * 16, v1);
      tmem_ld_32x16(tmem_base + lane_base + 320 + 32 + hf * 16, v2);
      tmem_ld_wait();
      const float posf = (float)(k0 + r);
#pragma unroll
      for (int j = 0; j < 16; j += 2) {
        const float2 f2 = __ldg(reinterpret_cast<const float2*>(rope_inv_freq + hf * 16 + j));
        const float a0 = posf * f2.x, a1 = posf * f2.y;
        const float c0 = __cosf(a0), s0 = __sinf(a0), c1 = __cosf(a1), s1 = __sinf(a1);
        const float g10 = __uint_as_float(v1[j]), g11 = __uint_as_float(v1[j + 1]);
        const float g20 = __uint_as_float(v2[j]), g21 = __uint_as_float(v2[j + 1]);
        v1[j] = __float_as_uint(g10 * c0 + g20 * s0);
        v1[j + 1] = __float_as_uint(g11 * c1 + g21 * s1);
        v2[j] = __float_as_uint(g20 * c0 - g10 * s0);
        v2[j + 1] = __float_as_uint(g21 * c1 - g11 * s1);
      }
#pragma unroll
      for (int q = 0; q < 2; ++q) {
        uint4 w1, w2;
        w1.x = pack_bf16x2(__uint_as_float(v1[8 * q + 0]), __uint_as_float(v1[8 * q + 1]));
        w1.y = pack_bf16x2(__uint_as_float(v1[8 * q + 2]), __uint_as_float(v1[8 * q + 3]));
        w1.z = pack_bf16x2(__uint_as_float(v1[8 * q + 4]), __uint_as_float(v1[8 * q + 5]));
        w1.w = pack_bf16x2(__uint_as_float(v1[8 * q + 6]), __uint_as_float(v1[8 * q + 7]));
        w2.x = pack_bf16x2(__uint_as_float(v2[8 * q + 0]), __uint_as_float(v2[8 * q + 1]));
        w2.y = pack_bf16x2(__uint_as_float(v2[8 * q + 2]), __uint_as_float(v2[8 * q + 3]));
        w2.z = pack_bf16x2(__uint_as_float(v2[8 * q + 4]), __uint_as_float(v2[8 * q + 5]));
        w2.w = pack_bf16x2(__uint_as_float(v2[8 * q + 6]), __uint_as_float(v2[8 * q + 7]));
        *reinterpret_cast<uint4*>(stg + r * 128 + (((hf * 2 + q) ^ (r & 7)) << 4)) = w1;
        *reinterpret_cast<uint4*>(stg + r * 128 + (((4 + hf * 2 + q) ^ (r & 7)) << 4)) = w2;
      }
    } else {
      uint32_t v[32];
      tmem_ld_32x32(tmem_base + lane_base + 256 + grp * 64 + hf * 32, v);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        uint4 w;
        w.x = pack_bf16x2(__uint_as_float(v[8 * q + 0]), __uint_as_float(v[8 * q + 1]));
        w.y = pack_bf16x2(__uint_as_float(v[8 * q + 2]), __uint_as_float(v[8 * q + 3]));
        w.z = pack_bf16x2(__uint_as_float(v[8 * q + 4]), __uint_as_float(v[8 * q + 5]));
        w.w = pack_bf16x2(__uint_as_float(v[8 * q + 6]), __uint_as_float(v[8 * q + 7]));
        *reinterpret_cast<uint4*>(stg + r * 128 + (((hf * 4 + q) ^ (r & 7)) << 4)) = w;
      }
    }
    named_bar_sync(2 + grp, 256);
    {
      const int tid = hf * 128 + r;
      const int rows_ok = min(128, len - k0);
      uint8_t* obase = reinterpret_cast<uint8_t*>(dqkv + ((size_t)(seq_begin + k0) * 3 + (grp == 0 ? 2 : 1)) * H * kDh + (size_t)head * kDh);
#pragma unroll
      for (int it = 0; it < 4; ++it) {
        const int idx = it * 256 + tid, row = idx >> 3, ch = idx & 7;
        if (row < rows_ok)
          *reinterpret_cast<uint4*>(obase + (size_t)row * 3 * H * kDh * 2 + ch * 16) =
              *reinterpret_cast<const uint4*>(stg + row * 128 + ((ch ^ (row & 7)) << 4));
      }
    }
    if (warp == 4) CX_TR(4, 0, 3);
  } else if (warp >= 20) {
    // ---------------------------------------------------------------- dQ drain: one thread per query row, 64 columns
    const int ew = warp & 3;
    const int r = ew * 32 + lane;
    const int etid = threadIdx.x - 640;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    uint8_t* stage = smem + Bwd4Smem::kDQ;  // two [128 x 32] fp32 boxes (128-byte rows, 128B swizzle)
    for (int i = 0; i < nq; ++i) {
      const bool row_ok = i * 128 + r < len;
      if (warp == 20) CX_TR(3, i, 0);
      mbar_wait(dq_full, i & 1);
      tc_fence_after();
      if (warp == 20) CX_TR(3, i, 1);
      uint32_t va[32], vb[32];
      tmem_ld_32x32(tmem_base + lane_base + 384, va);
      tmem_ld_32x32(tmem_base + lane_base + 384 + 32, vb);
      tmem_ld_wait();
      tc_fence_before();
      mbar_arrive(dq_free);  // the next tile's dQ MMA may overwrite the columns
      if (warp == 20) CX_TR(3, i, 2);
      if (etid == 0) tma_store_wait_read<0>();  // the previous reduce-add has finished reading the stage
      named_bar_sync(1, 128);
      if (warp == 20) CX_TR(3, i, 3);
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
      if (warp == 20) CX_TR(3, i, 4);
    }
    if (etid == 0) tma_store_wait_read<0>();
    if (warp == 20) CX_TR(4, 0, 5);
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) CX_TR(4, 0, 4);
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

