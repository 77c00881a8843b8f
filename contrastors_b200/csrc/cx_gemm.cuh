// Persistent warp-specialised tcgen05 GEMM for sm_100a with pluggable epilogues.
//
//   C[M,N] = A (x) B,  bf16 operands via TMA (128B swizzle), fp32 accumulators in TMEM (double-buffered),
//   one CTA per SM looping over 128 x BLOCK_N output tiles.
//
// Warp roles (384 threads):  warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//                            warps 4..11 = epilogue: warp w owns TMEM lanes 32*(w%4).. (rows) and column half (w-4)/4.
// Pipelines: smem ring full/empty (TMA <-> MMA), TMEM full/empty (MMA <-> epilogue).
//
// Operand layouts (CX_MAJOR_K / CX_MAJOR_MN, see include/contrastors_b200.h):
//   K-major  : smem tile = rows x 64 K-elements (128 B per row), descriptor SBO = 1024 B, K-step = +32 B
//   MN-major : smem tile = (BLOCK/64) atoms, each 64 K-rows x 64 MN-elements (128 B per row);
//              descriptor LBO = 8192 B (next MN atom), SBO = 1024 B (next 8 K rows), K-step (16 rows) = +2048 B
//
// Epilogue modes:
//   EPI_STORE      C = alpha * acc  as bf16 or fp32 through a swizzled smem stage + TMA store (or TMA reduce-add)
//   EPI_NCE_STATS  InfoNCE forward: per-row (max, sum-exp, first-argmax, label logit) partials per column tile
//   EPI_SWIGLU     gated MLP first layer: B tile = 128 rows of fc11 (y) + the matching 128 rows of fc12 (gate);
//                  epilogue writes out = y * silu(gate) (bf16 [M, N]) and optionally the pre-activations [y | gate]
//   EPI_SWIGLU_BWD gated MLP backward: acc = d(act) tile (the fc2 dgrad GEMM, N = gated width I); the epilogue TMA-loads the
//                  matching [y | gate] pre-activation tiles, forms dy = da silu(g), dg = da y silu'(g) in place and TMA-stores
//                  both into dyg [M, 2N]: the d(act) tensor never exists in HBM and the separate swiglu_bwd pass (1.0 GB of
//                  traffic per layer at T = 32768) is gone
//   EPI_NCE_DS     InfoNCE backward stage 1: dS = softmax - onehot stored UNSCALED as fp16, plus
//                  per-thread partial of sum dS*s (the logit-scale gradient)
#pragma once
#include "cx_host.h"
#include "cx_ptx.cuh"

namespace cx {

enum EpiMode { EPI_STORE = 0, EPI_NCE_STATS = 1, EPI_NCE_DS = 2, EPI_SWIGLU = 3, EPI_SWIGLU_BWD = 4 };

struct EpiParams {
  float alpha = 1.f;
  const float* alpha_dev = nullptr;  // optional device scalars multiplied into alpha (keep the host sync-free)
  const float* alpha_dev2 = nullptr;
  // InfoNCE
  float scale = 1.f;
  const float* scale_dev = nullptr;  // optional device scalar multiplied into scale
  const float* coef_dev = nullptr;   // optional device scalar multiplied into coef (autograd's grad_output)
  const float* rq = nullptr;
  const float* rd = nullptr;
  int label_offset = 0;
  int label_stride = 1;
  const float* lse = nullptr;
  float coef = 0.f;
  float* part_max = nullptr;   // [2 * n_col_tiles][M]  (one partial per column half of each tile, log2 domain)
  float* part_sum = nullptr;   // [2 * n_col_tiles][M]
  int* part_arg = nullptr;     // [2 * n_col_tiles][M]
  float* label_logit = nullptr;  // [M]
  float* dlogit_part = nullptr;  // [gridDim.x]
  // EPI_STORE bf16 with rotary embedding fused (QKV projection): columns [0, rope_cols) are heads of 64 whose halves
  // (x1 = first 32, x2 = last 32) are rotated by the angle rope_pos[row] * rope_inv_freq[j], j < 32 (layers/embedding.py:
  // 685-706).  The angle is formed in fp32 exactly as the reference's cos/sin cache does (outer(t, inv_freq)); cos/sin come
  // from the SFU (sin.approx / cos.approx: |err| < 1e-4 at positions <= 8192, far inside the bf16 rounding the reference applies to
  // its cached tables).  Round 1 gathered them from fp32 tables per row: 32 uncoalesced loads per thread per head, 219 us vs
  // 89 + 35 us for GEMM + standalone rotary; computing them costs 128 SFU ops per thread per tile under the next tile's MMAs.
  const float* bias = nullptr;  // EPI_STORE bf16: out = alpha * acc + bias[col]  (FusedDense: layers/attention.py:82-85, mlp.py:24)
  const int* rope_pos = nullptr;
  const float* rope_inv_freq = nullptr;  // [32] = base^(-2j/64)
  int rope_cols = 0;
  // EPI_SWIGLU (N = number of gated output columns)
  __nv_bfloat16* act_out = nullptr;  // [M, N]
  int64_t ld_act = 0;
  __nv_bfloat16* yg_out = nullptr;   // [M, 2N] = [y | gate], nullptr = do not keep the pre-activations
  int64_t ld_yg = 0;
  int ab_f16 = 0;  // A and B operands hold IEEE fp16 instead of bf16 (kind::f16 wants one input type): fp16 dS path
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 384;  // 4 control warps + 8 epilogue warps (2 per SMSP: TLP hides TMEM-load latency)
constexpr int kStageCBytes = 16384;  // 128 rows x 128 B

// NP = 2 (QUAD, opt-in): a cluster of FOUR CTAs = two such pairs stacked in M (a 512 x 256 super tile).  The pairs need the same
// B tile, so each B half (128 rows) is fetched ONCE per cluster: the two CTAs that hold the same half each load 64 of its rows
// and TMA-multicast them to both (48 instead of 64 KB from L2 per pair per k-block).  Hypothesis: the pair kernel's 6.3 KB/clk
// of L2->SM traffic (profiles/r01_gemm_final_ncu_full_raw.csv: 10.3 TB/s at 1.64 GHz) sits at the fabric's cap.  Measured in
// round 2: it does not -- the multicast form is 0-8 % slower (see gemm_use_quad in cx_gemm.cu), so pairs stay the default.
// PAIR: two CTAs of a cluster cooperate on one 256 x 256 tile (tcgen05 cta_group::2): each CTA stages its own 128 rows of
// A and 128 of the 256 B rows (so the B operand is read from shared memory once per SM pair), the leader CTA issues the
// MMAs for both, each CTA drains the 128 x 256 half of the accumulator that lives in its own TMEM.
template <int BLOCK_N, bool PAIR = false, int MODE = 0>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = (PAIR ? BLOCK_N / 2 : BLOCK_N) * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  // pair mode trades one pipeline stage (32 KB) for ping-pong epilogue staging (2 buffers per column half), so a TMA
  // store never has to drain before the next 64 columns are packed
  // EPI_SWIGLU_BWD (pairs only): the epilogue streams [y | gate] tiles through shared memory, two 64-column steps per group
  // double-buffered = 8 x 16 KB, paid for with two pipeline stages (the kernel is bound by that stream, not by the mainloop)
  static constexpr bool kStreamEpi = MODE == 4;
  static constexpr int kCBufs = kStreamEpi ? 8 : (PAIR ? 4 : 2);
  static constexpr int kStages = kStreamEpi ? 3 : (PAIR ? 5 : ((BLOCK_N == 256) ? 4 : 6));
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = 1024 + kStages * kStageBytes + kCBufs * kStageCBytes + kBarrierBytes;
};

template <int BLOCK_N, bool A_MN, bool B_MN, int MODE, bool OUT_F32, bool ACCUM, bool PAIR = false, int NP = 1>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, const __grid_constant__ CUtensorMap tmD, int M, int N, int K,
            int splits, EpiParams ep) {
  using S = GemmSmem<BLOCK_N, PAIR, MODE>;
  static_assert(MODE != EPI_SWIGLU_BWD || PAIR, "the swiglu-backward epilogue is sized for CTA pairs");
  static_assert(!PAIR || BLOCK_N == 256, "CTA pairs use 256-column tiles");
  static_assert(NP == 1 || (NP == 2 && PAIR), "a quad cluster is two CTA pairs");
  constexpr bool QUAD = NP == 2;
  constexpr int kStages = S::kStages;
  constexpr int TILE_M = (PAIR ? 2 * kBlockM : kBlockM) * NP;  // rows per work item (the whole cluster's)
  const uint32_t cluster_rank = PAIR ? cluster_ctarank() : 0u;
  const uint32_t cta_rank = cluster_rank & 1u;      // rank within the CTA pair (0 = the leader that issues the MMAs)
  const uint32_t pair_id = cluster_rank >> 1;       // which pair of the cluster (always 0 unless QUAD)
  const int pair_m = (int)pair_id * 2 * kBlockM;    // row offset of this pair inside the work item
  const int tile_start = PAIR ? (int)cluster_id_x() : (int)blockIdx.x;
  const int tile_step = PAIR ? (int)num_clusters_x() : (int)gridDim.x;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static_assert(kTmemCols <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + kStages * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + S::kCBufs * kStageCBytes);
  uint64_t* full_bar = bars;                   // [kStages]
  uint64_t* empty_bar = bars + kStages;        // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;    // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + TILE_M - 1) / TILE_M;
  constexpr int TILE_N = (MODE == EPI_SWIGLU) ? BLOCK_N / 2 : BLOCK_N;  // output columns per tile
  const int n_tiles = (N + TILE_N - 1) / TILE_N;
  const int num_tiles = m_tiles * n_tiles * splits;  // split-K slices are separate work items (reduce-add epilogue)
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE != EPI_NCE_STATS) tma_prefetch_desc(&tmC);
    if (MODE == EPI_SWIGLU || MODE == EPI_SWIGLU_BWD) tma_prefetch_desc(&tmD);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);   // pair: only the leader arrives (expect_tx covers both CTAs' bytes; a peer arrive per
                                    // k-block would put a cluster-scope release fence on the producer's critical path)
      mbar_init(&empty_bar[i], NP);  // one commit per pair leader: with multicast loads a stage is shared by both pairs
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], PAIR ? 257 : 256);  // pair: the leader's 256 epilogue threads + ONE forwarded arrive from the peer
    }
    if (MODE == EPI_SWIGLU_BWD) {  // "pre-activation tiles landed": one barrier per epilogue group and staging slot
      for (int i = 0; i < 4; ++i) mbar_init(bars + 16 + i, 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) {
    if (PAIR) tmem_alloc_2cta<kTmemCols>(tmem_ptr);
    else tmem_alloc<kTmemCols>(tmem_ptr);
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all();
  else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer (converged warp; one elected lane issues:
    // code under `lane == 0` makes ptxas wrap every TMA / MMA / commit in an elect-and-branch loop, see cx_attn_fwd.cuh)
    {
      int stage = 0;
      uint32_t phase = 0;
      // (an L2-prefetch cursor running 8 k-blocks ahead of the ring -- cp.async.bulk.prefetch.tensor -- was measured
      //  SLOWER on B200: 1.35 vs 1.46 PFLOP/s at 8192^3, so the ring loads are the only TMA traffic)
      for (int tile = tile_start; tile < num_tiles; tile += tile_step) {
        const int mn = tile / splits, ks = tile % splits;
        const int m0 = (mn / n_tiles) * TILE_M + pair_m + (int)cta_rank * kBlockM;
        const int n0 = (mn % n_tiles) * TILE_N;
        const int kb0 = (int)((long long)ks * num_kb / splits), kb1 = (int)((long long)(ks + 1) * num_kb / splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::kStageBytes;
          uint8_t* sB = sA + S::kABytes;
          if (elect_one()) {
          if (PAIR) {
            // both CTAs' TMA bytes are credited to the LEADER's full barrier
            const uint32_t lead_bar = smem_u32(&full_bar[stage]) & 0xFEFFFFFFu;
            if (cta_rank == 0) mbar_arrive_expect_tx(&full_bar[stage], 2 * S::kStageBytes);
            if (!A_MN) {
              tma_load_2d_2sm(sA, &tmA, lead_bar, kb * kBlockK, m0);
            } else {
#pragma unroll
              for (int i = 0; i < kBlockM / 64; ++i) tma_load_2d_2sm(sA + i * 8192, &tmA, lead_bar, m0 + i * 64, kb * kBlockK);
            }
            // this CTA's half of the B rows: pair rank r stages rows [n0 + 128 r, +128) (SwiGLU: r = 0 -> fc11, r = 1 -> fc12)
            const int nb = (MODE == EPI_SWIGLU) ? n0 + (int)cta_rank * N : n0 + (int)cta_rank * (BLOCK_N / 2);
            if (QUAD) {
              // this CTA and the CTA of the same pair rank in the other pair hold the same 128 B rows: each fetches 64 of them
              // (8 KB: rows [64 p, +64) K-major / the p-th 64-column atom MN-major) and multicasts to both
              const uint16_t mc = (uint16_t)((1u << cta_rank) | (1u << (2u + cta_rank)));
              if (!B_MN) tma_load_2d_2sm_mc(sB + pair_id * 8192, &tmB, lead_bar, kb * kBlockK, nb + (int)pair_id * 64, mc);
              else tma_load_2d_2sm_mc(sB + pair_id * 8192, &tmB, lead_bar, nb + (int)pair_id * 64, kb * kBlockK, mc);
            } else if (!B_MN) {
              tma_load_2d_2sm(sB, &tmB, lead_bar, kb * kBlockK, nb);
            } else {
#pragma unroll
              for (int i = 0; i < BLOCK_N / 128; ++i) tma_load_2d_2sm(sB + i * 8192, &tmB, lead_bar, nb + i * 64, kb * kBlockK);
            }
          } else {
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * kBlockK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i) tma_load_2d(sA + i * 8192, &tmA, &full_bar[stage], m0 + i * 64, kb * kBlockK);
          }
          if (MODE == EPI_SWIGLU) {  // rows [n0, n0+128) of fc11 and of fc12 (stored N rows further down)
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * kBlockK, n0);
            tma_load_2d(sB + (BLOCK_N / 2) * 128, &tmB, &full_bar[stage], kb * kBlockK, N + n0);
          } else if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * kBlockK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i) tma_load_2d(sB + i * 8192, &tmB, &full_bar[stage], n0 + i * 64, kb * kBlockK);
          }
          }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (converged warp, one elected lane issues)
    if (cta_rank == 0) {  // pair: only the leader CTA issues (for both)
      // a_format [7,10) / b_format [10,13): 1 = bf16, 0 = f16
      const uint32_t idesc = make_idesc_bf16(PAIR ? 2 * kBlockM : kBlockM, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u) & ~(ep.ab_f16 ? ((1u << 7) | (1u << 10)) : 0u);
      const uint16_t all_ctas = QUAD ? 0xF : 0x3;                     // a consumed stage is released in every CTA of the cluster
      const uint16_t my_pair = (uint16_t)(0x3u << (2u * pair_id));    // the accumulator-ready signal stays inside the pair
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = tile_start; tile < num_tiles; tile += tile_step, ++it) {
        const int acc = it & 1;
        const uint32_t acc_phase = (it >> 1) & 1;
        mbar_wait(&tempty_bar[acc], acc_phase ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        const int ks = tile % splits;
        const int kb0 = (int)((long long)ks * num_kb / splits), kb1 = (int)((long long)(ks + 1) * num_kb / splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * S::kStageBytes);
          const uint32_t b_addr = a_addr + S::kABytes;
          // descriptor of (base + off) = descriptor of base + (off >> 4): the address field never carries out of 14 bits
          const uint64_t adesc0 = A_MN ? make_smem_desc_sw128(a_addr, 8192, 1024) : make_smem_desc_sw128(a_addr, 0, 1024);
          const uint64_t bdesc0 = B_MN ? make_smem_desc_sw128(b_addr, 8192, 1024) : make_smem_desc_sw128(b_addr, 0, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < kBlockK / kUmmaK; ++k) {
              const uint64_t adesc = adesc0 + (uint64_t)((A_MN ? k * 2048 : k * 32) >> 4);
              const uint64_t bdesc = bdesc0 + (uint64_t)((B_MN ? k * 2048 : k * 32) >> 4);
              if (PAIR) umma_f16_ss_2cta(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
              else umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
            }
            if (PAIR) {
              umma_commit_2cta(&empty_bar[stage], all_ctas);
              if (kb == kb1 - 1) umma_commit_2cta(&tfull_bar[acc], my_pair);
            } else {
              umma_commit(&empty_bar[stage]);
              if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
            }
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (2 x 128 threads, one group per column half)
    constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
    constexpr int HALF_N = BLOCK_N / 2;
    constexpr int NC = HALF_N / 32;
    const int ew = warp & 3;            // TMEM lane quarter this warp may access
    const int hf = (warp - 4) >> 2;     // column half handled by this warp group
    const int row_in_tile = ew * 32 + lane;
    const int etid = (threadIdx.x - 128) & 127;  // thread index within the half's group
    // staging: one buffer per column half (single-CTA mode) or two that ping-pong (pair mode)
    constexpr int kPing = S::kCBufs / 2;
    uint8_t* const stage_base = smem_c + hf * kPing * kStageCBytes;
    int cb = 0;
#define CX_STAGE_ACQUIRE()                                         \
  uint8_t* stage_c = stage_base + cb * kStageCBytes;               \
  if (etid == 0) {                                                 \
    if (kPing == 2) tma_store_wait_read<1>();                      \
    else tma_store_wait_read<0>();                                 \
  }                                                                \
  cb = (kPing == 2) ? (cb ^ 1) : 0;
    int it = 0;
    uint32_t ld_phase = 0;  // EPI_SWIGLU_BWD: parity of this group's pre-activation-load barrier
    float dl0 = 0.f, dl1 = 0.f, dl2 = 0.f, dl3 = 0.f;  // sum p*t (log2 domain) for the logit-scale gradient
    const float ep_scale = ep.scale * (ep.scale_dev != nullptr ? *ep.scale_dev : 1.f);
    const float ep_coef = ep.coef * (ep.coef_dev != nullptr ? *ep.coef_dev : 1.f);
    const float ep_alpha = ep.alpha * (ep.alpha_dev != nullptr ? *ep.alpha_dev : 1.f) * (ep.alpha_dev2 != nullptr ? *ep.alpha_dev2 : 1.f);
    if (MODE == EPI_SWIGLU_BWD && etid == 0 && tile_start < num_tiles) {  // the first work item's [y | gate] tiles, both steps
      const int mn1 = tile_start / splits;
      const int m1 = (mn1 / n_tiles) * TILE_M + pair_m + (int)cta_rank * kBlockM, n1 = (mn1 % n_tiles) * BLOCK_N + hf * HALF_N;
#pragma unroll
      for (int pr = 0; pr < 2; ++pr) {
        mbar_arrive_expect_tx(bars + 16 + 2 * hf + pr, 2 * kStageCBytes);
        tma_load_2d(stage_base + (2 * pr) * kStageCBytes, &tmD, bars + 16 + 2 * hf + pr, n1 + pr * 64, m1);
        tma_load_2d(stage_base + (2 * pr + 1) * kStageCBytes, &tmD, bars + 16 + 2 * hf + pr, N + n1 + pr * 64, m1);
      }
    }
    for (int tile = tile_start; tile < num_tiles; tile += tile_step, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mn = tile / splits;
      const int mt = mn / n_tiles, nt = mn % n_tiles;
      const int m0 = mt * TILE_M + pair_m + (int)cta_rank * kBlockM, n0 = (MODE == EPI_SWIGLU) ? nt * TILE_N + hf * (TILE_N / 2) : nt * BLOCK_N + hf * HALF_N;
      const int row = m0 + row_in_tile;
      const bool row_ok = row < M;
      // fast path: no column masks and no per-column scale anywhere in this tile
      const bool plain = (n0 + HALF_N <= N) && ep.rd == nullptr;

      // per-row InfoNCE state, in the log2 domain: t = s * log2(e)
      float rs2 = 1.f, lse2 = 0.f;
      int label = -1;
      float run_max = -INFINITY, run_sum = 0.f;
      int run_arg = 0;
      if (MODE != EPI_STORE) {
        const float rqi = (ep.rq != nullptr && row_ok) ? ep.rq[row] : 1.f;
        rs2 = ep_scale * rqi * kLog2e;
        label = (row + ep.label_offset) * ep.label_stride;
        if (MODE == EPI_NCE_DS) lse2 = row_ok ? ep.lse[row] * kLog2e : INFINITY;  // +inf => p == 0 for rows past M
      }

      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N + hf * HALF_N;

      if (MODE == EPI_SWIGLU) {
        // this group owns output columns [n0, n0 + 64): y at TMEM cols hf*64 + .., gate at 128 + hf*64 + ..
        // Three passes over the same TMEM columns (act, then y, then gate when the pre-activations are kept); each
        // pass packs one 128 x 64 bf16 tile into this group's swizzled staging buffer and TMA-stores it.
        const uint32_t ty = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N + hf * (TILE_N / 2);
        const int npass = (ep.yg_out != nullptr) ? 3 : 1;
#pragma unroll 1
        for (int pass = 0; pass < npass; ++pass) {
          CX_STAGE_ACQUIRE();
          named_bar_sync(1 + hf, 128);
          uint8_t* dst = stage_c + row_in_tile * 128;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t vy[32], vg[32];
            if (pass != 2) tmem_ld_32x32(ty + c * 32, vy);
            if (pass != 1) tmem_ld_32x32(ty + TILE_N + c * 32, vg);
            tmem_ld_wait();
            uint32_t pk[16];
            if (pass == 0) {
#pragma unroll
              for (int j = 0; j < 16; ++j) {
                const float y0 = __uint_as_float(vy[2 * j]), y1 = __uint_as_float(vy[2 * j + 1]);
                const float g0 = __uint_as_float(vg[2 * j]), g1 = __uint_as_float(vg[2 * j + 1]);
                pk[j] = pack_bf16x2(y0 * __fdividef(g0, 1.f + fast_exp2(-g0 * kLog2e)),
                                    y1 * __fdividef(g1, 1.f + fast_exp2(-g1 * kLog2e)));
              }
            } else if (pass == 1) {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(vy[2 * j]), __uint_as_float(vy[2 * j + 1]));
            } else {
#pragma unroll
              for (int j = 0; j < 16; ++j) pk[j] = pack_bf16x2(__uint_as_float(vg[2 * j]), __uint_as_float(vg[2 * j + 1]));
            }
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const int chunk = c * 4 + q;
              *reinterpret_cast<uint4*>(dst + ((chunk ^ (row_in_tile & 7)) << 4)) =
                  make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
            }
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + hf, 128);
          if (etid == 0) {
            if (pass == 0) tma_store_2d(&tmC, stage_c, n0, m0);
            else tma_store_2d(&tmD, stage_c, (pass == 2 ? N : 0) + n0, m0);
            tma_store_commit();
          }
        }
      } else if (MODE == EPI_SWIGLU_BWD) {
        // this group owns d(act) columns [n0, n0 + 128): two 64-column steps, each transforming one y tile and one gate tile
        // (128 rows x 128 B, swizzled exactly as TMA wrote them) IN PLACE; a thread only ever touches its own row.  Two staging
        // slots per group (slot = step): the tiles of BOTH steps are in flight before the accumulator is waited for, and the
        // next work item's tiles are requested as soon as a slot's stores have been read (rolling prefetch, see the tail)
#pragma unroll 1
        for (int pr = 0; pr < 2; ++pr) {
          const int col0 = n0 + pr * 64;
          uint8_t* const buf_y = stage_base + (2 * pr) * kStageCBytes;
          uint8_t* const buf_g = buf_y + kStageCBytes;
          uint32_t v1[32], v2[32];
          tmem_ld_32x32(taddr + pr * 64, v1);
          tmem_ld_32x32(taddr + pr * 64 + 32, v2);
          mbar_wait(bars + 16 + 2 * hf + pr, ld_phase);
          tmem_ld_wait();
          uint8_t* const ry = buf_y + row_in_tile * 128;
          uint8_t* const rg = buf_g + row_in_tile * 128;
#pragma unroll
          for (int q = 0; q < 8; ++q) {
            const int off = (q ^ (row_in_tile & 7)) << 4;
            const uint4 y4 = *reinterpret_cast<const uint4*>(ry + off);
            const uint4 g4 = *reinterpret_cast<const uint4*>(rg + off);
            const uint32_t yw[4] = {y4.x, y4.y, y4.z, y4.w}, gw[4] = {g4.x, g4.y, g4.z, g4.w};
            uint32_t oy[4], og[4];
#pragma unroll
            for (int e = 0; e < 4; ++e) {
              const float2 yv = unpack_bf16x2(yw[e]), gv = unpack_bf16x2(gw[e]);
              const int j = (q & 3) * 8 + 2 * e;
              const float d0 = __uint_as_float(q < 4 ? v1[j] : v2[j]) * ep_alpha, d1 = __uint_as_float(q < 4 ? v1[j + 1] : v2[j + 1]) * ep_alpha;
              const float s0 = __fdividef(1.f, 1.f + fast_exp2(-gv.x * kLog2e)), s1 = __fdividef(1.f, 1.f + fast_exp2(-gv.y * kLog2e));
              oy[e] = pack_bf16x2(d0 * gv.x * s0, d1 * gv.y * s1);
              og[e] = pack_bf16x2(d0 * yv.x * (s0 * (1.f + gv.x * (1.f - s0))), d1 * yv.y * (s1 * (1.f + gv.y * (1.f - s1))));
            }
            *reinterpret_cast<uint4*>(ry + off) = make_uint4(oy[0], oy[1], oy[2], oy[3]);
            *reinterpret_cast<uint4*>(rg + off) = make_uint4(og[0], og[1], og[2], og[3]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + hf, 128);
          if (etid == 0) {
            tma_store_2d(&tmC, buf_y, col0, m0);
            tma_store_2d(&tmC, buf_g, N + col0, m0);
            tma_store_commit();
            if (tile + tile_step < num_tiles) {
              // rolling prefetch: as soon as this slot's stores have been read (a few hundred ns), request the NEXT work item's
              // tiles of the same step into it -- they then have a whole step of this group's arithmetic to arrive
              const int mn2 = (tile + tile_step) / splits;
              const int m2 = (mn2 / n_tiles) * TILE_M + pair_m + (int)cta_rank * kBlockM, n2 = (mn2 % n_tiles) * BLOCK_N + hf * HALF_N + pr * 64;
              tma_store_wait_read<0>();
              mbar_arrive_expect_tx(bars + 16 + 2 * hf + pr, 2 * kStageCBytes);
              tma_load_2d(buf_y, &tmD, bars + 16 + 2 * hf + pr, n2, m2);
              tma_load_2d(buf_g, &tmD, bars + 16 + 2 * hf + pr, N + n2, m2);
            }
          }
        }
        ld_phase ^= 1u;
      } else if (MODE == EPI_STORE && !OUT_F32) {
        // bf16 store path: one staging row (128 B) = 64 columns = two TMEM chunks = one attention head when RoPE is on
        float rpos = 0.f;
        const bool rope_row = ep.rope_pos != nullptr && row_ok;
        if (rope_row) rpos = (float)ep.rope_pos[row];
#pragma unroll 1
        for (int pr = 0; pr < NC / 2; ++pr) {
          uint32_t v1[32], v2[32];
          tmem_ld_32x32(taddr + pr * 64, v1);
          tmem_ld_32x32(taddr + pr * 64 + 32, v2);
          tmem_ld_wait();
          const int col0 = n0 + pr * 64;
          uint32_t pk1[16], pk2[16];
          if (rope_row && col0 < ep.rope_cols) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              const float2 f2 = __ldg(reinterpret_cast<const float2*>(ep.rope_inv_freq) + j);  // same address in every thread
              const float g0 = rpos * f2.x, g1 = rpos * f2.y;
              const float2 c2 = make_float2(__cosf(g0), __cosf(g1));
              const float2 s2 = make_float2(__sinf(g0), __sinf(g1));
              const float a0 = __uint_as_float(v1[2 * j]) * ep_alpha, a1 = __uint_as_float(v1[2 * j + 1]) * ep_alpha;
              const float b0 = __uint_as_float(v2[2 * j]) * ep_alpha, b1 = __uint_as_float(v2[2 * j + 1]) * ep_alpha;
              pk1[j] = pack_bf16x2(a0 * c2.x - b0 * s2.x, a1 * c2.y - b1 * s2.y);
              pk2[j] = pack_bf16x2(b0 * c2.x + a0 * s2.x, b1 * c2.y + a1 * s2.y);
            }
          } else if (ep.bias != nullptr) {
#pragma unroll
            for (int j = 0; j < 16; ++j) {  // every thread reads the same addresses: one broadcast wavefront per load
              const int c1 = col0 + 2 * j, c2 = c1 + 32;
              const float b10 = c1 < N ? ep.bias[c1] : 0.f, b11 = c1 + 1 < N ? ep.bias[c1 + 1] : 0.f;
              const float b20 = c2 < N ? ep.bias[c2] : 0.f, b21 = c2 + 1 < N ? ep.bias[c2 + 1] : 0.f;
              pk1[j] = pack_bf16x2(fmaf(__uint_as_float(v1[2 * j]), ep_alpha, b10), fmaf(__uint_as_float(v1[2 * j + 1]), ep_alpha, b11));
              pk2[j] = pack_bf16x2(fmaf(__uint_as_float(v2[2 * j]), ep_alpha, b20), fmaf(__uint_as_float(v2[2 * j + 1]), ep_alpha, b21));
            }
          } else {
#pragma unroll
            for (int j = 0; j < 16; ++j) {
              pk1[j] = pack_bf16x2(__uint_as_float(v1[2 * j]) * ep_alpha, __uint_as_float(v1[2 * j + 1]) * ep_alpha);
              pk2[j] = pack_bf16x2(__uint_as_float(v2[2 * j]) * ep_alpha, __uint_as_float(v2[2 * j + 1]) * ep_alpha);
            }
          }
          CX_STAGE_ACQUIRE();
          named_bar_sync(1 + hf, 128);
          uint8_t* dst = stage_c + row_in_tile * 128;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            *reinterpret_cast<uint4*>(dst + ((q ^ (row_in_tile & 7)) << 4)) =
                make_uint4(pk1[4 * q], pk1[4 * q + 1], pk1[4 * q + 2], pk1[4 * q + 3]);
            *reinterpret_cast<uint4*>(dst + (((4 + q) ^ (row_in_tile & 7)) << 4)) =
                make_uint4(pk2[4 * q], pk2[4 * q + 1], pk2[4 * q + 2], pk2[4 * q + 3]);
          }
          fence_proxy_async_smem();
          named_bar_sync(1 + hf, 128);
          if (etid == 0) {
            tma_store_2d(&tmC, stage_c, col0, m0);
            tma_store_commit();
          }
        }
      } else
#pragma unroll 1
      for (int c = 0; c < NC; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;

        if (MODE == EPI_NCE_STATS) {
          float t[32];
          if (plain) {
#pragma unroll
            for (int j = 0; j < 32; ++j) t[j] = __uint_as_float(v[j]) * rs2;
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              const float r = (ep.rd != nullptr) ? rs2 * (col < N ? ep.rd[col] : 0.f) : rs2;
              t[j] = (col < N) ? __uint_as_float(v[j]) * r : -INFINITY;
            }
          }
          float m0_ = t[0], m1_ = t[1], m2_ = t[2], m3_ = t[3];
#pragma unroll
          for (int j = 4; j < 32; j += 4) {
            m0_ = fmaxf(m0_, t[j]);
            m1_ = fmaxf(m1_, t[j + 1]);
            m2_ = fmaxf(m2_, t[j + 2]);
            m3_ = fmaxf(m3_, t[j + 3]);
          }
          const float cmax = fmaxf(fmaxf(m0_, m1_), fmaxf(m2_, m3_));
          if (cmax > run_max) {  // first max wins: only a strictly larger value moves the argmax
            int arg = 0;
#pragma unroll
            for (int j = 31; j >= 0; --j) arg = (t[j] == cmax) ? j : arg;
            run_arg = col0 + arg;
            run_sum *= fast_exp2(run_max - cmax);
            run_max = cmax;
          }
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
          const float mref = (run_max == -INFINITY) ? 0.f : run_max;  // a fully masked half tile contributes exactly 0
#pragma unroll
          for (int j = 0; j < 32; j += 4) {
            s0 += fast_exp2(t[j] - mref);
            s1 += fast_exp2(t[j + 1] - mref);
            s2 += fast_exp2(t[j + 2] - mref);
            s3 += fast_exp2(t[j + 3] - mref);
          }
          run_sum += (s0 + s1) + (s2 + s3);
          if (label >= col0 && label < col0 + 32 && row_ok) {
            float lv = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) lv = (col0 + j == label) ? t[j] : lv;
            ep.label_logit[row] = lv * kLn2;
          }
        } else {
          // value transform
          if (MODE == EPI_NCE_DS) {
            // stores (softmax - onehot) in [-1, 1] UNSCALED as fp16 (11-bit mantissa); coef, the logit scale and the per-row
            // inverse norms of the normalised-prefix losses are applied by the two contractions (alpha and pre-normalised
            // fp16 B operands), so one dS serves dQ and dD on every loss variant
            const bool has_label = (label >= col0 && label < col0 + 32 && row_ok);
            float tl = 0.f;
            if (plain) {
#pragma unroll
              for (int j = 0; j < 32; j += 4) {
                const float t0 = __uint_as_float(v[j]) * rs2, t1 = __uint_as_float(v[j + 1]) * rs2;
                const float t2 = __uint_as_float(v[j + 2]) * rs2, t3 = __uint_as_float(v[j + 3]) * rs2;
                const float p0 = fast_exp2(t0 - lse2), p1 = fast_exp2(t1 - lse2), p2 = fast_exp2(t2 - lse2), p3 = fast_exp2(t3 - lse2);
                dl0 = fmaf(p0, t0, dl0);
                dl1 = fmaf(p1, t1, dl1);
                dl2 = fmaf(p2, t2, dl2);
                dl3 = fmaf(p3, t3, dl3);
                if (has_label) {
                  tl = (col0 + j == label) ? t0 : tl;
                  tl = (col0 + j + 1 == label) ? t1 : tl;
                  tl = (col0 + j + 2 == label) ? t2 : tl;
                  tl = (col0 + j + 3 == label) ? t3 : tl;
                }
                v[j] = __float_as_uint(p0);
                v[j + 1] = __float_as_uint(p1);
                v[j + 2] = __float_as_uint(p2);
                v[j + 3] = __float_as_uint(p3);
              }
            } else {
#pragma unroll
              for (int j = 0; j < 32; ++j) {
                const int col = col0 + j;
                const float rdj = (ep.rd != nullptr) ? (col < N ? ep.rd[col] : 0.f) : 1.f;
                const float t = __uint_as_float(v[j]) * rs2 * rdj;
                const float p = (col < N) ? fast_exp2(t - lse2) : 0.f;
                dl0 = fmaf(p, t, dl0);
                tl = (col == label) ? t : tl;
                v[j] = __float_as_uint(p);
              }
            }
            if (has_label) {  // subtract the one-hot: rare (one chunk per row per pass)
              dl0 -= tl;
#pragma unroll
              for (int j = 0; j < 32; ++j)
                if (col0 + j == label) v[j] = __float_as_uint(__uint_as_float(v[j]) - 1.f);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * ep_alpha);
          }
          // stage + TMA store.  fp32: one 32-col chunk = 128 B per row; 16-bit: two chunks = 128 B per row.
          if (OUT_F32) {
            CX_STAGE_ACQUIRE();
            named_bar_sync(1 + hf, 128);
            uint8_t* dst = stage_c + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 w = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              *reinterpret_cast<uint4*>(dst + ((j ^ (row_in_tile & 7)) << 4)) = w;
            }
            fence_proxy_async_smem();
            named_bar_sync(1 + hf, 128);
            if (etid == 0) {
              if (ACCUM) tma_reduce_add_2d(&tmC, stage_c, col0, m0);
              else tma_store_2d(&tmC, stage_c, col0, m0);
              tma_store_commit();
            }
          } else {
            const int half = c & 1;
            if (half == 0) {
              cb = (kPing == 2) ? (cb ^ 1) : 0;
              if (etid == 0) {
                if (kPing == 2) tma_store_wait_read<1>();
                else tma_store_wait_read<0>();
              }
              named_bar_sync(1 + hf, 128);
            }
            uint8_t* stage_c = stage_base + cb * kStageCBytes;
            uint8_t* dst = stage_c + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              if (MODE == EPI_NCE_DS) {
                w.x = pack_f16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                w.y = pack_f16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                w.z = pack_f16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                w.w = pack_f16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              } else {
                w.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
                w.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
                w.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
                w.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              }
              const int chunk = half * 4 + j;
              *reinterpret_cast<uint4*>(dst + ((chunk ^ (row_in_tile & 7)) << 4)) = w;
            }
            if (half == 1) {
              fence_proxy_async_smem();
              named_bar_sync(1 + hf, 128);
              if (etid == 0) {
                tma_store_2d(&tmC, stage_c, col0 - 32, m0);
                tma_store_commit();
              }
            }
          }
        }
      }
      // accumulator drained -> hand the TMEM buffer back to the MMA warp (pair: the leader CTA's barrier)
      tc_fence_before();
      if (PAIR && cta_rank != 0) {
        // peer CTA: gather its 256 epilogue threads locally, then ONE remote arrive (256 remote arrives per tile would
        // serialise on the cluster interconnect)
        named_bar_sync(3, 256);
        if (threadIdx.x == 128) mbar_arrive_cluster(mapa_u32(smem_u32(&tempty_bar[acc]), cluster_rank & ~1u));
      } else if (PAIR) {
        mbar_arrive(&tempty_bar[acc]);
      } else {
        mbar_arrive(&tempty_bar[acc]);
      }

      if (MODE == EPI_NCE_STATS && row_ok) {  // partials stay in the log2 domain; the combine kernel converts
        const size_t o = static_cast<size_t>(nt * 2 + hf) * M + row;
        ep.part_max[o] = run_max;
        ep.part_sum[o] = run_sum;
        ep.part_arg[o] = run_arg;
      }
    }
    if (MODE != EPI_NCE_STATS && etid == 0) tma_store_wait<0>();
    if (MODE == EPI_NCE_DS) {
      // deterministic per-CTA reduction of the logit-scale gradient partial: d/dlog(scale) = coef * ln2 * sum (p - 1hot) t
      float dl = (dl0 + dl1) + (dl2 + dl3);
      for (int o = 16; o > 0; o >>= 1) dl += __shfl_xor_sync(0xffffffffu, dl, o);
      float* red = reinterpret_cast<float*>(bars) + 48;  // spare bytes behind the barriers
      if (lane == 0) red[warp - 4] = dl;
      named_bar_sync(3, 256);
      if (threadIdx.x == 128)
        ep.dlogit_part[blockIdx.x] = (((red[0] + red[1]) + (red[2] + red[3])) + ((red[4] + red[5]) + (red[6] + red[7]))) * ep_coef * kLn2;
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all();  // neither CTA may free TMEM / exit while its peer can still touch it
  else __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    if (PAIR) tmem_dealloc_2cta<kTmemCols>(tmem_base);
    else tmem_dealloc<kTmemCols>(tmem_base);
  }
}

// ---------------------------------------------------------------- host launcher
struct GemmArgs {
  const void* A;
  const void* B;
  void* C;  // may be null for EPI_NCE_STATS
  int M, N, K;
  bool a_mn, b_mn;
  int64_t lda, ldb, ldc;  // in elements
  bool out_f32;
  bool accumulate;
  int splits = 0;  // 0 = choose automatically (split-K only for fp32 outputs)
  int mode;
  EpiParams ep;
  cudaStream_t stream;
};

int launch_gemm(const GemmArgs& g);
// number of CTAs launch_gemm will use for (M, N): needed to size dlogit_part
int gemm_grid(int M, int N, int splits = 1, int tile_n = 0);
int gemm_block_n(int N);

}  // namespace cx
