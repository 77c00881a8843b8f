// Persistent warp-specialised tcgen05 GEMM for sm_100a with pluggable epilogues.
//
//   C[M,N] = A (x) B,  bf16 operands via TMA (128B swizzle), fp32 accumulators in TMEM (double-buffered),
//   one CTA per SM looping over 128 x BLOCK_N output tiles.
//
// Warp roles (256 threads):  warp 0 = TMA producer, warp 1 = MMA issuer (one thread), warp 2 = TMEM allocator,
//                            warps 4..7 = epilogue (thread t <-> accumulator row t of the tile).
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
//   EPI_NCE_DS     InfoNCE backward stage 1: dS = coef * (softmax - onehot) (x rq_i rd_j) stored as bf16, plus
//                  per-thread partial of sum dS*s (the logit-scale gradient)
#pragma once
#include "cx_host.h"
#include "cx_ptx.cuh"

namespace cx {

enum EpiMode { EPI_STORE = 0, EPI_NCE_STATS = 1, EPI_NCE_DS = 2 };

struct EpiParams {
  float alpha = 1.f;
  const float* alpha_dev = nullptr;  // optional device scalar multiplied into alpha (keeps the host sync-free)
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
  float* part_max = nullptr;   // [n_col_tiles][M]
  float* part_sum = nullptr;   // [n_col_tiles][M]
  int* part_arg = nullptr;     // [n_col_tiles][M]
  float* label_logit = nullptr;  // [M]
  float* dlogit_part = nullptr;  // [gridDim.x * 128]
};

constexpr int kBlockM = 128;
constexpr int kBlockK = 64;
constexpr int kUmmaK = 16;
constexpr int kGemmThreads = 256;
constexpr int kStageCBytes = 16384;  // 128 rows x 128 B

template <int BLOCK_N>
struct GemmSmem {
  static constexpr int kABytes = kBlockM * kBlockK * 2;
  static constexpr int kBBytes = BLOCK_N * kBlockK * 2;
  static constexpr int kStageBytes = kABytes + kBBytes;
  static constexpr int kStages = (BLOCK_N == 256) ? 4 : 6;
  static constexpr int kBarrierBytes = 256;
  static constexpr int kTotal = 1024 + kStages * kStageBytes + 2 * kStageCBytes + kBarrierBytes;
};

template <int BLOCK_N, bool A_MN, bool B_MN, int MODE, bool OUT_F32, bool ACCUM>
__global__ void __launch_bounds__(kGemmThreads, 1)
gemm_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB,
            const __grid_constant__ CUtensorMap tmC, int M, int N, int K, int splits, EpiParams ep) {
  using S = GemmSmem<BLOCK_N>;
  constexpr int kStages = S::kStages;
  constexpr uint32_t kTmemCols = 2 * BLOCK_N;
  static_assert(kTmemCols <= 512, "TMEM budget");

  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* smem_c = smem + kStages * S::kStageBytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem_c + 2 * kStageCBytes);
  uint64_t* full_bar = bars;                   // [kStages]
  uint64_t* empty_bar = bars + kStages;        // [kStages]
  uint64_t* tfull_bar = bars + 2 * kStages;    // [2]
  uint64_t* tempty_bar = bars + 2 * kStages + 2;  // [2]
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 2 * kStages + 4);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int m_tiles = (M + kBlockM - 1) / kBlockM;
  const int n_tiles = (N + BLOCK_N - 1) / BLOCK_N;
  const int num_tiles = m_tiles * n_tiles * splits;  // split-K slices are separate work items (reduce-add epilogue)
  const int num_kb = (K + kBlockK - 1) / kBlockK;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
    if (MODE != EPI_NCE_STATS) tma_prefetch_desc(&tmC);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&tfull_bar[i], 1);
      mbar_init(&tempty_bar[i], 128);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<kTmemCols>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ------------------------------------------------------------ TMA producer
    if (lane == 0) {
      int stage = 0;
      uint32_t phase = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
        const int mn = tile / splits, ks = tile % splits;
        const int m0 = (mn / n_tiles) * kBlockM;
        const int n0 = (mn % n_tiles) * BLOCK_N;
        const int kb0 = (int)((long long)ks * num_kb / splits), kb1 = (int)((long long)(ks + 1) * num_kb / splits);
        for (int kb = kb0; kb < kb1; ++kb) {
          mbar_wait(&empty_bar[stage], phase ^ 1);
          uint8_t* sA = smem + stage * S::kStageBytes;
          uint8_t* sB = sA + S::kABytes;
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          if (!A_MN) {
            tma_load_2d(sA, &tmA, &full_bar[stage], kb * kBlockK, m0);
          } else {
#pragma unroll
            for (int i = 0; i < kBlockM / 64; ++i) tma_load_2d(sA + i * 8192, &tmA, &full_bar[stage], m0 + i * 64, kb * kBlockK);
          }
          if (!B_MN) {
            tma_load_2d(sB, &tmB, &full_bar[stage], kb * kBlockK, n0);
          } else {
#pragma unroll
            for (int i = 0; i < BLOCK_N / 64; ++i) tma_load_2d(sB + i * 8192, &tmB, &full_bar[stage], n0 + i * 64, kb * kBlockK);
          }
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ------------------------------------------------------------ MMA issuer (single thread)
    if (lane == 0) {
      constexpr uint32_t idesc = make_idesc_bf16(kBlockM, BLOCK_N, A_MN ? 1u : 0u, B_MN ? 1u : 0u);
      int stage = 0;
      uint32_t phase = 0;
      int it = 0;
      for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
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
#pragma unroll
          for (int k = 0; k < kBlockK / kUmmaK; ++k) {
            const uint64_t adesc = A_MN ? make_smem_desc_sw128(a_addr + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(a_addr + k * 32, 0, 1024);
            const uint64_t bdesc = B_MN ? make_smem_desc_sw128(b_addr + k * 2048, 8192, 1024)
                                        : make_smem_desc_sw128(b_addr + k * 32, 0, 1024);
            umma_f16_ss(d_tmem, adesc, bdesc, idesc, (kb > kb0 || k > 0) ? 1u : 0u);
          }
          umma_commit(&empty_bar[stage]);
          if (kb == kb1 - 1) umma_commit(&tfull_bar[acc]);
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp >= 4) {
    // ------------------------------------------------------------ epilogue (128 threads)
    const int ew = warp - 4;            // == warp % 4: TMEM lane quarter this warp may access
    const int row_in_tile = ew * 32 + lane;
    const int etid = threadIdx.x - 128;
    int it = 0;
    int cbuf = 0;
    float dlogit_acc = 0.f;
    const float ep_scale = ep.scale * (ep.scale_dev != nullptr ? *ep.scale_dev : 1.f);
    const float ep_coef = ep.coef * (ep.coef_dev != nullptr ? *ep.coef_dev : 1.f);
    const float ep_alpha = ep.alpha * (ep.alpha_dev != nullptr ? *ep.alpha_dev : 1.f);
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int acc = it & 1;
      const uint32_t acc_phase = (it >> 1) & 1;
      const int mn = tile / splits;
      const int mt = mn / n_tiles, nt = mn % n_tiles;
      const int m0 = mt * kBlockM, n0 = nt * BLOCK_N;
      const int row = m0 + row_in_tile;
      const bool row_ok = row < M;
      mbar_wait(&tfull_bar[acc], acc_phase);
      tc_fence_after();
      const uint32_t taddr = tmem_base + (static_cast<uint32_t>(ew * 32) << 16) + acc * BLOCK_N;

      // per-row InfoNCE state
      float rscale = 1.f, row_lse = 0.f;
      int label = -1;
      float run_max = -INFINITY, run_sum = 0.f;
      int run_arg = 0;
      if (MODE != EPI_STORE) {
        rscale = ep_scale * ((ep.rq != nullptr && row_ok) ? ep.rq[row] : 1.f);
        label = (row + ep.label_offset) * ep.label_stride;
        if (MODE == EPI_NCE_DS) row_lse = row_ok ? ep.lse[row] : 0.f;
      }

#pragma unroll 1
      for (int c = 0; c < BLOCK_N / 32; ++c) {
        uint32_t v[32];
        tmem_ld_32x32(taddr + c * 32, v);
        tmem_ld_wait();
        const int col0 = n0 + c * 32;

        if (MODE == EPI_NCE_STATS) {
          float s[32];
          float cmax = -INFINITY;
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            float x = __uint_as_float(v[j]) * rscale;
            if (ep.rd != nullptr) x *= (col < N ? ep.rd[col] : 0.f);
            x = (col < N) ? x : -INFINITY;
            s[j] = x;
            cmax = fmaxf(cmax, x);
          }
          if (cmax > run_max) {  // first max wins: only a strictly larger value moves the argmax
            int arg = 0;
#pragma unroll
            for (int j = 31; j >= 0; --j) arg = (s[j] == cmax) ? j : arg;
            run_arg = col0 + arg;
            run_sum *= exp2f((run_max - cmax) * 1.4426950408889634f);
            run_max = cmax;
          }
          if (run_max > -INFINITY) {
            const float mb = run_max * 1.4426950408889634f;
            float acc_s = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) acc_s += exp2f(fmaf(s[j], 1.4426950408889634f, -mb));
            run_sum += acc_s;
          }
          if (label >= col0 && label < col0 + 32 && row_ok) {
            float lv = 0.f;
#pragma unroll
            for (int j = 0; j < 32; ++j) lv = (col0 + j == label) ? s[j] : lv;
            ep.label_logit[row] = lv;
          }
        } else {
          // value transform
          if (MODE == EPI_NCE_DS) {
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              const float rdj = (ep.rd != nullptr) ? (col < N ? ep.rd[col] : 0.f) : 1.f;
              const float s = __uint_as_float(v[j]) * rscale * rdj;
              float p = exp2f((s - row_lse) * 1.4426950408889634f);
              p = (col == label) ? p - 1.f : p;
              float ds = ep_coef * p;
              ds = (col < N && row_ok) ? ds : 0.f;
              dlogit_acc = fmaf(ds, s, dlogit_acc);
              if (ep.rq != nullptr || ep.rd != nullptr) ds *= (rscale / ep_scale) * rdj;
              v[j] = __float_as_uint(ds);
            }
          } else {
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) * ep_alpha);
          }
          // stage + TMA store.  fp32: one 32-col chunk = 128 B per row; bf16: two chunks = 128 B per row.
          if (OUT_F32) {
            if (etid == 0) tma_store_wait_read<1>();
            named_bar_sync(1, 128);
            uint8_t* dst = smem_c + cbuf * kStageCBytes + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 8; ++j) {
              uint4 w = make_uint4(v[4 * j], v[4 * j + 1], v[4 * j + 2], v[4 * j + 3]);
              *reinterpret_cast<uint4*>(dst + ((j ^ (row_in_tile & 7)) << 4)) = w;
            }
            fence_proxy_async_smem();
            named_bar_sync(1, 128);
            if (etid == 0) {
              if (ACCUM) tma_reduce_add_2d(&tmC, smem_c + cbuf * kStageCBytes, col0, m0);
              else tma_store_2d(&tmC, smem_c + cbuf * kStageCBytes, col0, m0);
              tma_store_commit();
            }
            cbuf ^= 1;
          } else {
            const int half = c & 1;
            if (half == 0) {
              if (etid == 0) tma_store_wait_read<1>();
              named_bar_sync(1, 128);
            }
            uint8_t* dst = smem_c + cbuf * kStageCBytes + row_in_tile * 128;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
              uint4 w;
              w.x = pack_bf16x2(__uint_as_float(v[8 * j + 0]), __uint_as_float(v[8 * j + 1]));
              w.y = pack_bf16x2(__uint_as_float(v[8 * j + 2]), __uint_as_float(v[8 * j + 3]));
              w.z = pack_bf16x2(__uint_as_float(v[8 * j + 4]), __uint_as_float(v[8 * j + 5]));
              w.w = pack_bf16x2(__uint_as_float(v[8 * j + 6]), __uint_as_float(v[8 * j + 7]));
              const int chunk = half * 4 + j;
              *reinterpret_cast<uint4*>(dst + ((chunk ^ (row_in_tile & 7)) << 4)) = w;
            }
            if (half == 1) {
              fence_proxy_async_smem();
              named_bar_sync(1, 128);
              if (etid == 0) {
                tma_store_2d(&tmC, smem_c + cbuf * kStageCBytes, col0 - 32, m0);
                tma_store_commit();
              }
              cbuf ^= 1;
            }
          }
        }
      }
      // accumulator drained -> hand the TMEM buffer back to the MMA warp
      tc_fence_before();
      mbar_arrive(&tempty_bar[acc]);

      if (MODE == EPI_NCE_STATS && row_ok) {
        const size_t o = static_cast<size_t>(nt) * M + row;
        ep.part_max[o] = run_max;
        ep.part_sum[o] = run_sum;
        ep.part_arg[o] = run_arg;
      }
    }
    if (MODE == EPI_NCE_DS) ep.dlogit_part[blockIdx.x * 128 + etid] = dlogit_acc;
    if (MODE != EPI_NCE_STATS && etid == 0) tma_store_wait<0>();
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<kTmemCols>(tmem_base);
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
int gemm_grid(int M, int N, int splits = 1);
int gemm_block_n(int N);

}  // namespace cx
