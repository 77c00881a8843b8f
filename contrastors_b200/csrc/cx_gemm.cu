// Host launcher + C entry point for the tcgen05 GEMM (see cx_gemm.cuh for the kernel).
#include "cx_gemm.cuh"

#include <stdlib.h>

#include <atomic>
#include <mutex>

namespace cx {

int gemm_block_n(int N) { return (N % 256 == 0) ? 256 : 128; }

int gemm_grid(int M, int N, int splits, int tile_n) {
  const int bn = tile_n > 0 ? tile_n : gemm_block_n(N);
  const int tiles = ((M + kBlockM - 1) / kBlockM) * ((N + bn - 1) / bn) * splits;
  const int sms = sm_count();
  return tiles < sms ? tiles : sms;  // an upper bound on the CTA count in every launch mode (single or paired)
}

// CTA-pair (tcgen05 cta_group::2, 256 x 256 tiles) launch mode: the default whenever the shape allows it.  A single SM
// with 128 x 256 tiles is shared-memory-bandwidth bound (48 KB read by the MMA + 48 KB written by TMA per 512-cycle
// k-block = 192 B/clk against a 128 B/clk port); the pair stages each B half once per SM pair (64 KB per k-block per SM =
// 128 B/clk), measured +8..12 % over the single-CTA kernel (8192^3: 1.46 vs 1.30 PFLOP/s).  CX_NO_PAIR=1 disables it.
static std::atomic<int> g_max_cluster{0};  // A/B switch (cx_gemm_select_cluster): 0 = default, else the largest cluster used
bool gemm_use_pair(const GemmArgs& g) {
  static const bool disabled = getenv("CX_NO_PAIR") != nullptr;
  const int cap = g_max_cluster.load(std::memory_order_relaxed);
  return !disabled && (cap == 0 || cap >= 2) && g.M >= 256 && ((g.mode == EPI_SWIGLU) || g.N % 256 == 0);
}

// QUAD (a cluster of two CTA pairs sharing their B tile through TMA multicast, see cx_gemm.cuh).  Measured on B200 in round 2
// (profiles/r02b_bench_kernels_cluster2_vs_cluster4.json): correct (bitwise equal to pair mode) but 0-8 % SLOWER on every encoder
// shape -- QKV 1231 vs 1288 TFLOP/s, out_proj 898 vs 944, fc2 1373 vs 1372, fc2-dgrad 1348 vs 1451, 8192^3 1447 vs 1523 -- so a
// quarter less L2->SM traffic buys nothing: the pair kernel is not bound by the L2 fabric, and coupling the two pairs' stage
// release costs more than the saved bytes.  Opt-in only (cx_gemm_select_cluster(4)); the default stays CTA pairs.
bool gemm_use_quad(const GemmArgs& g) {
  static const bool disabled = getenv("CX_NO_QUAD") != nullptr;
  const int cap = g_max_cluster.load(std::memory_order_relaxed);
  return !disabled && cap >= 4 && gemm_use_pair(g) && g.M % 512 == 0 &&
         (g.mode == EPI_STORE || g.mode == EPI_SWIGLU || g.mode == EPI_SWIGLU_BWD);
}

template <int MODE, bool OUT_F32, bool ACCUM, bool A_MN, bool B_MN, int NP>
static int launch_pair(const GemmArgs& g, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                       const CUtensorMap& tmD) {
  auto kern = gemm_kernel<256, A_MN, B_MN, MODE, OUT_F32, ACCUM, true, NP>;
  constexpr int smem = GemmSmem<256, true, MODE>::kTotal;
  CX_SET_SMEM_ONCE(kern, smem);
  const int tile_n = (MODE == EPI_SWIGLU) ? 128 : 256;
  const int tile_m = 256 * NP;
  const int tiles = ((g.M + tile_m - 1) / tile_m) * ((g.N + tile_n - 1) / tile_n) * g.splits;
  cudaLaunchConfig_t cfg{};
  cfg.blockDim = dim3(kGemmThreads);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = g.stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = 2 * NP;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  int clusters = sm_count() / (2 * NP);
  if (NP > 1) {
    // 4-CTA clusters must fit inside a GPC: ask the driver how many can be co-resident (once per instantiation)
    static std::once_flag once;
    static int max_clusters = 0;
    std::call_once(once, [&] {
      cfg.gridDim = dim3(2 * NP * clusters);
      int n = 0;
      if (cudaOccupancyMaxActiveClusters(&n, kern, &cfg) == cudaSuccess && n > 0) max_clusters = n;
      (void)cudaGetLastError();
    });
    if (max_clusters > 0 && max_clusters < clusters) clusters = max_clusters;
  }
  if (tiles < clusters) clusters = tiles;
  cfg.gridDim = dim3(2 * NP * clusters);
  CX_CUDA_CHECK(cudaLaunchKernelEx(&cfg, kern, tmA, tmB, tmC, tmD, g.M, g.N, g.K, g.splits, g.ep));
  CX_LAUNCH_CHECK();
  return 0;
}

template <int BLOCK_N, bool A_MN, bool B_MN, int MODE, bool OUT_F32, bool ACCUM>
static int launch_one(const GemmArgs& g, const CUtensorMap& tmA, const CUtensorMap& tmB, const CUtensorMap& tmC,
                      const CUtensorMap& tmD) {
  if (BLOCK_N == 256 && gemm_use_pair(g)) {
    if ((MODE == EPI_STORE || MODE == EPI_SWIGLU || MODE == EPI_SWIGLU_BWD) && gemm_use_quad(g)) return launch_pair<MODE, OUT_F32, ACCUM, A_MN, B_MN, 2>(g, tmA, tmB, tmC, tmD);
    return launch_pair<MODE, OUT_F32, ACCUM, A_MN, B_MN, 1>(g, tmA, tmB, tmC, tmD);
  }
  if constexpr (MODE == EPI_SWIGLU_BWD) {
    return fail(CX_ERR_UNSUPPORTED, "swiglu-bwd epilogue exists for CTA pairs only");
  } else {
    auto kern = gemm_kernel<BLOCK_N, A_MN, B_MN, MODE, OUT_F32, ACCUM>;
    constexpr int smem = GemmSmem<BLOCK_N>::kTotal;
    CX_SET_SMEM_ONCE(kern, smem);  // per instantiation
    const int grid = gemm_grid(g.M, g.N, g.splits, MODE == EPI_SWIGLU ? BLOCK_N / 2 : 0);
    kern<<<grid, kGemmThreads, smem, g.stream>>>(tmA, tmB, tmC, tmD, g.M, g.N, g.K, g.splits, g.ep);
    CX_LAUNCH_CHECK();
    return 0;
  }
}

template <int BLOCK_N, int MODE, bool OUT_F32, bool ACCUM>
static int launch_majors(const GemmArgs& g, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c,
                         const CUtensorMap& d) {
  if constexpr (MODE == EPI_SWIGLU_BWD) {  // the fc2 dgrad: A = d(out) K-major, B = fc2.weight [d, I] = MN-major; CTA pairs only
    if (g.a_mn || !g.b_mn || BLOCK_N != 256 || !gemm_use_pair(g)) return fail(CX_ERR_UNSUPPORTED, "swiglu-bwd epilogue: K-major A, MN-major B, M >= 256");
    return launch_one<256, false, true, EPI_SWIGLU_BWD, false, false>(g, a, b, c, d);
  } else if constexpr (MODE != EPI_STORE) {  // the other fused epilogues only exist for K-major operands
    if (g.a_mn || g.b_mn) return fail(CX_ERR_UNSUPPORTED, "fused epilogues need K-major operands");
    return launch_one<BLOCK_N, false, false, MODE, OUT_F32, ACCUM>(g, a, b, c, d);
  } else {
    if (!g.a_mn && !g.b_mn) return launch_one<BLOCK_N, false, false, EPI_STORE, OUT_F32, ACCUM>(g, a, b, c, d);
    if (!g.a_mn && g.b_mn) return launch_one<BLOCK_N, false, true, EPI_STORE, OUT_F32, ACCUM>(g, a, b, c, d);
    if (g.a_mn && !g.b_mn) return launch_one<BLOCK_N, true, false, EPI_STORE, OUT_F32, ACCUM>(g, a, b, c, d);
    return launch_one<BLOCK_N, true, true, EPI_STORE, OUT_F32, ACCUM>(g, a, b, c, d);
  }
}

template <int BLOCK_N>
static int launch_bn(const GemmArgs& g, const CUtensorMap& a, const CUtensorMap& b, const CUtensorMap& c, const CUtensorMap& d) {
  switch (g.mode) {
    case EPI_NCE_STATS:
      return launch_majors<BLOCK_N, EPI_NCE_STATS, false, false>(g, a, b, c, d);
    case EPI_NCE_DS:
      return launch_majors<BLOCK_N, EPI_NCE_DS, false, false>(g, a, b, c, d);
    case EPI_SWIGLU:
      if (BLOCK_N != 256) return fail(CX_ERR_INVALID, "swiglu epilogue uses 256-column accumulators");
      return launch_majors<256, EPI_SWIGLU, false, false>(g, a, b, c, d);
    case EPI_SWIGLU_BWD:
      if (BLOCK_N != 256) return fail(CX_ERR_INVALID, "swiglu-bwd epilogue uses 256-column accumulators");
      return launch_majors<256, EPI_SWIGLU_BWD, false, false>(g, a, b, c, d);
    case EPI_STORE:
      if (!g.out_f32) return launch_majors<BLOCK_N, EPI_STORE, false, false>(g, a, b, c, d);
      if (!g.accumulate) return launch_majors<BLOCK_N, EPI_STORE, true, false>(g, a, b, c, d);
      return launch_majors<BLOCK_N, EPI_STORE, true, true>(g, a, b, c, d);
  }
  return fail(CX_ERR_INVALID, "bad epilogue mode");
}

int launch_gemm(const GemmArgs& g_in) {
  GemmArgs g = g_in;
  if (g.M <= 0 || g.N <= 0 || g.K <= 0) return fail(CX_ERR_INVALID, "gemm: empty problem");
  if (g.accumulate && !g.out_f32) return fail(CX_ERR_INVALID, "gemm: accumulate needs an fp32 output");
  const int bn = (g.mode == EPI_SWIGLU) ? 256 : gemm_block_n(g.N);
  {
    // split-K: fp32 outputs only (partials are combined by TMA reduce-add at L2)
    const int tiles = ((g.M + kBlockM - 1) / kBlockM) * ((g.N + bn - 1) / bn);
    const int num_kb = (g.K + kBlockK - 1) / kBlockK;
    int splits = g.splits;
    if (splits <= 0) {
      splits = 1;
      if (g.mode == EPI_STORE && g.out_f32) {
        if (bn == 256 && gemm_use_pair(g)) {
          // CTA pairs work on 256 x 256 tiles: pick the split count whose work items fill whole waves of SM pairs (round 1 counted
          // 128-row tiles here: the Wqkv weight gradient, 9 x 3 pair tiles, got 2 splits = 54 items for 74 pairs)
          const int ptiles = ((g.M + 255) / 256) * ((g.N + 255) / 256), pairs = sm_count() / 2;
          int max_s = num_kb / 8;
          if (max_s > 16) max_s = 16;
          double best = 0.0;
          for (int sp = 1; sp <= (max_s < 1 ? 1 : max_s); ++sp) {
            const int items = ptiles * sp, waves = (items + pairs - 1) / pairs;
            const double eff = (double)items / ((double)waves * pairs) - 0.004 * sp;  // prefer fewer reduce-add passes on a tie
            if (eff > best + 1e-9) {
              best = eff;
              splits = sp;
            }
          }
        } else if (tiles * 2 <= sm_count()) {
          splits = sm_count() / tiles;
          if (splits > num_kb / 4) splits = num_kb / 4;
          if (splits < 1) splits = 1;
        }
      }
    }
    if (splits > 1 && !(g.mode == EPI_STORE && g.out_f32)) return fail(CX_ERR_INVALID, "gemm: split-K needs an fp32 output");
    if (splits > num_kb) splits = num_kb;
    g.splits = splits;
    if (splits > 1 && !g.accumulate) {
      CX_CUDA_CHECK(cudaMemset2DAsync(g.C, (size_t)g.ldc * 4, 0, (size_t)g.N * 4, (size_t)g.M, g.stream));
      g.accumulate = true;
    }
  }
  CUtensorMap tmA, tmB, tmC;
  const auto BF = CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  const auto SW = CU_TENSOR_MAP_SWIZZLE_128B;
  int rc;
  if (!g.a_mn) rc = make_tmap_2d(&tmA, BF, 2, g.A, (uint64_t)g.K, (uint64_t)g.M, (uint64_t)g.lda * 2, kBlockK, kBlockM, SW);
  else rc = make_tmap_2d(&tmA, BF, 2, g.A, (uint64_t)g.M, (uint64_t)g.K, (uint64_t)g.lda * 2, 64, kBlockK, SW);
  if (rc) return rc;
  const bool pair = (bn == 256) && gemm_use_pair(g);
  const bool quad = pair && gemm_use_quad(g);  // each CTA fetches 64 of its 128 B rows and multicasts them
  if (g.mode == EPI_SWIGLU) rc = make_tmap_2d(&tmB, BF, 2, g.B, (uint64_t)g.K, (uint64_t)2 * g.N, (uint64_t)g.ldb * 2, kBlockK, quad ? 64u : 128u, SW);
  else if (!g.b_mn) rc = make_tmap_2d(&tmB, BF, 2, g.B, (uint64_t)g.K, (uint64_t)g.N, (uint64_t)g.ldb * 2, kBlockK, quad ? 64u : (pair ? 128u : (uint32_t)bn), SW);
  else rc = make_tmap_2d(&tmB, BF, 2, g.B, (uint64_t)g.N, (uint64_t)g.K, (uint64_t)g.ldb * 2, 64, kBlockK, SW);
  if (rc) return rc;
  CUtensorMap tmD;
  if (g.mode == EPI_SWIGLU) {
    rc = make_tmap_2d(&tmC, BF, 2, g.ep.act_out, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ep.ld_act * 2, 64, kBlockM, SW);
    if (rc) return rc;
    if (g.ep.yg_out != nullptr)
      rc = make_tmap_2d(&tmD, BF, 2, g.ep.yg_out, (uint64_t)2 * g.N, (uint64_t)g.M, (uint64_t)g.ep.ld_yg * 2, 64, kBlockM, SW);
  } else if (g.mode == EPI_SWIGLU_BWD) {  // tmC: dyg [M, 2N] (stores), tmD: yg [M, 2N] (loads); 64-column x 128-row boxes
    rc = make_tmap_2d(&tmC, BF, 2, g.C, (uint64_t)2 * g.N, (uint64_t)g.M, (uint64_t)g.ldc * 2, 64, kBlockM, SW);
    if (rc) return rc;
    rc = make_tmap_2d(&tmD, BF, 2, g.ep.yg_out, (uint64_t)2 * g.N, (uint64_t)g.M, (uint64_t)g.ep.ld_yg * 2, 64, kBlockM, SW);
  } else if (g.mode == EPI_NCE_STATS) {
    tmC = tmA;
  } else if (g.out_f32) {
    rc = make_tmap_2d(&tmC, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 4, g.C, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc * 4, 32, kBlockM, SW);
  } else {
    rc = make_tmap_2d(&tmC, BF, 2, g.C, (uint64_t)g.N, (uint64_t)g.M, (uint64_t)g.ldc * 2, 64, kBlockM, SW);
  }
  if (rc) return rc;
  if (g.mode != EPI_SWIGLU_BWD && (g.mode != EPI_SWIGLU || g.ep.yg_out == nullptr)) tmD = tmC;
  return bn == 256 ? launch_bn<256>(g, tmA, tmB, tmC, tmD) : launch_bn<128>(g, tmA, tmB, tmC, tmD);
}

}  // namespace cx

extern "C" int cx_gemm_select_cluster(int max_cluster_ctas) {
  CX_REQUIRE(max_cluster_ctas == 0 || max_cluster_ctas == 1 || max_cluster_ctas == 2 || max_cluster_ctas == 4,
             "cx_gemm_select_cluster: 0 (default), 1 (single CTA), 2 (CTA pairs) or 4 (two pairs sharing B by multicast)");
  cx::g_max_cluster.store(max_cluster_ctas, std::memory_order_relaxed);
  return 0;
}

extern "C" int cx_gemm_bf16(const void* A, const void* B, void* C, int M, int N, int K, int a_major, int b_major,
                            int64_t lda, int64_t ldb, int64_t ldc, int c_dtype, int accumulate, float alpha,
                            cx_stream_t stream) {
  CX_REQUIRE(A && B && C, "cx_gemm_bf16: null pointer");
  CX_REQUIRE(c_dtype == CX_BF16 || c_dtype == CX_F32, "cx_gemm_bf16: bad c_dtype");
  cx::GemmArgs g{};
  g.A = A; g.B = B; g.C = C;
  g.M = M; g.N = N; g.K = K;
  g.a_mn = a_major == CX_MAJOR_MN;
  g.b_mn = b_major == CX_MAJOR_MN;
  g.lda = lda; g.ldb = ldb; g.ldc = ldc;
  g.out_f32 = c_dtype == CX_F32;
  g.accumulate = accumulate != 0;
  g.mode = cx::EPI_STORE;
  g.ep.alpha = alpha;
  g.stream = static_cast<cudaStream_t>(stream);
  return cx::launch_gemm(g);
}

extern "C" int cx_gemm_swiglu(const void* x, const void* w1, void* act_out, void* yg_out, int M, int I, int K, int64_t ldx,
                              int64_t ldw, int64_t ld_act, int64_t ld_yg, cx_stream_t stream) {
  CX_REQUIRE(x && w1 && act_out, "cx_gemm_swiglu: null pointer");
  CX_REQUIRE(I % 128 == 0, "cx_gemm_swiglu: the gated width must be a multiple of 128");
  CX_REQUIRE(ld_act % 8 == 0 && (yg_out == nullptr || ld_yg % 8 == 0), "cx_gemm_swiglu: output rows must be 16-byte aligned");
  cx::GemmArgs g{};
  g.A = x; g.B = w1; g.C = act_out;
  g.M = M; g.N = I; g.K = K;
  g.a_mn = false; g.b_mn = false;
  g.lda = ldx; g.ldb = ldw; g.ldc = ld_act;
  g.out_f32 = false; g.accumulate = false; g.splits = 1;
  g.mode = cx::EPI_SWIGLU;
  g.ep.act_out = reinterpret_cast<__nv_bfloat16*>(act_out);
  g.ep.ld_act = ld_act;
  g.ep.yg_out = reinterpret_cast<__nv_bfloat16*>(yg_out);
  g.ep.ld_yg = ld_yg;
  g.stream = static_cast<cudaStream_t>(stream);
  return cx::launch_gemm(g);
}

// d[y | gate] = SwiGLU'(d(out) fc2) with the activation backward in the fc2-dgrad epilogue (replaces the dgrad GEMM + the
// swiglu backward of flash-attn's `swiglu`, layers/mlp.py:75; autograd of y * silu(gate)).  dout [M, K = d], w2 = fc2.weight
// [K = d, I] row-major, yg / dyg [M, 2 I] = [y | gate] and their gradients.
extern "C" int cx_gemm_swiglu_bwd(const void* dout, const void* w2, const void* yg, void* dyg, int M, int I, int K, int64_t ld_dout,
                                  int64_t ld_w2, int64_t ld_yg, int64_t ld_dyg, cx_stream_t stream) {
  CX_REQUIRE(dout && w2 && yg && dyg, "cx_gemm_swiglu_bwd: null pointer");
  CX_REQUIRE(I % 256 == 0 && M >= 256, "cx_gemm_swiglu_bwd: gated width a multiple of 256, at least 256 rows (cx_gemm_bf16 + cx_swiglu_bwd otherwise)");
  CX_REQUIRE(ld_yg % 8 == 0 && ld_dyg % 8 == 0, "cx_gemm_swiglu_bwd: rows must be 16-byte aligned");
  cx::GemmArgs g{};
  g.A = dout; g.B = w2; g.C = dyg;
  g.M = M; g.N = I; g.K = K;
  g.a_mn = false; g.b_mn = true;
  g.lda = ld_dout; g.ldb = ld_w2; g.ldc = ld_dyg;
  g.out_f32 = false; g.accumulate = false; g.splits = 1;
  g.mode = cx::EPI_SWIGLU_BWD;
  g.ep.alpha = 1.f;
  g.ep.yg_out = reinterpret_cast<__nv_bfloat16*>(const_cast<void*>(yg));
  g.ep.ld_yg = ld_yg;
  g.stream = static_cast<cudaStream_t>(stream);
  return cx::launch_gemm(g);
}

// QKV projection with the rotary embedding applied in the epilogue (replaces Wqkv GEMM + apply_rotary_emb + torch.stack,
// layers/attention.py:112-133): qkv[T, 3*H*64] = x W^T, then q and k heads rotated by cos/sin[pos[t]].
extern "C" int cx_gemm_qkv_rope(const void* x, const void* w, void* qkv, int T, int n_out, int K, int64_t ldx, int64_t ldw,
                                int64_t ldo, const int32_t* pos, const float* inv_freq, int rope_cols, cx_stream_t stream) {
  CX_REQUIRE(x && w && qkv && pos && inv_freq, "cx_gemm_qkv_rope: null pointer");
  CX_REQUIRE(n_out % 64 == 0 && rope_cols % 64 == 0 && rope_cols <= n_out, "cx_gemm_qkv_rope: heads are 64 columns wide");
  cx::GemmArgs g{};
  g.A = x; g.B = w; g.C = qkv;
  g.M = T; g.N = n_out; g.K = K;
  g.a_mn = false; g.b_mn = false;
  g.lda = ldx; g.ldb = ldw; g.ldc = ldo;
  g.out_f32 = false; g.accumulate = false; g.splits = 1;
  g.mode = cx::EPI_STORE;
  g.ep.alpha = 1.f;
  g.ep.rope_pos = pos; g.ep.rope_inv_freq = inv_freq; g.ep.rope_cols = rope_cols;
  g.stream = static_cast<cudaStream_t>(stream);
  return cx::launch_gemm(g);
}

// y[M,N] = x[M,K] w[N,K]^T + bias[N]  (bf16 in/out, fp32 bias added in the epilogue; replaces flash-attn FusedDense)
extern "C" int cx_linear_bias_bf16(const void* x, const void* w, const float* bias, void* y, int M, int N, int K, int64_t ldx,
                                   int64_t ldw, int64_t ldy, cx_stream_t stream) {
  CX_REQUIRE(x && w && y, "cx_linear_bias_bf16: null pointer");
  cx::GemmArgs g{};
  g.A = x; g.B = w; g.C = y;
  g.M = M; g.N = N; g.K = K;
  g.a_mn = false; g.b_mn = false;
  g.lda = ldx; g.ldb = ldw; g.ldc = ldy;
  g.out_f32 = false; g.accumulate = false; g.splits = 1;
  g.mode = cx::EPI_STORE;
  g.ep.alpha = 1.f;
  g.ep.bias = bias;
  g.stream = static_cast<cudaStream_t>(stream);
  return cx::launch_gemm(g);
}
