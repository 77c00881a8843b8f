// Matryoshka InfoNCE in ONE accumulation over K (C ABI: cx_infonce_mat_fwd).
//
// Replaces the loop of the reference (/root/reference/src/contrastors/trainers/text_text.py:352-369: for every dim,
// F.normalize(q[:, :dim]), F.normalize(all_d[:, :dim]), clip_loss) -- P full logits passes with K = dim, 2*N*M*sum(dims) FLOPs --
// by one pass with K = max(dims): 2*N*M*K FLOPs.  The dot product of a prefix is the running sum of the SEGMENT products
//     <q[:k_s], d[:k_s]> = sum_{t <= s} <q[seg_t], d[seg_t]>,     seg_t = [k_{t-1}, k_t),
// so each 128 x 128 logits tile keeps its running prefix sum in tensor memory (P), the tensor cores accumulate every further
// segment into one of two alternating delta accumulators (D0 / D1), and the epilogue warps fold the delta into P as they read it
// (tcgen05.ld P + D, tcgen05.st P) while the next segment's MMAs already run into the other delta buffer.  At every prefix
// boundary the epilogue applies that prefix's per-row inverse norms (rq_s[i], rd_s[j]: the normalisation of the prefix, computed
// from the same bf16 rows) and the logit scale, and reduces the tile to the row statistics of the softmax cross-entropy exactly
// as cx_infonce_fwd's epilogue does (max, sum-exp, first argmax, label logit; log2 domain).
//
// Backward (cx_infonce_mat_bwd): the same accumulation once more with a dS epilogue.  With E_s = c_s (softmax_s - onehot) and
// F_s = diag(rq_s) E_s diag(rd_s), the gradient of column c in segment t (= columns [k_{t-1}, k_t)) is
//     dq[:, c] = scale * (T_t D)[:, c] - q[:, c] * A_t,     T_t = sum_{s >= t} F_s,   A_t[i] = sum_{s >= t} rq_s[i]^2 alpha_s[i],
//     alpha_s[i] = sum_j E_s[i,j] S_s[i,j]   (and symmetrically dd with T_t^T, Q, beta_s[j] = sum_i E_s S_s),
// i.e. the chain through F.normalize folds into one row / column scalar per prefix and ONE dS-like matrix per segment: the
// contractions cost 4 N M K FLOPs in total instead of 4 N M sum(dims).  The epilogue keeps the F_s tiles of the earlier prefixes in
// shared memory (fp16, own row per thread), forms the suffix sums T_t at the last boundary and writes them as fp16 (pre-scaled by
// a device scalar 1/gamma ~ 1 / (rq rd) so they sit in [-P, P]); alpha / beta are reduced in the same pass (beta: a 31-shuffle
// transpose-reduce per 32 x 32 block).  Up to 4 prefixes (3 tiles of 32 KB of F in shared memory); more fall back to the loop.
//
// TMEM (512 columns): P0 [0,128)  P1 [128,256)  (prefix sums of even / odd work items)   D0 [256,384)  D1 [384,512).
// One CTA per SM, 128 x 128 tiles, cta_group::1.  Warps: 0 TMA producer, 1 MMA issuer, 2 TMEM allocator, 4-11 epilogue
// (thread = row, two groups of 64 columns).  The kernel is bound by the SFU: P exponentials per logit.
#include <math.h>

#include "cx_gemm.cuh"

namespace cx {

int nce_rows_to_f16(const void* x, int64_t ldx, void* y, int64_t ldy, const float* inv_norm, int rows, int k, cudaStream_t stream);

constexpr int kMatMaxDims = 8;
constexpr int kMatStages = 6;
constexpr int kMatTile = 128 * 64 * 2;  // one operand tile of a k-block: 128 rows x 64 bf16
constexpr int kMatThreads = 384;

struct MatParams {
  int n_dims;
  int kb_end[kMatMaxDims];  // k-block index one past the end of each prefix (dims / 64), ascending
  const float* rq;          // [n_dims][M]
  const float* rd;          // [n_dims][N]
  float scale;
  const float* scale_dev;
  int label_offset, label_stride;
  float* part_max;          // [n_dims][2 * n_col_tiles][M]
  float* part_sum;
  int* part_arg;
  float* label_logit;       // [n_dims][M]
  // dS mode
  const float* lse;         // [n_dims][M]  natural log
  float wrel[kMatMaxDims];  // w_s / max_w
  const float* inv_gamma;   // device scalar: stored T = true T / (gamma * common coefficient)
  __half* T;                // [n_dims][M][ldT]
  int64_t ldT;
  float* alpha;             // [n_dims][M]   sum_j wrel_s (p - 1h) s      (natural-log logits)
  float* beta;              // [n_dims][N]
};

template <int MODE>
struct MatSmem {
  static constexpr int kStages = MODE == 0 ? kMatStages : 4;
  static constexpr int kStageBytes = 2 * kMatTile;
  static constexpr int kF = kStages * kStageBytes;          // dS mode: 3 prefixes x 2 groups x [128 rows x 128 B] fp16
  static constexpr int kBars = kF + (MODE == 0 ? 0 : 3 * 32768);
  static constexpr int kTotal = 1024 + kBars + 256;
};

// column sums of a 32 x 32 block held one row per lane (v[0..31] = the lane's row): after 31 shuffle-adds lane L holds the sum of
// column L (recursive halving: at every step a lane keeps the half of its columns that matches one more bit of its lane id)
__device__ __forceinline__ float warp_colsum32(float (&v)[32]) {
  const int lane = threadIdx.x & 31;
#pragma unroll
  for (int step = 0; step < 5; ++step) {
    const int half = 16 >> step;
    const bool upper = (lane & half) != 0;
#pragma unroll
    for (int j = 0; j < half; ++j) {
      const float send = upper ? v[j] : v[j + half];
      const float keep = upper ? v[j + half] : v[j];
      v[j] = keep + __shfl_xor_sync(0xffffffffu, send, half);
    }
  }
  return v[0];
}

template <int MODE>
__global__ void __launch_bounds__(kMatThreads, 1)
nce_mat_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, int M, int N, MatParams mp) {
  extern __shared__ uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  using S = MatSmem<MODE>;
  constexpr int kStages = S::kStages;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + S::kBars);
  uint64_t* full_bar = bars;                    // [kStages]  TMA -> MMA
  uint64_t* empty_bar = bars + kStages;         // [kStages]  MMA -> TMA
  uint64_t* p_full = bars + 2 * kStages;     // [2] segment 0 of a work item is in P[slot]
  uint64_t* p_free = p_full + 2;                // [2] the epilogue is done with P[slot] (256 arrivals)
  uint64_t* d_full = p_free + 2;                // [2] a later segment is in D[buf]
  uint64_t* d_free = d_full + 2;                // [2] the epilogue has folded D[buf] into P (256 arrivals)
  uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(d_free + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int m_tiles = (M + 127) / 128, n_tiles = (N + 127) / 128;
  const int num_tiles = m_tiles * n_tiles;
  const int nd = mp.n_dims;
  const int num_kb = mp.kb_end[nd - 1];

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&tmA);
    tma_prefetch_desc(&tmB);
  }
  if (warp == 1 && lane == 0) {
    for (int i = 0; i < kStages; ++i) {
      mbar_init(&full_bar[i], 1);
      mbar_init(&empty_bar[i], 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&p_full[i], 1);
      mbar_init(&p_free[i], 256);
      mbar_init(&d_full[i], 1);
      mbar_init(&d_free[i], 256);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc<512>(tmem_ptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_ptr;

  if (warp == 0) {
    // ---------------------------------------------------------------- TMA producer: the k-blocks of every work item, in order
    int stage = 0;
    uint32_t phase = 0;
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x) {
      const int m0 = (tile / n_tiles) * 128, n0 = (tile % n_tiles) * 128;
      for (int kb = 0; kb < num_kb; ++kb) {
        mbar_wait(&empty_bar[stage], phase ^ 1);
        if (elect_one()) {
          uint8_t* sA = smem + stage * S::kStageBytes;
          mbar_arrive_expect_tx(&full_bar[stage], S::kStageBytes);
          tma_load_2d(sA, &tmA, &full_bar[stage], kb * 64, m0);
          tma_load_2d(sA + kMatTile, &tmB, &full_bar[stage], kb * 64, n0);
        }
        __syncwarp();
        if (++stage == kStages) {
          stage = 0;
          phase ^= 1;
        }
      }
    }
  } else if (warp == 1) {
    // ---------------------------------------------------------------- MMA issuer: segment 0 -> P[slot], segment s >= 1 -> D[(s-1) & 1]
    constexpr uint32_t idesc = make_idesc_bf16(128, 128, 0, 0);
    int stage = 0;
    uint32_t phase = 0;
    int it = 0;
    uint32_t d_uses[2] = {0, 0};  // how many times each delta buffer has been handed to the epilogue so far
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int slot = it & 1;
      int kb = 0;
      for (int s = 0; s < nd; ++s) {
        uint32_t d_tmem;
        if (s == 0) {
          mbar_wait(&p_free[slot], ((it >> 1) & 1) ^ 1);  // the work item two back has left this prefix-sum slot
          d_tmem = tmem_base + slot * 128;
        } else {
          const int b = (s - 1) & 1;
          mbar_wait(&d_free[b], (d_uses[b] & 1) ^ 1);      // the previous delta in this buffer has been folded into its P
          d_tmem = tmem_base + 256 + b * 128;
        }
        tc_fence_after();
        const int kb_first = kb;
        for (; kb < mp.kb_end[s]; ++kb) {
          mbar_wait(&full_bar[stage], phase);
          tc_fence_after();
          const uint32_t a_addr = smem_u32(smem + stage * S::kStageBytes);
          const uint64_t adesc0 = make_smem_desc_sw128(a_addr, 0, 1024);
          const uint64_t bdesc0 = make_smem_desc_sw128(a_addr + kMatTile, 0, 1024);
          if (elect_one()) {
#pragma unroll
            for (int k = 0; k < 4; ++k)
              umma_f16_ss(d_tmem, adesc0 + (uint64_t)((k * 32) >> 4), bdesc0 + (uint64_t)((k * 32) >> 4), idesc,
                          (kb > kb_first || k > 0) ? 1u : 0u);
            umma_commit(&empty_bar[stage]);
            if (kb == mp.kb_end[s] - 1) umma_commit(s == 0 ? &p_full[slot] : &d_full[(s - 1) & 1]);
          }
          __syncwarp();
          if (++stage == kStages) {
            stage = 0;
            phase ^= 1;
          }
        }
        if (s > 0) ++d_uses[(s - 1) & 1];
      }
    }
  } else if (warp >= 4) {
    // ---------------------------------------------------------------- epilogue: thread = row, group hf = 64 columns
    constexpr float kLog2e = 1.4426950408889634f, kLn2 = 0.6931471805599453f;
    const int ew = warp & 3, hf = (warp - 4) >> 2;
    const int row_in_tile = ew * 32 + lane;
    const uint32_t lane_base = static_cast<uint32_t>(ew * 32) << 16;
    const float ep_scale = mp.scale * (mp.scale_dev != nullptr ? *mp.scale_dev : 1.f);
    int it = 0;
    uint32_t d_uses[2] = {0, 0};
    for (int tile = blockIdx.x; tile < num_tiles; tile += gridDim.x, ++it) {
      const int slot = it & 1;
      const int mt = tile / n_tiles, nt = tile % n_tiles;
      const int row = mt * 128 + row_in_tile;
      const bool row_ok = row < M;
      const int n0 = nt * 128 + hf * 64;
      const int label = (row + mp.label_offset) * mp.label_stride;
      const uint32_t t_p = tmem_base + lane_base + slot * 128 + hf * 64;
      for (int s = 0; s < nd; ++s) {
        const int b = (s - 1) & 1;
        const uint32_t t_d = tmem_base + lane_base + 256 + b * 128 + hf * 64;
        if (s == 0) mbar_wait(&p_full[slot], (it >> 1) & 1);
        else mbar_wait(&d_full[b], d_uses[b] & 1);
        tc_fence_after();
        const float rqi = row_ok ? mp.rq[(size_t)s * M + row] : 0.f;
        const float rs2 = ep_scale * kLog2e * rqi;
        const float* rd_s = mp.rd + (size_t)s * N;
        float run_max = -INFINITY, run_sum = 0.f;
        int run_arg = 0;
        if constexpr (MODE == 0) {
#pragma unroll 1
        for (int c = 0; c < 2; ++c) {
          uint32_t v[32];
          tmem_ld_32x32(t_p + c * 32, v);
          if (s > 0) {
            uint32_t dv[32];
            tmem_ld_32x32(t_d + c * 32, dv);
            tmem_ld_wait();
#pragma unroll
            for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(dv[j]));
            if (s < nd - 1) tmem_st_32x32(t_p + c * 32, v);  // the running prefix sum goes back for the next boundary
          } else {
            tmem_ld_wait();
          }
          const int col0 = n0 + c * 32;
          float t[32];
#pragma unroll
          for (int j = 0; j < 32; ++j) {
            const int col = col0 + j;
            t[j] = (col < N) ? __uint_as_float(v[j]) * rs2 * rd_s[col] : -INFINITY;
          }
          float cmax = t[0];
#pragma unroll
          for (int j = 1; j < 32; ++j) cmax = fmaxf(cmax, t[j]);
          if (cmax > run_max) {  // first max wins: only a strictly larger value moves the argmax
            int arg = 0;
#pragma unroll
            for (int j = 31; j >= 0; --j) arg = (t[j] == cmax) ? j : arg;
            run_arg = col0 + arg;
            run_sum *= fast_exp2(run_max - cmax);
            run_max = cmax;
          }
          const float mref = (run_max == -INFINITY) ? 0.f : run_max;
          float s0 = 0.f, s1 = 0.f, s2 = 0.f, s3 = 0.f;
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
            mp.label_logit[(size_t)s * M + row] = lv * kLn2;
          }
        }
        } else {
          // ---- dS mode: F_s = wrel_s rq_i (p - 1h) rd_j / gamma for this row's 64 columns; earlier prefixes park theirs in shared
          // memory (fp16, own row, 16-byte chunks swizzled by the row), the last one forms the suffix sums T_t and writes them out
          const float lse2 = row_ok ? mp.lse[(size_t)s * M + row] * kLog2e : INFINITY;
          const float fr = mp.wrel[s] * rqi * (*mp.inv_gamma);
          uint8_t* fbase = smem + S::kF + hf * 16384 + row_in_tile * 128;  // + prefix * 32768
          float a_row = 0.f;
#pragma unroll 1
          for (int c = 0; c < 2; ++c) {
            uint32_t v[32];
            tmem_ld_32x32(t_p + c * 32, v);
            if (s > 0) {
              uint32_t dv[32];
              tmem_ld_32x32(t_d + c * 32, dv);
              tmem_ld_wait();
#pragma unroll
              for (int j = 0; j < 32; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(dv[j]));
              if (s < nd - 1) tmem_st_32x32(t_p + c * 32, v);
            } else {
              tmem_ld_wait();
            }
            const int col0 = n0 + c * 32;
            float f[32], es[32];
#pragma unroll
            for (int j = 0; j < 32; ++j) {
              const int col = col0 + j;
              const float rdj = (col < N) ? rd_s[col] : 0.f;
              const float t = __uint_as_float(v[j]) * rs2 * rdj;              // logit in the log2 domain
              float p = (col < N) ? fast_exp2(t - lse2) : 0.f;
              if (col == label && row_ok) p -= 1.f;
              es[j] = p * t * (mp.wrel[s] * kLn2);                              // wrel (p - 1h) s, natural-log logit
              f[j] = p * fr * rdj;
              a_row += es[j];
            }
            // beta_s[col] += sum over this warp's 32 rows
            {
              const float cs = warp_colsum32(es);
              if (col0 + lane < N) atomicAdd(&mp.beta[(size_t)s * N + col0 + lane], cs);
            }
            if (s < nd - 1) {
              uint8_t* fr_ = fbase + s * 32768;
#pragma unroll
              for (int q = 0; q < 4; ++q) {
                uint4 w;
                w.x = pack_f16x2(f[8 * q + 0], f[8 * q + 1]);
                w.y = pack_f16x2(f[8 * q + 2], f[8 * q + 3]);
                w.z = pack_f16x2(f[8 * q + 4], f[8 * q + 5]);
                w.w = pack_f16x2(f[8 * q + 6], f[8 * q + 7]);
                *reinterpret_cast<uint4*>(fr_ + (((c * 4 + q) ^ (row_in_tile & 7)) << 4)) = w;
              }
            } else {
              // suffix sums from the last prefix down: T_{nd-1} = F_{nd-1}, T_t = F_t + T_{t+1}; each written as this row's 64 bytes
#pragma unroll 1
              for (int t = nd - 1; t >= 0; --t) {
                if (t < nd - 1) {
                  const uint8_t* fr_ = fbase + t * 32768;
#pragma unroll
                  for (int q = 0; q < 4; ++q) {
                    const uint4 w = *reinterpret_cast<const uint4*>(fr_ + (((c * 4 + q) ^ (row_in_tile & 7)) << 4));
                    const __half2* h2 = reinterpret_cast<const __half2*>(&w);
#pragma unroll
                    for (int e = 0; e < 4; ++e) {
                      const float2 x = __half22float2(h2[e]);
                      f[8 * q + 2 * e] += x.x;
                      f[8 * q + 2 * e + 1] += x.y;
                    }
                  }
                }
                if (row_ok) {
                  __half* dst = mp.T + ((size_t)t * M + row) * mp.ldT + col0;
                  if (col0 + 32 <= N) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                      uint4 w;
                      w.x = pack_f16x2(f[8 * q + 0], f[8 * q + 1]);
                      w.y = pack_f16x2(f[8 * q + 2], f[8 * q + 3]);
                      w.z = pack_f16x2(f[8 * q + 4], f[8 * q + 5]);
                      w.w = pack_f16x2(f[8 * q + 6], f[8 * q + 7]);
                      *reinterpret_cast<uint4*>(dst + 8 * q) = w;
                    }
                  } else {
                    for (int j = 0; j < 32 && col0 + j < N; ++j) dst[j] = __float2half_rn(f[j]);
                  }
                }
              }
            }
          }
          if (row_ok) atomicAdd(&mp.alpha[(size_t)s * M + row], a_row);
        }
        if (s > 0 && s < nd - 1) tmem_st_wait();
        tc_fence_before();
        if (s > 0) {
          mbar_arrive(&d_free[b]);
          ++d_uses[b];
        }
        if (s == nd - 1) mbar_arrive(&p_free[slot]);
        if (MODE == 0 && row_ok) {
          const size_t o = ((size_t)s * 2 * n_tiles + (size_t)(nt * 2 + hf)) * M + row;
          mp.part_max[o] = run_max;
          mp.part_sum[o] = run_sum;
          mp.part_arg[o] = run_arg;
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    tmem_dealloc<512>(tmem_base);
  }
}

// per prefix (blockIdx.y) and row: merge the per-column-half partials (same rules as nce_combine_kernel: first-maximum argmax,
// fixed-order sums); one warp per row; the loss / hit sums are accumulated with one atomicAdd per block (order-dependent in the
// last bits only: these two numbers are logging / the scalar loss, the gradients never read them)
__global__ void nce_mat_combine_kernel(const float* __restrict__ part_max, const float* __restrict__ part_sum,
                                       const int* __restrict__ part_arg, const float* __restrict__ label_logit, int n, int n_parts,
                                       int label_offset, int label_stride, float* __restrict__ lse, int* __restrict__ argmax,
                                       float* __restrict__ stats) {
  constexpr float kLn2 = 0.6931471805599453f;
  const int s = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int row = blockIdx.x * (blockDim.x >> 5) + warp;
  const float* pm = part_max + (size_t)s * n_parts * n;
  const float* ps = part_sum + (size_t)s * n_parts * n;
  const int* pa = part_arg + (size_t)s * n_parts * n;
  float loss_term = 0.f, hit = 0.f;
  if (row < n) {
    float gmax = -INFINITY;
    int garg = 0x7fffffff;
    for (int t = lane; t < n_parts; t += 32) {
      const float mx = pm[(size_t)t * n + row];
      const int ag = pa[(size_t)t * n + row];
      if (mx > gmax || (mx == gmax && ag < garg)) {
        gmax = mx;
        garg = ag;
      }
    }
    for (int o = 16; o > 0; o >>= 1) {
      const float om = __shfl_xor_sync(0xffffffffu, gmax, o);
      const int oa = __shfl_xor_sync(0xffffffffu, garg, o);
      if (om > gmax || (om == gmax && oa < garg)) {
        gmax = om;
        garg = oa;
      }
    }
    float sum = 0.f;
    for (int t = lane; t < n_parts; t += 32) {
      const float v = ps[(size_t)t * n + row];
      if (v > 0.f) sum += v * exp2f(pm[(size_t)t * n + row] - gmax);
    }
    for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
    const float l = (gmax + log2f(sum)) * kLn2;
    if (lane == 0) {
      lse[(size_t)s * n + row] = l;
      argmax[(size_t)s * n + row] = garg;
      loss_term = l - label_logit[(size_t)s * n + row];
      hit = (garg == (row + label_offset) * label_stride) ? 1.f : 0.f;
    }
  }
  __shared__ float s_loss[32], s_hit[32];
  if (lane == 0) {
    s_loss[warp] = loss_term;
    s_hit[warp] = hit;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f, b = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) {
      a += s_loss[w];
      b += s_hit[w];
    }
    atomicAdd(&stats[4 * s], a);
    atomicAdd(&stats[4 * s + 1], b);
  }
}

}  // namespace cx

using namespace cx;

static size_t mat_align(size_t x) { return (x + 255) & ~size_t(255); }

extern "C" size_t cx_infonce_mat_workspace_bytes(int n, int m, int n_dims) {
  if (n <= 0 || m <= 0 || n_dims <= 0) return 0;
  const size_t parts = 2 * ((size_t)(m + 127) / 128);
  return 3 * mat_align((size_t)n_dims * parts * n * 4) + 512;
}

extern "C" int cx_infonce_mat_fwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int n_dims, const int32_t* dims,
                                  float scale, const float* scale_dev, const float* rq, const float* rd, int label_offset,
                                  int label_stride, float* lse, int32_t* argmax, float* label_logit, float* stats, void* workspace,
                                  cx_stream_t stream_) {
  CX_REQUIRE(q && d && dims && rq && rd && lse && argmax && label_logit && stats && workspace, "cx_infonce_mat_fwd: null pointer");
  CX_REQUIRE(n > 0 && m > 0 && n_dims >= 1 && n_dims <= kMatMaxDims, "cx_infonce_mat_fwd: 1..8 prefix dims");
  CX_REQUIRE((long long)(n - 1 + label_offset) * label_stride < m && label_offset >= 0 && label_stride >= 1,
             "cx_infonce_mat_fwd: labels out of range");
  MatParams mp{};
  mp.n_dims = n_dims;
  int prev = 0;
  for (int i = 0; i < n_dims; ++i) {
    CX_REQUIRE(dims[i] > prev && dims[i] % 64 == 0, "cx_infonce_mat_fwd: dims must ascend in multiples of 64 (the k-block of the MMA pipeline)");
    mp.kb_end[i] = dims[i] / 64;
    prev = dims[i];
  }
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const size_t parts = 2 * ((size_t)(m + 127) / 128);
  uint8_t* w = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(workspace) + 255) & ~uintptr_t(255));
  const size_t chunk = mat_align((size_t)n_dims * parts * n * 4);
  mp.part_max = reinterpret_cast<float*>(w);
  mp.part_sum = reinterpret_cast<float*>(w + chunk);
  mp.part_arg = reinterpret_cast<int*>(w + 2 * chunk);
  mp.rq = rq; mp.rd = rd; mp.scale = scale; mp.scale_dev = scale_dev;
  mp.label_offset = label_offset; mp.label_stride = label_stride; mp.label_logit = label_logit;
  CUtensorMap tmA, tmB;
  const int K = dims[n_dims - 1];
  int rc = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q, (uint64_t)K, (uint64_t)n, (uint64_t)ldq * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, (uint64_t)K, (uint64_t)m, (uint64_t)ldd * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  CX_SET_SMEM_ONCE(nce_mat_kernel<0>, MatSmem<0>::kTotal);
  const int tiles = ((n + 127) / 128) * ((m + 127) / 128);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  CX_CUDA_CHECK(cudaMemsetAsync(stats, 0, (size_t)n_dims * 4 * sizeof(float), stream));
  nce_mat_kernel<0><<<grid, kMatThreads, MatSmem<0>::kTotal, stream>>>(tmA, tmB, n, m, mp);
  CX_LAUNCH_CHECK();
  dim3 cgrid((n + 7) / 8, n_dims);
  nce_mat_combine_kernel<<<cgrid, 256, 0, stream>>>(mp.part_max, mp.part_sum, mp.part_arg, label_logit, n, (int)parts, label_offset,
                                                    label_stride, lse, argmax, stats);
  CX_LAUNCH_CHECK();
  return 0;
}

// ---------------------------------------------------------------------------------------------------------------- backward
struct MatBwdWs {
  __half* T;
  int64_t ldT;
  __half* q16;
  __half* d16;
  int64_t ldh;
  size_t bytes;
};
static MatBwdWs mat_bwd_carve(void* base, int n, int m, int K, int n_dims) {
  MatBwdWs w{};
  uint8_t* p = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(base) + 255) & ~uintptr_t(255));
  size_t off = 0;
  w.ldT = (int64_t)((m + 7) / 8 * 8);
  w.T = reinterpret_cast<__half*>(p + off);
  off += mat_align((size_t)n_dims * n * w.ldT * 2);
  w.ldh = (int64_t)((K + 7) / 8 * 8);
  w.q16 = reinterpret_cast<__half*>(p + off);
  off += mat_align((size_t)n * w.ldh * 2);
  w.d16 = reinterpret_cast<__half*>(p + off);
  off += mat_align((size_t)m * w.ldh * 2);
  w.bytes = off + 256;
  return w;
}

extern "C" size_t cx_infonce_mat_bwd_workspace_bytes(int n, int m, int k_max, int n_dims) {
  if (n <= 0 || m <= 0 || k_max <= 0 || n_dims <= 0) return 0;
  return mat_bwd_carve(nullptr, n, m, k_max, n_dims).bytes;
}

extern "C" int cx_infonce_mat_bwd(const void* q, int64_t ldq, const void* d, int64_t ldd, int n, int m, int n_dims, const int32_t* dims,
                                  const float* wrel, float scale, const float* scale_dev, const float* rq, const float* rd,
                                  int label_offset, int label_stride, const float* lse, float coef, const float* coef_gamma_dev,
                                  const float* inv_gamma_dev, float* dq_raw, int64_t lddq, float* dd_raw, int64_t lddd, float* alpha,
                                  float* beta, void* workspace, cx_stream_t stream_) {
  CX_REQUIRE(q && d && dims && wrel && rq && rd && lse && coef_gamma_dev && inv_gamma_dev && dq_raw && dd_raw && alpha && beta && workspace,
             "cx_infonce_mat_bwd: null pointer");
  CX_REQUIRE(n > 0 && m > 0 && n_dims >= 2 && n_dims <= 4, "cx_infonce_mat_bwd: 2..4 prefix dims (F tiles of the earlier prefixes live in shared memory)");
  MatParams mp{};
  mp.n_dims = n_dims;
  int prev = 0;
  for (int i = 0; i < n_dims; ++i) {
    CX_REQUIRE(dims[i] > prev && dims[i] % 64 == 0, "cx_infonce_mat_bwd: dims must ascend in multiples of 64");
    mp.kb_end[i] = dims[i] / 64;
    mp.wrel[i] = wrel[i];
    prev = dims[i];
  }
  const int K = dims[n_dims - 1];
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  MatBwdWs w = mat_bwd_carve(workspace, n, m, K, n_dims);
  mp.rq = rq; mp.rd = rd; mp.scale = scale; mp.scale_dev = scale_dev;
  mp.label_offset = label_offset; mp.label_stride = label_stride;
  mp.lse = lse; mp.inv_gamma = inv_gamma_dev; mp.T = w.T; mp.ldT = w.ldT; mp.alpha = alpha; mp.beta = beta;
  CUtensorMap tmA, tmB;
  int rc = make_tmap_2d(&tmA, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, q, (uint64_t)K, (uint64_t)n, (uint64_t)ldq * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  rc = make_tmap_2d(&tmB, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d, (uint64_t)K, (uint64_t)m, (uint64_t)ldd * 2, 64, 128, CU_TENSOR_MAP_SWIZZLE_128B);
  if (rc) return rc;
  CX_CUDA_CHECK(cudaMemsetAsync(alpha, 0, (size_t)n_dims * n * sizeof(float), stream));
  CX_CUDA_CHECK(cudaMemsetAsync(beta, 0, (size_t)n_dims * m * sizeof(float), stream));
  if (w.ldT != m)  // the padding columns of T are read by the contractions' TMA boxes: keep them finite
    CX_CUDA_CHECK(cudaMemsetAsync(w.T, 0, (size_t)n_dims * n * w.ldT * 2, stream));
  CX_SET_SMEM_ONCE(nce_mat_kernel<1>, MatSmem<1>::kTotal);
  const int tiles = ((n + 127) / 128) * ((m + 127) / 128);
  const int grid = tiles < sm_count() ? tiles : sm_count();
  nce_mat_kernel<1><<<grid, kMatThreads, MatSmem<1>::kTotal, stream>>>(tmA, tmB, n, m, mp);
  CX_LAUNCH_CHECK();
  // fp16 copies of the RAW rows (the inverse norms already sit inside T)
  rc = nce_rows_to_f16(q, ldq, w.q16, w.ldh, nullptr, n, K, stream);
  if (rc) return rc;
  rc = nce_rows_to_f16(d, ldd, w.d16, w.ldh, nullptr, m, K, stream);
  if (rc) return rc;
  // one contraction pair per SEGMENT: dq[:, seg] = a T_t D16[:, seg],  dd[:, seg] = a T_t^T Q16[:, seg],  a = scale * coef * gamma * grad
  prev = 0;
  for (int t = 0; t < n_dims; ++t) {
    const int seg = dims[t] - prev;
    const __half* Tt = w.T + (size_t)t * n * w.ldT;
    GemmArgs a{};
    a.A = Tt; a.B = w.d16 + prev; a.C = dq_raw + prev;
    a.M = n; a.N = seg; a.K = m;
    a.a_mn = false; a.b_mn = true;
    a.lda = w.ldT; a.ldb = w.ldh; a.ldc = lddq;
    a.out_f32 = true; a.accumulate = false; a.splits = 0;
    a.mode = EPI_STORE; a.ep.alpha = scale * coef; a.ep.alpha_dev = scale_dev; a.ep.alpha_dev2 = coef_gamma_dev;
    a.ep.ab_f16 = 1;
    a.stream = stream;
    rc = launch_gemm(a);
    if (rc) return rc;
    GemmArgs b{};
    b.A = Tt; b.B = w.q16 + prev; b.C = dd_raw + prev;
    b.M = m; b.N = seg; b.K = n;
    b.a_mn = true; b.b_mn = true;
    b.lda = w.ldT; b.ldb = w.ldh; b.ldc = lddd;
    b.out_f32 = true; b.accumulate = false; b.splits = 0;
    b.mode = EPI_STORE; b.ep.alpha = scale * coef; b.ep.alpha_dev = scale_dev; b.ep.alpha_dev2 = coef_gamma_dev;
    b.ep.ab_f16 = 1;
    b.stream = stream;
    rc = launch_gemm(b);
    if (rc) return rc;
    prev = dims[t];
  }
  return 0;
}
