// HBM-bound encoder kernels (C ABI: cx_add_layernorm_*, cx_embed_layernorm_*, cx_rope_*, cx_swiglu_*, cx_mean_pool_*,
// cx_embed_head_*, cx_token_positions, cx_adamw_*, cx_sumsq, cx_cast_*).
//
// They replace the flash-attn side extensions the reference calls:
//   K2 dropout_add_layer_norm (layers/block.py:422-431,453-462; models/encoder/modeling_nomic_bert.py:531-535)
//   K4 swiglu (layers/mlp.py:73-75)      K5 rotary (layers/embedding.py:685-706)
//   K6 unpad/pad bookkeeping             K11 MeanPooling / hamming LN / F.normalize (modeling_biencoder.py:79-90,307-317)
// plus the optimizer tail (optimizer.py:7-47, trainers/base.py:372-385) as one fused pass.
// All are one-pass, 16-byte vectorised, one warp per row where rows are independent; roofline = HBM bytes.
#include <math.h>

#include "cx_host.h"
#include "cx_ptx.cuh"

namespace cx {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

struct BF8 {  // 8 bf16 = 16 bytes
  uint4 raw;
  __device__ __forceinline__ void unpack(float (&f)[8]) const {
    const __nv_bfloat162* p = reinterpret_cast<const __nv_bfloat162*>(&raw);
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      f[2 * i] = __low2float(p[i]);
      f[2 * i + 1] = __high2float(p[i]);
    }
  }
  __device__ __forceinline__ void pack(const float (&f)[8]) {
    raw.x = pack_bf16x2(f[0], f[1]);
    raw.y = pack_bf16x2(f[2], f[3]);
    raw.z = pack_bf16x2(f[4], f[5]);
    raw.w = pack_bf16x2(f[6], f[7]);
  }
};

constexpr int kLnMaxVec = 4;  // up to d = 32 lanes * 8 * 4 = 1024 columns per row

// Counter-based dropout: the keep mask of 8 consecutive columns is a pure function of (seed, row, col / 8), so the
// backward (and GradCache's second pass, which replays the CPU RNG that produced the seed) regenerates it for free.
struct DropParams {
  float p;       // drop probability (0 = off)
  float scale;   // 1 / (1 - p)
  unsigned long long seed;
};
__device__ __forceinline__ unsigned long long splitmix64(unsigned long long z) {
  z = (z ^ (z >> 30)) * 0xBF58476D1CE4E5B9ull;
  z = (z ^ (z >> 27)) * 0x94D049BB133111EBull;
  return z ^ (z >> 31);
}
// multiplies f[0..8) by keep * scale
__device__ __forceinline__ void apply_dropout8(float (&f)[8], const DropParams& dp, int64_t row, int col, int d) {
  const unsigned long long idx = (unsigned long long)row * (unsigned long long)(d / 8) + (unsigned long long)(col / 8);
  const unsigned long long x = dp.seed + idx * 0x9E3779B97F4A7C15ull;
  const unsigned long long r0 = splitmix64(x), r1 = splitmix64(x ^ 0xD1B54A32D192ED03ull);
  const unsigned int thr = (unsigned int)(dp.p * 65536.f);
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    const unsigned int u = (unsigned int)(((i < 4 ? r0 : r1) >> (16 * (i & 3))) & 0xFFFFull);
    f[i] = (u >= thr) ? f[i] * dp.scale : 0.f;
  }
}

__device__ __forceinline__ void load8f(const float* __restrict__ p, float (&f)[8]) {
  const float4 a = *reinterpret_cast<const float4*>(p);
  const float4 b = *reinterpret_cast<const float4*>(p + 4);
  f[0] = a.x; f[1] = a.y; f[2] = a.z; f[3] = a.w;
  f[4] = b.x; f[5] = b.y; f[6] = b.z; f[7] = b.w;
}

// ---------------------------------------------------------------------------------------------- LayerNorm forward
// z = a (+ b) ; y = (z - mean) * rstd * gamma + beta.  EMBED: a-row = word_emb[ids[r]] + type_emb[type_ids[r]].
// One warp per row; NV = ceil(d / 256) 16-byte vectors per lane, all loads issued before any math.
template <bool EMBED, int NV>
__global__ void __launch_bounds__(256)
add_layernorm_fwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                         const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids,
                         const __nv_bfloat16* __restrict__ type_emb, const float* __restrict__ gamma,
                         const float* __restrict__ beta, __nv_bfloat16* __restrict__ y, float* __restrict__ stats, int rows,
                         int d, float eps, __nv_bfloat16* __restrict__ z_out, DropParams dp) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __nv_bfloat16* ar = EMBED ? a + (size_t)ids[row] * d : a + (size_t)row * d;
  const __nv_bfloat16* br = EMBED ? type_emb + (size_t)(type_ids ? type_ids[row] : 0) * d : (b ? b + (size_t)row * d : nullptr);
  BF8 ra[NV], rb[NV];
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < d) {
      ra[c].raw = *reinterpret_cast<const uint4*>(ar + col);
      if (br != nullptr) rb[c].raw = *reinterpret_cast<const uint4*>(br + col);
    }
  }
  float z[NV][8];
  float sum = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < d) {
      ra[c].unpack(z[c]);
      if (!EMBED && dp.p > 0.f) apply_dropout8(z[c], dp, row, col, d);  // z = dropout(a) + b  (block.py:422-431)
      if (br != nullptr) {
        float t[8];
        rb[c].unpack(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[c][i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) sum += z[c][i];
      if (z_out != nullptr) {  // pre-norm blocks carry the residual stream z = a + b forward (layers/block.py:293-388)
        BF8 vz;
        vz.pack(z[c]);
        *reinterpret_cast<uint4*>(z_out + (size_t)row * d + col) = vz.raw;
      }
    }
  }
  const float mean = warp_sum(sum) / d;
  float var = 0.f;
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < d) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        const float t = z[c][i] - mean;
        var = fmaf(t, t, var);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(var) / d + eps);
  if (lane == 0 && stats != nullptr) {
    stats[2 * (size_t)row] = mean;
    stats[2 * (size_t)row + 1] = rstd;
  }
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < d) {
      float o[8], gm[8], bt[8];
      if (gamma != nullptr) {
        load8f(gamma + col, gm);
        load8f(beta + col, bt);
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        float v = (z[c][i] - mean) * rstd;
        if (gamma != nullptr) v = fmaf(v, gm[i], bt[i]);
        o[i] = v;
      }
      if (EMBED && dp.p > 0.f) apply_dropout8(o, dp, row, col, d);  // emb_drop AFTER emb_ln (modeling_nomic_bert.py:531-535)
      BF8 vo;
      vo.pack(o);
      *reinterpret_cast<uint4*>(y + (size_t)row * d + col) = vo.raw;
    }
  }
}

// ---------------------------------------------------------------------------------------------- LayerNorm backward
// g = g1 (+ g2); dz = rstd * (g*gamma - mean_d(g*gamma) - xhat * mean_d(g*gamma*xhat)); per-CTA partials of
// dgamma / dbeta (and, EMBED with type_ids == NULL, of the type-0 embedding row gradient) in a fixed order.
// EMBED: z is recomputed from the tables and dz is scattered (red.add fp32) into the word-embedding gradient.
template <bool EMBED, int NV>
__global__ void __launch_bounds__(256)
add_layernorm_bwd_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                         const int64_t* __restrict__ ids, const int64_t* __restrict__ type_ids,
                         const __nv_bfloat16* __restrict__ type_emb, const __nv_bfloat16* __restrict__ g1,
                         const __nv_bfloat16* __restrict__ g2, const float* __restrict__ gamma,
                         const float* __restrict__ stats, __nv_bfloat16* __restrict__ dz, float* __restrict__ dword,
                         float* __restrict__ dtype_emb, float* __restrict__ partials, int rows, int d, int64_t padding_idx,
                         const __nv_bfloat16* __restrict__ gres, DropParams dp, __nv_bfloat16* __restrict__ da_out) {
  extern __shared__ float sh[];  // [warps][NP][d]
  constexpr int NP = EMBED ? 3 : 2;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  float dg[NV][8], db[NV][8], dt[EMBED ? NV : 1][8];
#pragma unroll
  for (int c = 0; c < NV; ++c)
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      dg[c][i] = db[c][i] = 0.f;
      if (EMBED) dt[c][i] = 0.f;
    }
  float gm[NV][8];
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 8;
#pragma unroll
    for (int i = 0; i < 8; ++i) gm[c][i] = 1.f;
    if (col < d && gamma != nullptr) load8f(gamma + col, gm[c]);
  }

  for (int row = blockIdx.x * nwarps + warp; row < rows; row += gridDim.x * nwarps) {
    const __nv_bfloat16* ar = EMBED ? a + (size_t)ids[row] * d : a + (size_t)row * d;
    const int64_t tid = EMBED ? (type_ids ? type_ids[row] : 0) : 0;
    const __nv_bfloat16* br = EMBED ? type_emb + (size_t)tid * d : (b ? b + (size_t)row * d : nullptr);
    const float mean = stats[2 * (size_t)row], rstd = stats[2 * (size_t)row + 1];
    BF8 ra[NV], rb[NV], rg1[NV], rg2[NV];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < d) {
        ra[c].raw = *reinterpret_cast<const uint4*>(ar + col);
        if (br != nullptr) rb[c].raw = *reinterpret_cast<const uint4*>(br + col);
        rg1[c].raw = *reinterpret_cast<const uint4*>(g1 + (size_t)row * d + col);
        if (g2 != nullptr) rg2[c].raw = *reinterpret_cast<const uint4*>(g2 + (size_t)row * d + col);
      }
    }
    float xh[NV][8], wg[NV][8];
    float s1 = 0.f, s2 = 0.f;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < d) {
        float z[8], t[8], g[8];
        ra[c].unpack(z);
        if (!EMBED && dp.p > 0.f) apply_dropout8(z, dp, row, col, d);  // re-create z = dropout(a) + b
        if (br != nullptr) {
          rb[c].unpack(t);
#pragma unroll
          for (int i = 0; i < 8; ++i) z[i] += t[i];
        }
        rg1[c].unpack(g);
        if (g2 != nullptr) {
          rg2[c].unpack(t);
#pragma unroll
          for (int i = 0; i < 8; ++i) g[i] += t[i];
        }
        if (EMBED && dp.p > 0.f) apply_dropout8(g, dp, row, col, d);   // y = dropout(LN(z)): mask the upstream gradient
#pragma unroll
        for (int i = 0; i < 8; ++i) {
          const float x = (z[i] - mean) * rstd;
          const float w = g[i] * gm[c][i];
          xh[c][i] = x;
          wg[c][i] = w;
          s1 += w;
          s2 = fmaf(w, x, s2);
          dg[c][i] = fmaf(g[i], x, dg[c][i]);
          db[c][i] += g[i];
        }
      }
    }
    s1 = warp_sum(s1) / d;
    s2 = warp_sum(s2) / d;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
      const int col = (c * 32 + lane) * 8;
      if (col < d) {
        float o[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] = (wg[c][i] - s1 - xh[c][i] * s2) * rstd;
        if (EMBED) {
          float* dw = dword + (size_t)ids[row] * d + col;
          const bool pad = ids[row] == padding_idx;  // nn.Embedding(padding_idx=...) never receives a gradient there
#pragma unroll
          for (int i = 0; i < 8; ++i) {
            if (!pad) atomicAdd(dw + i, o[i]);
            if (type_ids == nullptr) dt[c][i] += o[i];  // all tokens are type 0: column sum, reduced below
            else atomicAdd(dtype_emb + (size_t)tid * d + col + i, o[i]);
          }
        } else {
          if (gres != nullptr) {  // pre-norm: z also feeds the residual stream, whose gradient adds here
            BF8 vr;
            float t[8];
            vr.raw = *reinterpret_cast<const uint4*>(gres + (size_t)row * d + col);
            vr.unpack(t);
#pragma unroll
            for (int i = 0; i < 8; ++i) o[i] += t[i];
          }
          BF8 vo;
          vo.pack(o);
          *reinterpret_cast<uint4*>(dz + (size_t)row * d + col) = vo.raw;
          if (da_out != nullptr) {  // gradient of the dropped branch a: dz * keep / (1 - p)
            if (dp.p > 0.f) apply_dropout8(o, dp, row, col, d);
            vo.pack(o);
            *reinterpret_cast<uint4*>(da_out + (size_t)row * d + col) = vo.raw;
          }
        }
      }
    }
  }
  if (partials == nullptr) return;
  // CTA reduction of the column-sum partials (fixed order => deterministic)
#pragma unroll
  for (int c = 0; c < NV; ++c) {
    const int col = (c * 32 + lane) * 8;
    if (col < d) {
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        sh[(warp * NP + 0) * d + col + i] = dg[c][i];
        sh[(warp * NP + 1) * d + col + i] = db[c][i];
        if (EMBED) sh[(warp * NP + 2) * d + col + i] = dt[c][i];
      }
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < NP * d; j += blockDim.x) {
    const int which = j / d, col = j % d;
    float acc = 0.f;
    for (int w = 0; w < nwarps; ++w) acc += sh[(w * NP + which) * d + col];
    partials[(size_t)blockIdx.x * NP * d + j] = acc;
  }
}


// LayerNorm backward, narrow form (the non-embedding case, 48 launches per GradCache chunk): a thread owns 8 columns of a row and
// a row spans WPR warps, so a thread holds 16 column-sum accumulators instead of 16 * d / 256 and four 16-byte loads instead of
// 4 * d / 256: ~70 registers against 219 for the warp-per-row form above at d = 768, i.e. three times the resident warps to
// cover the HBM latency (round 1: 65 us = 3.9 TB/s = 0.59 of the measured copy peak with one 256-thread block per SM).
// Block = RPB row slots x WPR warps (384 threads); the WPR warps of a slot combine their two row sums through shared memory and a
// named barrier of their own, so slots never wait for each other.
template <int WPR>
__global__ void __launch_bounds__(384, 2)
add_layernorm_bwd_narrow_kernel(const __nv_bfloat16* __restrict__ a, const __nv_bfloat16* __restrict__ b,
                                const __nv_bfloat16* __restrict__ g1, const __nv_bfloat16* __restrict__ g2,
                                const float* __restrict__ gamma, const float* __restrict__ stats, __nv_bfloat16* __restrict__ dz,
                                float* __restrict__ partials, int rows, int d, const __nv_bfloat16* __restrict__ gres, DropParams dp,
                                __nv_bfloat16* __restrict__ da_out) {
  constexpr int RPB = 12 / WPR;
  extern __shared__ float sh[];  // [RPB][2][d] at the end; the first RPB * WPR * 2 floats double as the row-sum exchange
  __shared__ float xch[2][12][2];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int slot = warp / WPR, wr = warp % WPR;
  const int col = (wr * 32 + lane) * 8;
  const bool col_ok = col < d;
  float gm[8], dg[8], db[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) {
    gm[i] = 1.f;
    dg[i] = db[i] = 0.f;
  }
  if (col_ok && gamma != nullptr) load8f(gamma + col, gm);
  const float inv_d = 1.f / (float)d;
  int it = 0;
  for (int row = blockIdx.x * RPB + slot; row < rows; row += gridDim.x * RPB, ++it) {
    const float mean = stats[2 * (size_t)row], rstd = stats[2 * (size_t)row + 1];
    float x[8], w[8];
    float s1 = 0.f, s2 = 0.f;
    if (col_ok) {
      BF8 ra, rb, r1, r2;
      const size_t off = (size_t)row * d + col;
      ra.raw = *reinterpret_cast<const uint4*>(a + off);
      if (b != nullptr) rb.raw = *reinterpret_cast<const uint4*>(b + off);
      r1.raw = *reinterpret_cast<const uint4*>(g1 + off);
      if (g2 != nullptr) r2.raw = *reinterpret_cast<const uint4*>(g2 + off);
      float z[8], t[8], g[8];
      ra.unpack(z);
      if (dp.p > 0.f) apply_dropout8(z, dp, row, col, d);  // re-create z = dropout(a) + b
      if (b != nullptr) {
        rb.unpack(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) z[i] += t[i];
      }
      r1.unpack(g);
      if (g2 != nullptr) {
        r2.unpack(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) g[i] += t[i];
      }
#pragma unroll
      for (int i = 0; i < 8; ++i) {
        x[i] = (z[i] - mean) * rstd;
        w[i] = g[i] * gm[i];
        s1 += w[i];
        s2 = fmaf(w[i], x[i], s2);
        dg[i] = fmaf(g[i], x[i], dg[i]);
        db[i] += g[i];
      }
    }
    s1 = warp_sum(s1);
    s2 = warp_sum(s2);
    if (WPR > 1) {  // combine the row's WPR warps: double-buffered exchange + a named barrier private to the slot
      float* xc = &xch[it & 1][0][0];
      if (lane == 0) {
        xc[(slot * WPR + wr) * 2] = s1;
        xc[(slot * WPR + wr) * 2 + 1] = s2;
      }
      named_bar_sync(1 + slot, WPR * 32);
      s1 = 0.f;
      s2 = 0.f;
#pragma unroll
      for (int k = 0; k < WPR; ++k) {
        s1 += xc[(slot * WPR + k) * 2];
        s2 += xc[(slot * WPR + k) * 2 + 1];
      }
    }
    s1 *= inv_d;
    s2 *= inv_d;
    if (col_ok) {
      float o[8];
#pragma unroll
      for (int i = 0; i < 8; ++i) o[i] = (w[i] - s1 - x[i] * s2) * rstd;
      const size_t off = (size_t)row * d + col;
      if (gres != nullptr) {  // pre-norm: z also feeds the residual stream, whose gradient adds here
        BF8 vr;
        float t[8];
        vr.raw = *reinterpret_cast<const uint4*>(gres + off);
        vr.unpack(t);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += t[i];
      }
      BF8 vo;
      vo.pack(o);
      *reinterpret_cast<uint4*>(dz + off) = vo.raw;
      if (da_out != nullptr) {  // gradient of the dropped branch a: dz * keep / (1 - p)
        if (dp.p > 0.f) apply_dropout8(o, dp, row, col, d);
        vo.pack(o);
        *reinterpret_cast<uint4*>(da_out + off) = vo.raw;
      }
    }
  }
  if (partials == nullptr) return;
  // CTA reduction of the column-sum partials over the row slots (fixed order => deterministic)
  if (col_ok) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      sh[(slot * 2 + 0) * d + col + i] = dg[i];
      sh[(slot * 2 + 1) * d + col + i] = db[i];
    }
  }
  __syncthreads();
  for (int j = threadIdx.x; j < 2 * d; j += blockDim.x) {
    const int which = j / d, c = j % d;
    float acc = 0.f;
#pragma unroll
    for (int r = 0; r < RPB; ++r) acc += sh[(r * 2 + which) * d + c];
    partials[(size_t)blockIdx.x * 2 * d + j] = acc;
  }
}

// out_k[j] += sum over blocks of partials[b][k][j], k < np (dgamma, dbeta, and optionally the type-0 embedding row).
// 8 threads per column split the blocks (fixed assignment + fixed shuffle tree => deterministic).
__global__ void ln_param_grad_reduce_kernel(const float* __restrict__ partials, int nblocks, int d, int np, float* __restrict__ o0,
                                            float* __restrict__ o1, float* __restrict__ o2) {
  const int t = blockIdx.x * blockDim.x + threadIdx.x;
  const int j = t >> 3, sub = t & 7;
  float acc = 0.f;
  if (j < np * d)
    for (int b = sub; b < nblocks; b += 8) acc += partials[(size_t)b * np * d + j];
  acc += __shfl_xor_sync(0xffffffffu, acc, 4);
  acc += __shfl_xor_sync(0xffffffffu, acc, 2);
  acc += __shfl_xor_sync(0xffffffffu, acc, 1);
  if (j >= np * d || sub != 0) return;
  const int which = j / d, col = j % d;
  float* dst = which == 0 ? o0 : (which == 1 ? o1 : o2);
  if (dst != nullptr) atomicAdd(dst + col, acc);  // GradCache chunks may run on two streams: accumulate atomically
}

// ---------------------------------------------------------------------------------------------- token bookkeeping
// pos[t] = t - cu_seqlens[seq(t)]  for packed (unpadded) tokens
__global__ void token_positions_kernel(const int* __restrict__ cu, int nseq, int* __restrict__ pos, int* __restrict__ seq_id) {
  const int s = blockIdx.x;
  if (s >= nseq) return;
  const int b = cu[s], e = cu[s + 1];
  for (int t = b + threadIdx.x; t < e; t += blockDim.x) {
    pos[t] = t - b;
    if (seq_id != nullptr) seq_id[t] = s;
  }
}

// ---------------------------------------------------------------------------------------------- RoPE (NeoX halves)
// In place on q and k of qkv [T, 3, H, Dh]: (x1, x2) -> (x1 c - x2 s, x2 c + x1 s); dir = -1 applies the transpose
// (the backward).  cos/sin tables are fp32 [max_pos, Dh/2].
__global__ void rope_kernel(__nv_bfloat16* __restrict__ qkv, const int* __restrict__ pos, const float* __restrict__ cos_t,
                            const float* __restrict__ sin_t, int T, int H, int Dh, float dir, int first_slot, int num_slots) {
  const int half = Dh / 2;
  const int per_tok = num_slots * H * (half / 8);  // 8 (x1, x2) pairs per thread
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * per_tok) return;
  const int t = (int)(i / per_tok);
  int r = (int)(i % per_tok);
  const int which = first_slot + r / (H * (half / 8));
  r %= H * (half / 8);
  const int h = r / (half / 8), j0 = (r % (half / 8)) * 8;
  __nv_bfloat16* base = qkv + ((size_t)t * 3 + which) * H * Dh + (size_t)h * Dh;
  const float* c = cos_t + (size_t)pos[t] * half + j0;
  const float* s = sin_t + (size_t)pos[t] * half + j0;
  BF8 v1, v2;
  float x1[8], x2[8], o1[8], o2[8];
  v1.raw = *reinterpret_cast<const uint4*>(base + j0);
  v2.raw = *reinterpret_cast<const uint4*>(base + half + j0);
  v1.unpack(x1);
  v2.unpack(x2);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float cs = c[k], sn = s[k] * dir;
    o1[k] = x1[k] * cs - x2[k] * sn;
    o2[k] = x2[k] * cs + x1[k] * sn;
  }
  v1.pack(o1);
  v2.pack(o2);
  *reinterpret_cast<uint4*>(base + j0) = v1.raw;
  *reinterpret_cast<uint4*>(base + half + j0) = v2.raw;
}

// fp32 dq accumulator [T, H*Dh] -> bf16 into dqkv's q slot, with the RoPE transpose fused (attention backward tail)
__global__ void dq_finalize_rope_kernel(const float* __restrict__ dq_acc, __nv_bfloat16* __restrict__ dqkv,
                                        const int* __restrict__ pos, const float* __restrict__ cos_t,
                                        const float* __restrict__ sin_t, int T, int H, int Dh) {
  const int half = Dh / 2;
  const int per_tok = H * (half / 8);
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)T * per_tok) return;
  const int t = (int)(i / per_tok);
  const int r = (int)(i % per_tok);
  const int h = r / (half / 8), j0 = (r % (half / 8)) * 8;
  const float* src = dq_acc + ((size_t)t * H + h) * Dh;
  const float* c = cos_t + (size_t)pos[t] * half + j0;
  const float* s = sin_t + (size_t)pos[t] * half + j0;
  float o1[8], o2[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float x1 = src[j0 + k], x2 = src[half + j0 + k];
    o1[k] = x1 * c[k] + x2 * s[k];
    o2[k] = x2 * c[k] - x1 * s[k];
  }
  BF8 v1, v2;
  v1.pack(o1);
  v2.pack(o2);
  __nv_bfloat16* dst = dqkv + ((size_t)t * 3) * H * Dh + (size_t)h * Dh;
  *reinterpret_cast<uint4*>(dst + j0) = v1.raw;
  *reinterpret_cast<uint4*>(dst + half + j0) = v2.raw;
}

// ---------------------------------------------------------------------------------------------- SwiGLU
// yg [T, 2I] = [y | gate];  out = y * silu(gate)   (mlp.py:68-75: fc11 -> y, fc12 -> gate)
__global__ void swiglu_fwd_kernel(const __nv_bfloat16* __restrict__ yg, __nv_bfloat16* __restrict__ out, int64_t T, int I) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = I / 8;
  if (i >= T * per_row) return;
  const int64_t t = i / per_row;
  const int c = (int)(i % per_row) * 8;
  BF8 vy, vg, vo;
  float y[8], g[8], o[8];
  vy.raw = *reinterpret_cast<const uint4*>(yg + t * 2 * I + c);
  vg.raw = *reinterpret_cast<const uint4*>(yg + t * 2 * I + I + c);
  vy.unpack(y);
  vg.unpack(g);
#pragma unroll
  for (int k = 0; k < 8; ++k) o[k] = y[k] * g[k] / (1.f + __expf(-g[k]));
  vo.pack(o);
  *reinterpret_cast<uint4*>(out + t * I + c) = vo.raw;
}

__global__ void swiglu_bwd_kernel(const __nv_bfloat16* __restrict__ dout, const __nv_bfloat16* __restrict__ yg,
                                  __nv_bfloat16* __restrict__ dyg, int64_t T, int I) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const int per_row = I / 8;
  if (i >= T * per_row) return;
  const int64_t t = i / per_row;
  const int c = (int)(i % per_row) * 8;
  BF8 vy, vg, vd, o1, o2;
  float y[8], g[8], d[8], dy[8], dg[8];
  vy.raw = *reinterpret_cast<const uint4*>(yg + t * 2 * I + c);
  vg.raw = *reinterpret_cast<const uint4*>(yg + t * 2 * I + I + c);
  vd.raw = *reinterpret_cast<const uint4*>(dout + t * I + c);
  vy.unpack(y);
  vg.unpack(g);
  vd.unpack(d);
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    const float sg = 1.f / (1.f + __expf(-g[k]));
    const float silu = g[k] * sg;
    dy[k] = d[k] * silu;
    dg[k] = d[k] * y[k] * (sg * (1.f + g[k] * (1.f - sg)));
  }
  o1.pack(dy);
  o2.pack(dg);
  *reinterpret_cast<uint4*>(dyg + t * 2 * I + c) = o1.raw;
  *reinterpret_cast<uint4*>(dyg + t * 2 * I + I + c) = o2.raw;
}

// ---------------------------------------------------------------------------------------------- pooling + head
// pooled[s, :] = mean over the tokens of sequence s (packed rows), fp32.  One CTA per (sequence, 256-column slab).
__global__ void mean_pool_fwd_kernel(const __nv_bfloat16* __restrict__ h, const int* __restrict__ cu, float* __restrict__ pooled,
                                     int d) {
  const int s = blockIdx.x;
  const int col = (blockIdx.y * 32 + (threadIdx.x & 31)) * 8;
  const int warp = threadIdx.x >> 5, nwarps = blockDim.x >> 5;
  const int b = cu[s], e = cu[s + 1];
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  if (col < d) {
    for (int t = b + warp; t < e; t += nwarps) {
      BF8 v;
      float f[8];
      v.raw = *reinterpret_cast<const uint4*>(h + (size_t)t * d + col);
      v.unpack(f);
#pragma unroll
      for (int k = 0; k < 8; ++k) acc[k] += f[k];
    }
  }
  __shared__ float sh[8][32][8];
#pragma unroll
  for (int k = 0; k < 8; ++k) sh[warp][threadIdx.x & 31][k] = acc[k];
  __syncthreads();
  if (warp == 0 && col < d) {
    const float inv = 1.f / (float)max(e - b, 1);
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      float a = 0.f;
      for (int w = 0; w < nwarps; ++w) a += sh[w][threadIdx.x][k];
      pooled[(size_t)s * d + col + k] = a * inv;
    }
  }
}

// dh[t, :] = dpooled[seq(t), :] / len(seq(t))   (bf16)
__global__ void mean_pool_bwd_kernel(const float* __restrict__ dpooled, const int* __restrict__ cu, __nv_bfloat16* __restrict__ dh,
                                     int d) {
  const int s = blockIdx.x;
  const int b = cu[s], e = cu[s + 1];
  const float inv = 1.f / (float)max(e - b, 1);
  const int per_row = d / 8;
  for (int64_t i = threadIdx.x; i < (int64_t)(e - b) * per_row; i += blockDim.x) {
    const int t = b + (int)(i / per_row), c = (int)(i % per_row) * 8;
    float o[8];
#pragma unroll
    for (int k = 0; k < 8; ++k) o[k] = dpooled[(size_t)s * d + c + k] * inv;
    BF8 v;
    v.pack(o);
    *reinterpret_cast<uint4*>(dh + (size_t)t * d + c) = v.raw;
  }
}

// BiEncoder tail on [B, d] (modeling_biencoder.py:307-317): optional affine-free LayerNorm ("hamming"), cast to the trunk
// dtype (bf16), optional F.normalize; fp32 output.  One warp per row.  Saves (mean, rstd, inv_norm) for the backward.
__global__ void embed_head_fwd_kernel(const float* __restrict__ pooled, float* __restrict__ out, float* __restrict__ save,
                                      int rows, int d, int hamming, int normalize) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = pooled + (size_t)row * d;
  float mean = 0.f, rstd = 1.f;
  if (hamming) {
    float s = 0.f;
    for (int j = lane; j < d; j += 32) s += x[j];
    mean = warp_sum(s) / d;
    float v = 0.f;
    for (int j = lane; j < d; j += 32) v = fmaf(x[j] - mean, x[j] - mean, v);
    rstd = rsqrtf(warp_sum(v) / d + 1e-5f);
  }
  float ss = 0.f;
  for (int j = lane; j < d; j += 32) {
    const float r = __bfloat162float(__float2bfloat16_rn((x[j] - mean) * rstd));
    ss = fmaf(r, r, ss);
  }
  const float inv = normalize ? 1.f / fmaxf(sqrtf(warp_sum(ss)), 1e-12f) : 1.f;
  for (int j = lane; j < d; j += 32) {
    const float r = __bfloat162float(__float2bfloat16_rn((x[j] - mean) * rstd));
    out[(size_t)row * d + j] = r * inv;
  }
  if (lane == 0) {
    save[3 * (size_t)row] = mean;
    save[3 * (size_t)row + 1] = rstd;
    save[3 * (size_t)row + 2] = inv;
  }
}

__global__ void embed_head_bwd_kernel(const float* __restrict__ pooled, const float* __restrict__ gout, const float* __restrict__ save,
                                      float* __restrict__ gpooled, int rows, int d, int hamming, int normalize) {
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (row >= rows) return;
  const float* x = pooled + (size_t)row * d;
  const float* g = gout + (size_t)row * d;
  const float mean = save[3 * (size_t)row], rstd = save[3 * (size_t)row + 1], inv = save[3 * (size_t)row + 2];
  // through F.normalize: gr = inv * (g - y (g.y)), y = r * inv  (the bf16 cast is treated as identity, as autograd does)
  float dot = 0.f;
  if (normalize)
    for (int j = lane; j < d; j += 32) {
      const float r = __bfloat162float(__float2bfloat16_rn((x[j] - mean) * rstd));
      dot = fmaf(g[j], r * inv, dot);
    }
  dot = warp_sum(dot);
  float s1 = 0.f, s2 = 0.f;
  if (hamming)
    for (int j = lane; j < d; j += 32) {
      const float r = __bfloat162float(__float2bfloat16_rn((x[j] - mean) * rstd));
      const float gr = normalize ? inv * (g[j] - r * inv * dot) : g[j];
      const float xh = (x[j] - mean) * rstd;
      s1 += gr;
      s2 = fmaf(gr, xh, s2);
    }
  s1 = warp_sum(s1) / d;
  s2 = warp_sum(s2) / d;
  for (int j = lane; j < d; j += 32) {
    const float r = __bfloat162float(__float2bfloat16_rn((x[j] - mean) * rstd));
    const float gr = normalize ? inv * (g[j] - r * inv * dot) : g[j];
    float o = gr;
    if (hamming) o = (gr - s1 - (x[j] - mean) * rstd * s2) * rstd;
    gpooled[(size_t)row * d + j] = o;
  }
}

// ---------------------------------------------------------------------------------------------- optimizer tail
__global__ void sumsq_kernel(const float* __restrict__ x, int64_t n, float* __restrict__ partial) {
  float acc = 0.f;
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i * 4 + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
      acc += v.x * v.x + v.y * v.y + v.z * v.z + v.w * v.w;
    } else {
      for (int64_t j = i * 4; j < n; ++j) acc = fmaf(x[j], x[j], acc);
    }
  }
  acc = warp_sum(acc);
  __shared__ float sh[32];
  if ((threadIdx.x & 31) == 0) sh[threadIdx.x >> 5] = acc;
  __syncthreads();
  if (threadIdx.x == 0) {
    float a = 0.f;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) a += sh[w];
    partial[blockIdx.x] = a;
  }
}
// out[0] = ||g||, out[1] = clip coefficient = min(1, max_norm / (||g|| + 1e-6))   (torch.nn.utils.clip_grad_norm_)
__global__ void clip_coef_kernel(const float* __restrict__ partial, int n, float max_norm, float* __restrict__ out) {
  float a = 0.f;
  for (int i = threadIdx.x; i < n; i += 32) a += partial[i];
  a = warp_sum(a);
  if (threadIdx.x == 0) {
    const float norm = sqrtf(a);
    out[0] = norm;
    out[1] = max_norm > 0.f ? fminf(1.f, max_norm / (norm + 1e-6f)) : 1.f;
  }
}

// AdamW (decoupled decay, torch.optim.AdamW semantics) over a flat fp32 master buffer; also refreshes the bf16 shadow
// the GEMMs read and zeroes the gradient.  grad_scale_dev: optional device scalar (the clip coefficient).
__global__ void adamw_kernel(float* __restrict__ p, float* __restrict__ g, float* __restrict__ m, float* __restrict__ v,
                             __nv_bfloat16* __restrict__ shadow, int64_t n, float lr, float beta1, float beta2, float eps,
                             float wd, float bc1, float bc2, const float* __restrict__ grad_scale_dev, float grad_scale,
                             int zero_grad) {
  const float gs = grad_scale * (grad_scale_dev != nullptr ? grad_scale_dev[0] : 1.f);
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (int64_t)gridDim.x * blockDim.x) {
    const float gi = g[i] * gs;
    float pi = p[i];
    pi *= (1.f - lr * wd);
    const float mi = beta1 * m[i] + (1.f - beta1) * gi;
    const float vi = beta2 * v[i] + (1.f - beta2) * gi * gi;
    m[i] = mi;
    v[i] = vi;
    const float denom = sqrtf(vi) / sqrtf(bc2) + eps;
    pi -= (lr / bc1) * (mi / denom);
    p[i] = pi;
    if (shadow != nullptr) shadow[i] = __float2bfloat16_rn(pi);
    if (zero_grad) g[i] = 0.f;
  }
}

__global__ void cast_f32_bf16_kernel(const float* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n) {
  for (int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; i * 4 < n; i += (int64_t)gridDim.x * blockDim.x) {
    if (i * 4 + 4 <= n) {
      const float4 v = *reinterpret_cast<const float4*>(x + i * 4);
      uint2 o;
      o.x = pack_bf16x2(v.x, v.y);
      o.y = pack_bf16x2(v.z, v.w);
      *reinterpret_cast<uint2*>(y + i * 4) = o;
    } else {
      for (int64_t j = i * 4; j < n; ++j) y[j] = __float2bfloat16_rn(x[j]);
    }
  }
}

// ---------------------------------------------------------------------------------------------- ViT pieces
// out[j] (+)= sum_t x[t, j]   (bias gradients of FusedDense); grid.y slabs of rows, fp32 atomics into out
__global__ void colsum_kernel(const __nv_bfloat16* __restrict__ x, int64_t T, int N, float* __restrict__ out, int rows_per_block) {
  const int col = (blockIdx.x * blockDim.x + threadIdx.x) * 8;
  if (col >= N) return;
  const int64_t r0 = (int64_t)blockIdx.y * rows_per_block;
  const int64_t r1 = min(T, r0 + rows_per_block);
  float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  for (int64_t r = r0; r < r1; ++r) {
    BF8 v;
    float f[8];
    v.raw = *reinterpret_cast<const uint4*>(x + r * N + col);
    v.unpack(f);
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] += f[i];
  }
#pragma unroll
  for (int i = 0; i < 8; ++i) atomicAdd(out + col + i, acc[i]);
}

// kind 0: gelu (erf), 1: quick_gelu x*sigmoid(1.702x) (layers/activations.py), elementwise over bf16
__device__ __forceinline__ float act_fwd(float x, int kind) {
  if (kind == 1) return x / (1.f + __expf(-1.702f * x));
  return 0.5f * x * (1.f + erff(x * 0.70710678118654752f));
}
__device__ __forceinline__ float act_grad(float x, int kind) {
  if (kind == 1) {
    const float s = 1.f / (1.f + __expf(-1.702f * x));
    return s * (1.f + 1.702f * x * (1.f - s));
  }
  return 0.5f * (1.f + erff(x * 0.70710678118654752f)) + x * 0.3989422804014327f * __expf(-0.5f * x * x);
}
__global__ void act_fwd_kernel(const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ y, int64_t n8, int kind) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  BF8 v, o;
  float f[8], g[8];
  v.raw = *reinterpret_cast<const uint4*>(x + i * 8);
  v.unpack(f);
#pragma unroll
  for (int k = 0; k < 8; ++k) g[k] = act_fwd(f[k], kind);
  o.pack(g);
  *reinterpret_cast<uint4*>(y + i * 8) = o.raw;
}
__global__ void act_bwd_kernel(const __nv_bfloat16* __restrict__ dy, const __nv_bfloat16* __restrict__ x, __nv_bfloat16* __restrict__ dx,
                               int64_t n8, int kind) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n8) return;
  BF8 v, w, o;
  float f[8], g[8], r[8];
  v.raw = *reinterpret_cast<const uint4*>(x + i * 8);
  w.raw = *reinterpret_cast<const uint4*>(dy + i * 8);
  v.unpack(f);
  w.unpack(g);
#pragma unroll
  for (int k = 0; k < 8; ++k) r[k] = g[k] * act_grad(f[k], kind);
  o.pack(r);
  *reinterpret_cast<uint4*>(dx + i * 8) = o.raw;
}

// pixels [B, C, Himg, Wimg] fp32 -> patch matrix [B*gh*gw, C*p*p] bf16, column order (c p1 p2) (embedding.py:465-476)
__global__ void patchify_kernel(const float* __restrict__ px, __nv_bfloat16* __restrict__ out, int B, int C, int Himg, int Wimg, int p) {
  const int gw = Wimg / p, gh = Himg / p;
  const int K = C * p * p;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * gh * gw * K) return;
  const int k = (int)(i % K);
  const int64_t row = i / K;
  const int c = k / (p * p), p1 = (k / p) % p, p2 = k % p;
  const int b = (int)(row / (gh * gw)), hw = (int)(row % (gh * gw));
  const int h = hw / gw, w = hw % gw;
  out[i] = __float2bfloat16_rn(px[(((size_t)b * C + c) * Himg + h * p + p1) * Wimg + w * p + p2]);
}

// z[b, 0] = cls + pos[0];  z[b, 1+i] = proj[b*nP + i] + pos[1+i]   (cls token + learned position embedding)
__global__ void vit_assemble_fwd_kernel(const __nv_bfloat16* __restrict__ proj, const float* __restrict__ cls, const float* __restrict__ pos,
                                        __nv_bfloat16* __restrict__ z, int B, int nP, int d) {
  const int per_row = d / 8;
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * (nP + 1) * per_row) return;
  const int64_t row = i / per_row;
  const int col = (int)(i % per_row) * 8;
  const int b = (int)(row / (nP + 1)), tok = (int)(row % (nP + 1));
  float f[8];
  if (tok == 0) {
#pragma unroll
    for (int k = 0; k < 8; ++k) f[k] = cls[col + k];
  } else {
    BF8 v;
    v.raw = *reinterpret_cast<const uint4*>(proj + ((size_t)b * nP + tok - 1) * d + col);
    v.unpack(f);
  }
#pragma unroll
  for (int k = 0; k < 8; ++k) f[k] += pos[(size_t)tok * d + col + k];
  BF8 o;
  o.pack(f);
  *reinterpret_cast<uint4*>(z + row * d + col) = o.raw;
}
// dproj rows = dz rows of the patch tokens (bf16 copy); dcls += sum_b dz[b,0]; dpos[tok] += sum_b dz[b,tok]
__global__ void vit_assemble_bwd_kernel(const __nv_bfloat16* __restrict__ dz, __nv_bfloat16* __restrict__ dproj, float* __restrict__ dcls,
                                        float* __restrict__ dpos, int B, int nP, int d) {
  const int tok = blockIdx.x;  // one block per token position
  for (int col = threadIdx.x; col < d; col += blockDim.x) {
    float acc = 0.f;
    for (int b = 0; b < B; ++b) {
      const __nv_bfloat16 v = dz[((size_t)b * (nP + 1) + tok) * d + col];
      acc += __bfloat162float(v);
      if (tok > 0) dproj[((size_t)b * nP + tok - 1) * d + col] = v;
    }
    dpos[(size_t)tok * d + col] += acc;
    if (tok == 0) dcls[col] += acc;
  }
}
// CLS selector: out[b, :] = h[b*S, :] (fp32), and its backward (zeros elsewhere)
__global__ void cls_select_fwd_kernel(const __nv_bfloat16* __restrict__ h, float* __restrict__ out, int B, int S, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * d) return;
  out[i] = __bfloat162float(h[(size_t)(i / d) * S * d + (i % d)]);
}
__global__ void cls_select_bwd_kernel(const float* __restrict__ g, __nv_bfloat16* __restrict__ dh, int B, int S, int d) {
  const int64_t i = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= (int64_t)B * S * d) return;
  const int64_t row = i / d;
  const int b = (int)(row / S), tok = (int)(row % S);
  dh[i] = tok == 0 ? __float2bfloat16_rn(g[(size_t)b * d + (i % d)]) : __float2bfloat16_rn(0.f);
}

// ---------------------------------------------------------------------------------------------- attention pooling (MAP)
// Single-query multi-head attention over each sequence's keys/values (MultiHeadAttentionPooling's FlashAttentionPooling,
// models/biencoder/modeling_biencoder.py:93-152, layers/attention.py:313-440): one learned latent query per head, shared by
// every sequence.  kv [T, 2, H, 64] bf16 (packed tokens), q [H, 64] fp32, out [nseq, H, 64] fp32, lse [nseq, H].
// One CTA per (sequence, head), 128 threads.  HBM-bound: kv is read once (twice in the backward).
constexpr int kPoolThreads = 128;

__device__ __forceinline__ float pool_block_reduce(float v, float* red, bool is_max) {
  for (int o = 16; o > 0; o >>= 1) {
    const float w = __shfl_xor_sync(0xffffffffu, v, o);
    v = is_max ? fmaxf(v, w) : v + w;
  }
  __syncthreads();
  if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = v;
  __syncthreads();
  v = red[0];
  for (int i = 1; i < kPoolThreads / 32; ++i) v = is_max ? fmaxf(v, red[i]) : v + red[i];
  return v;
}

// scores of this sequence's keys against the head's query -> sc[len] (shared), returns nothing; dot over 64 dims per key
__device__ __forceinline__ void pool_scores(const __nv_bfloat16* __restrict__ kv, const float* qh, int begin, int len, int H, int head,
                                            float scale, float* sc) {
  for (int s = threadIdx.x; s < len; s += kPoolThreads) {
    const __nv_bfloat16* kr = kv + ((size_t)(begin + s) * 2 * H + head) * 64;
    float acc = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      BF8 v;
      float f[8];
      v.raw = *reinterpret_cast<const uint4*>(kr + c * 8);
      v.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) acc = fmaf(f[j], qh[c * 8 + j], acc);
    }
    sc[s] = acc * scale;
  }
}

__global__ void __launch_bounds__(kPoolThreads)
attn_pool_fwd_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kv, const int* __restrict__ cu, float* __restrict__ out,
                     float* __restrict__ lse, int H, float scale) {
  extern __shared__ float pool_smem[];
  float* sc = pool_smem;  // [max_len]
  __shared__ float red[kPoolThreads / 32];
  __shared__ float qh[64];
  __shared__ float part[2][64];
  const int seq = blockIdx.x, head = blockIdx.y;
  const int begin = cu[seq], len = cu[seq + 1] - begin;
  if (threadIdx.x < 64) qh[threadIdx.x] = q[head * 64 + threadIdx.x];
  __syncthreads();
  pool_scores(kv, qh, begin, len, H, head, scale, sc);
  __syncthreads();
  float m = -INFINITY;
  for (int s = threadIdx.x; s < len; s += kPoolThreads) m = fmaxf(m, sc[s]);
  m = pool_block_reduce(m, red, true);
  float l = 0.f;
  for (int s = threadIdx.x; s < len; s += kPoolThreads) {
    const float p = __expf(sc[s] - m);
    sc[s] = p;
    l += p;
  }
  l = pool_block_reduce(l, red, false);
  __syncthreads();
  // out[d] = sum_s p_s v[s][d] / l: thread = (key parity, dim)
  const int d = threadIdx.x & 63, par = threadIdx.x >> 6;
  float acc = 0.f;
  for (int s = par; s < len; s += 2) acc = fmaf(sc[s], __bfloat162float(kv[((size_t)(begin + s) * 2 * H + H + head) * 64 + d]), acc);
  part[par][d] = acc;
  __syncthreads();
  if (threadIdx.x < 64) out[((size_t)seq * H + head) * 64 + d] = (part[0][d] + part[1][d]) / l;
  if (threadIdx.x == 0) lse[seq * H + head] = m + __logf(l);
}

// dq [H, 64] fp32 is ACCUMULATED with atomics over sequences (zero it first); dkv [T, 2, H, 64] bf16 is written.
__global__ void __launch_bounds__(kPoolThreads)
attn_pool_bwd_kernel(const float* __restrict__ q, const __nv_bfloat16* __restrict__ kv, const int* __restrict__ cu,
                     const float* __restrict__ dout, const float* __restrict__ lse, float* __restrict__ dq,
                     __nv_bfloat16* __restrict__ dkv, int H, float scale) {
  extern __shared__ float pool_smem[];
  const int seq = blockIdx.x, head = blockIdx.y;
  const int begin = cu[seq], len = cu[seq + 1] - begin;
  float* sc = pool_smem;          // [max_len]  scores, then d(score)
  float* dp = pool_smem + ((len + 3) & ~3);  // [len]  p_s * <dout, v_s>, then d(score)
  __shared__ float red[kPoolThreads / 32];
  __shared__ float qh[64], go[64], dqh[64];
  if (threadIdx.x < 64) {
    qh[threadIdx.x] = q[head * 64 + threadIdx.x];
    go[threadIdx.x] = dout[((size_t)seq * H + head) * 64 + threadIdx.x];
    dqh[threadIdx.x] = 0.f;
  }
  __syncthreads();
  pool_scores(kv, qh, begin, len, H, head, scale, sc);
  __syncthreads();
  const float L = lse[seq * H + head];
  float dsum = 0.f;
  for (int s = threadIdx.x; s < len; s += kPoolThreads) {
    const float p = __expf(sc[s] - L);
    const __nv_bfloat16* vr = kv + ((size_t)(begin + s) * 2 * H + H + head) * 64;
    __nv_bfloat16* dvr = dkv + ((size_t)(begin + s) * 2 * H + H + head) * 64;
    float dot = 0.f;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      BF8 v, o;
      float f[8], g[8];
      v.raw = *reinterpret_cast<const uint4*>(vr + c * 8);
      v.unpack(f);
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dot = fmaf(f[j], go[c * 8 + j], dot);
        g[j] = p * go[c * 8 + j];  // dv_s = p_s * dout
      }
      o.pack(g);
      *reinterpret_cast<uint4*>(dvr + c * 8) = o.raw;
    }
    sc[s] = p;
    dp[s] = p * dot;
    dsum += p * dot;
  }
  dsum = pool_block_reduce(dsum, red, false);
  __syncthreads();
  for (int s = threadIdx.x; s < len; s += kPoolThreads) {
    const float ds = (dp[s] - sc[s] * dsum) * scale;  // d loss / d (q . k_s)
    const __nv_bfloat16* kr = kv + ((size_t)(begin + s) * 2 * H + head) * 64;
    __nv_bfloat16* dkr = dkv + ((size_t)(begin + s) * 2 * H + head) * 64;
#pragma unroll
    for (int c = 0; c < 8; ++c) {
      BF8 o;
      float g[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) g[j] = ds * qh[c * 8 + j];
      o.pack(g);
      *reinterpret_cast<uint4*>(dkr + c * 8) = o.raw;
    }
    dp[s] = ds;
  }
  __syncthreads();
  // dq[d] = sum_s ds_s k[s][d]
  const int d = threadIdx.x & 63, par = threadIdx.x >> 6;
  float acc = 0.f;
  for (int s = par; s < len; s += 2) acc = fmaf(dp[s], __bfloat162float(kv[((size_t)(begin + s) * 2 * H + head) * 64 + d]), acc);
  atomicAdd(&dqh[d], acc);
  __syncthreads();
  if (threadIdx.x < 64) atomicAdd(&dq[head * 64 + threadIdx.x], dqh[threadIdx.x]);
}

}  // namespace cx

using namespace cx;
#define STREAM static_cast<cudaStream_t>(stream)

static int ln_rows_grid(int rows, int warps) { return (rows + warps - 1) / warps; }

static DropParams make_drop(float p, unsigned long long seed) {
  DropParams dp;
  dp.p = (p > 0.f && p < 1.f) ? p : 0.f;
  dp.scale = dp.p > 0.f ? 1.f / (1.f - dp.p) : 1.f;
  dp.seed = seed;
  return dp;
}

extern "C" int cx_add_layernorm_fwd(const void* a, const void* b, const float* gamma, const float* beta, void* y, float* stats,
                                    int rows, int d, float eps, void* z_out, float p_drop, unsigned long long seed,
                                    cx_stream_t stream) {
  const DropParams dp = make_drop(p_drop, seed);
  CX_REQUIRE(a && y, "cx_add_layernorm_fwd: null pointer");
  CX_REQUIRE(d % 8 == 0 && d <= 256 * kLnMaxVec, "cx_add_layernorm_fwd: d must be a multiple of 8 and <= 1024");
  if (rows <= 0) return 0;
#define CX_LN_FWD(NV_)                                                                                              \
  add_layernorm_fwd_kernel<false, NV_><<<ln_rows_grid(rows, 8), 256, 0, STREAM>>>(                                  \
      (const __nv_bfloat16*)a, (const __nv_bfloat16*)b, nullptr, nullptr, nullptr, gamma, beta, (__nv_bfloat16*)y, stats, rows, d, eps, \
      (__nv_bfloat16*)z_out, dp)
  switch ((d + 255) / 256) {
    case 1: CX_LN_FWD(1); break;
    case 2: CX_LN_FWD(2); break;
    case 3: CX_LN_FWD(3); break;
    default: CX_LN_FWD(4); break;
  }
#undef CX_LN_FWD
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_embed_layernorm_fwd(const int64_t* ids, const int64_t* type_ids, const void* word_emb, const void* type_emb,
                                      const float* gamma, const float* beta, void* y, float* stats, int rows, int d, float eps,
                                      float p_drop, unsigned long long seed, cx_stream_t stream) {
  const DropParams dp = make_drop(p_drop, seed);
  CX_REQUIRE(ids && word_emb && type_emb && y, "cx_embed_layernorm_fwd: null pointer");
  CX_REQUIRE(d % 8 == 0 && d <= 256 * kLnMaxVec, "cx_embed_layernorm_fwd: d must be a multiple of 8 and <= 1024");
  if (rows <= 0) return 0;
#define CX_LN_FWD(NV_)                                                                                              \
  add_layernorm_fwd_kernel<true, NV_><<<ln_rows_grid(rows, 8), 256, 0, STREAM>>>(                                   \
      (const __nv_bfloat16*)word_emb, nullptr, ids, type_ids, (const __nv_bfloat16*)type_emb, gamma, beta, (__nv_bfloat16*)y, \
      stats, rows, d, eps, nullptr, dp)
  switch ((d + 255) / 256) {
    case 1: CX_LN_FWD(1); break;
    case 2: CX_LN_FWD(2); break;
    case 3: CX_LN_FWD(3); break;
    default: CX_LN_FWD(4); break;
  }
#undef CX_LN_FWD
  CX_LAUNCH_CHECK();
  return 0;
}

static int ln_bwd_grid(int rows) {
  int g = (rows + 7) / 8;
  const int cap = 2 * sm_count();  // 4x measured slower (72 vs 63 us at T = 32768)
  return g < cap ? (g < 1 ? 1 : g) : cap;
}

extern "C" size_t cx_layernorm_bwd_workspace_bytes(int d) { return (size_t)4 * sm_count() * 3 * d * sizeof(float); }

template <bool EMBED, int NV>
static int ln_bwd_launch(int grid, size_t smem, cudaStream_t st, const __nv_bfloat16* a, const __nv_bfloat16* b, const int64_t* ids,
                         const int64_t* type_ids, const __nv_bfloat16* type_emb, const __nv_bfloat16* g1, const __nv_bfloat16* g2,
                         const float* gamma, const float* stats, __nv_bfloat16* dz, float* dword, float* dtype_emb, float* partials,
                         int rows, int d, int64_t padding_idx, const __nv_bfloat16* gres, DropParams dp, __nv_bfloat16* da_out) {
  auto kern = add_layernorm_bwd_kernel<EMBED, NV>;
  CX_SET_SMEM_ONCE(kern, 8 * 3 * 1024 * 4);
  add_layernorm_bwd_kernel<EMBED, NV><<<grid, 256, smem, st>>>(a, b, ids, type_ids, type_emb, g1, g2, gamma, stats, dz, dword, dtype_emb,
                                                               partials, rows, d, padding_idx, gres, dp, da_out);
  CX_LAUNCH_CHECK();
  return 0;
}
template <bool EMBED, typename... Args>
static int ln_bwd_dispatch(int d, Args... args) {
  switch ((d + 255) / 256) {
    case 1: return ln_bwd_launch<EMBED, 1>(args...);
    case 2: return ln_bwd_launch<EMBED, 2>(args...);
    case 3: return ln_bwd_launch<EMBED, 3>(args...);
    default: return ln_bwd_launch<EMBED, 4>(args...);
  }
}

extern "C" int cx_add_layernorm_bwd(const void* a, const void* b, const void* g1, const void* g2, const float* gamma,
                                    const float* stats, void* dz, float* dgamma, float* dbeta, void* workspace, int rows, int d,
                                    const void* gres, float p_drop, unsigned long long seed, void* da_out, cx_stream_t stream) {
  const DropParams dp = make_drop(p_drop, seed);
  CX_REQUIRE(a && g1 && stats && dz, "cx_add_layernorm_bwd: null pointer");
  CX_REQUIRE(d % 8 == 0 && d <= 256 * kLnMaxVec, "cx_add_layernorm_bwd: d must be a multiple of 8 and <= 1024");
  CX_REQUIRE((dgamma == nullptr) == (dbeta == nullptr), "cx_add_layernorm_bwd: dgamma/dbeta go together");
  CX_REQUIRE(dgamma == nullptr || workspace != nullptr, "cx_add_layernorm_bwd: workspace required for parameter grads");
  if (rows <= 0) return 0;
  // narrow form: 384-thread blocks of 12 / WPR row slots, two blocks per SM (the workspace holds one partial per block and is
  // sized for 4 * SMs blocks)
  const int wpr = (d + 255) / 256;  // 1..4; 3 at d = 768
  const int rpb = 12 / wpr;
  int grid = (rows + rpb - 1) / rpb;
  if (grid > 2 * sm_count()) grid = 2 * sm_count();
  const size_t smem = (size_t)rpb * 2 * d * sizeof(float);
  float* part = dgamma ? (float*)workspace : (float*)nullptr;
#define CX_LN_NARROW(W)                                                                                                          \
  add_layernorm_bwd_narrow_kernel<W><<<grid, 384, smem, STREAM>>>((const __nv_bfloat16*)a, (const __nv_bfloat16*)b,               \
      (const __nv_bfloat16*)g1, (const __nv_bfloat16*)g2, gamma, stats, (__nv_bfloat16*)dz, part, rows, d,                          \
      (const __nv_bfloat16*)gres, dp, (__nv_bfloat16*)da_out)
  switch (wpr) {
    case 1: CX_LN_NARROW(1); break;
    case 2: CX_LN_NARROW(2); break;
    case 3: CX_LN_NARROW(3); break;
    default: CX_LN_NARROW(4); break;
  }
#undef CX_LN_NARROW
  CX_LAUNCH_CHECK();
  if (dgamma) {
    ln_param_grad_reduce_kernel<<<(2 * d * 8 + 255) / 256, 256, 0, STREAM>>>((const float*)workspace, grid, d, 2, dgamma, dbeta, nullptr);
    CX_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int cx_embed_layernorm_bwd(const int64_t* ids, const int64_t* type_ids, const void* word_emb, const void* type_emb,
                                      const void* g1, const void* g2, const float* gamma, const float* stats, float* dword,
                                      float* dtype_emb, float* dgamma, float* dbeta, void* workspace, int rows, int d,
                                      int64_t padding_idx, float p_drop, unsigned long long seed, cx_stream_t stream) {
  const DropParams dp = make_drop(p_drop, seed);
  CX_REQUIRE(ids && word_emb && type_emb && g1 && stats && dword && dtype_emb && dgamma && dbeta && workspace,
             "cx_embed_layernorm_bwd: null pointer");
  CX_REQUIRE(d % 8 == 0 && d <= 256 * kLnMaxVec, "cx_embed_layernorm_bwd: d must be a multiple of 8 and <= 1024");
  if (rows <= 0) return 0;
  const int grid = ln_bwd_grid(rows);
  const size_t smem = (size_t)8 * 3 * d * sizeof(float);
  int rc = ln_bwd_dispatch<true>(d, grid, smem, STREAM, (const __nv_bfloat16*)word_emb, (const __nv_bfloat16*)nullptr, ids, type_ids,
                                 (const __nv_bfloat16*)type_emb, (const __nv_bfloat16*)g1, (const __nv_bfloat16*)g2, gamma, stats,
                                 (__nv_bfloat16*)nullptr, dword, dtype_emb, (float*)workspace, rows, d, padding_idx,
                                 (const __nv_bfloat16*)nullptr, dp, (__nv_bfloat16*)nullptr);
  if (rc) return rc;
  // type_ids == NULL: every token is type 0, its embedding-row gradient is the third column-sum partial
  ln_param_grad_reduce_kernel<<<(3 * d * 8 + 255) / 256, 256, 0, STREAM>>>((const float*)workspace, grid, d, 3, dgamma, dbeta,
                                                                       type_ids == nullptr ? dtype_emb : nullptr);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_token_positions(const int32_t* cu_seqlens, int nseq, int32_t* pos, int32_t* seq_id, cx_stream_t stream) {
  CX_REQUIRE(cu_seqlens && pos, "cx_token_positions: null pointer");
  if (nseq <= 0) return 0;
  token_positions_kernel<<<nseq, 128, 0, STREAM>>>(cu_seqlens, nseq, pos, seq_id);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_rope_inplace(void* qkv, const int32_t* pos, const float* cos_t, const float* sin_t, int T, int H, int Dh,
                               int backward, int first_slot, int num_slots, cx_stream_t stream) {
  CX_REQUIRE(qkv && pos && cos_t && sin_t, "cx_rope_inplace: null pointer");
  CX_REQUIRE(Dh % 16 == 0, "cx_rope_inplace: head dim must be a multiple of 16");
  if (T <= 0) return 0;
  CX_REQUIRE(first_slot >= 0 && num_slots >= 1 && first_slot + num_slots <= 2, "cx_rope_inplace: slots are q (0) and k (1)");
  const int64_t n = (int64_t)T * num_slots * H * (Dh / 16);
  rope_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>((__nv_bfloat16*)qkv, pos, cos_t, sin_t, T, H, Dh, backward ? -1.f : 1.f,
                                                               first_slot, num_slots);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_dq_finalize_rope(const float* dq_acc, void* dqkv, const int32_t* pos, const float* cos_t, const float* sin_t,
                                   int T, int H, int Dh, cx_stream_t stream) {
  CX_REQUIRE(dq_acc && dqkv && pos && cos_t && sin_t, "cx_dq_finalize_rope: null pointer");
  CX_REQUIRE(Dh % 16 == 0, "cx_dq_finalize_rope: head dim must be a multiple of 16");
  if (T <= 0) return 0;
  const int64_t n = (int64_t)T * H * (Dh / 16);
  dq_finalize_rope_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(dq_acc, (__nv_bfloat16*)dqkv, pos, cos_t, sin_t, T, H, Dh);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_swiglu_fwd(const void* yg, void* out, int64_t T, int I, cx_stream_t stream) {
  CX_REQUIRE(yg && out, "cx_swiglu_fwd: null pointer");
  CX_REQUIRE(I % 8 == 0, "cx_swiglu_fwd: inner dim must be a multiple of 8");
  if (T <= 0) return 0;
  const int64_t n = T * (I / 8);
  swiglu_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>((const __nv_bfloat16*)yg, (__nv_bfloat16*)out, T, I);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_swiglu_bwd(const void* dout, const void* yg, void* dyg, int64_t T, int I, cx_stream_t stream) {
  CX_REQUIRE(dout && yg && dyg, "cx_swiglu_bwd: null pointer");
  CX_REQUIRE(I % 8 == 0, "cx_swiglu_bwd: inner dim must be a multiple of 8");
  if (T <= 0) return 0;
  const int64_t n = T * (I / 8);
  swiglu_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>((const __nv_bfloat16*)dout, (const __nv_bfloat16*)yg,
                                                                     (__nv_bfloat16*)dyg, T, I);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_mean_pool_fwd(const void* h, const int32_t* cu_seqlens, float* pooled, int nseq, int d, cx_stream_t stream) {
  CX_REQUIRE(h && cu_seqlens && pooled, "cx_mean_pool_fwd: null pointer");
  CX_REQUIRE(d % 8 == 0, "cx_mean_pool_fwd: d must be a multiple of 8");
  if (nseq <= 0) return 0;
  dim3 grid(nseq, (d + 255) / 256);
  mean_pool_fwd_kernel<<<grid, 256, 0, STREAM>>>((const __nv_bfloat16*)h, cu_seqlens, pooled, d);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_mean_pool_bwd(const float* dpooled, const int32_t* cu_seqlens, void* dh, int nseq, int d, cx_stream_t stream) {
  CX_REQUIRE(dpooled && cu_seqlens && dh, "cx_mean_pool_bwd: null pointer");
  CX_REQUIRE(d % 8 == 0, "cx_mean_pool_bwd: d must be a multiple of 8");
  if (nseq <= 0) return 0;
  mean_pool_bwd_kernel<<<nseq, 256, 0, STREAM>>>(dpooled, cu_seqlens, (__nv_bfloat16*)dh, d);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_embed_head_fwd(const float* pooled, float* out, float* save, int rows, int d, int hamming, int normalize,
                                 cx_stream_t stream) {
  CX_REQUIRE(pooled && out && save, "cx_embed_head_fwd: null pointer");
  if (rows <= 0) return 0;
  embed_head_fwd_kernel<<<(rows + 7) / 8, 256, 0, STREAM>>>(pooled, out, save, rows, d, hamming, normalize);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_embed_head_bwd(const float* pooled, const float* gout, const float* save, float* gpooled, int rows, int d,
                                 int hamming, int normalize, cx_stream_t stream) {
  CX_REQUIRE(pooled && gout && save && gpooled, "cx_embed_head_bwd: null pointer");
  if (rows <= 0) return 0;
  embed_head_bwd_kernel<<<(rows + 7) / 8, 256, 0, STREAM>>>(pooled, gout, save, gpooled, rows, d, hamming, normalize);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_grad_clip_coef(const float* grad, int64_t n, float max_norm, float* out2, void* workspace, cx_stream_t stream) {
  CX_REQUIRE(grad && out2 && workspace, "cx_grad_clip_coef: null pointer");
  const int blocks = 4 * sm_count();
  sumsq_kernel<<<blocks, 256, 0, STREAM>>>(grad, n, (float*)workspace);
  CX_LAUNCH_CHECK();
  clip_coef_kernel<<<1, 32, 0, STREAM>>>((const float*)workspace, blocks, max_norm, out2);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" size_t cx_grad_clip_workspace_bytes(void) { return (size_t)4 * sm_count() * sizeof(float); }

extern "C" int cx_adamw_step(float* param, float* grad, float* exp_avg, float* exp_avg_sq, void* shadow_bf16, int64_t n, float lr,
                             float beta1, float beta2, float eps, float weight_decay, int step, const float* grad_scale_dev,
                             float grad_scale, int zero_grad, cx_stream_t stream) {
  CX_REQUIRE(param && grad && exp_avg && exp_avg_sq, "cx_adamw_step: null pointer");
  CX_REQUIRE(step >= 1, "cx_adamw_step: step counts from 1");
  if (n <= 0) return 0;
  const float bc1 = 1.f - powf(beta1, (float)step), bc2 = 1.f - powf(beta2, (float)step);
  adamw_kernel<<<8 * sm_count(), 256, 0, STREAM>>>(param, grad, exp_avg, exp_avg_sq, (__nv_bfloat16*)shadow_bf16, n, lr, beta1, beta2,
                                                  eps, weight_decay, bc1, bc2, grad_scale_dev, grad_scale, zero_grad);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_cast_f32_bf16(const float* x, void* y, int64_t n, cx_stream_t stream) {
  CX_REQUIRE(x && y, "cx_cast_f32_bf16: null pointer");
  if (n <= 0) return 0;
  cast_f32_bf16_kernel<<<8 * sm_count(), 256, 0, STREAM>>>(x, (__nv_bfloat16*)y, n);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_colsum_bf16(const void* x, int64_t T, int N, float* out, cx_stream_t stream) {
  CX_REQUIRE(x && out, "cx_colsum_bf16: null pointer");
  CX_REQUIRE(N % 8 == 0, "cx_colsum_bf16: N must be a multiple of 8");
  if (T <= 0) return 0;
  const int rows_per_block = 256;
  dim3 grid((N / 8 + 127) / 128, (unsigned)((T + rows_per_block - 1) / rows_per_block));
  colsum_kernel<<<grid, 128, 0, STREAM>>>((const __nv_bfloat16*)x, T, N, out, rows_per_block);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" int cx_act_fwd(const void* x, void* y, int64_t n, int kind, cx_stream_t stream) {
  CX_REQUIRE(x && y, "cx_act_fwd: null pointer");
  CX_REQUIRE(n % 8 == 0 && (kind == 0 || kind == 1), "cx_act_fwd: n % 8 == 0, kind in {0 gelu, 1 quick_gelu}");
  if (n <= 0) return 0;
  act_fwd_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, STREAM>>>((const __nv_bfloat16*)x, (__nv_bfloat16*)y, n / 8, kind);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" int cx_act_bwd(const void* dy, const void* x, void* dx, int64_t n, int kind, cx_stream_t stream) {
  CX_REQUIRE(dy && x && dx, "cx_act_bwd: null pointer");
  CX_REQUIRE(n % 8 == 0 && (kind == 0 || kind == 1), "cx_act_bwd: n % 8 == 0, kind in {0 gelu, 1 quick_gelu}");
  if (n <= 0) return 0;
  act_bwd_kernel<<<(unsigned)((n / 8 + 255) / 256), 256, 0, STREAM>>>((const __nv_bfloat16*)dy, (const __nv_bfloat16*)x, (__nv_bfloat16*)dx, n / 8, kind);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" int cx_patchify(const float* pixels, void* out, int B, int C, int Himg, int Wimg, int patch, cx_stream_t stream) {
  CX_REQUIRE(pixels && out, "cx_patchify: null pointer");
  CX_REQUIRE(Himg % patch == 0 && Wimg % patch == 0, "cx_patchify: image size must be a multiple of the patch size");
  const int64_t n = (int64_t)B * C * Himg * Wimg;
  if (n <= 0) return 0;
  patchify_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(pixels, (__nv_bfloat16*)out, B, C, Himg, Wimg, patch);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" int cx_vit_assemble_fwd(const void* proj, const float* cls, const float* pos, void* z, int B, int nP, int d, cx_stream_t stream) {
  CX_REQUIRE(proj && cls && pos && z, "cx_vit_assemble_fwd: null pointer");
  CX_REQUIRE(d % 8 == 0, "cx_vit_assemble_fwd: d must be a multiple of 8");
  const int64_t n = (int64_t)B * (nP + 1) * (d / 8);
  if (n <= 0) return 0;
  vit_assemble_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>((const __nv_bfloat16*)proj, cls, pos, (__nv_bfloat16*)z, B, nP, d);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" int cx_vit_assemble_bwd(const void* dz, void* dproj, float* dcls, float* dpos, int B, int nP, int d, cx_stream_t stream) {
  CX_REQUIRE(dz && dproj && dcls && dpos, "cx_vit_assemble_bwd: null pointer");
  if (B <= 0) return 0;
  vit_assemble_bwd_kernel<<<nP + 1, 256, 0, STREAM>>>((const __nv_bfloat16*)dz, (__nv_bfloat16*)dproj, dcls, dpos, B, nP, d);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" int cx_cls_select_fwd(const void* h, float* out, int B, int S, int d, cx_stream_t stream) {
  CX_REQUIRE(h && out, "cx_cls_select_fwd: null pointer");
  const int64_t n = (int64_t)B * d;
  if (n <= 0) return 0;
  cls_select_fwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>((const __nv_bfloat16*)h, out, B, S, d);
  CX_LAUNCH_CHECK();
  return 0;
}
extern "C" int cx_cls_select_bwd(const float* g, void* dh, int B, int S, int d, cx_stream_t stream) {
  CX_REQUIRE(g && dh, "cx_cls_select_bwd: null pointer");
  const int64_t n = (int64_t)B * S * d;
  if (n <= 0) return 0;
  cls_select_bwd_kernel<<<(unsigned)((n + 255) / 256), 256, 0, STREAM>>>(g, (__nv_bfloat16*)dh, B, S, d);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_attn_pool_fwd(const float* q, const void* kv, const int32_t* cu_seqlens, float* out, float* lse, int nseq,
                                int max_seqlen, int H, int Dh, float softmax_scale, cx_stream_t stream) {
  CX_REQUIRE(q && kv && cu_seqlens && out && lse, "cx_attn_pool_fwd: null pointer");
  CX_REQUIRE(Dh == 64, "cx_attn_pool_fwd: only head_dim 64 is implemented");
  CX_REQUIRE(max_seqlen > 0 && max_seqlen <= 8192, "cx_attn_pool_fwd: sequences of 1..8192 tokens");
  if (nseq <= 0) return 0;
  const size_t smem = (size_t)((max_seqlen + 3) & ~3) * sizeof(float);
  attn_pool_fwd_kernel<<<dim3(nseq, H), kPoolThreads, smem, STREAM>>>(q, (const __nv_bfloat16*)kv, cu_seqlens, out, lse, H, softmax_scale);
  CX_LAUNCH_CHECK();
  return 0;
}

extern "C" int cx_attn_pool_bwd(const float* q, const void* kv, const int32_t* cu_seqlens, const float* dout, const float* lse,
                                float* dq, void* dkv, int nseq, int max_seqlen, int H, int Dh, float softmax_scale, cx_stream_t stream) {
  CX_REQUIRE(q && kv && cu_seqlens && dout && lse && dq && dkv, "cx_attn_pool_bwd: null pointer");
  CX_REQUIRE(Dh == 64, "cx_attn_pool_bwd: only head_dim 64 is implemented");
  CX_REQUIRE(max_seqlen > 0 && max_seqlen <= 4096, "cx_attn_pool_bwd: sequences of 1..4096 tokens");
  if (nseq <= 0) return 0;
  const size_t smem = (size_t)2 * ((max_seqlen + 3) & ~3) * sizeof(float);
  attn_pool_bwd_kernel<<<dim3(nseq, H), kPoolThreads, smem, STREAM>>>(q, (const __nv_bfloat16*)kv, cu_seqlens, dout, lse, dq,
                                                                     (__nv_bfloat16*)dkv, H, softmax_scale);
  CX_LAUNCH_CHECK();
  return 0;
}
