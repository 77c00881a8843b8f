// Micro-benchmarks of the SM-local resources the attention kernels lean on (B200): tcgen05.ld / tcgen05.st throughput per SM
// as a function of the number of warps, and MUFU.EX2 throughput.  Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3
#include <cstdio>
#include <cstdint>
#include <cuda_runtime.h>
#include "../../contrastors_b200/csrc/cx_ptx.cuh"
using namespace cx;

template <int kInFlight>
__global__ void ldtm_kernel(int iters, long long* out, float* sink) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  float acc = 0.f;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    uint32_t v[kInFlight][32];
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) tmem_ld_32x32(base + ((it * kInFlight + k) & 15) * 32, v[k]);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < kInFlight; ++k) acc += __uint_as_float(v[k][0]) + __uint_as_float(v[k][13]) + __uint_as_float(v[k][31]);
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (acc == 123.456f) sink[0] = acc;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tptr);
}

__global__ void sttm_kernel(int iters, long long* out) {
  __shared__ uint32_t tptr;
  const int warp = threadIdx.x >> 5;
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t base = tptr + (static_cast<uint32_t>((warp & 3) * 32) << 16);
  uint32_t v[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) v[i] = threadIdx.x + i;
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
    tmem_st_32x32(base + (it & 15) * 32, v);
    tmem_st_32x32(base + ((it + 7) & 15) * 32, v);
    tmem_st_wait();
  }
  __syncthreads();
  const long long t1 = clock64();
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tptr);
}

__global__ void mufu_kernel(int iters, long long* out, float* sink) {
  float x[16];
#pragma unroll
  for (int i = 0; i < 16; ++i) x[i] = -0.001f * (threadIdx.x + i);
  __syncthreads();
  const long long t0 = clock64();
  for (int it = 0; it < iters; ++it) {
#pragma unroll
    for (int i = 0; i < 16; ++i) x[i] = fast_exp2(x[i]) - 1.0f;
  }
  __syncthreads();
  const long long t1 = clock64();
  float s = 0.f;
#pragma unroll
  for (int i = 0; i < 16; ++i) s += x[i];
  if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  if (s == 123.456f) sink[0] = s;
}

// UTCHMMA issue->completion rate: warp 0 (converged, one elected lane) issues `n` back-to-back MMAs on garbage operands.
// variant 0: SS N=128 (A,B K-major)  1: SS N=64 (A K-major, B MN-major)  2: TS N=64 (A from TMEM, B MN-major)
//         3: SS N=64 (A,B MN-major)  4: SS N=256 (A,B K-major)          5: SS N=64 (A,B K-major)
template <int variant, int nacc>
__global__ void mma_kernel(int n, long long* out) {
  extern __shared__ __align__(1024) uint8_t smem[];
  __shared__ uint32_t tptr;
  __shared__ uint64_t bar;
  const int warp = threadIdx.x >> 5;
  if (threadIdx.x == 0) {
    mbar_init(&bar, 1);
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc<512>(&tptr);
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  
  if (warp == 0) {
    const uint64_t ak = make_smem_desc_sw128(smem_u32(smem), 0, 1024);
    const uint64_t bk = make_smem_desc_sw128(smem_u32(smem + 32768), 0, 1024);
    const uint64_t am = make_smem_desc_sw128(smem_u32(smem), 16384, 1024);
    const uint64_t bm = make_smem_desc_sw128(smem_u32(smem + 32768), 8192, 1024);
    const long long t0 = clock64();
    if (elect_one()) {
      for (int i = 0; i < n; i += 4) {
#pragma unroll
        for (int kk = 0; kk < 4; ++kk) {
          const uint32_t tb = tptr + ((kk % nacc) * (variant == 4 ? 256 : (variant == 0 ? 128 : 64)));  // rotate among nacc independent accumulators (<= 256 columns)
          if (variant == 0) umma_f16_ss(tb, ak + ((kk * 32) >> 4), bk + ((kk * 32) >> 4), make_idesc_bf16(128, 128, 0, 0), 1u);
          else if (variant == 1) umma_f16_ss(tb, ak + ((kk * 32) >> 4), bm + ((kk * 2048) >> 4), make_idesc_bf16(128, 64, 0, 1), 1u);
          else if (variant == 2) umma_f16_ts(tb, tptr + 384 + kk * 8, bm + ((kk * 2048) >> 4), make_idesc_bf16(128, 64, 0, 1), 1u);
          else if (variant == 3) umma_f16_ss(tb, am + ((kk * 2048) >> 4), bm + ((kk * 2048) >> 4), make_idesc_bf16(128, 64, 1, 1), 1u);
          else if (variant == 4) umma_f16_ss(tb, ak + ((kk * 32) >> 4), bk + ((kk * 32) >> 4), make_idesc_bf16(128, 256, 0, 0), 1u);
          else umma_f16_ss(tb, ak + ((kk * 32) >> 4), bk + ((kk * 32) >> 4), make_idesc_bf16(128, 64, 0, 0), 1u);
        }
      }
      umma_commit(&bar);
    }
    __syncwarp();
    mbar_wait(&bar, 0);
    const long long t1 = clock64();
    if (threadIdx.x == 0) out[blockIdx.x] = t1 - t0;
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 0) tmem_dealloc<512>(tptr);
}

static double med(long long* h, int n) {
  double s = 0;
  for (int i = 0; i < n; ++i) s += (double)h[i];
  return s / n;
}

int main() {
  const int nb = 148;
  long long *d, h[nb];
  float* sink;
  cudaMalloc(&d, nb * sizeof(long long));
  cudaMalloc(&sink, 4);
  const int iters = 2000;
  for (int threads : {256}) {
    ldtm_kernel<1><<<nb, threads>>>(iters, d, sink);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c1 = med(h, nb);
    ldtm_kernel<2><<<nb, threads>>>(iters, d, sink);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c2 = med(h, nb);
    ldtm_kernel<3><<<nb, threads>>>(iters, d, sink);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    double c3 = med(h, nb);
    const double bytes = (double)iters * (threads / 32) * 32 * 32 * 4;  // per CTA, one x32 load per warp per iteration
    printf("LDTM 32x32b.x32, %2d warps/SM: 1 in flight %.1f B/clk/SM, 2 in flight %.1f, 3 in flight %.1f   (err %s)\n", threads / 32,
           bytes / c1, 2 * bytes / c2, 3 * bytes / c3, cudaGetErrorString(cudaGetLastError()));
    sttm_kernel<<<nb, threads>>>(iters, d);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("STTM 32x32b.x32, %2d warps/SM: %.1f B/clk/SM\n", threads / 32, 2 * bytes / med(h, nb));
    mufu_kernel<<<nb, threads>>>(iters, d, sink);
    cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);
    printf("MUFU.EX2 (+FADD), %2d warps/SM: %.2f ex2/clk/SM\n", threads / 32, (double)iters * 16 * threads / med(h, nb));
  }
  const char* names[6] = {"SS M128 N128 K16 (A,B K-major)", "SS M128 N64 (A K-major, B MN-major)", "TS M128 N64 (A tmem, B MN-major)",
                          "SS M128 N64 (A,B MN-major)", "SS M128 N256 (A,B K-major)", "SS M128 N64 (A,B K-major)"};
  const int n = 4096;
#define RUN_MMA(V, A)                                                                                   \
  cudaFuncSetAttribute(mma_kernel<V, A>, cudaFuncAttributeMaxDynamicSharedMemorySize, 98304);           \
  mma_kernel<V, A><<<nb, 128, 98304>>>(n, d);                                                            \
  cudaMemcpy(h, d, sizeof(h), cudaMemcpyDeviceToHost);                                                   \
  printf("UTCHMMA %-40s: %.1f clk per MMA with %d accumulator(s) (%s)\n", names[V], med(h, nb) / n, A, \
         cudaGetErrorString(cudaGetLastError()));
  RUN_MMA(0, 1) RUN_MMA(0, 2) RUN_MMA(1, 1) RUN_MMA(1, 2) RUN_MMA(1, 4) RUN_MMA(2, 1) RUN_MMA(2, 2) RUN_MMA(3, 1) RUN_MMA(3, 2)
  RUN_MMA(3, 4) RUN_MMA(4, 1) RUN_MMA(5, 1) RUN_MMA(5, 2)
  cudaDeviceSynchronize();
  printf("done: %s\n", cudaGetErrorString(cudaGetLastError()));
  return 0;
}
