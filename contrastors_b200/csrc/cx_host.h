// Host-side helpers shared by the C-ABI translation units: error reporting, launch accounting, TMA tensor maps.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <mutex>
#include <string>

#include "../../include/contrastors_b200.h"

namespace cx {

void set_error(const std::string& msg);
int fail(int code, const std::string& msg);
extern std::atomic<unsigned long long> g_launches;
inline void count_launch(int n = 1) { g_launches.fetch_add(n, std::memory_order_relaxed); }

#define CX_CUDA_CHECK(expr)                                                                          \
  do {                                                                                               \
    cudaError_t _e = (expr);                                                                         \
    if (_e != cudaSuccess) return ::cx::fail(CX_ERR_CUDA, std::string(#expr) + ": " + cudaGetErrorString(_e)); \
  } while (0)

#define CX_REQUIRE(cond, msg)                                          \
  do {                                                                 \
    if (!(cond)) return ::cx::fail(CX_ERR_INVALID, std::string(msg));  \
  } while (0)

#define CX_LAUNCH_CHECK()                                                                              \
  do {                                                                                                 \
    cudaError_t _e = cudaGetLastError();                                                               \
    if (_e != cudaSuccess) return ::cx::fail(CX_ERR_CUDA, std::string("kernel launch: ") + cudaGetErrorString(_e)); \
    ::cx::count_launch();                                                                              \
  } while (0)

// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per process and per kernel instantiation, safe from any thread
// (the forward runs on the Python thread, the backward on autograd's worker thread): a `static bool` here was a data race.
#define CX_SET_SMEM_ONCE(kern, bytes)                                                                                 \
  do {                                                                                                                \
    static std::once_flag _once;                                                                                      \
    static cudaError_t _err = cudaSuccess;                                                                            \
    std::call_once(_once, [&] { _err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes); }); \
    CX_CUDA_CHECK(_err);                                                                                              \
  } while (0)

// 2D row-major tensor map: `inner` contiguous elements per row, `outer` rows, `row_stride_bytes` between rows.
// Returns 0 on success (sets the error string otherwise).
int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                 CUtensorMapSwizzle swizzle);

int sm_count();

}  // namespace cx
