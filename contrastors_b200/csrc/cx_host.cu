// Error reporting, launch accounting, and TMA tensor-map encoding (driver entry point fetched at run time so the
// library links and loads on machines without libcuda -- compute calls still fail loudly there).
#include "cx_host.h"

#include <mutex>

namespace cx {

static thread_local std::string t_last_error;
std::atomic<unsigned long long> g_launches{0};

void set_error(const std::string& msg) { t_last_error = msg; }
int fail(int code, const std::string& msg) {
  t_last_error = msg;
  return code;
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  static std::once_flag once;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  });
  return fn;
}

int make_tmap_2d(CUtensorMap* out, CUtensorMapDataType dtype, size_t elem_bytes, const void* base, uint64_t inner,
                 uint64_t outer, uint64_t row_stride_bytes, uint32_t box_inner, uint32_t box_outer,
                 CUtensorMapSwizzle swizzle) {
  EncodeTiledFn fn = get_encode_fn();
  if (!fn) return fail(CX_ERR_CUDA, "cuTensorMapEncodeTiled unavailable (no CUDA driver / no GPU)");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return fail(CX_ERR_INVALID, "tensor base must be 16-byte aligned");
  if ((row_stride_bytes & 15) != 0) return fail(CX_ERR_INVALID, "row stride must be a multiple of 16 bytes");
  if (box_inner * elem_bytes > 128 && swizzle == CU_TENSOR_MAP_SWIZZLE_128B)
    return fail(CX_ERR_INVALID, "box inner extent exceeds the 128B swizzle span");
  cuuint64_t dims[2] = {inner, outer};
  cuuint64_t strides[1] = {row_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer};
  cuuint32_t estr[2] = {1, 1};
  CUresult r = fn(out, dtype, 2, const_cast<void*>(base), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                  swizzle, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) return fail(CX_ERR_CUDA, "cuTensorMapEncodeTiled failed with code " + std::to_string((int)r));
  return 0;
}

int sm_count() {
  static int n = 0;
  if (n == 0) {
    int dev = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    cudaDeviceProp p;
    if (cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 148;
    n = p.multiProcessorCount;
  }
  return n;
}

}  // namespace cx

extern "C" {
const char* cx_last_error(void) { return cx::t_last_error.c_str(); }
int cx_version(void) { return 100; }
unsigned long long cx_launch_count(void) { return cx::g_launches.load(); }
}
