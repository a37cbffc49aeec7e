// Host-side plumbing of the C ABI: thread-local error string, device check, TMA descriptor encoding
// (cuTensorMapEncodeTiled resolved through the runtime's driver entry point, so libcuda is not a link dependency).
#include "ir_host.h"

#include <cudaTypedefs.h>
#include <stdarg.h>
#include <stdio.h>
#include <stdlib.h>

#include <atomic>
#include <mutex>

namespace ir {

static thread_local char g_err[512] = "ok";
static std::atomic<unsigned long long> g_launches{0};

// Programmatic dependent launch is OFF by default: with several requests in flight on separate streams (the throughput
// configuration) dependents that are resident but blocked in griddepcontrol.wait hold SM resources another stream's
// kernels would use (same-box A/B: 56.8 vs 58.4 images/s at B=1 x 3 streams, 66.9 vs 67.4 at B=8). It shortens the
// critical path of a single request in flight; ir_set_pdl(1) or IR_PDL=1 turns it on for the launches (and graph
// captures) that follow.
static std::atomic<int> g_pdl{-1};
bool pdl_enabled() {
  int v = g_pdl.load(std::memory_order_relaxed);
  if (v < 0) {
    const char* e = getenv("IR_PDL");
    v = (e && e[0] == '1') ? 1 : 0;
    g_pdl.store(v, std::memory_order_relaxed);
  }
  return v != 0;
}

void count_launch() { g_launches.fetch_add(1, std::memory_order_relaxed); }

int set_error(int code, const char* fmt, ...) {
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(g_err, sizeof(g_err), fmt, ap);
  va_end(ap);
  return code;
}

int check_arch() {
  static std::mutex mu;
  static int cached_dev = -1, cached_rc = 0;
  int dev = 0;
  cudaError_t e = cudaGetDevice(&dev);
  if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaGetDevice: %s", cudaGetErrorString(e));
  std::lock_guard<std::mutex> lk(mu);
  if (dev == cached_dev) return cached_rc;
  int major = 0, minor = 0;
  e = cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev);
  if (e == cudaSuccess) e = cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev);
  if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaDeviceGetAttribute: %s", cudaGetErrorString(e));
  cached_dev = dev;
  cached_rc = (major == 10) ? 0 : IR_ERR_ARCH;
  if (cached_rc) set_error(IR_ERR_ARCH, "device %d is sm_%d%d; these kernels are built for sm_100a only", dev, major, minor);
  return cached_rc;
}

typedef CUresult (*encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static encode_tiled_fn get_encode() {
  static std::once_flag once;
  static encode_tiled_fn fn = nullptr;
  std::call_once(once, [] {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
        q == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<encode_tiled_fn>(p);
  });
  return fn;
}

int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box) {
  encode_tiled_fn fn = get_encode();
  if (!fn) return set_error(IR_ERR_CUDA, "cuTensorMapEncodeTiled entry point unavailable");
  if ((reinterpret_cast<uintptr_t>(base) & 15) != 0) return set_error(IR_ERR_ALIGN, "TMA base pointer not 16-byte aligned");
  cuuint64_t gdim[5];
  cuuint64_t gstr[4];
  cuuint32_t bdim[5], estr[5];
  for (int i = 0; i < rank; ++i) {
    gdim[i] = dims[i];
    bdim[i] = box[i];
    estr[i] = 1;
    if (box[i] == 0 || box[i] > 256) return set_error(IR_ERR_SHAPE, "TMA box dim %d = %u out of range", i, box[i]);
  }
  for (int i = 0; i + 1 < rank; ++i) {
    gstr[i] = strides_bytes[i];
    if (strides_bytes[i] % 16 != 0) return set_error(IR_ERR_ALIGN, "TMA stride %d = %llu not a multiple of 16 bytes", i,
                                                      (unsigned long long)strides_bytes[i]);
  }
  CUresult r = fn(out, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, static_cast<cuuint32_t>(rank), const_cast<void*>(base), gdim,
                  gstr, bdim, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                  CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) {
    return set_error(IR_ERR_CUDA,
                     "cuTensorMapEncodeTiled failed (%d): rank %d dims [%llu,%llu,%llu,%llu] box [%u,%u,%u,%u]", (int)r,
                     rank, (unsigned long long)dims[0], (unsigned long long)(rank > 1 ? dims[1] : 0),
                     (unsigned long long)(rank > 2 ? dims[2] : 0), (unsigned long long)(rank > 3 ? dims[3] : 0), box[0],
                     rank > 1 ? box[1] : 0, rank > 2 ? box[2] : 0, rank > 3 ? box[3] : 0);
  }
  return 0;
}

}  // namespace ir

extern "C" const char* ir_last_error_string(void) { return ir::g_err; }
extern "C" int ir_version(void) { return 104; }  // 104: ir_conv_gemm_params += col_partial, col_begin; ir_adain_coeffs_params += own_partial, ref_partial; 103: ir_groupnorm_params += fused, ir_groupnorm_fused_supported; 102: ir_conv_gemm_params += cta_pair, gn_partial, gn_groups, halo; ir_groupnorm_params += partial_in
extern "C" unsigned long long ir_launch_count(void) { return ir::g_launches.load(std::memory_order_relaxed); }
extern "C" int ir_check_device(void) { return ir::check_arch(); }
extern "C" int ir_set_pdl(int enabled) {
  const int prev = ir::pdl_enabled() ? 1 : 0;
  ir::g_pdl.store(enabled ? 1 : 0, std::memory_order_relaxed);
  return prev;
}
