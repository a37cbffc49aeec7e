// Host-side helpers shared by the C-ABI translation units: error reporting and TMA descriptor encoding.
#pragma once
#include <cuda.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/instantrestore_b200.h"

namespace ir {

// Records a message retrievable through ir_last_error_string() and returns `code`.
int set_error(int code, const char* fmt, ...);

// Checks the device behind the current context is sm_100 (B200). Cached.
int check_arch();

// Encodes a tiled fp16 tensor map (rank 2..5) with 128-byte swizzle and zero OOB fill.
// dims/box are innermost-first; strides_bytes[i] is the byte stride of dimension i+1.
int make_tmap_f16(CUtensorMap* out, const void* base, int rank, const uint64_t* dims, const uint64_t* strides_bytes,
                  const uint32_t* box);

// GroupNorm pass A as a stand-alone launch (ir_norm.cu): partial[(b * slabs + slab) * groups + g] = (mean, M2) over
// `rows_per_slab` pixels; slabs = ceil(hw / rows_per_slab).
int launch_gn_partial(const void* x, int row_stride, int batch, int hw, int channels, int groups, int rows_per_slab,
                      void* partial, cudaStream_t stream);

// Counts kernel launches issued (or captured into a CUDA graph) through the C ABI; read by ir_launch_count().
void count_launch();

// cudaFuncSetAttribute state is per device: one flag per device ordinal (benign race: setting it twice is idempotent).
struct PerDeviceOnce {
  bool done[64] = {};
  bool& slot() {
    int d = 0;
    cudaGetDevice(&d);
    return done[d & 63];
  }
};

// Programmatic dependent launch for the launches that follow (ir_set_pdl / IR_PDL=1; default off, see ir_host.cu).
bool pdl_enabled();

// Kernel launch with an optional thread-block cluster and the programmatic-stream-serialization attribute.
template <typename... KArgs, typename... Args>
inline cudaError_t launch_kernel(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, dim3 cluster,
                                 Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[2];
  unsigned n = 0;
  if (cluster.x * cluster.y * cluster.z > 1) {
    attr[n].id = cudaLaunchAttributeClusterDimension;
    attr[n].val.clusterDim.x = cluster.x;
    attr[n].val.clusterDim.y = cluster.y;
    attr[n].val.clusterDim.z = cluster.z;
    ++n;
  }
  if (pdl_enabled()) {
    attr[n].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[n].val.programmaticStreamSerializationAllowed = 1;
    ++n;
  }
  cfg.attrs = attr;
  cfg.numAttrs = n;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#define IR_LAUNCH(kernel, grid, block, smem, stream, ...) \
  (void)ir::launch_kernel(kernel, dim3(grid), dim3(block), smem, stream, dim3(1, 1, 1), __VA_ARGS__)

#define IR_CUDA_LAUNCH_CHECK(what)                                                      \
  do {                                                                                  \
    ir::count_launch();                                                                 \
    cudaError_t e__ = cudaGetLastError();                                               \
    if (e__ != cudaSuccess) return ir::set_error(IR_ERR_CUDA, "%s: %s", what, cudaGetErrorString(e__)); \
  } while (0)

}  // namespace ir
