// Diagnostic entry point: one 128 x N x 64 tcgen05.mma tile with every descriptor field supplied by the caller.
// Used by tools/gpu_probe.py to pin the UMMA shared-memory descriptor conventions (K-major vs MN-major operands,
// LBO/SBO, per-K-step address advance) on real hardware. Not on the product path.
#include "ir_host.h"
#include "ir_ptx.cuh"

namespace ir {

struct DebugParams {
  CUtensorMap tma_a, tma_b;
  int N, b_mn_major;
  uint32_t lbo_a, sbo_a, lbo_b, sbo_b, kadv_a, kadv_b;
  float* d;
};

__global__ void __launch_bounds__(128) debug_umma_kernel(const __grid_constant__ DebugParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;              // 16 KB
  uint8_t* sB = smem + 16384;      // up to 32 KB
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 16384 + 32768);
  uint32_t* slot = reinterpret_cast<uint32_t*>(bars + 2);
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    mbar_init(&bars[0], 1);
    mbar_init(&bars[1], 1);
    fence_mbar_init();
  }
  if (warp == 1) {
    tmem_alloc(slot, 256);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem = *slot;
  if (threadIdx.x == 0) {
    const uint32_t bytes = 16384 + p.N * 64 * 2;
    mbar_arrive_expect_tx(&bars[0], bytes);
    tma_load_2d(sA, &p.tma_a, &bars[0], 0, 0);
    if (p.b_mn_major) {
      for (int a = 0; a < p.N / 64; ++a) tma_load_2d(sB + a * 8192, &p.tma_b, &bars[0], a * 64, 0);
    } else {
      tma_load_2d(sB, &p.tma_b, &bars[0], 0, 0);
    }
    mbar_wait(&bars[0], 0);
    tc_fence_after();
    const uint32_t idesc = umma_idesc_f16(128, p.N, 0, p.b_mn_major ? 1 : 0);
    for (int k = 0; k < 4; ++k) {
      const uint64_t ad = umma_smem_desc(smem_u32(sA) + k * p.kadv_a, p.lbo_a, p.sbo_a);
      const uint64_t bd = umma_smem_desc(smem_u32(sB) + k * p.kadv_b, p.lbo_b, p.sbo_b);
      umma_f16_ss(tmem, ad, bd, idesc, k != 0);
    }
    umma_commit(&bars[1]);
  }
  __syncwarp();
  mbar_wait(&bars[1], 0);
  tc_fence_after();
  const int row = warp * 32 + lane;
  for (int c0 = 0; c0 < p.N; c0 += 32) {
    uint32_t r[32];
    tmem_ld32(tmem + (static_cast<uint32_t>(warp * 32) << 16) + c0, r);
    tmem_ld_wait();
    for (int i = 0; i < 32; ++i) p.d[row * p.N + c0 + i] = __uint_as_float(r[i]);
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) tmem_dealloc(tmem, 256);
}

}  // namespace ir

// a: fp16 [128, 64] (K contiguous). b: fp16 [N, 64] when b_mn_major == 0, else fp16 [64, N] (N contiguous).
// d: fp32 [128, N]. N in {64, 128, 256}.
extern "C" int ir_debug_umma(const void* a, const void* b, float* d, int N, int b_mn_major, unsigned lbo_a, unsigned sbo_a,
                             unsigned lbo_b, unsigned sbo_b, unsigned kadv_a, unsigned kadv_b, ir_stream_t stream) {
  using namespace ir;
  if (int rc = check_arch()) return rc;
  if (N != 64 && N != 128 && N != 256) return set_error(IR_ERR_SHAPE, "ir_debug_umma: N=%d", N);
  DebugParams p;
  memset(&p, 0, sizeof(p));
  {
    uint64_t dims[2] = {64, 128};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, 128};
    if (int rc = make_tmap_f16(&p.tma_a, a, 2, dims, str, box)) return rc;
  }
  if (b_mn_major) {
    uint64_t dims[2] = {static_cast<uint64_t>(N), 64};
    uint64_t str[1] = {static_cast<uint64_t>(N) * 2};
    uint32_t box[2] = {64, 64};
    if (int rc = make_tmap_f16(&p.tma_b, b, 2, dims, str, box)) return rc;
  } else {
    uint64_t dims[2] = {64, static_cast<uint64_t>(N)};
    uint64_t str[1] = {128};
    uint32_t box[2] = {64, static_cast<uint32_t>(N)};
    if (int rc = make_tmap_f16(&p.tma_b, b, 2, dims, str, box)) return rc;
  }
  p.N = N;
  p.b_mn_major = b_mn_major;
  p.lbo_a = lbo_a; p.sbo_a = sbo_a; p.lbo_b = lbo_b; p.sbo_b = sbo_b; p.kadv_a = kadv_a; p.kadv_b = kadv_b;
  p.d = d;
  const int smem = 16384 + 32768 + 64 + 1024;
  cudaError_t e = cudaFuncSetAttribute(debug_umma_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
  if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaFuncSetAttribute(debug_umma): %s", cudaGetErrorString(e));
  debug_umma_kernel<<<1, 128, smem, static_cast<cudaStream_t>(stream)>>>(p);
  IR_CUDA_LAUNCH_CHECK("debug_umma launch");
  return 0;
}
