// ir_conv_gemm — im2col-free implicit GEMM on tcgen05 (sm_100a).
//
// One CTA computes a 128 (pixels/tokens) x BN (output channels) tile:
//   warp 0 (one elected lane): TMA producer. Per K step it loads a 128x64 fp16 activation box straight from the
//       channel-last tensor — for a 3x3 tap the box is shifted by (dy, dx) and TMA's out-of-bounds zero fill IS the
//       convolution padding; a stride-2 conv reads one of four phase-strided tensor maps — plus a BNx64 weight box.
//       Both land in 128-byte-swizzled shared memory, exactly the K-major layout tcgen05.mma consumes.
//   warp 1 (one elected lane): issues 4 x tcgen05.mma (M=128, N=BN, K=16) per stage, accumulating in TMEM, and
//       releases the stage with tcgen05.commit.
//   all 4 warps: epilogue — tcgen05.ld the accumulator (thread == row), + bias, fp16 round, + residual,
//       optional GEGLU / SiLU, 16-byte stores.
// Two CTAs are resident per SM (<= 113 KB smem, <= 256 TMEM columns each) so one CTA's epilogue overlaps the
// other's main loop.
//
// Persistent variant (conv_gemm_persistent_kernel, used whenever K is not split): one CTA per SM loops over output
// tiles; the TMEM accumulator is double-buffered (2 x BN columns) and a dedicated epilogue warpgroup pair (8 warps)
// drains tile i while the producer / MMA warps are already deep into tile i+1, so the epilogue (bias, fp16 round,
// residual, GEGLU with a polynomial erf) and the per-CTA set-up (barrier init, TMEM allocation, pipeline fill)
// disappear from the critical path of short-K layers (VAE 128-channel convolutions, K = 320 linears, GEGLU).
//
// Split-K over a thread-block cluster (small-M layers: 8x8 / 16x16 resolution convs and the B=1 linears launch only
// 10-80 output tiles on 148 SMs and are weight-bandwidth bound): gridDim.z = cluster size S in {2,4,8}; CTA z
// accumulates K-blocks [z*kb, (z+1)*kb) in its own TMEM, then the S partial tiles are reduce-scattered through
// distributed shared memory — CTA z pushes column slice j of its accumulator into CTA j's (now idle) pipeline
// buffers with st.shared::cluster, and CTA j sums the S slices in rank order (deterministic) and runs the epilogue
// for its BN/S columns. No global workspace, no atomics, no second kernel.
//
// Roofline: tensor-core bound for C_in*taps >= ~600 (AI = 2*128*BN*K / ((128+BN)*K*2 B)); algorithmic FLOPs per
// launch = 2 * M * c_out * taps * c_in.
#include "ir_host.h"
#include "ir_ptx.cuh"

namespace ir {

constexpr int kBM = 128;
constexpr int kBK = 64;
constexpr int kABytes = kBM * kBK * 2;  // 16 KB

struct GemmKParams {
  CUtensorMap tma_a[4];
  CUtensorMap tma_b;
  CUtensorMap tma_out;   // TMA-store epilogue: out as [M rows][N columns]; persistent kernel (TSTORE) 64-column x 128-row boxes,
                         // 128-wide halo pair 64-column x 32-row boxes (one per epilogue warp)
  CUtensorMap tma_res;   // 128-wide halo pair: the residual, same boxes (TMA load into the warp's staging tile)
  int M, N, n_out;
  int kc_per_tap, taps;
  int tiles_w, tiles_h;
  int bw, bh, bn;
  int8_t tap_map[16], tap_dx[16], tap_dy[16];   // upsample mode: [phase * 4 + tap]
  int split, kb_per_split;
  int m_tiles;
  // nearest-2x upsample folded into the convolution (ir_conv_gemm_params.upsample2x): M tiles [phase * mtp, (phase+1) * mtp)
  // belong to output sub-pixel phase (py, px) = (phase / 2, phase % 2); M, the tile boxes and `grow` are low-resolution
  // quantities, the weight rows of phase ph start at ph * N, and low-resolution pixel row grow = (b*H + y)*W + x is stored
  // at output pixel (b, 2y + py, 2x + px) = row 4*grow - 2*x + py*2W + px.
  int up, mtp, up_wmask, up_w2;
  int img_h, img_w;     // halo kernel: output (= input) image size
  float2* gn_partial;   // optional: per-(image, 32-row slab, group) (mean, M2) of the stored outputs (GroupNorm pass A)
  int gn_cpg, gn_hw, gn_groups;
  float2* col_partial;  // optional: per-(32-row slab, column >= col_begin) (mean, M2) of the stored outputs (AdaIN statistics)
  int col_begin;
  int wide_io;          // 256-bit epilogue loads / stores (rows are 32-byte aligned)
  const float* bias;
  const __half* residual;
  int res_stride;
  __half* out;
  int out_stride;
  int act;
};

__device__ __forceinline__ float gelu_erf(float x) { return 0.5f * x * (1.0f + erff(x * 0.70710678118654752f)); }
__device__ __forceinline__ float silu(float x) { return x / (1.0f + __expf(-x)); }
// gelu_erf with erf from Abramowitz-Stegun 7.1.26 (|error| <= 1.5e-7, far below one fp16 ulp): 2 MUFU + ~12 FMA
__device__ __forceinline__ float gelu_erf_fast(float x) {
  const float z = fabsf(x) * 0.70710678118654752f;
  const float t = __fdividef(1.0f, fmaf(0.3275911f, z, 1.0f));
  float poly = fmaf(1.061405429f, t, -1.453152027f);
  poly = fmaf(poly, t, 1.421413741f);
  poly = fmaf(poly, t, -0.284496736f);
  poly = fmaf(poly, t, 0.254829592f);
  const float erf_abs = 1.0f - poly * t * fast_exp2(-1.4426950408889634f * z * z);
  return 0.5f * x * (1.0f + copysignf(erf_abs, x));
}
__device__ __forceinline__ float round_h(float x) { return __half2float(__float2half_rn(x)); }


// Epilogue for NC (8, 16 or 32) consecutive output columns of one row: + bias, fp16 round, + residual, activation,
// 16-byte stores. v[] holds the fp32 accumulators.
__device__ __forceinline__ long out_row(const GemmKParams& p, long grow, int phase) {
  if (!p.up) return grow;
  return 4 * grow - 2 * (grow & p.up_wmask) + (phase >> 1) * p.up_w2 + (phase & 1);
}

template <int NC>
__device__ __forceinline__ void epilogue_store(float (&v)[NC], const GemmKParams& p, long grow, int gcol, int phase = 0) {
  if (gcol >= p.N) return;
  const bool full = (gcol + NC <= p.N);
  if (p.bias) {
#pragma unroll
    for (int i = 0; i < NC; ++i)
      if (full || gcol + i < p.N) v[i] += __ldg(p.bias + gcol + i);
  }
#pragma unroll
  for (int i = 0; i < NC; ++i) v[i] = round_h(v[i]);
  if (p.residual) {
    const __half* rp = p.residual + grow * p.res_stride + gcol;
    if (full) {
#pragma unroll
      for (int q = 0; q < NC / 8; ++q) {
        uint4 u = __ldg(reinterpret_cast<const uint4*>(rp) + q);
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __half22float2(h2[j]);
          v[q * 8 + 2 * j] += f.x;
          v[q * 8 + 2 * j + 1] += f.y;
        }
      }
    } else {
      for (int i = 0; i < NC; ++i)
        if (gcol + i < p.N) v[i] += __half2float(rp[i]);
    }
  }
  if (p.act == IR_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < NC; ++i) v[i] = silu(v[i]);
  }
  __half* op = p.out + out_row(p, grow, phase) * p.out_stride + gcol;
  if (full) {
#pragma unroll
    for (int q = 0; q < NC / 8; ++q) {
      uint4 u = make_uint4(pack_half2(v[q * 8], v[q * 8 + 1]), pack_half2(v[q * 8 + 2], v[q * 8 + 3]),
                           pack_half2(v[q * 8 + 4], v[q * 8 + 5]), pack_half2(v[q * 8 + 6], v[q * 8 + 7]));
      reinterpret_cast<uint4*>(op)[q] = u;
    }
  } else {
    for (int i = 0; i < NC; ++i)
      if (gcol + i < p.N) op[i] = __float2half_rn(v[i]);
  }
}

// AdaIN statistics fused into the GEMM epilogue (the V third of a fused QKV projection): per-column (mean, M2) over
// the warp's 32 rows (= 32 consecutive tokens of one image) of the fp16-rounded outputs. A reduce-scatter butterfly over
// the 32 lanes (16 + 8 + 4 + 2 + 1 exchanges per quantity, fixed tree: deterministic) leaves column `lane` of the chunk
// in lane `lane`:  col_partial[(row / 32) * (N - col_begin) + (col - col_begin)] = (mean, M2).
__device__ __forceinline__ float col_reduce_scatter32(float (&t)[32], int lane) {
  int off = 16;
#pragma unroll
  for (int h = 16; h >= 1; h >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float send = up ? t[i] : t[i + h], keep = up ? t[i + h] : t[i];
      t[i] = keep + __shfl_xor_sync(0xffffffffu, send, off);
    }
    off >>= 1;
  }
  return t[0];
}
__device__ __forceinline__ void col_stats32(const float (&v)[32], const GemmKParams& p, long grow, int gcol) {
  const int lane = threadIdx.x & 31;
  float t[32];
#pragma unroll
  for (int i = 0; i < 32; ++i) t[i] = round_h(v[i]);
  const float s = col_reduce_scatter32(t, lane);
#pragma unroll
  for (int i = 0; i < 32; ++i) { const float r = round_h(v[i]); t[i] = r * r; }
  const float q = col_reduce_scatter32(t, lane);
  const float mean = s * (1.0f / 32.0f);
  p.col_partial[((grow - lane) >> 5) * (p.N - p.col_begin) + (gcol - p.col_begin) + lane] = make_float2(mean, fmaxf(q - s * mean, 0.f));
}
// the same for 8 consecutive columns (split-K epilogue): 4 + 2 + 1 exchanges scatter the columns over lane bits 4..2,
// two more butterfly steps finish the sum over the remaining 4 lanes
__device__ __forceinline__ void col_stats8(const float (&v)[8], const GemmKParams& p, long grow, int gcol) {
  const int lane = threadIdx.x & 31;
  float s[8], q[8];
#pragma unroll
  for (int i = 0; i < 8; ++i) { s[i] = round_h(v[i]); q[i] = s[i] * s[i]; }
  int off = 16;
#pragma unroll
  for (int h = 4; h >= 1; h >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float s_send = up ? s[i] : s[i + h], s_keep = up ? s[i + h] : s[i];
      const float q_send = up ? q[i] : q[i + h], q_keep = up ? q[i + h] : q[i];
      s[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, off);
      q[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (int o = 2; o >= 1; o >>= 1) {
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
  }
  if ((lane & 3) == 0) {
    const float mean = s[0] * (1.0f / 32.0f);
    p.col_partial[((grow - lane) >> 5) * (p.N - p.col_begin) + (gcol - p.col_begin) + (lane >> 2)] =
        make_float2(mean, fmaxf(q[0] - s[0] * mean, 0.f));
  }
}

// GroupNorm pass A fused into the conv epilogue: (mean, M2) of one warp's 32 rows x 32 columns (= 32 / CPG whole groups)
// of the fp16-rounded outputs. Per-lane group sums, then a reduce-scatter butterfly over the 32 lanes (each halving
// step keeps half of the groups) — fixed summation tree, deterministic. The slab format is the one gn_merge_kernel
// reads (rows_per_slab = 32).
template <int CPG>
__device__ __forceinline__ void gn_stats_chunk(const float (&v)[32], int lane, float2* __restrict__ dst) {
  constexpr int G = 32 / CPG;
  constexpr int LOG2G = G == 8 ? 3 : G == 4 ? 2 : 1;
  float s[G], q[G];
#pragma unroll
  for (int g = 0; g < G; ++g) {
    s[g] = 0.f;
    q[g] = 0.f;
#pragma unroll
    for (int j = 0; j < CPG; ++j) {
      const float r = round_h(v[g * CPG + j]);
      s[g] += r;
      q[g] = fmaf(r, r, q[g]);
    }
  }
  int off = 16;
#pragma unroll
  for (int h = G / 2; h >= 1; h >>= 1) {
    const bool up = (lane & off) != 0;
#pragma unroll
    for (int i = 0; i < h; ++i) {
      const float s_send = up ? s[i] : s[i + h], s_keep = up ? s[i + h] : s[i];
      const float q_send = up ? q[i] : q[i + h], q_keep = up ? q[i + h] : q[i];
      s[i] = s_keep + __shfl_xor_sync(0xffffffffu, s_send, off);
      q[i] = q_keep + __shfl_xor_sync(0xffffffffu, q_send, off);
    }
    off >>= 1;
  }
#pragma unroll
  for (int o = 16 >> LOG2G; o >= 1; o >>= 1) {
    s[0] += __shfl_xor_sync(0xffffffffu, s[0], o);
    q[0] += __shfl_xor_sync(0xffffffffu, q[0], o);
  }
  if ((lane & ((32 >> LOG2G) - 1)) == 0) {
    const float mean = s[0] * (1.0f / (32 * CPG));
    dst[lane >> (5 - LOG2G)] = make_float2(mean, fmaxf(q[0] - s[0] * mean, 0.f));
  }
}

template <bool UPOK = true>
__device__ __forceinline__ void gn_stats32(const float (&v)[32], const GemmKParams& p, long grow, int gcol, int phase) {
  const int lane = threadIdx.x & 31;
  const int w0 = static_cast<int>(grow) - lane;          // first row of this warp's 32-row slab (all of one image)
  const int img = w0 / p.gn_hw;
  const int spi = p.gn_hw >> 5;                          // slabs per image (upsample mode: per image and phase)
  const int slab = ((w0 - img * p.gn_hw) >> 5) + (UPOK ? phase * spi : 0);    // any bijection onto the image's slab slots will do
  float2* dst = p.gn_partial + (static_cast<size_t>(img) * (UPOK && p.up ? 4 * spi : spi) + slab) * p.gn_groups + gcol / p.gn_cpg;
  if (p.gn_cpg == 4) gn_stats_chunk<4>(v, lane, dst);
  else if (p.gn_cpg == 8) gn_stats_chunk<8>(v, lane, dst);
  else gn_stats_chunk<16>(v, lane, dst);
}

// 64 bytes of one output row (32 fp16 columns) per thread: four 16-byte accesses, or two 32-byte ones (sm_100
// LDG/STG.256) — the thread-per-row epilogue touches one sector per thread per access, so halving the number of
// accesses halves the L1TEX LSU wavefronts (the busiest unit of the short-K 128-channel layers).
__device__ __forceinline__ void load_row64(const __half* ptr, uint4 (&r)[4], int wide) {
  if (wide) {
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[0].x), "=r"(r[0].y), "=r"(r[0].z), "=r"(r[0].w), "=r"(r[1].x), "=r"(r[1].y), "=r"(r[1].z), "=r"(r[1].w)
                 : "l"(ptr));
    asm volatile("ld.global.nc.v8.b32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
                 : "=r"(r[2].x), "=r"(r[2].y), "=r"(r[2].z), "=r"(r[2].w), "=r"(r[3].x), "=r"(r[3].y), "=r"(r[3].z), "=r"(r[3].w)
                 : "l"(ptr + 16));
  } else {
    const uint4* rp = reinterpret_cast<const uint4*>(ptr);
#pragma unroll
    for (int q = 0; q < 4; ++q) r[q] = __ldg(rp + q);
  }
}
__device__ __forceinline__ void store_row64(__half* ptr, const uint4 (&r)[4], int wide) {
  if (wide) {
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr), "r"(r[0].x), "r"(r[0].y), "r"(r[0].z),
                 "r"(r[0].w), "r"(r[1].x), "r"(r[1].y), "r"(r[1].z), "r"(r[1].w)
                 : "memory");
    asm volatile("st.global.v8.b32 [%0], {%1,%2,%3,%4,%5,%6,%7,%8};" ::"l"(ptr + 16), "r"(r[2].x), "r"(r[2].y), "r"(r[2].z),
                 "r"(r[2].w), "r"(r[3].x), "r"(r[3].y), "r"(r[3].z), "r"(r[3].w)
                 : "memory");
  } else {
#pragma unroll
    for (int q = 0; q < 4; ++q) reinterpret_cast<uint4*>(ptr)[q] = r[q];
  }
}

// epilogue_store<32> with the residual already in registers (prefetched before the accumulator was ready)
// `sbias`: optional shared-memory copy of the 32 bias values of this chunk (halo kernel: staged once per N tile; the
// broadcast LDG.128s otherwise queue behind the thread-per-row residual loads and output stores in L1TEX).
// everything of epilogue_store32_pre except the store: + bias, fp16 round, + residual, activation, statistics, pack
template <bool COLS = false, bool UPOK = true>
__device__ __forceinline__ void epilogue_pack32(float (&v)[32], const uint4 (&res)[4], bool has_res, const GemmKParams& p,
                                                long grow, int gcol, const float* sbias, int phase, uint4 (&packed)[4]) {
  if (p.bias) {
#pragma unroll
    for (int q = 0; q < 8; ++q) {
      const float4 b4 = sbias ? reinterpret_cast<const float4*>(sbias)[q] : __ldg(reinterpret_cast<const float4*>(p.bias + gcol) + q);
      v[4 * q] += b4.x; v[4 * q + 1] += b4.y; v[4 * q + 2] += b4.z; v[4 * q + 3] += b4.w;
    }
  }
  if (has_res) {
#pragma unroll
    for (int q = 0; q < 4; ++q) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&res[q]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        v[q * 8 + 2 * j] = round_h(v[q * 8 + 2 * j]) + f.x;
        v[q * 8 + 2 * j + 1] = round_h(v[q * 8 + 2 * j + 1]) + f.y;
      }
    }
  }
  if (p.act == IR_ACT_SILU) {
#pragma unroll
    for (int i = 0; i < 32; ++i) v[i] = silu(round_h(v[i]));
  }
  if (p.gn_partial) gn_stats32<UPOK>(v, p, grow, gcol, phase);
  if (COLS && p.col_partial && gcol >= p.col_begin) col_stats32(v, p, grow, gcol);   // only instantiated without a residual
#pragma unroll
  for (int q = 0; q < 4; ++q)
    packed[q] = make_uint4(pack_half2(v[q * 8], v[q * 8 + 1]), pack_half2(v[q * 8 + 2], v[q * 8 + 3]),
                           pack_half2(v[q * 8 + 4], v[q * 8 + 5]), pack_half2(v[q * 8 + 6], v[q * 8 + 7]));
}

template <bool COLS = false, bool UPOK = true>
__device__ __forceinline__ void epilogue_store32_pre(float (&v)[32], const uint4 (&res)[4], bool has_res, const GemmKParams& p,
                                                     long grow, int gcol, const float* sbias = nullptr, int phase = 0) {
  uint4 packed[4];
  epilogue_pack32<COLS, UPOK>(v, res, has_res, p, grow, gcol, sbias, phase, packed);
  __half* op = p.out + (UPOK ? out_row(p, grow, phase) : grow) * p.out_stride + gcol;
  store_row64(op, packed, p.wide_io);
}

template <int BN, int STAGES>
__global__ void __launch_bounds__(128) conv_gemm_kernel(const __grid_constant__ GemmKParams p) {
  constexpr int B_BYTES = BN * kBK * 2;
  constexpr int STAGE_BYTES = kABytes + B_BYTES;
  constexpr uint32_t TMEM_COLS = BN <= 32 ? 32 : BN <= 64 ? 64 : BN <= 128 ? 128 : 256;
  constexpr uint32_t IDESC = umma_idesc_f16(kBM, BN, 0, 0);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* accum_bar = empty_bar + STAGES;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(accum_bar + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a[0]);
    tma_prefetch_desc(&p.tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    mbar_init(accum_bar, 1);
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();                 // set-up above overlapped the previous kernel's tail; no global access before this point

  int mt = blockIdx.x, phase = 0;
  if (p.up) {                 // upsample mode: blockIdx.x runs over phase-major M tiles
    phase = mt / p.mtp;
    mt -= phase * p.mtp;
  }
  const int nt = blockIdx.y;
  const int num_k_total = p.taps * p.kc_per_tap;
  const int k_begin = blockIdx.z * p.kb_per_split;                    // split-K: this CTA's K-block range
  const int k_end = min(k_begin + p.kb_per_split, num_k_total);
  const int num_k = k_end - k_begin;

  if (warp == 0) {
    if (elect_one()) {   // elect.sync (not lane == 0): the compiler then keeps TMA / MMA operands in uniform registers
      const int w0 = (mt % p.tiles_w) * p.bw;
      const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
      const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
      int tap = k_begin / p.kc_per_tap + phase * 4, kc = k_begin % p.kc_per_tap;
      const int b_row = nt * BN + phase * p.N;
      int s = 0;
      uint32_t ph = 1;
      for (int ks = 0; ks < num_k; ++ks) {
        mbar_wait(&empty_bar[s], ph);
        mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
        tma_load_4d(sA + s * kABytes, &p.tma_a[p.tap_map[tap]], &full_bar[s], kc * kBK, w0 + p.tap_dx[tap],
                    h0 + p.tap_dy[tap], n0);
        tma_load_2d(sB + s * B_BYTES, &p.tma_b, &full_bar[s], (k_begin + ks) * kBK, b_row);
        if (++kc == p.kc_per_tap) {
          kc = 0;
          ++tap;
        }
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 0;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA)), b_lo0 = umma_desc_lo(smem_u32(sB));
      for (int ks = 0; ks < num_k; ++ks) {
        mbar_wait(&full_bar[s], ph);
        tc_fence_after();
        const uint32_t a_lo = a_lo0 + s * (kABytes >> 4), b_lo = b_lo0 + s * (B_BYTES >> 4);
#pragma unroll
        for (int k = 0; k < kBK / 16; ++k) umma_f16_ss_lo(tmem_base, a_lo + 2 * k, b_lo + 2 * k, IDESC, (ks | k) != 0 ? 1u : 0u);
        umma_commit(&empty_bar[s]);
        if (++s == STAGES) { s = 0; ph ^= 1; }
      }
      umma_commit(accum_bar);
    }
    __syncwarp();
  }

  // ------------------------------------------------------------------ epilogue (all 128 threads, thread == row)
  const int row = warp * 32 + lane;
  const long grow = static_cast<long>(mt) * kBM + row;
  const bool row_ok = grow < p.M;
  // Vector epilogue (16-byte bias loads, residual already in registers) whenever the rows allow it. The residual of
  // every chunk is fetched BEFORE waiting for the accumulator: the per-instruction stall samples of the single-identity
  // GEMMs (tools/ncu_source.sh) had 40 % of the kernel's samples on the bias / residual loads of a chunk-by-chunk
  // epilogue — four dependent HBM round trips after the MMAs were done.
  constexpr int kEpiChunks = (BN + 31) / 32;
  const bool fast = p.N % 32 == 0 && p.act != IR_ACT_GEGLU && (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
  const bool pre_res = fast && p.split == 1 && p.residual != nullptr && row_ok;
  uint4 resid[kEpiChunks][4];
  if (pre_res) {
#pragma unroll
    for (int ci = 0; ci < kEpiChunks; ++ci)
      if (ci * 32 < BN && nt * BN + ci * 32 < p.N) load_row64(p.residual + grow * p.res_stride + nt * BN + ci * 32, resid[ci], p.wide_io);
  }
  mbar_wait(accum_bar, 0);
  tc_fence_after();
  pdl_launch_dependents();    // only the epilogue is left: the next kernel may set itself up

  const uint32_t taddr = tmem_base + (static_cast<uint32_t>(warp * 32) << 16);

  if (p.act == IR_ACT_GEGLU) {
    // tile columns come in blocks of 128: [64 value | 64 gate]
    if constexpr (BN % 128 == 0) {
#pragma unroll 1
      for (int blk = 0; blk < BN / 128; ++blk) {
#pragma unroll 1
        for (int half = 0; half < 2; ++half) {
          uint32_t rv[32], rg[32];
          tmem_ld32(taddr + blk * 128 + half * 32, rv);
          tmem_ld32(taddr + blk * 128 + 64 + half * 32, rg);
          tmem_ld_wait();
          const int wcol = nt * BN + blk * 128 + half * 32;           // weight-row index of the value columns
          const int ocol = (nt * BN + blk * 128) / 2 + half * 32;     // output column
          if (row_ok) {
            uint32_t packed[16];
#pragma unroll
            for (int i = 0; i < 32; i += 2) {
              float v0 = __uint_as_float(rv[i]), v1 = __uint_as_float(rv[i + 1]);
              float g0 = __uint_as_float(rg[i]), g1 = __uint_as_float(rg[i + 1]);
              if (p.bias) {
                v0 += __ldg(p.bias + wcol + i);
                v1 += __ldg(p.bias + wcol + i + 1);
                g0 += __ldg(p.bias + wcol + 64 + i);
                g1 += __ldg(p.bias + wcol + 64 + i + 1);
              }
              v0 = round_h(v0); v1 = round_h(v1); g0 = round_h(g0); g1 = round_h(g1);
              packed[i / 2] = pack_half2(v0 * round_h(gelu_erf(g0)), v1 * round_h(gelu_erf(g1)));
            }
            uint4* dst = reinterpret_cast<uint4*>(p.out + grow * p.out_stride + ocol);
#pragma unroll
            for (int v = 0; v < 4; ++v) dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
          }
        }
      }
    }
  } else if (p.split > 1) {
    // ---- split-K: reduce-scatter the S partial tiles through distributed shared memory
    const int S = p.split;
    const int W = BN / S;                       // columns owned by each CTA of the cluster (multiple of 8)
    const uint32_t my_rank = cluster_ctarank();
    const bool w_pow2 = (W & (W - 1)) == 0;
    const int wshift = __ffs(W) - 1;
    // This CTA's slice of the residual row and of the bias, fetched now and consumed after the exchange (they were
    // eight dependent round trips at the very end of the kernel: one per 8-column group).
    constexpr int kPreV = 8;                    // W <= 64
    const int gcol0 = nt * BN + static_cast<int>(my_rank) * W;
    const bool pre = fast && row_ok && W <= 8 * kPreV && gcol0 + W <= p.N && (gcol0 & 7) == 0;
    uint4 rres[kPreV];
    float4 rb[2 * kPreV];
    if (pre) {
#pragma unroll
      for (int j = 0; j < kPreV; ++j) {
        if (j * 8 < W) {
          if (p.residual) rres[j] = __ldg(reinterpret_cast<const uint4*>(p.residual + grow * p.res_stride + gcol0 + j * 8));
          if (p.bias) {
            rb[2 * j] = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0 + j * 8));
            rb[2 * j + 1] = __ldg(reinterpret_cast<const float4*>(p.bias + gcol0 + j * 8 + 4));
          }
        }
      }
    }
    cluster_sync_all();                          // every CTA of the cluster has drained its pipeline buffers
    const uint32_t recv_local = smem_u32(smem);  // [src][W/4][128 rows] float4, reuses the stage buffers
#pragma unroll 1
    for (int c0 = 0; c0 < BN; c0 += 32) {
      uint32_t r[32];
      tmem_ld32(taddr + c0, r);
      tmem_ld_wait();
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        const int col = c0 + q * 4;
        const uint32_t owner = static_cast<uint32_t>(w_pow2 ? col >> wshift : col / W);
        const int c4 = (col - owner * W) >> 2;
        const uint32_t off = ((my_rank * (W >> 2) + c4) * 128 + row) * 16;
        dsmem_st_f4(dsmem_addr(recv_local + off, owner), __uint_as_float(r[q * 4]), __uint_as_float(r[q * 4 + 1]),
                    __uint_as_float(r[q * 4 + 2]), __uint_as_float(r[q * 4 + 3]));
      }
    }
    cluster_sync_all();                          // all partial slices have landed in their owner's shared memory
    const float4* recv = reinterpret_cast<const float4*>(smem);
    if (pre) {
#pragma unroll
      for (int j = 0; j < kPreV; ++j) {
        if (j * 8 < W) {
          const int c8 = j * 8;
          float v[8];
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = 0.f;
          for (int src = 0; src < S; ++src) {       // fixed order: deterministic sum
            const float4 a = recv[(src * (W >> 2) + (c8 >> 2)) * 128 + row];
            const float4 b = recv[(src * (W >> 2) + (c8 >> 2) + 1) * 128 + row];
            v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
            v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
          }
          // epilogue_store<8> with its operands already in registers: + bias, fp16 round, + residual, activation
          if (p.bias) {
            v[0] += rb[2 * j].x; v[1] += rb[2 * j].y; v[2] += rb[2 * j].z; v[3] += rb[2 * j].w;
            v[4] += rb[2 * j + 1].x; v[5] += rb[2 * j + 1].y; v[6] += rb[2 * j + 1].z; v[7] += rb[2 * j + 1].w;
          }
#pragma unroll
          for (int i = 0; i < 8; ++i) v[i] = round_h(v[i]);
          if (p.residual) {
            const __half2* h2 = reinterpret_cast<const __half2*>(&rres[j]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
              const float2 f = __half22float2(h2[i]);
              v[2 * i] += f.x;
              v[2 * i + 1] += f.y;
            }
          }
          if (p.act == IR_ACT_SILU) {
#pragma unroll
            for (int i = 0; i < 8; ++i) v[i] = silu(v[i]);
          }
          const int gcol = gcol0 + c8;
          *reinterpret_cast<uint4*>(p.out + out_row(p, grow, phase) * p.out_stride + gcol) =
              make_uint4(pack_half2(v[0], v[1]), pack_half2(v[2], v[3]), pack_half2(v[4], v[5]), pack_half2(v[6], v[7]));
          if (p.col_partial && gcol >= p.col_begin) col_stats8(v, p, grow, gcol);   // M % 128 == 0: row_ok is uniform
        }
      }
    } else if (row_ok) {
#pragma unroll 1
      for (int c8 = 0; c8 < W; c8 += 8) {
        float v[8];
#pragma unroll
        for (int i = 0; i < 8; ++i) v[i] = 0.f;
        for (int src = 0; src < S; ++src) {       // fixed order: deterministic sum
          const float4 a = recv[(src * (W >> 2) + (c8 >> 2)) * 128 + row];
          const float4 b = recv[(src * (W >> 2) + (c8 >> 2) + 1) * 128 + row];
          v[0] += a.x; v[1] += a.y; v[2] += a.z; v[3] += a.w;
          v[4] += b.x; v[5] += b.y; v[6] += b.z; v[7] += b.w;
        }
        const int gcol = nt * BN + static_cast<int>(my_rank) * W + c8;
        epilogue_store<8>(v, p, grow, gcol, phase);   // leaves the stored (bias added, fp16-rounded) values in v
        if (p.col_partial && gcol >= p.col_begin) col_stats8(v, p, grow, gcol);   // M % 128 == 0: row_ok is uniform
      }
    }
  } else {
#pragma unroll
    for (int ci = 0; ci < kEpiChunks; ++ci) {       // unrolled: the prefetched residual stays in registers
      const int c0 = ci * 32;
      if (c0 < BN) {
        uint32_t r[32];
        tmem_ld32(taddr + c0, r);
        tmem_ld_wait();
        if (row_ok) {
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          if (fast && nt * BN + c0 < p.N) {
            epilogue_store32_pre<true>(v, resid[ci], pre_res, p, grow, nt * BN + c0, nullptr, phase);
          } else {
            epilogue_store<32>(v, p, grow, nt * BN + c0, phase);  // leaves the stored (bias added, fp16-rounded) values in v
            if (p.col_partial && nt * BN + c0 >= p.col_begin && nt * BN + c0 < p.N) col_stats32(v, p, grow, nt * BN + c0);
          }
        }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}


// ------------------------------------------------------------------------------------------------ persistent kernel
constexpr int kPersistThreads = 320;   // warp 0: TMA, warp 1: MMA, warps 2-9: epilogue (two warps per TMEM lane quarter)

// GEGLU epilogue for 32 outputs: values at accumulator columns vcol.., gates 64 columns further
__device__ __forceinline__ void geglu_store32(uint32_t taddr_v, const GemmKParams& p, long grow, int wcol, int ocol, bool row_ok) {
  uint32_t rv[32], rg[32];
  tmem_ld32(taddr_v, rv);
  tmem_ld32(taddr_v + 64, rg);
  tmem_ld_wait();
  if (!row_ok) return;
  uint32_t packed[16];
#pragma unroll
  for (int i = 0; i < 32; i += 2) {
    float v0 = __uint_as_float(rv[i]), v1 = __uint_as_float(rv[i + 1]);
    float g0 = __uint_as_float(rg[i]), g1 = __uint_as_float(rg[i + 1]);
    if (p.bias) {
      v0 += __ldg(p.bias + wcol + i);
      v1 += __ldg(p.bias + wcol + i + 1);
      g0 += __ldg(p.bias + wcol + 64 + i);
      g1 += __ldg(p.bias + wcol + 64 + i + 1);
    }
    v0 = round_h(v0); v1 = round_h(v1); g0 = round_h(g0); g1 = round_h(g1);
    packed[i / 2] = pack_half2(v0 * round_h(gelu_erf_fast(g0)), v1 * round_h(gelu_erf_fast(g1)));
  }
  uint4* dst = reinterpret_cast<uint4*>(p.out + grow * p.out_stride + ocol);
#pragma unroll
  for (int v = 0; v < 4; ++v) dst[v] = make_uint4(packed[4 * v], packed[4 * v + 1], packed[4 * v + 2], packed[4 * v + 3]);
}

// MSUB = 2: a CTA tile is two stacked 128-row tiles (256 x BN) that share every weight box: the operand traffic per
// MMA cycle drops from 128 to 96 B/clk for BN = 128, the width of the 128-channel VAE layers.
// RESID: the launch carries a residual (and no GEGLU) — its values are prefetched into registers.
// TSTORE (BN = 128, MSUB = 1, no residual): the epilogue warps write their packed fp16 rows into a shared-memory tile in
// the 128-byte-swizzled box layout (16-byte pieces of 8 consecutive rows land in 8 different bank groups: conflict-free)
// and one thread per 64-column half issues ONE TMA store per tile (cp.async.bulk.tensor...global.shared::cta) instead of
// 128 threads x two 32-byte STG each to a different line: the thread-per-row stores are what keeps the L1TEX LSU pipe
// the busiest unit of the short-K layers. Two staging buffers: tile t+1 is packed while the store
// of tile t drains. Partial last M tiles are clipped by the tensor map.
template <int BN, int STAGES, int MSUB, bool RESID, bool TSTORE = false>
__global__ void __launch_bounds__(kPersistThreads, 1) conv_gemm_persistent_kernel(const __grid_constant__ GemmKParams p) {
  static_assert(!TSTORE || (BN == 128 && MSUB == 1 && !RESID), "TMA-store epilogue: 128-wide tiles without residual");
  constexpr int B_BYTES = BN * kBK * 2;
  constexpr int STAGE_BYTES = MSUB * kABytes + B_BYTES;
  constexpr int OUT_BYTES = TSTORE ? 2 * 2 * 16384 : 0;     // [buffer][column half] x (128 rows x 128 B)
  constexpr uint32_t ACC_COLS = BN <= 64 ? 64 : BN <= 128 ? 128 : 256;   // columns per accumulator
  constexpr uint32_t TMEM_COLS = 2 * MSUB * ACC_COLS;                    // two stages x MSUB sub-tiles
  static_assert(TMEM_COLS <= 512, "accumulators do not fit in tensor memory");
  constexpr uint32_t IDESC = umma_idesc_f16(kBM, BN, 0, 0);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;                                   // [STAGES][MSUB] x 16 KB
  uint8_t* sB = smem + STAGES * MSUB * kABytes;
  uint8_t* sOut = smem + STAGES * STAGE_BYTES;          // 1024-aligned (STAGE_BYTES is a multiple of 16 KB)
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES + OUT_BYTES);
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;    // accumulator stage ready for the epilogue
  uint64_t* tempty_bar = tfull_bar + 2;        // accumulator stage drained
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int n_tiles_n = (p.N + BN - 1) / BN;
  const int m_super = (p.m_tiles + MSUB - 1) / MSUB;
  const int total_tiles = m_super * n_tiles_n;
  const int num_k = p.taps * p.kc_per_tap;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a[0]);
    tma_prefetch_desc(&p.tma_b);
    if (TSTORE) tma_prefetch_desc(&p.tma_out);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 256);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc(tmem_slot, TMEM_COLS);
    tmem_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int s = 0;          // stage ring position and phase, carried across tiles
      uint32_t ph = 1;
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x) {
        const int ms = tile / n_tiles_n, nt = tile - ms * n_tiles_n;
        int phase = 0;                          // upsample mode (MSUB == 1 there): M tiles are phase-major
        if constexpr (MSUB == 1) { if (p.up) phase = ms / p.mtp; }
        const int b_row = nt * BN + phase * p.N;
        int w0[MSUB], h0[MSUB], n0[MSUB];
#pragma unroll
        for (int sub = 0; sub < MSUB; ++sub) {
          const int mt = ms * MSUB + sub - phase * p.mtp;   // may run past m_tiles: out-of-range boxes are zero fill
          w0[sub] = (mt % p.tiles_w) * p.bw;
          h0[sub] = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
          n0[sub] = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
        }
        int tap = phase * 4, kc = 0;
        for (int ks = 0; ks < num_k; ++ks) {
          mbar_wait(&empty_bar[s], ph);
          mbar_arrive_expect_tx(&full_bar[s], STAGE_BYTES);
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub)
            tma_load_4d(sA + (s * MSUB + sub) * kABytes, &p.tma_a[p.tap_map[tap]], &full_bar[s], kc * kBK,
                        w0[sub] + p.tap_dx[tap], h0[sub] + p.tap_dy[tap], n0[sub]);
          tma_load_2d(sB + s * B_BYTES, &p.tma_b, &full_bar[s], ks * kBK, b_row);
          if (++kc == p.kc_per_tap) {
            kc = 0;
            ++tap;
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      int s = 0, local = 0;
      uint32_t ph = 0;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA)), b_lo0 = umma_desc_lo(smem_u32(sB));
      for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
        const int as = local & 1;
        mbar_wait(&tempty_bar[as], ((local >> 1) & 1) ^ 1);       // the epilogue has drained this accumulator stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * MSUB * ACC_COLS;
        for (int ks = 0; ks < num_k; ++ks) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t b_lo = b_lo0 + s * (B_BYTES >> 4);
#pragma unroll
          for (int sub = 0; sub < MSUB; ++sub) {
            const uint32_t a_lo = a_lo0 + (s * MSUB + sub) * (kABytes >> 4);
#pragma unroll
            for (int k = 0; k < kBK / 16; ++k)
              umma_f16_ss_lo(d_tmem + sub * ACC_COLS, a_lo + 2 * k, b_lo + 2 * k, IDESC, (ks | k) != 0 ? 1u : 0u);
          }
          umma_commit(&empty_bar[s]);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit(&tfull_bar[as]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (thread == row, half the columns)
    const int quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const int half = (warp - 2) >> 2;             // which half of the tile's columns
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    int local = 0;
    for (int tile = blockIdx.x; tile < total_tiles; tile += gridDim.x, ++local) {
      const int ms = tile / n_tiles_n, nt = tile - ms * n_tiles_n;
      const int as = local & 1;
      if (tile + static_cast<int>(gridDim.x) >= total_tiles) pdl_launch_dependents();   // this CTA's last tile
      // prefetch this thread's residual values while the tile is still being accumulated: a short-K tile would
      // otherwise pay one HBM round trip per 32-column chunk after the accumulator is ready
      constexpr int kMaxChunks = (BN == 160 ? 3 : (BN / 2 + 31) / 32);
      uint4 resid[RESID ? MSUB : 1][RESID ? kMaxChunks : 1][4];
      if (RESID && p.N % 32 == 0) {
        const int cb = half == 0 ? 0 : (BN == 160 ? 96 : BN / 2);
        const int ce = half == 0 ? (BN == 160 ? 96 : BN / 2) : BN;
#pragma unroll
        for (int sub = 0; sub < MSUB; ++sub) {
          const long gr = (static_cast<long>(ms) * MSUB + sub) * kBM + row;
#pragma unroll
          for (int ci = 0; ci < kMaxChunks; ++ci) {
            const int gcol = nt * BN + cb + ci * 32;
            if (gr < p.M && cb + ci * 32 < ce && gcol < p.N) {
              load_row64(p.residual + gr * p.res_stride + gcol, resid[sub][ci], p.wide_io);
            }
          }
        }
      }
      mbar_wait(&tfull_bar[as], (local >> 1) & 1);
      tc_fence_after();
      if constexpr (TSTORE) {
        // (host guarantees: no GEGLU / residual / upsample mode, N % 128 == 0, 16-byte aligned bias)
        const long grow = static_cast<long>(ms) * kBM + row;
        const bool row_ok = grow < p.M;
        uint8_t* stg = sOut + ((local & 1) * 2 + half) * 16384;
        const uint32_t taddr = tmem_base + as * ACC_COLS + lane_addr;
        const uint4 no_res[4] = {};
#pragma unroll
        for (int ci = 0; ci < 2; ++ci) {
          const int c0 = half * 64 + ci * 32;
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          if (row_ok) {
            float v[32];
#pragma unroll
            for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
            uint4 packed[4];
            epilogue_pack32<true, false>(v, no_res, false, p, grow, nt * BN + c0, nullptr, 0, packed);
#pragma unroll
            for (int q = 0; q < 4; ++q)     // 16-byte piece j = ci*4+q of this row, at its swizzled slot j ^ (row & 7)
              *reinterpret_cast<uint4*>(stg + row * 128 + (((ci * 4 + q) ^ (row & 7)) << 4)) = packed[q];
          }
        }
        tc_fence_before();
        mbar_arrive(&tempty_bar[as]);               // the accumulator stage is drained: the next tile's MMAs may start
        fence_proxy_async_smem();                   // my generic-proxy writes are visible to the TMA engine
        // One warp of the half issues (elect.sync, not lane == 0: the tensor-map operands then stay in uniform registers;
        // the elected lane of a converged warp is the same every time, and bulk groups are per thread).
        // The previous tile's store (the OTHER buffer) has finished reading shared memory: after the barrier every
        // thread of this half may overwrite it (the buffer written above was released one barrier earlier).
        if (quarter == 0) {
          if (elect_one()) tma_store_wait_read<0>();
          __syncwarp();
        }
        named_bar_sync(1 + half, 128);
        if (quarter == 0) {
          if (elect_one()) {
            tma_store_2d(&p.tma_out, stg, nt * BN + half * 64, ms * kBM);
            tma_store_commit();
          }
          __syncwarp();
        }
        continue;
      }
      if (p.act == IR_ACT_GEGLU) {
        if constexpr (BN % 128 == 0) {
#pragma unroll 1
          for (int sub = 0; sub < MSUB; ++sub) {
            const long grow = (static_cast<long>(ms) * MSUB + sub) * kBM + row;
            const uint32_t taddr = tmem_base + (as * MSUB + sub) * ACC_COLS + lane_addr;
#pragma unroll 1
            for (int blk = 0; blk < BN / 128; ++blk) {
              const int wcol = nt * BN + blk * 128 + half * 32;            // weight-row index of the value columns
              const int ocol = (nt * BN + blk * 128) / 2 + half * 32;      // output column
              geglu_store32(taddr + blk * 128 + half * 32, p, grow, wcol, ocol, grow < p.M);
            }
          }
        }
      } else {
        // BN = 160 splits 96 | 64 so both halves work in 32-column chunks
        const int c_begin = half == 0 ? 0 : (BN == 160 ? 96 : BN / 2);
        const int c_end = half == 0 ? (BN == 160 ? 96 : BN / 2) : BN;
        const bool fast = p.N % 32 == 0 && (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
        int phase = 0;
        if constexpr (MSUB == 1) { if (p.up) phase = ms / p.mtp; }   // never MSUB == 2 in upsample mode
#pragma unroll
        for (int sub = 0; sub < MSUB; ++sub) {
          const long grow = (static_cast<long>(ms) * MSUB + sub - phase * p.mtp) * kBM + row;
          const bool row_ok = grow < p.M;
          const uint32_t taddr = tmem_base + (as * MSUB + sub) * ACC_COLS + lane_addr;
#pragma unroll
          for (int ci = 0; ci < kMaxChunks; ++ci) {      // fully unrolled: the prefetched residual stays in registers
            const int c0 = c_begin + ci * 32;
            if (c0 < c_end) {
              uint32_t r[32];
              tmem_ld32(taddr + c0, r);
              tmem_ld_wait();
              if (row_ok) {
                float v[32];
#pragma unroll
                for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
                if (fast && nt * BN + c0 < p.N) epilogue_store32_pre<!RESID, !RESID && MSUB == 1>(v, resid[RESID ? sub : 0][RESID ? ci : 0], RESID, p, grow, nt * BN + c0, nullptr, phase);
                else epilogue_store<32>(v, p, grow, nt * BN + c0, phase);
              }
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive(&tempty_bar[as]);
    }
    if (TSTORE && quarter == 0) {      // shared memory stays valid until the last store has read it
      if (elect_one()) tma_store_wait_read<0>();
      __syncwarp();
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) tmem_dealloc(tmem_base, TMEM_COLS);
}

// ------------------------------------------------------------------------------------------------ CTA-pair kernel
// Persistent kernel on tcgen05.mma.cta_group::2: a cluster of two CTAs (one TPC) computes 256 x 256 output tiles as ONE
// M = 256, N = 256 MMA. CTA r stages its own 128 rows of A and rows [r*128, r*128+128) of the weight tile (32 KB per
// stage instead of 48 KB), so each SM reads 64 B/clk of operands from shared memory instead of 96 and every weight
// box is fetched from L2 once per 256 output rows. Accumulator: 128 lanes x 256 fp32 columns in EACH CTA's tensor
// memory (rows r*128.. of the pair tile), double-buffered (512 columns). Roles per CTA as in the persistent kernel;
// only the leader's (cluster rank 0) warp 1 issues MMAs; its commits are multicast to both CTAs' barriers. Both
// CTAs' TMA loads signal the leader's full barrier (expect_tx covers both halves).
template <int BN, int STAGES, bool RESID>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(kPersistThreads, 1) conv_gemm_pair_kernel(const __grid_constant__ GemmKParams p) {
  static_assert(BN == 128 || BN == 160 || BN == 256, "pair tile N");
  constexpr int B_BYTES = (BN / 2) * kBK * 2;   // this CTA's half of the weight box
  constexpr int STAGE_BYTES = kABytes + B_BYTES;
  constexpr uint32_t ACC_COLS = BN <= 128 ? 128 : 256;
  constexpr uint32_t TMEM_COLS = 2 * ACC_COLS;
  constexpr uint32_t IDESC = umma_idesc_f16(256, BN, 0, 0);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + STAGES * kABytes;
  uint64_t* full_bar = reinterpret_cast<uint64_t*>(smem + STAGES * STAGE_BYTES);   // used in the leader only
  uint64_t* empty_bar = full_bar + STAGES;
  uint64_t* tfull_bar = empty_bar + STAGES;
  uint64_t* tempty_bar = tfull_bar + 2;                                            // used in the leader only
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(tempty_bar + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();
  const int n_tiles_n = p.N / BN;
  const int m_pairs = (p.m_tiles + 1) / 2;
  const int total_tiles = m_pairs * n_tiles_n;
  const int num_k = p.taps * p.kc_per_tap;
  const int tile0 = blockIdx.x >> 1, tile_step = gridDim.x >> 1;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a[0]);
    tma_prefetch_desc(&p.tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < STAGES; ++s) {
      mbar_init(&full_bar[s], 1);
      mbar_init(&empty_bar[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], 512);           // 256 epilogue threads of each CTA
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    tmem_alloc_pair(tmem_slot, TMEM_COLS);
    tmem_relinquish_pair();
  }
  tc_fence_before();
  cluster_sync_all();                           // the peer's barriers are initialised before anything signals them
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer (both CTAs)
    if (elect_one()) {
      int s = 0;
      uint32_t ph = 1;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int mp = tile / n_tiles_n, nt = tile - mp * n_tiles_n;
        int mt = mp * 2 + static_cast<int>(rank);            // may run past m_tiles: out-of-range boxes are zero fill
        int phase = 0;                                       // upsample mode: mtp is even, both CTAs share the phase
        if (p.up) {
          phase = mt / p.mtp;
          mt -= phase * p.mtp;
        }
        const int b_row = nt * BN + phase * p.N + static_cast<int>(rank) * (BN / 2);
        const int w0 = (mt % p.tiles_w) * p.bw;
        const int h0 = ((mt / p.tiles_w) % p.tiles_h) * p.bh;
        const int n0 = (mt / (p.tiles_w * p.tiles_h)) * p.bn;
        int tap = phase * 4, kc = 0;
        for (int ks = 0; ks < num_k; ++ks) {
          mbar_wait(&empty_bar[s], ph);
          if (rank == 0) mbar_arrive_expect_tx(&full_bar[s], 2 * STAGE_BYTES);
          const uint32_t leader_full = dsmem_addr(smem_u32(&full_bar[s]), 0);
          tma_load_4d_pair(sA + s * kABytes, &p.tma_a[p.tap_map[tap]], leader_full, kc * kBK, w0 + p.tap_dx[tap],
                           h0 + p.tap_dy[tap], n0);
          tma_load_2d_pair(sB + s * B_BYTES, &p.tma_b, leader_full, ks * kBK, b_row);
          if (++kc == p.kc_per_tap) {
            kc = 0;
            ++tap;
          }
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (leader CTA only)
    if (rank == 0 && elect_one()) {
      // one elected thread; ring position and phase are carried (no div/mod), descriptors advance by adding to the
      // low word: the issue loop must stay well under the 512 clk of tensor work per K block
      int s = 0, local = 0;
      uint32_t ph = 0;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA)), b_lo0 = umma_desc_lo(smem_u32(sB));
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++local) {
        const int as = local & 1;
        mbar_wait(&tempty_bar[as], ((local >> 1) & 1) ^ 1);       // both CTAs' epilogues have drained this stage
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * ACC_COLS;
        for (int ks = 0; ks < num_k; ++ks) {
          mbar_wait(&full_bar[s], ph);
          tc_fence_after();
          const uint32_t a_lo = a_lo0 + s * (kABytes >> 4), b_lo = b_lo0 + s * (B_BYTES >> 4);
#pragma unroll
          for (int k = 0; k < kBK / 16; ++k)
            umma_f16_ss_pair_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, IDESC, (ks | k) != 0 ? 1u : 0u);
          umma_commit_pair(&empty_bar[s], 3);
          if (++s == STAGES) { s = 0; ph ^= 1; }
        }
        umma_commit_pair(&tfull_bar[as], 3);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (thread == row, half the columns)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    // 160-wide tiles split 96 | 64 so both halves work in 32-column chunks
    const int c_begin = half == 0 ? 0 : (BN == 160 ? 96 : BN / 2);
    const int c_end = half == 0 ? (BN == 160 ? 96 : BN / 2) : BN;
    constexpr int kChunks = BN == 160 ? 3 : BN / 2 / 32;
    const bool fast = p.N % 32 == 0 && (p.bias == nullptr || (reinterpret_cast<uintptr_t>(p.bias) & 15) == 0);
    int local = 0;
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++local) {
      const int mp = tile / n_tiles_n, nt = tile - mp * n_tiles_n;
      const int as = local & 1;
      if (tile + tile_step >= total_tiles) pdl_launch_dependents();                     // this cluster's last tile
      int phase = 0;
      if (p.up) phase = (mp * 2) / p.mtp;
      const long grow = (static_cast<long>(mp) * 2 + rank - phase * p.mtp) * kBM + row;
      const bool row_ok = grow < p.M;
      uint4 resid[RESID ? kChunks : 1][4];
      if (RESID && row_ok) {
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          if (c_begin + ci * 32 < c_end) {
            load_row64(p.residual + grow * p.res_stride + nt * BN + c_begin + ci * 32, resid[ci], p.wide_io);
          }
        }
      }
      mbar_wait(&tfull_bar[as], (local >> 1) & 1);
      tc_fence_after();
      const uint32_t taddr = tmem_base + as * ACC_COLS + lane_addr;
      if (p.act == IR_ACT_GEGLU) {
        if constexpr (BN % 128 == 0) {
#pragma unroll 1
          for (int blk = 0; blk < BN / 128; ++blk) {
            const int wcol = nt * BN + blk * 128 + half * 32;
            const int ocol = (nt * BN + blk * 128) / 2 + half * 32;
            geglu_store32(taddr + blk * 128 + half * 32, p, grow, wcol, ocol, row_ok);
          }
        }
      } else {
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          const int c0 = c_begin + ci * 32;
          if (c0 < c_end) {
            uint32_t r[32];
            tmem_ld32(taddr + c0, r);
            tmem_ld_wait();
            if (row_ok) {
              float v[32];
#pragma unroll
              for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
              if (fast) epilogue_store32_pre<!RESID, !RESID>(v, resid[RESID ? ci : 0], RESID, p, grow, nt * BN + c0, nullptr, phase);
              else epilogue_store<32>(v, p, grow, nt * BN + c0, phase);
            }
          }
        }
      }
      tc_fence_before();
      mbar_arrive_cluster(dsmem_addr(smem_u32(&tempty_bar[as]), 0));
    }
  }

  tc_fence_before();
  cluster_sync_all();     // the leader's MMAs read the peer's shared memory and write its tensor memory until here
  if (warp == 2) tmem_dealloc_pair(tmem_base, TMEM_COLS);
}

template <int BN, int STAGES, bool RESID>
static int launch_pair_r(const GemmKParams& kp, cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + (BN / 2) * kBK * 2) + 1024 + 256;
  static PerDeviceOnce attr_once;   // function attributes are per device
  bool& attr_done = attr_once.slot();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_pair_kernel<BN, STAGES, RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaFuncSetAttribute(conv_gemm_pair): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  const long tiles = static_cast<long>((kp.m_tiles + 1) / 2) * (kp.N / BN);
  const int grid = 2 * static_cast<int>(tiles < 74 ? tiles : 74);
  IR_LAUNCH((conv_gemm_pair_kernel<BN, STAGES, RESID>), grid, kPersistThreads, smem, stream, kp);   // cluster dims are compiled in
  IR_CUDA_LAUNCH_CHECK("conv_gemm_pair launch");
  return 0;
}

static int launch_pair(const GemmKParams& kp, int bn_pair, cudaStream_t stream) {
  const bool resid = kp.residual != nullptr && kp.act != IR_ACT_GEGLU && kp.N % 32 == 0 &&
                     (kp.bias == nullptr || (reinterpret_cast<uintptr_t>(kp.bias) & 15) == 0);
  if (bn_pair == 128) return resid ? launch_pair_r<128, 8, true>(kp, stream) : launch_pair_r<128, 8, false>(kp, stream);
  if (bn_pair == 160) return resid ? launch_pair_r<160, 7, true>(kp, stream) : launch_pair_r<160, 7, false>(kp, stream);
  return resid ? launch_pair_r<256, 6, true>(kp, stream) : launch_pair_r<256, 6, false>(kp, stream);
}

// ------------------------------------------------------------------------------------------------ halo kernel
// 3x3 stride-1 convolutions on images >= 128 pixels wide. ncu on the kernels above shows the big VAE convolutions pinned
// at ~50 B/clk of L2->SM traffic per SM (l1tex__m_xbar2l1tex_read_bytes ~ 11.4 TB/s over the chip) with the tensor pipe
// 46 % (128-wide outputs) to 77 % (CTA pair) busy: an im2col-free conv still re-fetches every input pixel nine times,
// once per tap. Here the 64-channel slice of the input a tile needs is fetched ONCE as a (rows + 2) x 130-pixel halo
// box (TMA zero fill = padding) and the nine taps are nine shifted views of it: the A descriptor of tap (ky, kx) starts
// (ky * 130 + kx) * 128 bytes into the box — 128-byte swizzling is a function of the shared-memory address bits, so a
// view that starts on any 128-byte row reads back exactly what TMA wrote. Weights stream through their own ring
// (one 128 x 64 box per tap). Bytes into the SM per 64-channel slice and 256 output pixels: 66 KB + 144 KB instead of
// 288 KB + 144 KB (single CTA, 128-wide outputs, two image rows stacked per CTA), 2 x (50 KB + 144 KB) instead of
// 2 x (144 KB + 144 KB) for the CTA pair (256-wide N tiles, one image row per CTA).
// K order: channel slice, then tap (the kernels above run tap, then slice): same products, different fp32 summation
// order.
constexpr int kHaloW = 130;
// weight-ring depth: 16 KB slots on the single CTA (5) and the 256-wide pair (7), 8 KB slots on the 128-wide pair (10)
constexpr int halo_b_slots(int bn, bool pair) { return pair ? (bn == 256 ? 7 : 10) : 5; }

template <int BN, bool PAIR, bool RESID>
__global__ void __launch_bounds__(kPersistThreads, 1) conv3_halo_kernel(const __grid_constant__ GemmKParams p) {
  static_assert(BN == 128 || (BN == 256 && PAIR), "halo kernel: 128-wide single CTA, 128- or 256-wide CTA pair");
  constexpr int MSUB = PAIR ? 1 : 2;                       // output image rows per CTA
  constexpr int HROWS = MSUB + 2;
  constexpr uint32_t HALO_BYTES = HROWS * kHaloW * 128;    // one 64-channel halo box
  constexpr int HALO_SLOT = (HALO_BYTES + 1023) / 1024 * 1024;
  constexpr int NA = 2;
  constexpr int NB = halo_b_slots(BN, PAIR);
  constexpr int B_ROWS = PAIR ? BN / 2 : BN;               // weight rows of one tap in this CTA (pair: its half of the N tile)
  constexpr int B_BYTES = B_ROWS * kBK * 2;
  constexpr uint32_t STAGE_COLS = 256;                     // accumulator columns per stage (2 x 128, 1 x 256, or 128 of them)
  constexpr uint32_t TMEM_COLS = 512;
  constexpr uint32_t IDESC = PAIR ? umma_idesc_f16(256, BN, 0, 0) : umma_idesc_f16(128, 128, 0, 0);

  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sA = smem;
  uint8_t* sB = smem + NA * HALO_SLOT;
  uint64_t* a_full = reinterpret_cast<uint64_t*>(sB + NB * B_BYTES);
  uint64_t* a_empty = a_full + NA;
  uint64_t* b_full = a_empty + NA;
  uint64_t* b_empty = b_full + NB;
  uint64_t* tfull_bar = b_empty + NB;
  uint64_t* tempty_bar = tfull_bar + 2;
  // TIO (128-wide pair: the shared memory the 8 KB weight slots leave free): every epilogue warp owns a 32-row x 64-column
  // staging tile in the tensor map's 128-byte-swizzled box layout. The residual box is TMA-LOADED into it while the tile's
  // MMAs run, the thread (= pixel row) adds its accumulator in place, and ONE elected lane TMA-STORES the box: no
  // thread-per-row global access is left in the epilogue (ncu on the single-CTA kernel: L1TEX LSU pipe 68 %, the busiest
  // unit of the 128-channel layers, 32 sectors per warp-wide 256-bit access).
  constexpr bool TIO = PAIR && BN == 128;
  uint64_t* r_full = tempty_bar + 2;                          // TIO: 8 (per epilogue warp): residual box landed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(r_full + (TIO ? 8 : 0));
  float* sBias = reinterpret_cast<float*>(tmem_slot + 4);     // BN bias values of the current N tile (16-byte aligned)
  uint8_t* sStage = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sBias + BN) + 1023) & ~uintptr_t(1023));   // TIO: 8 x 4 KB

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const uint32_t rank = PAIR ? cluster_ctarank() : 0u;
  const int tiles_w = p.img_w >> 7;
  const int per_img = (p.img_h >> 1) * tiles_w;            // tiles of 128 pixels x 2 rows per image
  const int n_tiles_n = p.N / BN;
  const int total_tiles = p.bn * per_img * n_tiles_n;      // p.bn carries the batch size here
  const int chunks = p.kc_per_tap;
  const int c_in = chunks * kBK;
  const int tile0 = PAIR ? (blockIdx.x >> 1) : blockIdx.x, tile_step = PAIR ? (gridDim.x >> 1) : gridDim.x;

  if (warp == 0 && lane == 0) {
    tma_prefetch_desc(&p.tma_a[0]);
    tma_prefetch_desc(&p.tma_b);
  }
  if (warp == 1 && lane == 0) {
    for (int s = 0; s < NA; ++s) {
      mbar_init(&a_full[s], 1);
      mbar_init(&a_empty[s], 1);
    }
    for (int s = 0; s < NB; ++s) {
      mbar_init(&b_full[s], 1);
      mbar_init(&b_empty[s], 1);
    }
    for (int a = 0; a < 2; ++a) {
      mbar_init(&tfull_bar[a], 1);
      mbar_init(&tempty_bar[a], PAIR ? 512 : 256);
    }
    if (TIO) {
      for (int w = 0; w < 8; ++w) mbar_init(&r_full[w], 1);
      tma_prefetch_desc(&p.tma_out);
      if (RESID) tma_prefetch_desc(&p.tma_res);
    }
    fence_mbar_init();
  }
  if (warp == 2) {
    if (PAIR) {
      tmem_alloc_pair(tmem_slot, TMEM_COLS);
      tmem_relinquish_pair();
    } else {
      tmem_alloc(tmem_slot, TMEM_COLS);
      tmem_relinquish();
    }
  }
  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_wait();

  if (warp == 0) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {
      int sa = 0, sb = 0;
      uint32_t pha = 1, phb = 1;
      for (int tile = tile0; tile < total_tiles; tile += tile_step) {
        const int ms = tile / n_tiles_n, nt = tile - ms * n_tiles_n;
        const int img = ms / per_img, rem = ms - img * per_img;
        const int h0 = (rem / tiles_w) * 2 + static_cast<int>(rank), w0 = (rem % tiles_w) << 7;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&a_empty[sa], pha);
          if (PAIR) {
            if (rank == 0) mbar_arrive_expect_tx(&a_full[sa], 2 * HALO_BYTES);
            tma_load_4d_pair(sA + sa * HALO_SLOT, &p.tma_a[0], dsmem_addr(smem_u32(&a_full[sa]), 0), c * kBK, w0 - 1, h0 - 1, img);
          } else {
            mbar_arrive_expect_tx(&a_full[sa], HALO_BYTES);
            tma_load_4d(sA + sa * HALO_SLOT, &p.tma_a[0], &a_full[sa], c * kBK, w0 - 1, h0 - 1, img);
          }
          for (int tap = 0; tap < 9; ++tap) {
            mbar_wait(&b_empty[sb], phb);
            if (PAIR) {
              if (rank == 0) mbar_arrive_expect_tx(&b_full[sb], 2 * B_BYTES);
              tma_load_2d_pair(sB + sb * B_BYTES, &p.tma_b, dsmem_addr(smem_u32(&b_full[sb]), 0), tap * c_in + c * kBK,
                               nt * BN + static_cast<int>(rank) * B_ROWS);
            } else {
              mbar_arrive_expect_tx(&b_full[sb], B_BYTES);
              tma_load_2d(sB + sb * B_BYTES, &p.tma_b, &b_full[sb], tap * c_in + c * kBK, nt * BN);
            }
            if (++sb == NB) { sb = 0; phb ^= 1; }
          }
          if (++sa == NA) { sa = 0; pha ^= 1; }
        }
      }
    }
    __syncwarp();
  } else if (warp == 1) {
    // ------------------------------------------------------------------ MMA issuer (pair: leader CTA only)
    if (rank == 0 && elect_one()) {
      int sa = 0, sb = 0, local = 0;
      uint32_t pha = 0, phb = 0;
      const uint32_t a_lo0 = umma_desc_lo(smem_u32(sA)), b_lo0 = umma_desc_lo(smem_u32(sB));
      for (int tile = tile0; tile < total_tiles; tile += tile_step, ++local) {
        const int as = local & 1;
        mbar_wait(&tempty_bar[as], ((local >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + as * STAGE_COLS;
        for (int c = 0; c < chunks; ++c) {
          mbar_wait(&a_full[sa], pha);
          const uint32_t halo_lo = a_lo0 + sa * (HALO_SLOT >> 4);
#pragma unroll
          for (int tap = 0; tap < 9; ++tap) {      // unrolled: the tap offsets are immediates
            mbar_wait(&b_full[sb], phb);
            tc_fence_after();
            const uint32_t b_lo = b_lo0 + sb * (B_BYTES >> 4);
#pragma unroll
            for (int sub = 0; sub < MSUB; ++sub) {
              const uint32_t a_lo = halo_lo + ((sub + tap / 3) * kHaloW + tap % 3) * 8;     // 128-byte rows = 8 x 16 B
#pragma unroll
              for (int k = 0; k < kBK / 16; ++k) {
                const uint32_t acc = (tap | k) != 0 ? 1u : (c != 0 ? 1u : 0u);
                if (PAIR) umma_f16_ss_pair_lo(d_tmem, a_lo + 2 * k, b_lo + 2 * k, IDESC, acc);
                else umma_f16_ss_lo(d_tmem + sub * 128, a_lo + 2 * k, b_lo + 2 * k, IDESC, acc);
              }
            }
            if (PAIR) umma_commit_pair(&b_empty[sb], 3); else umma_commit(&b_empty[sb]);
            if (++sb == NB) { sb = 0; phb ^= 1; }
          }
          if (PAIR) umma_commit_pair(&a_empty[sa], 3); else umma_commit(&a_empty[sa]);
          if (++sa == NA) { sa = 0; pha ^= 1; }
        }
        if (PAIR) umma_commit_pair(&tfull_bar[as], 3); else umma_commit(&tfull_bar[as]);
      }
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ epilogue warps (thread == pixel, half the columns)
    const int quarter = warp & 3;
    const int half = (warp - 2) >> 2;
    const int row = quarter * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
    constexpr int kChunks = BN / 2 / 32;
    const int c_begin = half * (BN / 2);
    int local = 0;
    int bias_nt = -1;                                         // N tile whose bias is staged in shared memory
    for (int tile = tile0; tile < total_tiles; tile += tile_step, ++local) {
      const int ms = tile / n_tiles_n, nt = tile - ms * n_tiles_n;
      const int img = ms / per_img, rem = ms - img * per_img;
      const int h0 = (rem / tiles_w) * 2 + static_cast<int>(rank), w0 = (rem % tiles_w) << 7;
      const int as = local & 1;
      if (tile + tile_step >= total_tiles) pdl_launch_dependents();                     // last tile of this CTA / pair
      if (p.bias && nt != bias_nt) {                          // once per kernel for the 128- / 256-channel layers
        const int e = static_cast<int>(threadIdx.x) - 64;     // 0 .. 255 over the eight epilogue warps
        if (bias_nt >= 0) named_bar_sync(1, 256);             // every warp is done with the previous tile's values
        if (e < BN / 4) reinterpret_cast<float4*>(sBias)[e] = __ldg(reinterpret_cast<const float4*>(p.bias + nt * BN) + e);
        named_bar_sync(1, 256);
        bias_nt = nt;
      }
      const long grow0 = (static_cast<long>(img) * p.img_h + h0) * p.img_w + w0 + row;
      if constexpr (TIO) {
        uint8_t* stg = sStage + (warp - 2) * 4096;
        uint64_t* rbar = &r_full[warp - 2];
        const int orow0 = static_cast<int>(grow0) - lane, ocol0 = nt * BN + c_begin;
        // the previous tile's store has finished reading the staging tile; the residual box of this tile lands in it while
        // the MMAs of the tile run (this point is reached right after the previous tile's epilogue)
        if (elect_one()) {
          tma_store_wait_read<0>();
          if (RESID) {
            mbar_arrive_expect_tx(rbar, 4096);
            tma_load_2d(stg, &p.tma_res, rbar, ocol0, orow0);
          }
        }
        __syncwarp();
        mbar_wait(&tfull_bar[as], (local >> 1) & 1);
        tc_fence_after();
        if (RESID) mbar_wait(rbar, local & 1);
        const uint32_t taddr = tmem_base + as * STAGE_COLS + lane_addr;
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          const int c0 = c_begin + ci * 32;
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          uint4 res[4] = {};
          if (RESID) {
#pragma unroll
            for (int q = 0; q < 4; ++q)     // 16-byte piece j = ci*4+q of this row, at its swizzled slot j ^ (row & 7)
              res[q] = *reinterpret_cast<const uint4*>(stg + lane * 128 + (((ci * 4 + q) ^ (lane & 7)) << 4));
          }
          uint4 packed[4];
          epilogue_pack32<false, false>(v, res, RESID, p, grow0, nt * BN + c0, p.bias ? sBias + c0 : nullptr, 0, packed);
#pragma unroll
          for (int q = 0; q < 4; ++q)
            *reinterpret_cast<uint4*>(stg + lane * 128 + (((ci * 4 + q) ^ (lane & 7)) << 4)) = packed[q];
        }
        tc_fence_before();
        mbar_arrive_cluster(dsmem_addr(smem_u32(&tempty_bar[as]), 0));   // the accumulator stage is drained
        fence_proxy_async_smem();                   // my generic-proxy writes are visible to the TMA engine
        __syncwarp();
        if (elect_one()) {
          tma_store_2d(&p.tma_out, stg, ocol0, orow0);
          tma_store_commit();
        }
        __syncwarp();
        continue;
      }
      uint4 resid[RESID ? MSUB : 1][RESID ? kChunks : 1][4];
      if (RESID) {
#pragma unroll
        for (int sub = 0; sub < MSUB; ++sub)
#pragma unroll
          for (int ci = 0; ci < kChunks; ++ci) {
            load_row64(p.residual + (grow0 + static_cast<long>(sub) * p.img_w) * p.res_stride + nt * BN + c_begin + ci * 32,
                       resid[sub][ci], p.wide_io);
          }
      }
      mbar_wait(&tfull_bar[as], (local >> 1) & 1);
      tc_fence_after();
#pragma unroll
      for (int sub = 0; sub < MSUB; ++sub) {
        const long grow = grow0 + static_cast<long>(sub) * p.img_w;
        const uint32_t taddr = tmem_base + as * STAGE_COLS + (PAIR ? 0 : sub * 128) + lane_addr;
#pragma unroll
        for (int ci = 0; ci < kChunks; ++ci) {
          const int c0 = c_begin + ci * 32;
          uint32_t r[32];
          tmem_ld32(taddr + c0, r);
          tmem_ld_wait();
          float v[32];
#pragma unroll
          for (int i = 0; i < 32; ++i) v[i] = __uint_as_float(r[i]);
          epilogue_store32_pre<false, false>(v, resid[RESID ? sub : 0][RESID ? ci : 0], RESID, p, grow, nt * BN + c0, p.bias ? sBias + c0 : nullptr);
        }
      }
      tc_fence_before();
      if (PAIR) mbar_arrive_cluster(dsmem_addr(smem_u32(&tempty_bar[as]), 0));
      else mbar_arrive(&tempty_bar[as]);
    }
    if (TIO) {                         // shared memory stays valid until the last store has read it
      if (elect_one()) tma_store_wait_read<0>();
      __syncwarp();
    }
  }

  tc_fence_before();
  if (PAIR) cluster_sync_all(); else __syncthreads();
  if (warp == 2) {
    if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS);
    else tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

template <int BN, bool PAIR, bool RESID>
static int launch_halo_r(const GemmKParams& kp, cudaStream_t stream) {
  constexpr int msub = PAIR ? 1 : 2;
  constexpr int halo_slot = ((msub + 2) * kHaloW * 128 + 1023) / 1024 * 1024;
  constexpr int smem = 2 * halo_slot + halo_b_slots(BN, PAIR) * (PAIR ? BN / 2 : BN) * kBK * 2 + 1024 + 256 + 1024 +   // + the staged bias tile
                       (PAIR && BN == 128 ? 1024 + 8 * 4096 : 0);                                                      // + the epilogue warps' staging tiles
  static PerDeviceOnce attr_once;   // function attributes are per device
  bool& attr_done = attr_once.slot();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv3_halo_kernel<BN, PAIR, RESID>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaFuncSetAttribute(conv3_halo<%d>): %s", BN, cudaGetErrorString(e));
    attr_done = true;
  }
  const long tiles = static_cast<long>(kp.bn) * (kp.img_h / 2) * (kp.img_w / 128) * (kp.N / BN);
  const unsigned grid = PAIR ? 2 * static_cast<unsigned>(tiles < 74 ? tiles : 74) : static_cast<unsigned>(tiles < 148 ? tiles : 148);
  cudaError_t e = launch_kernel(conv3_halo_kernel<BN, PAIR, RESID>, dim3(grid), dim3(kPersistThreads), smem, stream,
                                dim3(PAIR ? 2 : 1, 1, 1), kp);
  if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "conv3_halo launch: %s", cudaGetErrorString(e));
  IR_CUDA_LAUNCH_CHECK("conv3_halo launch");
  return 0;
}

static int launch_halo(const GemmKParams& kp, int pair_bn, cudaStream_t stream) {
  const bool resid = kp.residual != nullptr;
  if (pair_bn == 256) return resid ? launch_halo_r<256, true, true>(kp, stream) : launch_halo_r<256, true, false>(kp, stream);
  if (pair_bn == 128) return resid ? launch_halo_r<128, true, true>(kp, stream) : launch_halo_r<128, true, false>(kp, stream);
  return resid ? launch_halo_r<128, false, true>(kp, stream) : launch_halo_r<128, false, false>(kp, stream);
}

template <int BN, int STAGES, int MSUB, bool RESID, bool TSTORE = false>
static int launch_persistent_r(const GemmKParams& kp, cudaStream_t stream) {
  constexpr int smem = STAGES * (MSUB * kABytes + BN * kBK * 2) + (TSTORE ? 65536 : 0) + 1024 + 256;
  static PerDeviceOnce attr_once;   // function attributes are per device
  bool& attr_done = attr_once.slot();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_persistent_kernel<BN, STAGES, MSUB, RESID, TSTORE>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaFuncSetAttribute(conv_gemm_persistent<%d,%d>): %s", BN, MSUB, cudaGetErrorString(e));
    attr_done = true;
  }
  const long tiles = static_cast<long>((kp.m_tiles + MSUB - 1) / MSUB) * ((kp.N + BN - 1) / BN);
  const int grid = static_cast<int>(tiles < 148 ? tiles : 148);
  IR_LAUNCH((conv_gemm_persistent_kernel<BN, STAGES, MSUB, RESID, TSTORE>), grid, kPersistThreads, smem, stream, kp);
  IR_CUDA_LAUNCH_CHECK("conv_gemm_persistent launch");
  return 0;
}

template <int BN, int STAGES, int MSUB>
static int launch_persistent(const GemmKParams& kp, cudaStream_t stream) {
  const bool resid = kp.residual != nullptr && kp.act != IR_ACT_GEGLU && kp.N % 32 == 0 &&
                     (kp.bias == nullptr || (reinterpret_cast<uintptr_t>(kp.bias) & 15) == 0);
  return resid ? launch_persistent_r<BN, STAGES, MSUB, true>(kp, stream) : launch_persistent_r<BN, STAGES, MSUB, false>(kp, stream);
}

template <int BN, int STAGES>
static int launch(const GemmKParams& kp, int m_tiles, cudaStream_t stream) {
  constexpr int smem = STAGES * (kABytes + BN * kBK * 2) + 1024 + 256;
  static PerDeviceOnce attr_once;   // function attributes are per device
  bool& attr_done = attr_once.slot();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(conv_gemm_kernel<BN, STAGES>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaFuncSetAttribute(conv_gemm<%d>): %s", BN, cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid(m_tiles, (kp.N + BN - 1) / BN, kp.split);
  cudaError_t e = launch_kernel(conv_gemm_kernel<BN, STAGES>, grid, dim3(128), smem, stream, dim3(1, 1, kp.split), kp);
  if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "conv_gemm launch (split %d): %s", kp.split, cudaGetErrorString(e));
  IR_CUDA_LAUNCH_CHECK("conv_gemm launch");
  return 0;
}

constexpr int kTmaStoreMaxK = 4;   // auto mode: TMA-store epilogue up to this many 64-wide K blocks (tools/misc_bench.py)

static int pow2_floor(int x) {
  int p = 1;
  while (p * 2 <= x) p *= 2;
  return p;
}

}  // namespace ir

static int conv_gemm_dispatch(const ir_conv_gemm_params* p, ir_stream_t stream_, bool* stats_fused);

extern "C" int ir_conv_gemm(const ir_conv_gemm_params* p, ir_stream_t stream_) {
  using namespace ir;
  bool stats_fused = false;
  const int rc = conv_gemm_dispatch(p, stream_, &stats_fused);
  if (rc != 0 || !p->gn_partial || stats_fused) return rc;
  // the kernel chosen for this shape (one tile per CTA / split-K) has no fused statistics: separate pass A
  const int hw = (p->h_in / p->stride) * (p->w_in / p->stride) * (p->upsample2x ? 4 : 1);
  return launch_gn_partial(p->out, p->out_row_stride, p->batch, hw, p->c_out, p->gn_groups, 32, p->gn_partial,
                           static_cast<cudaStream_t>(stream_));
}

static int conv_gemm_dispatch(const ir_conv_gemm_params* p, ir_stream_t stream_, bool* stats_fused) {
  using namespace ir;
  if (!p || !p->a || !p->w || !p->out) return set_error(IR_ERR_ARG, "ir_conv_gemm: NULL argument");
  if (int rc = check_arch()) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (p->c_in <= 0 || p->c_in % 64 != 0) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: c_in=%d must be a positive multiple of 64", p->c_in);
  if (!((p->ksize == 1 && p->stride == 1) || (p->ksize == 3 && (p->stride == 1 || p->stride == 2))))
    return set_error(IR_ERR_SHAPE, "ir_conv_gemm: unsupported ksize=%d stride=%d", p->ksize, p->stride);
  if (p->pad_hi_only && !(p->ksize == 3 && p->stride == 2))
    return set_error(IR_ERR_ARG, "ir_conv_gemm: pad_hi_only applies to 3x3 stride-2 convolutions");
  if (p->a_row_stride < p->c_in || p->a_row_stride % 8 != 0) return set_error(IR_ERR_ALIGN, "ir_conv_gemm: a_row_stride=%d", p->a_row_stride);
  if (p->batch <= 0 || p->h_in <= 0 || p->w_in <= 0 || p->c_out <= 0) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: non-positive dims");
  const bool geglu = p->act == IR_ACT_GEGLU;
  const int n_out = geglu ? p->c_out / 2 : p->c_out;
  if (p->out_row_stride < n_out) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: out_row_stride=%d < %d", p->out_row_stride, n_out);
  if (p->c_out % 8 == 0 && (p->out_row_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(p->out) & 15)))
    return set_error(IR_ERR_ALIGN, "ir_conv_gemm: out pointer/stride not 16-byte aligned");
  if (p->residual && p->c_out % 8 == 0 && (p->res_row_stride % 8 != 0 || (reinterpret_cast<uintptr_t>(p->residual) & 15)))
    return set_error(IR_ERR_ALIGN, "ir_conv_gemm: residual pointer/stride not 16-byte aligned");
  if (geglu && (p->c_out % 128 != 0 || p->residual)) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: GEGLU needs c_out %% 128 == 0 and no residual");
  const bool up = p->upsample2x != 0;
  if (up && (p->ksize != 3 || p->stride != 1 || p->residual || geglu || p->col_partial || p->halo == 2 || p->m_sub > 1))
    return set_error(IR_ERR_ARG, "ir_conv_gemm: upsample2x needs a 3x3 stride-1 convolution without residual / GEGLU / col_partial / forced halo");

  GemmKParams kp;
  memset(&kp, 0, sizeof(kp));
  const int taps = up ? 4 : p->ksize * p->ksize;   // upsample mode: four phases of 2x2 taps on the low-resolution input
  const int h_out = p->h_in / p->stride, w_out = p->w_in / p->stride;
  if (p->stride == 2 && ((p->h_in | p->w_in) & 1)) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: stride 2 needs even h, w");
  const long M = static_cast<long>(p->batch) * h_out * w_out;
  if (M > 0x7fffffffL) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: M too large");
  kp.M = static_cast<int>(M);
  kp.N = p->c_out;
  kp.n_out = n_out;
  kp.kc_per_tap = p->c_in / 64;
  kp.taps = taps;
  kp.bias = p->bias;
  kp.residual = static_cast<const __half*>(p->residual);
  kp.res_stride = p->res_row_stride;
  kp.out = static_cast<__half*>(p->out);
  kp.out_stride = p->out_row_stride;
  kp.act = p->act;
  // 256-bit epilogue accesses whenever every row the epilogue touches starts on a 32-byte boundary (bit-identical
  // results; +14 % on the 128-channel VAE layers, tools/wide_bench.py)
  if (p->wide_io != 1 && (reinterpret_cast<uintptr_t>(p->out) & 31) == 0 && p->out_row_stride % 16 == 0 &&
      (!p->residual || ((reinterpret_cast<uintptr_t>(p->residual) & 31) == 0 && p->res_row_stride % 16 == 0)))
    kp.wide_io = 1;
  // GroupNorm pass A on the outputs (optional): whole groups per 32-column chunk, whole 32-row slabs per image
  const bool fast_epilogue = p->c_out % 32 == 0 && (p->bias == nullptr || (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0);
  if (p->gn_partial) {
    const int hw_out = h_out * w_out;             // upsample mode: per phase (the low-resolution image)
    const int cpg = p->gn_groups > 0 && p->c_out % p->gn_groups == 0 ? p->c_out / p->gn_groups : 0;
    // upsample mode: the epilogue's 32-row slabs are low-resolution pixels of ONE phase; any slab that stays inside an image will do
    if (geglu || (cpg != 4 && cpg != 8 && cpg != 16) || (up ? hw_out % 32 != 0 : hw_out % 128 != 0) || (reinterpret_cast<uintptr_t>(p->gn_partial) & 7))
      return set_error(IR_ERR_SHAPE, "ir_conv_gemm: gn_partial needs c_out / gn_groups in {4, 8, 16}, h_out*w_out %% 128 == 0, no GEGLU (c_out=%d groups=%d hw=%d)",
                       p->c_out, p->gn_groups, hw_out);
    if (fast_epilogue) {
      kp.gn_partial = static_cast<float2*>(p->gn_partial);
      kp.gn_cpg = cpg;
      kp.gn_hw = hw_out;
      kp.gn_groups = p->gn_groups;
    }
  }

  // AdaIN column statistics of the outputs (optional): plain GEMM epilogue only, whole 32-row slabs, 32-column chunks
  if (p->col_partial) {
    if (geglu || p->residual || p->act != IR_ACT_NONE || kp.M % 128 != 0 || p->c_out % 32 != 0 || p->col_begin < 0 ||
        p->col_begin % 32 != 0 || p->col_begin >= p->c_out || (reinterpret_cast<uintptr_t>(p->col_partial) & 7) ||
        (p->bias && (reinterpret_cast<uintptr_t>(p->bias) & 15)))
      return set_error(IR_ERR_SHAPE, "ir_conv_gemm: col_partial needs no residual / activation, M %% 128 == 0, c_out %% 32 == 0, col_begin %% 32 == 0 (M=%d c_out=%d col_begin=%d)",
                       kp.M, p->c_out, p->col_begin);
    kp.col_partial = static_cast<float2*>(p->col_partial);
    kp.col_begin = p->col_begin;
  }

  const uint64_t rs = static_cast<uint64_t>(p->a_row_stride) * 2;  // pixel stride in bytes

  // Halo kernel: 3x3 stride-1 convolutions on images >= 128 pixels wide (every input pixel enters the SM once per
  // 64-channel slice instead of once per tap). 256-wide N tiles run on the CTA pair, 128-wide ones on a single CTA.
  if (p->halo < 0 || p->halo > 2) return set_error(IR_ERR_ARG, "ir_conv_gemm: halo=%d (0 = auto, 1 = off, 2 = force)", p->halo);
  {
    const bool eligible = !up && p->ksize == 3 && p->stride == 1 && p->w_in % 128 == 0 && p->h_in % 2 == 0 && !geglu &&
                          p->c_out % 128 == 0 && p->split_k <= 1 && p->tile_n == 0 && p->out_row_stride % 8 == 0 &&
                          (p->bias == nullptr || (reinterpret_cast<uintptr_t>(p->bias) & 15) == 0);
    if (p->halo == 2 && !eligible)
      return set_error(IR_ERR_SHAPE, "ir_conv_gemm: halo needs a 3x3 stride-1 conv, w %% 128 == 0, even h, c_out %% 128 == 0, no K split / tile_n");
    if (eligible && p->halo != 1) {
      // CTA pair: 256-wide N tiles, or 128-wide ones for the 128-channel layers (each CTA loads one image row and HALF of the
      // weight tile: 96 B/clk of shared-memory operand reads per SM instead of the 128 B/clk -- the port limit -- of the
      // single-CTA 128 x 128 MMA). cta_pair = 1 keeps the single-CTA kernel (A/B).
      // The 128-wide pair moves its outputs / residual with TMA: rows must be 16-byte multiples from a 16-byte-aligned base.
      const bool tio_ok = (reinterpret_cast<uintptr_t>(p->out) & 15) == 0 &&
                          (!p->residual || ((reinterpret_cast<uintptr_t>(p->residual) & 15) == 0 && p->res_row_stride % 8 == 0));
      const int pair_bn = p->cta_pair == 1 ? 0 : (p->c_out % 256 == 0 ? 256 : (tio_ok ? 128 : 0));
      const bool pair = pair_bn != 0;
      const long tiles = static_cast<long>(p->batch) * (p->h_in / 2) * (p->w_in / 128) * (p->c_out / (pair_bn == 256 ? 256 : 128));
      if (p->halo == 2 || (p->no_persistent != 1 && tiles >= (pair ? 74 : 148))) {
        kp.img_h = p->h_in;
        kp.img_w = p->w_in;
        kp.bn = p->batch;
        uint64_t dims[4] = {static_cast<uint64_t>(p->c_in), static_cast<uint64_t>(p->w_in), static_cast<uint64_t>(p->h_in),
                            static_cast<uint64_t>(p->batch)};
        uint64_t str[3] = {rs, rs * p->w_in, rs * p->w_in * p->h_in};
        uint32_t box[4] = {64, 130, pair ? 3u : 4u, 1};
        if (int rc = make_tmap_f16(&kp.tma_a[0], p->a, 4, dims, str, box)) return rc;
        uint64_t wdims[2] = {static_cast<uint64_t>(taps) * p->c_in, static_cast<uint64_t>(p->c_out)};
        uint64_t wstr[1] = {static_cast<uint64_t>(taps) * p->c_in * 2};
        uint32_t wbox[2] = {64, pair_bn == 128 ? 64u : 128u};
        if (int rc = make_tmap_f16(&kp.tma_b, p->w, 2, wdims, wstr, wbox)) return rc;
        if (pair_bn == 128) {
          uint64_t odims[2] = {static_cast<uint64_t>(p->c_out), static_cast<uint64_t>(kp.M)};
          uint64_t ostr[1] = {static_cast<uint64_t>(p->out_row_stride) * 2};
          uint32_t obox[2] = {64, 32};
          if (int rc = make_tmap_f16(&kp.tma_out, p->out, 2, odims, ostr, obox)) return rc;
          if (p->residual) {
            uint64_t rstr[1] = {static_cast<uint64_t>(kp.res_stride) * 2};
            if (int rc = make_tmap_f16(&kp.tma_res, p->residual, 2, odims, rstr, obox)) return rc;
          }
        }
        *stats_fused = kp.gn_partial != nullptr;
        return launch_halo(kp, pair_bn, stream);
      }
    }
  }
  int m_tiles;
  if (p->ksize == 1) {
    // flattened token-major GEMM: dims (c_in, M, 1, 1)
    kp.bw = 128; kp.bh = 1; kp.bn = 1;
    kp.tiles_w = (kp.M + 127) / 128; kp.tiles_h = 1;
    m_tiles = kp.tiles_w;
    uint64_t dims[4] = {static_cast<uint64_t>(p->c_in), static_cast<uint64_t>(kp.M), 1, 1};
    uint64_t str[3] = {rs, rs * kp.M, rs * kp.M};
    uint32_t box[4] = {64, 128, 1, 1};
    if (int rc = make_tmap_f16(&kp.tma_a[0], p->a, 4, dims, str, box)) return rc;
  } else {
    if ((w_out & (w_out - 1)) || (h_out & (h_out - 1)))
      return set_error(IR_ERR_SHAPE, "ir_conv_gemm: 3x3 conv needs power-of-two output h,w (got %dx%d)", h_out, w_out);
    kp.bw = w_out < 128 ? w_out : 128;
    kp.bh = (128 / kp.bw) < h_out ? (128 / kp.bw) : h_out;
    kp.bn = 128 / (kp.bw * kp.bh);
    kp.tiles_w = w_out / kp.bw;
    kp.tiles_h = h_out / kp.bh;
    m_tiles = kp.tiles_w * kp.tiles_h * ((p->batch + kp.bn - 1) / kp.bn);
    uint32_t box[4] = {64, static_cast<uint32_t>(kp.bw), static_cast<uint32_t>(kp.bh), static_cast<uint32_t>(kp.bn)};
    if (p->stride == 1) {
      uint64_t dims[4] = {static_cast<uint64_t>(p->c_in), static_cast<uint64_t>(p->w_in), static_cast<uint64_t>(p->h_in),
                          static_cast<uint64_t>(p->batch)};
      uint64_t str[3] = {rs, rs * p->w_in, rs * p->w_in * p->h_in};
      if (int rc = make_tmap_f16(&kp.tma_a[0], p->a, 4, dims, str, box)) return rc;
      if (up) {
        // output pixel (2y + py, 2x + px) of conv3x3(nearest2x(in)) reads input rows {y - 1, y, y} (py = 0) or {y, y, y + 1}
        // (py = 1) for ky = 0, 1, 2: two distinct rows py - 1 + a, a = 0, 1, whose taps were summed on the host (likewise
        // in x). The upsampled image's zero border coincides with the low-resolution one: TMA zero fill is the padding.
        for (int ph = 0; ph < 4; ++ph)
          for (int a = 0; a < 2; ++a)
            for (int b = 0; b < 2; ++b) {
              kp.tap_map[ph * 4 + a * 2 + b] = 0;
              kp.tap_dy[ph * 4 + a * 2 + b] = static_cast<int8_t>((ph >> 1) - 1 + a);
              kp.tap_dx[ph * 4 + a * 2 + b] = static_cast<int8_t>((ph & 1) - 1 + b);
            }
        kp.up = 1;
        kp.mtp = m_tiles;
        kp.up_wmask = p->w_in - 1;
        kp.up_w2 = 2 * p->w_in;
        m_tiles *= 4;
      } else
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          kp.tap_map[ky * 3 + kx] = 0;
          kp.tap_dx[ky * 3 + kx] = static_cast<int8_t>(kx - 1);
          kp.tap_dy[ky * 3 + kx] = static_cast<int8_t>(ky - 1);
        }
    } else {
      // pad_lo = 1 (symmetric padding 1): input row 2i + ky - 1: ky=0 -> odd phase, offset -1; ky=1 -> even phase;
      // ky=2 -> odd phase, offset 0.  pad_lo = 0 (diffusers Downsample2D(padding=0): F.pad (0,1,0,1) then a
      // valid conv): input row 2i + ky: ky=0 -> even phase; ky=1 -> odd phase; ky=2 -> even phase, offset +1
      // (the row past the bottom/right edge is TMA zero fill).
      uint64_t dims[4] = {static_cast<uint64_t>(p->c_in), static_cast<uint64_t>(w_out), static_cast<uint64_t>(h_out),
                          static_cast<uint64_t>(p->batch)};
      uint64_t str[3] = {rs * 2, rs * p->w_in * 2, rs * p->w_in * p->h_in};
      for (int pr = 0; pr < 2; ++pr)
        for (int pc = 0; pc < 2; ++pc) {
          const char* base = static_cast<const char*>(p->a) + (static_cast<uint64_t>(pr) * p->w_in + pc) * rs;
          if (int rc = make_tmap_f16(&kp.tma_a[pr * 2 + pc], base, 4, dims, str, box)) return rc;
        }
      for (int ky = 0; ky < 3; ++ky)
        for (int kx = 0; kx < 3; ++kx) {
          if (p->pad_hi_only) {
            const int pr = (ky == 1) ? 1 : 0, pc = (kx == 1) ? 1 : 0;
            kp.tap_map[ky * 3 + kx] = static_cast<int8_t>(pr * 2 + pc);
            kp.tap_dx[ky * 3 + kx] = static_cast<int8_t>(kx == 2 ? 1 : 0);
            kp.tap_dy[ky * 3 + kx] = static_cast<int8_t>(ky == 2 ? 1 : 0);
          } else {
            const int pr = (ky == 1) ? 0 : 1, pc = (kx == 1) ? 0 : 1;
            kp.tap_map[ky * 3 + kx] = static_cast<int8_t>(pr * 2 + pc);
            kp.tap_dx[ky * 3 + kx] = static_cast<int8_t>(kx == 0 ? -1 : 0);
            kp.tap_dy[ky * 3 + kx] = static_cast<int8_t>(ky == 0 ? -1 : 0);
          }
        }
    }
  }

  // tile-N selection: exact divisors first, fewest wasted columns
  int bn_tile = p->tile_n;
  bool wide_persistent = false;   // 128 x 256 tiles on the persistent kernel: 96 B/clk of operand reads instead of 128
  // (from 111 tiles = 3/4 of the SMs: m4096_k4608_n512 is 64 such tiles, 29.4 us; as 128-wide tiles with a 2-way K split 27.0)
  if (bn_tile == 0 && !geglu && p->c_out % 256 == 0 && p->split_k <= 1 && p->no_persistent != 1 &&
      static_cast<long>(m_tiles) * (p->c_out / 256) >= 111) {
    bn_tile = 256;
    wide_persistent = true;
  }
  if (bn_tile == 0) {
    if (geglu) bn_tile = (p->c_out % 256 == 0 && static_cast<long>(m_tiles) * (p->c_out / 256) >= 296) ? 256 : 128;
    else if (p->c_out <= 64) bn_tile = 64;
    // 160-wide tiles (115 B/clk of operand reads, fewer A re-reads) when they fill the machine without a K split
    else if (p->c_out % 160 == 0 && (p->c_out % 128 != 0 || static_cast<long>(m_tiles) * (p->c_out / 160) >= 148)) bn_tile = 160;
    else if (p->c_out % 128 == 0) bn_tile = 128;
    else if (p->c_out % 160 == 0) bn_tile = 160;
    else if (p->c_out % 64 == 0) bn_tile = 64;
    else bn_tile = 128;
  }
  if (geglu && bn_tile % 128 != 0) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: GEGLU needs tile_n 128 or 256");
  (void)pow2_floor;

  const int num_k = taps * kp.kc_per_tap;
  // CTA-pair kernel, automatic choice (made before the K split: a launch that pairs is never split). tools/pair_sweep.py
  // times every (op, shape) of a step with the pair forced at 256 / 160 / 128-wide tiles against the previous choice
  // (profiles/r02aj_pair_sweep_b{1,8}.txt). What it showed: the one-tile kernel's 128 x 128 tiles with a 2-way K split are
  // bound by L2 -> SM operand traffic on the mid-size long-K layers (m4096_k4608_n512: 256 CTAs x 36 K-blocks x 32 KB =
  // 295 MB per launch at the 11 TB/s the crossbar delivers); 256 x 128 pair tiles move 25 % less (each CTA loads HALF of
  // the weight box) and need no exchange: 27.5 -> 19.2 us (18 launches per step), m1024_k5120_n1280 26.1 -> 19.2,
  // m1024_k11520_n1280 41.9 -> 34.8. 160-wide pair tiles win on the 640 / 1280-channel linears from 64 pair tiles
  // (m4096_k2560_n640 20.8 -> 15.8, m4096_k640_n640 11.2 -> 9.7; at 8 identities m2048_k11520_n1280 66.1 -> 43.6,
  // m2048_k5120_n1280 34.0 -> 23.9) and are ahead of the 256-wide ones whenever both apply below 296 tiles.
  int bn_pair_auto = 0;
  // upsample mode: both CTAs of a pair must share the phase (even tiles per phase); with GroupNorm statistics in the
  // epilogue the short-K folded convolutions (K = 4 c_in) are epilogue-bound and the single-CTA 128 x 256 persistent tiles
  // are ahead of the pair's coupled epilogues (tools/up_bench.py: 512 -> 512 @128^2 -> 256^2 134.6 vs 149.7 us, 256 -> 256
  // @256^2 -> 512^2 157.6 vs 164.0)
  const bool up_no_pair = up && ((kp.mtp & 1) || (kp.gn_partial && p->cta_pair != 2));
  if (p->cta_pair == 0 && p->tile_n == 0 && p->split_k == 0 && p->no_persistent != 1 && !geglu && m_tiles >= 2 && !up_no_pair) {
    const long m_pairs = (m_tiles + 1) / 2;
    const bool conv3 = p->ksize == 3;
    const long pt256 = p->c_out % 256 == 0 ? m_pairs * (p->c_out / 256) : 0;
    const long pt160 = p->c_out % 160 == 0 ? m_pairs * (p->c_out / 160) : 0;
    const long pt128 = p->c_out % 128 == 0 ? m_pairs * (p->c_out / 128) : 0;
    const bool ok256 = pt256 && ((conv3 && num_k >= 18 && pt256 >= 64) || (num_k >= 16 && pt256 >= 296));
    const bool ok160 = pt160 && ((conv3 && num_k >= 18 && pt160 >= 64) || (num_k >= 16 && pt160 >= 296) ||
                                 (!conv3 && pt160 >= 64 && pt160 <= 296 && (num_k >= 20 || (num_k >= 10 && pt160 <= 74))));
    const bool ok128 = pt128 && num_k >= 40 && pt128 >= 40 && pt128 <= 148;
    if (ok256 && !(ok160 && pt256 < 296)) bn_pair_auto = 256;
    else if (ok160) bn_pair_auto = 160;
    else if (ok128) bn_pair_auto = 128;
  }

  // split-K over a cluster when the output tiles alone cannot fill the 148 SMs (see the header comment)
  int split = 1;
  if (p->split_k < 0 || p->split_k > 8 || (p->split_k & (p->split_k - 1)))
    return set_error(IR_ERR_ARG, "ir_conv_gemm: split_k=%d (0 = auto, 1, 2, 4 or 8)", p->split_k);
  const bool can_split = !geglu && p->c_out % 64 == 0 && p->c_out % 8 == 0;
  if (p->split_k > 1) {
    if (!can_split) return set_error(IR_ERR_SHAPE, "ir_conv_gemm: split_k needs c_out %% 64 == 0 and no GEGLU");
    split = p->split_k;
  }
  if (can_split && !wide_persistent && !bn_pair_auto && (p->split_k == 0 || p->split_k > 1) && p->tile_n == 0) {
    // narrower N tile first when even 8-way split of 128-wide tiles leaves most SMs idle
    if (static_cast<long>(m_tiles) * ((p->c_out + bn_tile - 1) / bn_tile) * 8 < 148 && bn_tile > 64) bn_tile = 64;
  }
  if (can_split && !wide_persistent && !bn_pair_auto && p->split_k == 0 && p->c_out % bn_tile == 0 && (bn_tile == 64 || bn_tile == 128)) {
    // K split only while every CTA of the cluster keeps >= 16 K-blocks: below that the cluster launch, the two cluster
    // barriers and the DSMEM exchange cost more than the idle SMs (tools/small_gemm_bench.py, device time per launch:
    // m1024_k1280_n1280 15.5 us with a 2-way split vs 9.3 without, m1024_k640_n640 12.0 vs 7.3, m256_k1280_n3840 14.5 vs 8.5;
    // the long-K 3x3 convolutions keep their split: m256_k11520_n1280 47.8 us unsplit vs 19.9 with 4 ways)
    const long tiles = static_cast<long>(m_tiles) * (p->c_out / bn_tile);
    while (split < 8 && tiles * split < 148 && tiles * split * 2 <= 296 && num_k / (split * 2) >= 16) split *= 2;
  }
  // unsplit launches that fill less than half of the SMs: 64-wide tiles double the CTA count (m1024_k640_n640 7.3 -> 5.9 us,
  // m256_k1280_n1280 8.6 -> 7.0, m256_k1280_n3840 8.5 -> 7.4, m4096_k320_n320 7.5 -> 6.6)
  if (split == 1 && !bn_pair_auto && p->split_k == 0 && p->tile_n == 0 && !geglu && !wide_persistent && bn_tile > 64 && p->c_out % 64 == 0 &&
      static_cast<long>(m_tiles) * ((p->c_out + bn_tile - 1) / bn_tile) < 74)
    bn_tile = 64;
  if (split > 1 && (p->c_out % bn_tile != 0 || (bn_tile / split) % 8 != 0 || num_k < split))
    return set_error(IR_ERR_SHAPE, "ir_conv_gemm: split_k=%d incompatible with tile_n=%d, c_out=%d, k-blocks=%d", split, bn_tile, p->c_out, num_k);
  kp.split = split;
  kp.kb_per_split = (num_k + split - 1) / split;
  if (split > 1 && kp.kb_per_split * (split - 1) >= num_k) {   // every CTA of the cluster must own >= 1 K-block
    kp.split = split = 1;
    kp.kb_per_split = num_k;
  }

  // CTA-pair kernel (tcgen05 cta_group::2, 256 x 256 or 256 x 128 tiles over the two SMs of a TPC). Measured
  // (tools/pair_bench.py): ahead of the single-CTA kernels once there are >= 4 tiles per cluster and K >= 1024; behind
  // them on short-K / GEGLU launches (the accumulator hand-over couples the two epilogues), on few-tile launches and
  // for 128-wide outputs with plenty of tiles (256 x 128 pair tiles: 798 vs 905 TFLOP/s of the stacked-M single-CTA
  // kernel). The automatic choice (bn_pair_auto) is made above, before the K split.
  int bn_pair = 0;
  if (p->cta_pair < 0 || p->cta_pair > 2) return set_error(IR_ERR_ARG, "ir_conv_gemm: cta_pair=%d (0 = auto, 1 = off, 2 = force)", p->cta_pair);
  if (p->cta_pair == 2 && split == 1 && m_tiles >= 2 && !up_no_pair) {       // forced: the widest tile the arguments allow
    if (p->c_out % 256 == 0 && (p->tile_n == 0 || p->tile_n == 256)) bn_pair = 256;
    else if (p->c_out % 160 == 0 && !geglu && (p->tile_n == 0 || p->tile_n == 160)) bn_pair = 160;
    else if (p->c_out % 128 == 0 && (p->tile_n == 0 || p->tile_n == 128)) bn_pair = 128;
  } else if (split == 1) {
    bn_pair = bn_pair_auto;
  }
  if (p->cta_pair == 2 && !bn_pair)
    return set_error(IR_ERR_SHAPE, "ir_conv_gemm: cta_pair needs c_out %% 128 == 0 or %% 160 == 0, no K split, >= 2 M tiles (c_out=%d split=%d m_tiles=%d)", p->c_out, split, m_tiles);
  const bool use_pair = bn_pair != 0;

  {
    uint64_t dims[2] = {static_cast<uint64_t>(taps) * p->c_in, static_cast<uint64_t>(p->c_out) * (up ? 4 : 1)};
    uint64_t str[1] = {static_cast<uint64_t>(taps) * p->c_in * 2};
    uint32_t box[2] = {64, static_cast<uint32_t>(use_pair ? bn_pair / 2 : bn_tile)};
    if (int rc = make_tmap_f16(&kp.tma_b, p->w, 2, dims, str, box)) return rc;
  }

  kp.m_tiles = m_tiles;
  if (use_pair) {
    *stats_fused = kp.gn_partial != nullptr;
    return launch_pair(kp, bn_pair, stream);
  }
  // Persistent kernel when the epilogue / per-CTA set-up is a visible fraction of a tile (short K, GEGLU) and the
  // tile count quantises well over 148 SMs; long-K layers keep two one-tile CTAs per SM (measured: tools/gemm_bench.py).
  bool persistent = false;
  if (split == 1 && (p->no_persistent == 2 || wide_persistent)) persistent = true;
  else if (split == 1 && !p->no_persistent) {
    const long tiles = static_cast<long>(m_tiles) * ((p->c_out + bn_tile - 1) / bn_tile);
    const long rounds = (tiles + 147) / 148;
    const double eff = tiles < 148 ? 1.0 : static_cast<double>(tiles) / (rounds * 148.0);
    persistent = geglu || (num_k <= 40 && (tiles >= 592 || eff >= 0.85));
  }
  if (persistent) {
    // 256 x 128 CTA tiles (two stacked M tiles sharing the weight boxes) when there are plenty of M tiles
    const bool tall = bn_tile == 128 && !geglu && m_tiles >= 2 * 148 && num_k >= 16 && p->m_sub != 1 && !up;
    *stats_fused = kp.gn_partial != nullptr;
    // TMA-store epilogue: 128-wide full tiles without residual / GEGLU. Auto mode: the short-K launches whose epilogue
    // is the bottleneck (the variant keeps 4 operand stages instead of 6 to make room for the staging tiles)
    if (p->tma_store < 0 || p->tma_store > 2) return set_error(IR_ERR_ARG, "ir_conv_gemm: tma_store=%d (0 = auto, 1 = off, 2 = force)", p->tma_store);
    const bool ts_ok = bn_tile == 128 && !tall && !geglu && !p->residual && !up && p->c_out % 128 == 0 && fast_epilogue &&
                       p->out_row_stride % 8 == 0;
    if (p->tma_store == 2 && !ts_ok)
      return set_error(IR_ERR_SHAPE, "ir_conv_gemm: tma_store needs the persistent kernel with 128-wide full tiles, no residual / GEGLU / upsample2x");
    if (ts_ok && (p->tma_store == 2 || (p->tma_store == 0 && num_k <= kTmaStoreMaxK))) {
      uint64_t odims[2] = {static_cast<uint64_t>(p->c_out), static_cast<uint64_t>(kp.M)};
      uint64_t ostr[1] = {static_cast<uint64_t>(p->out_row_stride) * 2};
      uint32_t obox[2] = {64, 128};
      if (int rc = make_tmap_f16(&kp.tma_out, p->out, 2, odims, ostr, obox)) return rc;
      return launch_persistent_r<128, 4, 1, false, true>(kp, stream);
    }
    switch (bn_tile) {
      case 64: return launch_persistent<64, 8, 1>(kp, stream);
      case 128: return tall ? launch_persistent<128, 4, 2>(kp, stream) : launch_persistent<128, 6, 1>(kp, stream);
      case 160: return launch_persistent<160, 5, 1>(kp, stream);
      case 256: return launch_persistent<256, 4, 1>(kp, stream);
      default: return set_error(IR_ERR_SHAPE, "ir_conv_gemm: tile_n=%d unsupported", bn_tile);
    }
  }
  kp.gn_partial = nullptr;   // one tile per CTA / split-K: statistics come from the separate pass (ir_conv_gemm)
  switch (bn_tile) {
    case 64: return launch<64, 4>(kp, m_tiles, stream);
    case 128: return launch<128, 3>(kp, m_tiles, stream);
    case 160: return launch<160, 3>(kp, m_tiles, stream);
    case 256: return launch<256, 4>(kp, m_tiles, stream);
    default: return set_error(IR_ERR_SHAPE, "ir_conv_gemm: tile_n=%d unsupported", bn_tile);
  }
}
