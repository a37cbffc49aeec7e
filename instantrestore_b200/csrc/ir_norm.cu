// GroupNorm(+SiLU) and LayerNorm on channel-last fp16 with fp32 statistics (HBM-bound; sized for L2-resident
// second passes). Algorithmic bytes: one fp16 read + one fp16 write of the activation.
#include "ir_host.h"
#include "ir_ptx.cuh"

namespace ir {

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// block-wide sum, result broadcast to every thread. `red` holds >= 33 floats.
__device__ __forceinline__ float block_sum(float v, float* red) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[warp] = v;
  __syncthreads();
  if (warp == 0) {
    float t = lane < (blockDim.x >> 5) ? red[lane] : 0.f;
    t = warp_sum(t);
    if (lane == 0) red[32] = t;
  }
  __syncthreads();
  return red[32];
}

// ---- GroupNorm pass A: per (batch, group) mean and rstd (two passes, centred variance).
// grid = batch * groups; the group's slice is cpg (even) contiguous channels per pixel.
__global__ void __launch_bounds__(512) gn_stats_kernel(const __half* __restrict__ x, int row_stride, int hw, int channels,
                                                       int groups, float eps, float2* __restrict__ stats) {
  __shared__ float red[33];
  const int b = blockIdx.x / groups, g = blockIdx.x % groups;
  const int cpg = channels / groups;
  const int hpg = cpg >> 1;  // half2 per pixel in this group
  const __half* base = x + static_cast<size_t>(b) * hw * row_stride + g * cpg;
  const int n2 = hw * hpg;
  float s = 0.f;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int px = i / hpg, c2 = i - px * hpg;
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(base + static_cast<size_t>(px) * row_stride + 2 * c2));
    s += f.x + f.y;
  }
  const float n = static_cast<float>(hw) * cpg;
  const float mean = block_sum(s, red) / n;
  float q = 0.f;
  for (int i = threadIdx.x; i < n2; i += blockDim.x) {
    const int px = i / hpg, c2 = i - px * hpg;
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(base + static_cast<size_t>(px) * row_stride + 2 * c2));
    const float dx = f.x - mean, dy = f.y - mean;
    q += dx * dx + dy * dy;
  }
  const float var = block_sum(q, red) / n;
  if (threadIdx.x == 0) stats[blockIdx.x] = make_float2(mean, rsqrtf(var + eps));
}

// ---- GroupNorm pass B: y = (x - mean) * rstd * gamma + beta, optional SiLU; 8 channels (16 B) per thread.
__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x, int x_stride, int hw, int channels,
                                                       int groups, const float2* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int silu, __half* __restrict__ out, int out_stride, long total_vec) {
  const int cpg = channels / groups;
  const int vec_per_row = channels >> 3;
  for (long v = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; v < total_vec;
       v += static_cast<long>(gridDim.x) * blockDim.x) {
    const long rowi = v / vec_per_row;
    const int c0 = static_cast<int>(v - rowi * vec_per_row) << 3;
    const int b = static_cast<int>(rowi / hw);
    const uint4 u = *reinterpret_cast<const uint4*>(x + rowi * x_stride + c0);
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h2[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int c = c0 + j;
      const float2 st = __ldg(&stats[b * groups + c / cpg]);
      float y = (f[j] - st.x) * st.y * __ldg(gamma + c) + __ldg(beta + c);
      if (silu) y = y / (1.0f + __expf(-y));
      f[j] = y;
    }
    *reinterpret_cast<uint4*>(out + rowi * out_stride + c0) =
        make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
  }
}

// ---- LayerNorm: one warp per row.
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, int x_stride, int rows, int channels,
                                                        float eps, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __half* __restrict__ out,
                                                        int out_stride) {
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __half* xr = x + static_cast<size_t>(row) * x_stride;
  const int nvec = channels >> 3;
  float s = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + (v << 3));
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h2[j]);
      s += t.x + t.y;
    }
  }
  const float mean = warp_sum(s) / channels;
  float q = 0.f;
  for (int v = lane; v < nvec; v += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + (v << 3));
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h2[j]);
      q += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / channels + eps);
  __half* orow = out + static_cast<size_t>(row) * out_stride;
  for (int v = lane; v < nvec; v += 32) {
    const int c0 = v << 3;
    const uint4 u = *reinterpret_cast<const uint4*>(xr + c0);
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
    float f[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 t = __half22float2(h2[j]);
      f[2 * j] = t.x;
      f[2 * j + 1] = t.y;
    }
    const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
    const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
    const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
    const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
    const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
    for (int j = 0; j < 8; ++j) f[j] = (f[j] - mean) * rstd * gg[j] + bb[j];
    *reinterpret_cast<uint4*>(orow + c0) =
        make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
  }
}

}  // namespace ir

extern "C" size_t ir_groupnorm_workspace_bytes(int batch, int groups) {
  return static_cast<size_t>(batch) * groups * sizeof(float2);
}

extern "C" int ir_groupnorm(const ir_groupnorm_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->x || !p->out || !p->gamma || !p->beta || !p->workspace) return set_error(IR_ERR_ARG, "ir_groupnorm: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->groups <= 0 || p->channels % p->groups != 0 || ((p->channels / p->groups) & 1) || p->channels % 8 != 0)
    return set_error(IR_ERR_SHAPE, "ir_groupnorm: channels=%d groups=%d (need channels %% 8 == 0, even channels/group)", p->channels, p->groups);
  if (p->x_row_stride % 8 || p->out_row_stride % 8 || (reinterpret_cast<uintptr_t>(p->x) & 15) || (reinterpret_cast<uintptr_t>(p->out) & 15))
    return set_error(IR_ERR_ALIGN, "ir_groupnorm: pointers/strides must be 16-byte aligned");
  if (p->batch <= 0 || p->hw <= 0) return set_error(IR_ERR_SHAPE, "ir_groupnorm: non-positive dims");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float2* stats = static_cast<float2*>(p->workspace);
  gn_stats_kernel<<<p->batch * p->groups, 512, 0, stream>>>(static_cast<const __half*>(p->x), p->x_row_stride, p->hw,
                                                             p->channels, p->groups, p->eps, stats);
  IR_CUDA_LAUNCH_CHECK("gn_stats launch");
  const long total_vec = static_cast<long>(p->batch) * p->hw * (p->channels >> 3);
  long blocks = (total_vec + 255) / 256;
  if (blocks > 148 * 16) blocks = 148 * 16;
  gn_apply_kernel<<<static_cast<int>(blocks), 256, 0, stream>>>(static_cast<const __half*>(p->x), p->x_row_stride, p->hw,
                                                                 p->channels, p->groups, stats, p->gamma, p->beta, p->silu,
                                                                 static_cast<__half*>(p->out), p->out_row_stride, total_vec);
  IR_CUDA_LAUNCH_CHECK("gn_apply launch");
  return 0;
}

extern "C" int ir_layernorm(const ir_layernorm_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->x || !p->out || !p->gamma || !p->beta) return set_error(IR_ERR_ARG, "ir_layernorm: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->channels % 8 != 0 || p->channels <= 0 || p->rows <= 0) return set_error(IR_ERR_SHAPE, "ir_layernorm: rows=%d channels=%d", p->rows, p->channels);
  if (p->x_row_stride % 8 || p->out_row_stride % 8 || (reinterpret_cast<uintptr_t>(p->x) & 15) || (reinterpret_cast<uintptr_t>(p->out) & 15) ||
      (reinterpret_cast<uintptr_t>(p->gamma) & 15) || (reinterpret_cast<uintptr_t>(p->beta) & 15))
    return set_error(IR_ERR_ALIGN, "ir_layernorm: pointers/strides must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int rows_per_block = 8;
  layernorm_kernel<<<(p->rows + rows_per_block - 1) / rows_per_block, 256, 0, stream>>>(
      static_cast<const __half*>(p->x), p->x_row_stride, p->rows, p->channels, p->eps, p->gamma, p->beta,
      static_cast<__half*>(p->out), p->out_row_stride);
  IR_CUDA_LAUNCH_CHECK("layernorm launch");
  return 0;
}
