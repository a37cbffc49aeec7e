// GroupNorm(+SiLU) and LayerNorm on channel-last fp16 with fp32 statistics (HBM-bound: slab-parallel partial moments
// with coalesced 16-byte loads, Chan merge, then one fused affine(+SiLU) pass that hits L2 for the small tensors). Algorithmic bytes: one fp16 read + one fp16 write of the activation.
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

// (count, mean, M2) merge of two partial moments (Chan et al.); n_a, n_b > 0.
__device__ __forceinline__ void merge_moments(float& n_a, float& mean_a, float& m2_a, float n_b, float mean_b, float m2_b) {
  const float n = n_a + n_b;
  const float d = mean_b - mean_a;
  const float f = n_b / n;
  mean_a = fmaf(d, f, mean_a);
  m2_a = m2_a + m2_b + d * d * n_a * f;
  n_a = n;
}

// ---- GroupNorm pass A: partial moments. grid = (slabs, batch); a CTA reads `rows_per_slab` full pixel rows with
// coalesced 16-byte loads (every thread owns fixed 8-channel vectors, so per-channel sums live in registers), folds
// channels into groups through shared memory and writes one (mean, M2) per (batch, slab, group).
// partial[((b * slabs + slab) * groups + g)] = (mean, M2); the element count follows from the slab index.
__global__ void __launch_bounds__(256) gn_partial_kernel(const __half* __restrict__ x, int row_stride, int hw, int channels,
                                                         int groups, int rows_per_slab, float2* __restrict__ partial) {
  extern __shared__ float sm[];                 // [row group][channels][2]: per-channel (sum, sumsq) of this slab
  const int slab = blockIdx.x, b = blockIdx.y, slabs = gridDim.x;
  const int vpr = channels >> 3;                // 16-byte vectors per row
  const int r0 = slab * rows_per_slab;
  const int r1 = min(r0 + rows_per_slab, hw);
  const __half* base = x + static_cast<size_t>(b) * hw * row_stride;
  // thread layout: vpr >= 256 -> thread owns vectors t, t+256, ... over all rows; else (row lane, vector)
  const int lanes = vpr < 256 ? vpr : 256;
  const int rgroups = 256 / lanes;              // rows processed concurrently
  const int vec0 = threadIdx.x % lanes;
  const int rg = threadIdx.x / lanes;
  if (rg < rgroups) {
    for (int v = vec0; v < vpr; v += lanes) {
      float s[8], q[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
      int r = r0 + rg;
      // 4 independent 16-byte loads in flight per thread
      for (; r + 3 * rgroups < r1; r += 4 * rgroups) {
        uint4 u[4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
          u[k] = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(r + k * rgroups) * row_stride + (v << 3));
#pragma unroll
        for (int k = 0; k < 4; ++k) {
          const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
          for (int j = 0; j < 4; ++j) {
            const float2 f = __half22float2(h2[j]);
            s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
            s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
          }
        }
      }
      for (; r < r1; r += rgroups) {
        const uint4 u = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(r) * row_stride + (v << 3));
        const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
          s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
        }
      }
      float* dst = sm + (static_cast<size_t>(rg) * channels + (v << 3)) * 2;   // one owner per slot: no atomics
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        dst[2 * j] = s[j];
        dst[2 * j + 1] = q[j];
      }
    }
  }
  __syncthreads();
  const int cpg = channels / groups;
  const float n = static_cast<float>(r1 - r0) * cpg;
  for (int g = threadIdx.x; g < groups; g += blockDim.x) {
    float s = 0.f, q = 0.f;
    for (int k = 0; k < rgroups; ++k)            // fixed order: deterministic
      for (int c = g * cpg; c < (g + 1) * cpg; ++c) {
        s += sm[(static_cast<size_t>(k) * channels + c) * 2];
        q += sm[(static_cast<size_t>(k) * channels + c) * 2 + 1];
      }
    const float mean = s / n;
    partial[(static_cast<size_t>(b) * slabs + slab) * groups + g] = make_float2(mean, fmaxf(q - s * mean, 0.f));
  }
}

// ---- GroupNorm pass B: merge the slab moments of this batch entry (prologue), fold them with gamma/beta into a
// per-channel affine in shared memory, then y = a[c] * x + b[c] (+ SiLU); 8 channels (16 B) per thread.
// grid = (blocks_per_batch, batch).
__global__ void __launch_bounds__(256) gn_apply_kernel(const __half* __restrict__ x, int x_stride, int hw, int channels,
                                                       int groups, int slabs, int rows_per_slab, float eps,
                                                       const float2* __restrict__ partial, const float* __restrict__ gamma,
                                                       const float* __restrict__ beta, int silu, __half* __restrict__ out,
                                                       int out_stride) {
  extern __shared__ float sm[];                 // [channels][2] (a, b) then [groups][2] (mean, rstd)
  float* ab = sm;
  float* st = sm + 2 * channels;
  const int b = blockIdx.y;
  const int cpg = channels / groups;
  // 8 threads per group merge slabs/8 partials each, then a 3-step shuffle merge
  for (int g0 = 0; g0 < groups; g0 += 32) {
    const int g = g0 + (threadIdx.x >> 3), sub = threadIdx.x & 7;
    float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
    if (g < groups) {
      for (int sl = sub; sl < slabs; sl += 8) {
        const float2 pm = __ldg(&partial[(static_cast<size_t>(b) * slabs + sl) * groups + g]);
        const float n_b = static_cast<float>(min(rows_per_slab, hw - sl * rows_per_slab)) * cpg;
        if (n_a == 0.f) { n_a = n_b; mean_a = pm.x; m2_a = pm.y; }
        else merge_moments(n_a, mean_a, m2_a, n_b, pm.x, pm.y);
      }
    }
#pragma unroll
    for (int o = 1; o < 8; o <<= 1) {
      const float n_b = __shfl_xor_sync(0xffffffffu, n_a, o);
      const float mean_b = __shfl_xor_sync(0xffffffffu, mean_a, o);
      const float m2_b = __shfl_xor_sync(0xffffffffu, m2_a, o);
      if (n_b > 0.f) {
        if (n_a == 0.f) { n_a = n_b; mean_a = mean_b; m2_a = m2_b; }
        else merge_moments(n_a, mean_a, m2_a, n_b, mean_b, m2_b);
      }
    }
    if (g < groups && sub == 0) {
      st[2 * g] = mean_a;
      st[2 * g + 1] = rsqrtf(m2_a / n_a + eps);
    }
  }
  __syncthreads();
  for (int c = threadIdx.x; c < channels; c += blockDim.x) {
    const int g = c / cpg;
    const float a = st[2 * g + 1] * __ldg(gamma + c);
    ab[2 * c] = a;
    ab[2 * c + 1] = fmaf(-st[2 * g], a, __ldg(beta + c));
  }
  __syncthreads();
  const int vec_per_row = channels >> 3;
  const long total_vec = static_cast<long>(hw) * vec_per_row;
  const __half* xb = x + static_cast<size_t>(b) * hw * x_stride;
  __half* ob = out + static_cast<size_t>(b) * hw * out_stride;
  const long stride = static_cast<long>(gridDim.x) * blockDim.x;
  auto apply8 = [&](const uint4& u, int c0) -> uint4 {
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
      const float2 cf = *reinterpret_cast<const float2*>(&ab[2 * (c0 + j)]);
      float y = fmaf(f[j], cf.x, cf.y);
      if (silu) y = y / (1.0f + __expf(-y));
      f[j] = y;
    }
    return make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
  };
  long v = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  // 4 independent 16-byte loads in flight per thread
  for (; v + 3 * stride < total_vec; v += 4 * stride) {
    uint4 u[4];
    long rowi[4];
    int c0[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const long vv = v + k * stride;
      rowi[k] = vv / vec_per_row;
      c0[k] = static_cast<int>(vv - rowi[k] * vec_per_row) << 3;
      u[k] = *reinterpret_cast<const uint4*>(xb + rowi[k] * x_stride + c0[k]);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) *reinterpret_cast<uint4*>(ob + rowi[k] * out_stride + c0[k]) = apply8(u[k], c0[k]);
  }
  for (; v < total_vec; v += stride) {
    const long rowi = v / vec_per_row;
    const int c0 = static_cast<int>(v - rowi * vec_per_row) << 3;
    const uint4 u = *reinterpret_cast<const uint4*>(xb + rowi * x_stride + c0);
    *reinterpret_cast<uint4*>(ob + rowi * out_stride + c0) = apply8(u, c0);
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

static void gn_plan(int batch, int hw, int* slabs, int* rows_per_slab) {
  // <= 32 slabs per batch entry: the apply kernel's prologue merges them (8 lanes x 4 dependent steps); even at
  // batch 1 thirty-two CTAs pull a <= 8 MB activation out of L2/HBM in a couple of microseconds.
  int want = hw >= 65536 ? 256 : (hw >= 16384 ? 128 : 32);   // x batch CTAs   // VAE-resolution tensors need more CTAs to reach HBM speed
  int max_slabs = hw / 8 > 0 ? hw / 8 : 1;         // >= 8 rows per slab
  if (want > max_slabs) want = max_slabs;
  (void)batch;
  const int rps = (hw + want - 1) / want;
  *rows_per_slab = rps;
  *slabs = (hw + rps - 1) / rps;
}

extern "C" size_t ir_groupnorm_workspace_bytes(int batch, int groups) {
  return static_cast<size_t>(batch) * 256 * groups * sizeof(float2);   // up to 256 slabs per batch entry
}

extern "C" int ir_groupnorm(const ir_groupnorm_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->x || !p->out || !p->gamma || !p->beta || !p->workspace) return set_error(IR_ERR_ARG, "ir_groupnorm: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->groups <= 0 || p->channels % p->groups != 0 || p->channels % 8 != 0 || p->channels > 4096)
    return set_error(IR_ERR_SHAPE, "ir_groupnorm: channels=%d groups=%d (need channels %% 8 == 0, channels %% groups == 0, channels <= 4096)", p->channels, p->groups);
  if (p->x_row_stride % 8 || p->out_row_stride % 8 || (reinterpret_cast<uintptr_t>(p->x) & 15) || (reinterpret_cast<uintptr_t>(p->out) & 15))
    return set_error(IR_ERR_ALIGN, "ir_groupnorm: pointers/strides must be 16-byte aligned");
  if (p->batch <= 0 || p->hw <= 0) return set_error(IR_ERR_SHAPE, "ir_groupnorm: non-positive dims");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  float2* partial = static_cast<float2*>(p->workspace);
  int slabs, rps;
  gn_plan(p->batch, p->hw, &slabs, &rps);
  const int vpr = p->channels >> 3;
  const int rgroups = vpr < 256 ? 256 / vpr : 1;
  gn_partial_kernel<<<dim3(slabs, p->batch), 256, static_cast<size_t>(rgroups) * p->channels * 2 * sizeof(float), stream>>>(
      static_cast<const __half*>(p->x), p->x_row_stride, p->hw, p->channels, p->groups, rps, partial);
  IR_CUDA_LAUNCH_CHECK("gn_partial launch");
  const long total_vec = static_cast<long>(p->hw) * (p->channels >> 3);
  long blocks = (total_vec + 256 * 4 - 1) / (256 * 4);                   // ~4 vectors per thread
  const long cap = (148 * 8 + p->batch - 1) / p->batch;
  if (blocks > cap) blocks = cap;
  if (blocks < 1) blocks = 1;
  const size_t smem = (static_cast<size_t>(p->channels) * 2 + p->groups * 2) * sizeof(float);
  gn_apply_kernel<<<dim3(static_cast<unsigned>(blocks), p->batch), 256, smem, stream>>>(
      static_cast<const __half*>(p->x), p->x_row_stride, p->hw, p->channels, p->groups, slabs, rps, p->eps, partial,
      p->gamma, p->beta, p->silu, static_cast<__half*>(p->out), p->out_row_stride);
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
