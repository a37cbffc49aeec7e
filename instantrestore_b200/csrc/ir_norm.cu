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
  pdl_launch_dependents();
  pdl_wait();
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

// ---- GroupNorm pass A2: merge the slab moments (Chan) into per-(batch, group) mean / rstd. grid = (groups, batch):
// one CTA per statistic, every thread takes <= 4 slabs per sweep (loads issued ahead of the merge arithmetic), then a
// shuffle tree and an 8-entry shared-memory merge in fixed order (deterministic). The serial chain is ~20 merges
// instead of one per slab — this kernel used to be 11 us of pure latency per GroupNorm.
__device__ __forceinline__ void merge_opt(float& n_a, float& mean_a, float& m2_a, float n_b, float mean_b, float m2_b) {
  if (n_b > 0.f) {
    if (n_a == 0.f) { n_a = n_b; mean_a = mean_b; m2_a = m2_b; }
    else merge_moments(n_a, mean_a, m2_a, n_b, mean_b, m2_b);
  }
}

__global__ void __launch_bounds__(256) gn_merge_kernel(const float2* __restrict__ partial, int groups, int slabs,
                                                       int rows_per_slab, int hw, int cpg, float eps,
                                                       float2* __restrict__ stats) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_n[8], s_mean[8], s_m2[8];
  const int g = blockIdx.x, b = blockIdx.y;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const float2* src = partial + static_cast<size_t>(b) * slabs * groups + g;
  float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
  for (int sl0 = threadIdx.x; sl0 < slabs; sl0 += 1024) {
    float2 pm[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int sl = sl0 + 256 * k;
      pm[k] = sl < slabs ? __ldg(src + static_cast<size_t>(sl) * groups) : make_float2(0.f, 0.f);
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
      const int sl = sl0 + 256 * k;
      if (sl < slabs) merge_opt(n_a, mean_a, m2_a, static_cast<float>(min(rows_per_slab, hw - sl * rows_per_slab)) * cpg, pm[k].x, pm[k].y);
    }
  }
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    const float n_b = __shfl_xor_sync(0xffffffffu, n_a, o);
    const float mean_b = __shfl_xor_sync(0xffffffffu, mean_a, o);
    const float m2_b = __shfl_xor_sync(0xffffffffu, m2_a, o);
    merge_opt(n_a, mean_a, m2_a, n_b, mean_b, m2_b);
  }
  if (lane == 0) { s_n[warp] = n_a; s_mean[warp] = mean_a; s_m2[warp] = m2_a; }
  __syncthreads();
  if (threadIdx.x == 0) {
    for (int w = 1; w < 8; ++w) merge_opt(n_a, mean_a, m2_a, s_n[w], s_mean[w], s_m2[w]);
    stats[b * groups + g] = make_float2(mean_a, rsqrtf(m2_a / n_a + eps));
  }
}

constexpr int kGnApplyThreads = 256;

// ---- GroupNorm pass B: y = a[c] * x + b[c] (+ SiLU) with a = rstd * gamma, b = beta - mean * a. Every thread owns
// fixed 8-channel vectors (its 16 affine coefficients live in registers) and walks the rows of its CTA's row range
// with 4 sixteen-byte loads in flight: no per-element index math, no shared-memory lookups.
// grid = (row blocks, batch).
template <int NT>
__global__ void __launch_bounds__(NT, 1024 / NT) gn_apply_kernel(const __half* __restrict__ x, int x_stride, int hw, int channels,
                                                       int groups, int rows_per_block, const float2* __restrict__ stats,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int silu, __half* __restrict__ out, int out_stride) {
  const int b = blockIdx.y;
  pdl_launch_dependents();
  pdl_wait();
  const int cpg = channels / groups;
  const int vpr = channels >> 3;
  const int lanes = vpr < NT ? vpr : NT;
  const int rgroups = NT / lanes;
  const int vec0 = threadIdx.x % lanes, rg = threadIdx.x / lanes;
  if (rg >= rgroups) return;
  const int r0 = blockIdx.x * rows_per_block;
  const int r1 = min(r0 + rows_per_block, hw);
  const __half* xb = x + static_cast<size_t>(b) * hw * x_stride;
  __half* ob = out + static_cast<size_t>(b) * hw * out_stride;
  for (int v = vec0; v < vpr; v += lanes) {
    const int c0 = v << 3;
    float ca[8], cb[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 st = __ldg(&stats[b * groups + (c0 + j) / cpg]);
      ca[j] = st.y * __ldg(gamma + c0 + j);
      cb[j] = fmaf(-st.x, ca[j], __ldg(beta + c0 + j));
    }
    auto apply8 = [&](const uint4& u) -> uint4 {
      const __half2* h2 = reinterpret_cast<const __half2*>(&u);
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h2[j]);
        f[2 * j] = fmaf(t.x, ca[2 * j], cb[2 * j]);
        f[2 * j + 1] = fmaf(t.y, ca[2 * j + 1], cb[2 * j + 1]);
      }
      if (silu) {
#pragma unroll
        for (int j = 0; j < 8; ++j) f[j] = __fdividef(f[j], 1.0f + fast_exp2(-1.4426950408889634f * f[j]));
      }
      return make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
    };
    int r = r0 + rg;
    // software pipeline: the four loads of the NEXT sweep are in flight while this sweep runs its 64 MUFU ops and its
    // stores (ncu r02h: long-scoreboard stalls 16 per issue, 0.65 of the copy bandwidth without it)
    if (r + 3 * rgroups < r1) {
      uint4 u[4];
#pragma unroll
      for (int k = 0; k < 4; ++k)
        u[k] = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(r + k * rgroups) * x_stride + c0);
      for (;;) {
        const int rn = r + 4 * rgroups;
        const bool more = rn + 3 * rgroups < r1;
        uint4 nx[4];
        if (more) {
#pragma unroll
          for (int k = 0; k < 4; ++k)
            nx[k] = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(rn + k * rgroups) * x_stride + c0);
        }
#pragma unroll
        for (int k = 0; k < 4; ++k)
          *reinterpret_cast<uint4*>(ob + static_cast<size_t>(r + k * rgroups) * out_stride + c0) = apply8(u[k]);
        r = rn;
        if (!more) break;
#pragma unroll
        for (int k = 0; k < 4; ++k) u[k] = nx[k];
      }
    }
    for (; r < r1; r += rgroups) {
      const uint4 u = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(r) * x_stride + c0);
      *reinterpret_cast<uint4*>(ob + static_cast<size_t>(r) * out_stride + c0) = apply8(u);
    }
  }
}

// ---- GroupNorm(+SiLU) in ONE launch with ONE read of the activation (every UNet norm and the VAE norms up to 128 x 128).
// grid = (CL, slices, batch), thread-block cluster (CL, 1, 1). A CTA owns `rows` pixels x `slice_ch` channels of one
// image (slice_ch = a multiple of lcm(8, channels per group): whole groups AND whole 16-byte vectors); its tile lives in
// registers (<= KMAX 16-byte vectors per thread, every thread keeps one fixed channel vector so its per-channel sums and
// later its 16 affine coefficients stay in registers). Per-CTA (mean, M2) of each group of the slice are exchanged
// through distributed shared memory: after one cluster barrier every CTA merges the CL partials in rank order (Chan;
// deterministic and identical in every CTA), applies a*x+b (+SiLU) to its registers and stores. The second cluster
// barrier (no CTA may exit while a peer still reads its shared memory) is split: arrive right after the remote reads,
// wait after the stores.
__device__ __forceinline__ float2 dsmem_ld_f2(uint32_t addr) {
  float2 v;
  asm volatile("ld.shared::cluster.v2.f32 {%0, %1}, [%2];" : "=f"(v.x), "=f"(v.y) : "r"(addr) : "memory");
  return v;
}
__device__ __forceinline__ void cluster_arrive() { asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory"); }
__device__ __forceinline__ void cluster_wait() { asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory"); }

template <int KMAX>
__global__ void __launch_bounds__(256) gn_fused_kernel(const __half* __restrict__ x, int x_stride, int hw, int channels,
                                                       int groups, int slice_ch, int rows, int cl, float eps,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta,
                                                       int silu, __half* __restrict__ out, int out_stride) {
  __shared__ float2 s_acc[256 * 8];      // [row group][slice channel] (sum, sumsq) of the thread-level partials
  __shared__ float2 s_ch[256];           // stage A: [T][slice channel], T * slice_ch <= 256
  __shared__ float2 s_part[64];          // this CTA's (mean, M2) per group of the slice (read by the cluster)
  __shared__ float2 s_stat[64];          // merged (mean, rstd) per group of the slice
  const int tid = threadIdx.x;
  pdl_launch_dependents();
  pdl_wait();
  const int V = slice_ch >> 3;                    // 16-byte vectors per row of the slice
  const int rgroups = 256 / V;                    // rows in flight per sweep
  const int v = tid % V, rg = tid / V;
  const bool active = rg < rgroups;
  const int b = blockIdx.z, c_base = blockIdx.y * slice_ch;
  const int r0 = blockIdx.x * rows;               // cluster rank == blockIdx.x (cluster spans the x dimension)
  const int cpg = channels / groups;
  const int gs = slice_ch / cpg;                  // groups in this slice
  const __half* xb = x + (static_cast<size_t>(b) * hw + r0) * x_stride + c_base + (v << 3);

  uint4 u[KMAX];
  float s[8], q[8];
  float4 gm[2], bt[2];                              // gamma / beta of this thread's channel vector, fetched ahead of the statistics
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  if (active) {
    const int c0 = c_base + (v << 3);
    gm[0] = __ldg(reinterpret_cast<const float4*>(gamma + c0)); gm[1] = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
    bt[0] = __ldg(reinterpret_cast<const float4*>(beta + c0)); bt[1] = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int r = rg + k * rgroups;
      if (r < rows) u[k] = *reinterpret_cast<const uint4*>(xb + static_cast<size_t>(r) * x_stride);
    }
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int r = rg + k * rgroups;
      if (r < rows) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 f = __half22float2(h2[j]);
          s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
          s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
        }
      }
    }
    float2* dst = s_acc + rg * slice_ch + (v << 3);
#pragma unroll
    for (int j = 0; j < 8; ++j) dst[j] = make_float2(s[j], q[j]);
  }
  __syncthreads();
  // stage A: T threads per channel, each sums a strided subset of the row groups (fixed order)
  const int T = 256 / slice_ch > 0 ? 256 / slice_ch : 1;
  for (int i = tid; i < T * slice_ch; i += 256) {
    const int ch = i % slice_ch, t = i / slice_ch;
    float a = 0.f, c = 0.f;
    for (int k = t; k < rgroups; k += T) {
      const float2 e = s_acc[k * slice_ch + ch];
      a += e.x; c += e.y;
    }
    s_ch[t * slice_ch + ch] = make_float2(a, c);
  }
  __syncthreads();
  // stage B: one thread per group of the slice
  const float n_cta = static_cast<float>(rows) * cpg;
  if (tid < gs) {
    float a = 0.f, c = 0.f;
    for (int t = 0; t < T; ++t)
      for (int ch = tid * cpg; ch < (tid + 1) * cpg; ++ch) {
        const float2 e = s_ch[t * slice_ch + ch];
        a += e.x; c += e.y;
      }
    const float mean = a / n_cta;
    s_part[tid] = make_float2(mean, fmaxf(c - a * mean, 0.f));
  }
  if (cl > 1) {
    cluster_arrive();
    cluster_wait();
  } else {
    __syncthreads();
  }
  if (tid < gs) {
    float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
    const uint32_t local = smem_u32(&s_part[tid]);
    for (int r = 0; r < cl; ++r) {
      const float2 pm = cl > 1 ? dsmem_ld_f2(dsmem_addr(local, static_cast<uint32_t>(r))) : s_part[tid];
      if (r == 0) { n_a = n_cta; mean_a = pm.x; m2_a = pm.y; }
      else merge_moments(n_a, mean_a, m2_a, n_cta, pm.x, pm.y);
    }
    s_stat[tid] = make_float2(mean_a, rsqrtf(m2_a / n_a + eps));
  }
  if (cl > 1) cluster_arrive();      // my remote reads are done
  __syncthreads();
  if (active) {
    const int c0 = c_base + (v << 3);
    float ca[8], cb[8];
    const float gg[8] = {gm[0].x, gm[0].y, gm[0].z, gm[0].w, gm[1].x, gm[1].y, gm[1].z, gm[1].w};
    const float bb[8] = {bt[0].x, bt[0].y, bt[0].z, bt[0].w, bt[1].x, bt[1].y, bt[1].z, bt[1].w};
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const float2 st = s_stat[((v << 3) + j) / cpg];
      ca[j] = st.y * gg[j];
      cb[j] = fmaf(-st.x, ca[j], bb[j]);
    }
    __half* ob = out + (static_cast<size_t>(b) * hw + r0) * out_stride + c0;
#pragma unroll
    for (int k = 0; k < KMAX; ++k) {
      const int r = rg + k * rgroups;
      if (r < rows) {
        const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
        float f[8];
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          const float2 t = __half22float2(h2[j]);
          f[2 * j] = fmaf(t.x, ca[2 * j], cb[2 * j]);
          f[2 * j + 1] = fmaf(t.y, ca[2 * j + 1], cb[2 * j + 1]);
        }
        if (silu) {
#pragma unroll
          for (int j = 0; j < 8; ++j) f[j] = __fdividef(f[j], 1.0f + fast_exp2(-1.4426950408889634f * f[j]));
        }
        *reinterpret_cast<uint4*>(ob + static_cast<size_t>(r) * out_stride) =
            make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
      }
    }
  }
  if (cl > 1) cluster_wait();
}

// ---- LayerNorm: one warp per row; the row (<= 2560 channels = 10 sixteen-byte vectors per lane) is read from
// global memory once and stays in registers for the mean, the centred variance and the affine.
template <int kMaxVec>
__global__ void __launch_bounds__(256) layernorm_kernel(const __half* __restrict__ x, int x_stride, int rows, int channels,
                                                        float eps, const float* __restrict__ gamma,
                                                        const float* __restrict__ beta, __half* __restrict__ out,
                                                        int out_stride) {
  pdl_launch_dependents();
  pdl_wait();
  const int lane = threadIdx.x & 31;
  const int row = blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
  if (row >= rows) return;
  const __half* xr = x + static_cast<size_t>(row) * x_stride;
  const int nvec = channels >> 3;
  uint4 u[kMaxVec];
  float s = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) {
      u[k] = *reinterpret_cast<const uint4*>(xr + (v << 3));
      const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h2[j]);
        s += t.x + t.y;
      }
    }
  }
  const float mean = warp_sum(s) / channels;
  float q = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    if (lane + 32 * k < nvec) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h2[j]);
        q += (t.x - mean) * (t.x - mean) + (t.y - mean) * (t.y - mean);
      }
    }
  }
  const float rstd = rsqrtf(warp_sum(q) / channels + eps);
  __half* orow = out + static_cast<size_t>(row) * out_stride;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int v = lane + 32 * k;
    if (v < nvec) {
      const int c0 = v << 3;
      const __half2* h2 = reinterpret_cast<const __half2*>(&u[k]);
      const float4 g0 = __ldg(reinterpret_cast<const float4*>(gamma + c0));
      const float4 g1 = __ldg(reinterpret_cast<const float4*>(gamma + c0 + 4));
      const float4 b0 = __ldg(reinterpret_cast<const float4*>(beta + c0));
      const float4 b1 = __ldg(reinterpret_cast<const float4*>(beta + c0 + 4));
      const float gg[8] = {g0.x, g0.y, g0.z, g0.w, g1.x, g1.y, g1.z, g1.w};
      const float bb[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
      float f[8];
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 t = __half22float2(h2[j]);
        f[2 * j] = (t.x - mean) * rstd * gg[2 * j] + bb[2 * j];
        f[2 * j + 1] = (t.y - mean) * rstd * gg[2 * j + 1] + bb[2 * j + 1];
      }
      *reinterpret_cast<uint4*>(orow + c0) =
          make_uint4(pack_half2(f[0], f[1]), pack_half2(f[2], f[3]), pack_half2(f[4], f[5]), pack_half2(f[6], f[7]));
    }
  }
}

}  // namespace ir

constexpr int kGnMaxSlabs = 1024;

int ir::launch_gn_partial(const void* x, int row_stride, int batch, int hw, int channels, int groups, int rows_per_slab,
                          void* partial, cudaStream_t stream) {
  const int slabs = (hw + rows_per_slab - 1) / rows_per_slab;
  const int vpr = channels >> 3;
  const int rgroups = vpr < 256 ? 256 / vpr : 1;
  IR_LAUNCH(gn_partial_kernel, dim3(slabs, batch), 256, static_cast<size_t>(rgroups) * channels * 2 * sizeof(float), stream,
            static_cast<const __half*>(x), row_stride, hw, channels, groups, rows_per_slab, static_cast<float2*>(partial));
  IR_CUDA_LAUNCH_CHECK("gn_partial launch");
  return 0;
}

static void gn_plan(int batch, int hw, int* slabs, int* rows_per_slab) {
  // enough CTAs for ~10 MB of loads in flight (592 = 4 per SM), >= 8 rows per slab
  int want = (592 + batch - 1) / batch;
  if (want < 32) want = 32;
  if (want > kGnMaxSlabs) want = kGnMaxSlabs;
  int max_slabs = hw / 8 > 0 ? hw / 8 : 1;
  if (want > max_slabs) want = max_slabs;
  const int rps = (hw + want - 1) / want;
  *rows_per_slab = rps;
  *slabs = (hw + rps - 1) / rps;
}

// ---- plan of the single-launch GroupNorm: slice width, cluster size, rows per CTA, register tile depth
struct GnFusedPlan { int slice_ch, cl, rows, kmax; };
constexpr long kGnFusedMaxBytes = 12l << 20;

static int gcd_i(int a, int b) { while (b) { int t = a % b; a = b; b = t; } return a; }

static bool gn_fused_plan(int batch, int hw, int channels, int groups, GnFusedPlan* pl) {
  if (groups <= 0 || channels % groups || channels % 8) return false;
  // latency-bound tensors only: above ~12 MB the three-kernel path streams faster (same-box A/B at 8 identities per step:
  // 69.4 vs 66.9 images/s) — a register-resident tile serialises its load, reduce and store phases
  if (static_cast<long>(batch) * hw * channels * 2 > kGnFusedMaxBytes) return false;
  const int cpg = channels / groups;
  const int wmin = cpg / gcd_i(cpg, 8) * 8;               // lcm(8, channels per group)
  if (wmin > 256 || channels % wmin) return false;
  int w = wmin;
  while (w < 32 && channels % (2 * w) == 0) w *= 2;        // >= 64-byte row segments where the width allows
  if (w / cpg > 64) return false;
  const int rgroups = 256 / (w / 8);
  // smallest power-of-two cluster (<= 8: portable) whose per-CTA tile fits 32 vectors per thread
  auto depth = [&](int c) { return (hw / c + rgroups - 1) / rgroups; };
  int cl = 1;
  while (cl <= 8 && (hw % cl != 0 || depth(cl) > 32)) cl *= 2;
  if (cl > 8) return false;
  // prefer <= 16 vectors per thread; then more CTAs while the grid is small (latency-bound tensors) and every thread
  // keeps >= 2 vectors
  const long slices = channels / w;
  while (cl < 8 && hw % (2 * cl) == 0 && depth(cl) > 16) cl *= 2;
  while (cl < 8 && hw % (2 * cl) == 0 && slices * cl * batch < 296 && hw / (2 * cl) >= 2 * rgroups) cl *= 2;
  const int rows = hw / cl;
  const int k = (rows + rgroups - 1) / rgroups;
  pl->slice_ch = w; pl->cl = cl; pl->rows = rows;
  pl->kmax = k <= 4 ? 4 : k <= 8 ? 8 : k <= 16 ? 16 : 32;
  return true;
}

extern "C" int ir_groupnorm_fused_supported(int batch, int hw, int channels, int groups) {
  GnFusedPlan pl;
  return batch > 0 && hw > 0 && gn_fused_plan(batch, hw, channels, groups, &pl) ? 1 : 0;
}

template <int KMAX>
static cudaError_t launch_gn_fused(const ir_groupnorm_params* p, const GnFusedPlan& pl, cudaStream_t stream) {
  return ir::launch_kernel(ir::gn_fused_kernel<KMAX>, dim3(pl.cl, p->channels / pl.slice_ch, p->batch), dim3(256), 0, stream,
                       dim3(pl.cl, 1, 1), static_cast<const __half*>(p->x), p->x_row_stride, p->hw,
                       p->channels, p->groups, pl.slice_ch, pl.rows, pl.cl, p->eps, p->gamma, p->beta, p->silu,
                       static_cast<__half*>(p->out), p->out_row_stride);
}

extern "C" size_t ir_groupnorm_workspace_bytes(int batch, int groups) {
  // per-(batch, slab, group) partial moments + per-(batch, group) mean / rstd
  return static_cast<size_t>(batch) * (kGnMaxSlabs + 1) * groups * sizeof(float2);
}

extern "C" int ir_groupnorm(const ir_groupnorm_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->x || !p->out || !p->gamma || !p->beta) return set_error(IR_ERR_ARG, "ir_groupnorm: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->groups <= 0 || p->channels % p->groups != 0 || p->channels % 8 != 0 || p->channels > 4096)
    return set_error(IR_ERR_SHAPE, "ir_groupnorm: channels=%d groups=%d (need channels %% 8 == 0, channels %% groups == 0, channels <= 4096)", p->channels, p->groups);
  if (p->x_row_stride % 8 || p->out_row_stride % 8 || (reinterpret_cast<uintptr_t>(p->x) & 15) || (reinterpret_cast<uintptr_t>(p->out) & 15))
    return set_error(IR_ERR_ALIGN, "ir_groupnorm: pointers/strides must be 16-byte aligned");
  if (p->batch <= 0 || p->hw <= 0) return set_error(IR_ERR_SHAPE, "ir_groupnorm: non-positive dims");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  GnFusedPlan pl;
  const bool fused_ok = p->fused != 1 && gn_fused_plan(p->batch, p->hw, p->channels, p->groups, &pl) &&
                        ((reinterpret_cast<uintptr_t>(p->gamma) | reinterpret_cast<uintptr_t>(p->beta)) & 15) == 0;
  if (p->fused == 2 && !fused_ok)
    return set_error(IR_ERR_SHAPE, "ir_groupnorm: fused=2 but batch=%d hw=%d channels=%d groups=%d has no single-launch plan", p->batch, p->hw, p->channels, p->groups);
  if (fused_ok) {
    cudaError_t e = pl.kmax == 4 ? launch_gn_fused<4>(p, pl, stream) : pl.kmax == 8 ? launch_gn_fused<8>(p, pl, stream)
                  : pl.kmax == 16 ? launch_gn_fused<16>(p, pl, stream) : launch_gn_fused<32>(p, pl, stream);
    count_launch();
    if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "gn_fused launch: %s", cudaGetErrorString(e));
    return 0;
  }
  if (!p->workspace) return set_error(IR_ERR_ARG, "ir_groupnorm: workspace is NULL (required when the single-launch path does not apply)");
  float2* partial = static_cast<float2*>(p->workspace);
  float2* stats = partial + static_cast<size_t>(p->batch) * kGnMaxSlabs * p->groups;
  int slabs, rps;
  const int vpr = p->channels >> 3;
  // Threads per apply CTA (64 registers each). 128: a CTA takes 8192 registers, so one fits NEXT TO a resident 320-thread x
  // 168-register conv / attention CTA of another request (11776 registers free per SM) and the HBM-bound pass overlaps the
  // tensor-bound one; 256-thread CTAs (16384 registers) have to wait for the SM. IR_GN_APPLY_THREADS=256|128 for the A/B.
  static const int apply_nt = [] { const char* e = getenv("IR_GN_APPLY_THREADS"); const int v = e ? atoi(e) : 0; return v == 128 || v == 256 ? v : kGnApplyThreads; }();
  const int rgroups = vpr < apply_nt ? apply_nt / vpr : 1;
  if (p->partial_in) {
    // pass A already ran in the epilogue of the convolution that produced x: 32-pixel slabs
    if (p->hw % 32 != 0) return set_error(IR_ERR_SHAPE, "ir_groupnorm: partial_in needs hw %% 32 == 0 (hw=%d)", p->hw);
    slabs = p->hw / 32;
    rps = 32;
    partial = const_cast<float2*>(static_cast<const float2*>(p->partial_in));
  } else {
    gn_plan(p->batch, p->hw, &slabs, &rps);
    if (int rc = launch_gn_partial(p->x, p->x_row_stride, p->batch, p->hw, p->channels, p->groups, rps, partial, stream)) return rc;
  }
  IR_LAUNCH(gn_merge_kernel, dim3(p->groups, p->batch), 256, 0, stream, partial, p->groups, slabs, rps, p->hw, p->channels / p->groups, p->eps, stats);
  IR_CUDA_LAUNCH_CHECK("gn_merge launch");
  // apply: one wave of 4 resident CTAs per SM (64 registers: two sweeps of four 16-byte loads per thread in flight),
  // >= 4 * rgroups rows per CTA so the pipelined loop is used
  int row_blocks = (148 * (1024 / apply_nt) + p->batch - 1) / p->batch;
  int min_rows = 4 * rgroups;
  int rpb = (p->hw + row_blocks - 1) / row_blocks;
  if (rpb < min_rows) rpb = min_rows;
  row_blocks = (p->hw + rpb - 1) / rpb;
  if (apply_nt == 128)
    IR_LAUNCH(gn_apply_kernel<128>, dim3(row_blocks, p->batch), 128, 0, stream,
              static_cast<const __half*>(p->x), p->x_row_stride, p->hw, p->channels, p->groups, rpb, stats, p->gamma, p->beta, p->silu,
              static_cast<__half*>(p->out), p->out_row_stride);
  else
    IR_LAUNCH(gn_apply_kernel<256>, dim3(row_blocks, p->batch), 256, 0, stream,
              static_cast<const __half*>(p->x), p->x_row_stride, p->hw, p->channels, p->groups, rpb, stats, p->gamma, p->beta, p->silu,
              static_cast<__half*>(p->out), p->out_row_stride);
  IR_CUDA_LAUNCH_CHECK("gn_apply launch");
  return 0;
}

extern "C" int ir_layernorm(const ir_layernorm_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->x || !p->out || !p->gamma || !p->beta) return set_error(IR_ERR_ARG, "ir_layernorm: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->channels % 8 != 0 || p->channels <= 0 || p->channels > 2560 || p->rows <= 0)
    return set_error(IR_ERR_SHAPE, "ir_layernorm: rows=%d channels=%d (channels %% 8 == 0, <= 2560)", p->rows, p->channels);
  if (p->x_row_stride % 8 || p->out_row_stride % 8 || (reinterpret_cast<uintptr_t>(p->x) & 15) || (reinterpret_cast<uintptr_t>(p->out) & 15) ||
      (reinterpret_cast<uintptr_t>(p->gamma) & 15) || (reinterpret_cast<uintptr_t>(p->beta) & 15))
    return set_error(IR_ERR_ALIGN, "ir_layernorm: pointers/strides must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int rows_per_block = 8;
  const int blocks = (p->rows + rows_per_block - 1) / rows_per_block;
  const __half* xp = static_cast<const __half*>(p->x);
  __half* op = static_cast<__half*>(p->out);
  if (p->channels <= 512)
    IR_LAUNCH(layernorm_kernel<2>, blocks, 256, 0, stream, xp, p->x_row_stride, p->rows, p->channels, p->eps, p->gamma, p->beta, op, p->out_row_stride);
  else if (p->channels <= 1280)
    IR_LAUNCH(layernorm_kernel<5>, blocks, 256, 0, stream, xp, p->x_row_stride, p->rows, p->channels, p->eps, p->gamma, p->beta, op, p->out_row_stride);
  else
    IR_LAUNCH(layernorm_kernel<10>, blocks, 256, 0, stream, xp, p->x_row_stride, p->rows, p->channels, p->eps, p->gamma, p->beta, op, p->out_row_stride);
  IR_CUDA_LAUNCH_CHECK("layernorm launch");
  return 0;
}
