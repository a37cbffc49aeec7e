// HBM-bound helpers of the hot path: AdaIN statistics -> affine, FreeU (closed form) + concat, nearest 2x
// upsample, scheduler AXPYs fused with the NCHW <-> channel-last layout change.
#include "ir_host.h"
#include "ir_ptx.cuh"

namespace ir {

// ------------------------------------------------------------------------------------------------ AdaIN
// Column statistics (over tokens) of token-major fp16 matrices. grid = (channels/64, slabs, (1 + n_ref) * batch),
// block = 256 = 32 row lanes x 8 sixteen-byte vectors: a CTA reads a [rows_per_slab x 64-channel] slab with coalesced
// 128-byte rows, 4 loads in flight per thread, and writes per-channel partial (mean, M2) moments:
//   ws[(((b * (1 + n_ref) + chunk) * slabs + slab) * channels + c)] = (mean, M2)
__global__ void __launch_bounds__(256) colstats_kernel(const __half* __restrict__ v_own, int own_stride, int own_col_off,
                                                       int s_own, const __half* __restrict__ v_ref, int ref_stride,
                                                       int ref_col_off, int n_ref, int s_ref, int channels, int slabs,
                                                       float2* __restrict__ ws) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[2][32][64 + 1];
  const int cb = blockIdx.x, slab = blockIdx.y;
  const int chunk = blockIdx.z % (1 + n_ref), b = blockIdx.z / (1 + n_ref);
  const int vec = threadIdx.x & 7, rl = threadIdx.x >> 3;
  const __half* base;
  int stride, rows;
  if (chunk == 0) {
    base = v_own + static_cast<size_t>(b) * s_own * own_stride + own_col_off + cb * 64;
    stride = own_stride;
    rows = s_own;
  } else {
    base = v_ref + (static_cast<size_t>(b) * n_ref + (chunk - 1)) * s_ref * ref_stride + ref_col_off + cb * 64;
    stride = ref_stride;
    rows = s_ref;
  }
  const int rps = (rows + slabs - 1) / slabs;
  const int r0 = slab * rps, r1 = min(r0 + rps, rows);
  float s[8], q[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) s[j] = q[j] = 0.f;
  int r = r0 + rl;
  for (; r + 96 < r1; r += 128) {
    uint4 u[4];
#pragma unroll
    for (int k = 0; k < 4; ++k) u[k] = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(r + 32 * k) * stride + (vec << 3));
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
  for (; r < r1; r += 32) {
    const uint4 u = *reinterpret_cast<const uint4*>(base + static_cast<size_t>(r) * stride + (vec << 3));
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      s[2 * j] += f.x; q[2 * j] = fmaf(f.x, f.x, q[2 * j]);
      s[2 * j + 1] += f.y; q[2 * j + 1] = fmaf(f.y, f.y, q[2 * j + 1]);
    }
  }
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    red[0][rl][vec * 8 + j] = s[j];
    red[1][rl][vec * 8 + j] = q[j];
  }
  __syncthreads();
  if (threadIdx.x < 64) {
    float ts = 0.f, tq = 0.f;
#pragma unroll
    for (int w = 0; w < 32; ++w) {
      ts += red[0][w][threadIdx.x];
      tq += red[1][w][threadIdx.x];
    }
    const float n = static_cast<float>(r1 - r0);
    const float mean = n > 0.f ? ts / n : 0.f;
    ws[((static_cast<size_t>(blockIdx.z) * slabs + slab) * channels) + cb * 64 + threadIdx.x] =
        make_float2(mean, fmaxf(tq - ts * mean, 0.f));
  }
}

// merges the slab moments of channel c of chunk `chunk_idx` (Chan), returns mean and the UNBIASED std
__device__ __forceinline__ void colstats_merge(const float2* __restrict__ ws, size_t chunk_idx, int slabs, int channels,
                                               int c, int rows, float* mean_out, float* std_out) {
  const int rps = (rows + slabs - 1) / slabs;
  float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
  for (int sl = 0; sl < slabs; ++sl) {
    const float n_b = static_cast<float>(min(rps, rows - sl * rps));
    if (n_b <= 0.f) break;
    const float2 pm = ws[(chunk_idx * slabs + sl) * channels + c];
    if (n_a == 0.f) {
      n_a = n_b; mean_a = pm.x; m2_a = pm.y;
    } else {
      const float n = n_a + n_b, d = pm.x - mean_a, f = n_b / n;
      mean_a = fmaf(d, f, mean_a);
      m2_a = m2_a + pm.y + d * d * n_a * f;
      n_a = n;
    }
  }
  *mean_out = mean_a;
  *std_out = rows > 1 ? sqrtf(m2_a / (rows - 1)) : 0.f;
}

__global__ void adain_finalize_kernel(const float2* __restrict__ ws, int n_ref, int channels, int slabs, int s_own,
                                      int s_ref, float eps, float* __restrict__ scale, float* __restrict__ shift,
                                      int total) {
  pdl_launch_dependents();
  pdl_wait();
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % channels;
  const int r = (i / channels) % n_ref;
  const int b = i / (channels * n_ref);
  float style_mean, style_std, cm, cs;
  colstats_merge(ws, static_cast<size_t>(b) * (1 + n_ref), slabs, channels, c, s_own, &style_mean, &style_std);
  colstats_merge(ws, static_cast<size_t>(b) * (1 + n_ref) + 1 + r, slabs, channels, c, s_ref, &cm, &cs);
  const float a = (style_std + eps) / (cs + eps);
  scale[i] = a;
  shift[i] = style_mean - cm * a;
}

// AdaIN affine from the 32-row slab moments the QKV GEMM epilogues wrote (ir_conv_gemm.col_partial): no pass over V.
// grid = (channels / 32, batch), block = (32 channels, `lanes` slab lanes, 1 + n_ref chunks): thread (c, l, k) merges
// every lanes-th slab of chunk k (k = 0: the image's own V = the style; k >= 1: reference k - 1 = the content) for its
// channel (Chan), lane 0 merges the lanes in fixed order, the style statistics go through shared memory, threads
// (c, 0, k >= 1) write scale / shift. The serial chain is slabs / lanes merges (32 for 4096 tokens).
//   own_partial [batch, s_own / 32, channels], ref_partial [batch, n_ref, s_ref / 32, channels]  (mean, M2)
__global__ void __launch_bounds__(1024) adain_from_partials_kernel(const float2* __restrict__ own_partial,
                                                                   const float2* __restrict__ ref_partial, int n_ref, int channels,
                                                                   int s_own, int s_ref, float eps, float* __restrict__ scale,
                                                                   float* __restrict__ shift) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float s_n[1024], s_mean[1024], s_m2[1024];
  __shared__ float2 chunk_stat[32][32];          // [chunk][channel] (mean, std)
  const int tx = threadIdx.x, lane = threadIdx.y, k = threadIdx.z, lanes = blockDim.y;
  const int c = blockIdx.x * 32 + tx, b = blockIdx.y;
  const int rows = k == 0 ? s_own : s_ref;
  const int slabs = rows >> 5;
  const float2* src = (k == 0 ? own_partial + static_cast<size_t>(b) * slabs * channels
                              : ref_partial + (static_cast<size_t>(b) * n_ref + (k - 1)) * slabs * channels) + c;
  float n_a = 0.f, mean_a = 0.f, m2_a = 0.f;
  auto merge = [&](float n_b, float mean_b, float m2_b) {
    if (n_b <= 0.f) return;
    if (n_a == 0.f) { n_a = n_b; mean_a = mean_b; m2_a = m2_b; return; }
    const float n = n_a + n_b, d = mean_b - mean_a, f = __fdividef(n_b, n);
    mean_a = fmaf(d, f, mean_a);
    m2_a = m2_a + m2_b + d * d * n_a * f;
    n_a = n;
  };
  int sl = lane;
  for (; sl + 3 * lanes < slabs; sl += 4 * lanes) {          // four loads in flight ahead of the merge chain
    float2 pm[4];
#pragma unroll
    for (int j = 0; j < 4; ++j) pm[j] = __ldg(src + static_cast<size_t>(sl + j * lanes) * channels);
#pragma unroll
    for (int j = 0; j < 4; ++j) merge(32.f, pm[j].x, pm[j].y);
  }
  for (; sl < slabs; sl += lanes) {
    const float2 pm = __ldg(src + static_cast<size_t>(sl) * channels);
    merge(32.f, pm.x, pm.y);
  }
  const int slot = (k * lanes + lane) * 32 + tx;
  s_n[slot] = n_a; s_mean[slot] = mean_a; s_m2[slot] = m2_a;
  __syncthreads();
  if (lane == 0) {
    for (int l = 1; l < lanes; ++l) {                         // fixed order: deterministic
      const int o = (k * lanes + l) * 32 + tx;
      merge(s_n[o], s_mean[o], s_m2[o]);
    }
    chunk_stat[k][tx] = make_float2(mean_a, rows > 1 ? sqrtf(m2_a / (rows - 1)) : 0.f);   // unbiased, like torch.std
  }
  __syncthreads();
  if (lane == 0 && k > 0) {
    const float2 st = chunk_stat[0][tx], ct = chunk_stat[k][tx];
    const float a = (st.y + eps) / (ct.y + eps);
    const size_t i = (static_cast<size_t>(b) * n_ref + (k - 1)) * channels + c;
    scale[i] = a;
    shift[i] = st.x - ct.x * a;
  }
}

// ------------------------------------------------------------------------------------------------ concat / FreeU
// out[:, :, 0:c_hidden] = hidden * (c < c_hidden/2 ? bscale : 1);  out[:, :, c_hidden:] = skip  (when copy_skip)
__global__ void __launch_bounds__(256) concat_kernel(const __half* __restrict__ hidden, const __half* __restrict__ skip,
                                                     int c_hidden, int c_skip, float bscale, int copy_skip,
                                                     __half* __restrict__ out, long rows) {
  pdl_launch_dependents();
  pdl_wait();
  const int c_tot = c_hidden + c_skip;
  const int vec_per_row = (copy_skip ? c_tot : c_hidden) >> 3;
  const long total = rows * vec_per_row;
  const int half_c = c_hidden >> 1;
  for (long v = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; v < total; v += static_cast<long>(gridDim.x) * blockDim.x) {
    const long row = v / vec_per_row;
    const int c0 = static_cast<int>(v - row * vec_per_row) << 3;
    uint4 u;
    if (c0 < c_hidden) {
      u = *reinterpret_cast<const uint4*>(hidden + row * c_hidden + c0);
      if (bscale != 1.0f && c0 < half_c) {
        __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __half22float2(h2[j]);
          // channels are processed 8 at a time; c_hidden/2 is a multiple of 8 for every supported width
          f.x *= bscale;
          f.y *= bscale;
          h2[j] = __floats2half2_rn(f.x, f.y);
        }
      }
    } else {
      u = *reinterpret_cast<const uint4*>(skip + row * c_skip + (c0 - c_hidden));
    }
    *reinterpret_cast<uint4*>(out + row * c_tot + c0) = u;
  }
}

// FreeU fourier_filter(threshold=1, scale=s) in closed form. The filter scales the 2x2 block of lowest
// frequencies {-1,0}x{-1,0} of the 2-D DFT by s, so
//   y[m,n] = x[m,n] - (1-s)/(HW) * Re( X00 + X(-1,0) e^{-2 pi i m/H} + X(0,-1) e^{-2 pi i n/W}
//                                     + X(-1,-1) e^{-2 pi i (m/H + n/W)} ),
//   X(u,v) = sum x[m,n] e^{-2 pi i (u m/H + v n/W)}.
// grid = (c_skip/32, batch); block = 256 = 8 pixel groups x 32 channels. fp32 throughout.
__global__ void __launch_bounds__(256) freeu_skip_kernel(const __half* __restrict__ skip, int h, int w, int c_skip,
                                                         int c_hidden, float s, __half* __restrict__ out) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8][32][7];
  __shared__ float coef[32][7];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int c = blockIdx.x * 32 + lane;
  const int b = blockIdx.y;
  const int hw = h * w;
  const int c_tot = c_hidden + c_skip;
  const __half* xp = skip + static_cast<size_t>(b) * hw * c_skip + c;
  // X00 (real), X(-1,0), X(0,-1), X(-1,-1) (complex)
  float a00 = 0.f, a10r = 0.f, a10i = 0.f, a01r = 0.f, a01i = 0.f, a11r = 0.f, a11i = 0.f;
  for (int px = warp; px < hw; px += 8) {
    const int m = px / w, n = px - m * w;
    const float x = __half2float(xp[static_cast<size_t>(px) * c_skip]);
    float sm, cm, sn, cn;
    sincospif(2.0f * m / h, &sm, &cm);   // e^{+2 pi i m/H} for u = -1
    sincospif(2.0f * n / w, &sn, &cn);
    a00 += x;
    a10r += x * cm;
    a10i += x * sm;
    a01r += x * cn;
    a01i += x * sn;
    a11r += x * (cm * cn - sm * sn);
    a11i += x * (sm * cn + cm * sn);
  }
  float* rr = red[warp][lane];
  rr[0] = a00; rr[1] = a10r; rr[2] = a10i; rr[3] = a01r; rr[4] = a01i; rr[5] = a11r; rr[6] = a11i;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += red[q][lane][k];
      coef[lane][k] = t;
    }
  }
  __syncthreads();
  const float k = (1.0f - s) / hw;
  const float* cf = coef[lane];
  __half* op = out + static_cast<size_t>(b) * hw * c_tot + c_hidden + c;
  for (int px = warp; px < hw; px += 8) {
    const int m = px / w, n = px - m * w;
    const float x = __half2float(xp[static_cast<size_t>(px) * c_skip]);
    float sm, cm, sn, cn;
    sincospif(2.0f * m / h, &sm, &cm);
    sincospif(2.0f * n / w, &sn, &cn);
    // Re( X * e^{-i t} ) = Xr cos t + Xi sin t
    const float corr = cf[0] + (cf[1] * cm + cf[2] * sm) + (cf[3] * cn + cf[4] * sn) +
                       (cf[5] * (cm * cn - sm * sn) + cf[6] * (sm * cn + cm * sn));
    op[static_cast<size_t>(px) * c_tot] = __float2half_rn(x - k * corr);
  }
}

// The same in ONE launch for the two FreeU stages of the step (8 x 8 and 16 x 16 skips): blocks [0, n_filter) filter 32 skip
// channels each with the thread's PPT = hw / 8 pixels held in registers (one read; every load in flight at once) and the
// twiddles in shared-memory tables (the two-pass kernel above evaluates two sincospif per pixel and pass); the remaining
// blocks copy / scale image b's hidden half. Same operations in the same order as concat_kernel + freeu_skip_kernel:
// results equal up to fp32 contraction (an fp16 ulp). 29.5 -> <= 14 us per call on the 16 x 16 x (1280 + 1280) stage (tools/misc_bench.py, host-launch bound).
template <int PPT>
__global__ void __launch_bounds__(256) freeu_concat_kernel(const __half* __restrict__ hidden, const __half* __restrict__ skip, int h, int w,
                                                           int c_skip, int c_hidden, float bscale, float s, __half* __restrict__ out,
                                                           int n_filter) {
  pdl_launch_dependents();
  pdl_wait();
  const int b = blockIdx.y;
  const int hw = h * w;
  const int c_tot = c_hidden + c_skip;
  if (static_cast<int>(blockIdx.x) >= n_filter) {
    const int nb = gridDim.x - n_filter, cb = blockIdx.x - n_filter;
    const int vec_per_row = c_hidden >> 3, total = hw * vec_per_row, half_c = c_hidden >> 1;
    const __half* hp = hidden + static_cast<size_t>(b) * hw * c_hidden;
    __half* op = out + static_cast<size_t>(b) * hw * c_tot;
    for (int v = cb * 256 + threadIdx.x; v < total; v += nb * 256) {
      const int row = v / vec_per_row;
      const int c0 = (v - row * vec_per_row) << 3;
      uint4 u = *reinterpret_cast<const uint4*>(hp + static_cast<size_t>(row) * c_hidden + c0);
      if (bscale != 1.0f && c0 < half_c) {
        __half2* h2 = reinterpret_cast<__half2*>(&u);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
          float2 f = __half22float2(h2[j]);
          f.x *= bscale;
          f.y *= bscale;
          h2[j] = __floats2half2_rn(f.x, f.y);
        }
      }
      *reinterpret_cast<uint4*>(op + static_cast<size_t>(row) * c_tot + c0) = u;
    }
    return;
  }
  __shared__ float red[8][32][7];
  __shared__ float coef[32][7];
  __shared__ float tab[4][64];       // cos, sin of 2 pi m / h; cos, sin of 2 pi n / w
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  if (static_cast<int>(threadIdx.x) < h) sincospif(2.0f * threadIdx.x / h, &tab[1][threadIdx.x], &tab[0][threadIdx.x]);
  else if (threadIdx.x >= 64 && static_cast<int>(threadIdx.x) - 64 < w) sincospif(2.0f * (threadIdx.x - 64) / w, &tab[3][threadIdx.x - 64], &tab[2][threadIdx.x - 64]);
  const int c = blockIdx.x * 32 + lane;
  const __half* xp = skip + static_cast<size_t>(b) * hw * c_skip + c;
  float x[PPT];
#pragma unroll
  for (int j = 0; j < PPT; ++j) x[j] = __half2float(xp[static_cast<size_t>(warp + 8 * j) * c_skip]);
  __syncthreads();
  float a00 = 0.f, a10r = 0.f, a10i = 0.f, a01r = 0.f, a01i = 0.f, a11r = 0.f, a11i = 0.f;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int px = warp + 8 * j;
    const int m = px / w, n = px - m * w;
    const float cm = tab[0][m], sm = tab[1][m], cn = tab[2][n], sn = tab[3][n];
    a00 += x[j];
    a10r += x[j] * cm;
    a10i += x[j] * sm;
    a01r += x[j] * cn;
    a01i += x[j] * sn;
    a11r += x[j] * (cm * cn - sm * sn);
    a11i += x[j] * (sm * cn + cm * sn);
  }
  float* rr = red[warp][lane];
  rr[0] = a00; rr[1] = a10r; rr[2] = a10i; rr[3] = a01r; rr[4] = a01i; rr[5] = a11r; rr[6] = a11i;
  __syncthreads();
  if (warp == 0) {
#pragma unroll
    for (int k = 0; k < 7; ++k) {
      float t = 0.f;
#pragma unroll
      for (int q = 0; q < 8; ++q) t += red[q][lane][k];
      coef[lane][k] = t;
    }
  }
  __syncthreads();
  const float k = (1.0f - s) / hw;
  const float* cf = coef[lane];
  __half* op = out + static_cast<size_t>(b) * hw * c_tot + c_hidden + c;
#pragma unroll
  for (int j = 0; j < PPT; ++j) {
    const int px = warp + 8 * j;
    const int m = px / w, n = px - m * w;
    const float cm = tab[0][m], sm = tab[1][m], cn = tab[2][n], sn = tab[3][n];
    const float corr = cf[0] + (cf[1] * cm + cf[2] * sm) + (cf[3] * cn + cf[4] * sn) +
                       (cf[5] * (cm * cn - sm * sn) + cf[6] * (sm * cn + cm * sn));
    op[static_cast<size_t>(px) * c_tot] = __float2half_rn(x[j] - k * corr);
  }
}

// ------------------------------------------------------------------------------------------------ upsample
__global__ void __launch_bounds__(256) upsample2x_kernel(const __half* __restrict__ x, __half* __restrict__ out, int h,
                                                         int w, int c, long total_vec) {
  pdl_launch_dependents();
  pdl_wait();
  const int vec_per_px = c >> 3;
  const int ow = 2 * w, oh = 2 * h;
  for (long v = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x; v < total_vec; v += static_cast<long>(gridDim.x) * blockDim.x) {
    const int cv = static_cast<int>(v % vec_per_px);
    long px = v / vec_per_px;
    const int ox = static_cast<int>(px % ow);
    px /= ow;
    const int oy = static_cast<int>(px % oh);
    const long b = px / oh;
    const uint4 u = *reinterpret_cast<const uint4*>(x + ((b * h + (oy >> 1)) * w + (ox >> 1)) * c + (cv << 3));
    *reinterpret_cast<uint4*>(out + ((b * oh + oy) * ow + ox) * static_cast<long>(c) + (cv << 3)) = u;
  }
}

// ------------------------------------------------------------------------------------------------ latents
__global__ void latent_in_kernel(const float* __restrict__ x, const float* __restrict__ noise, float a, float s,
                                 __half* __restrict__ out, int c, int hw, int c_pad, long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (i >= total) return;
  const int ch = static_cast<int>(i % c_pad);
  const long px = i / c_pad;  // b * hw + p
  float v = 0.f;
  if (ch < c) {
    const long b = px / hw, p = px - b * hw;
    const long src = (b * c + ch) * hw + p;
    v = a * x[src] + (noise ? s * noise[src] : 0.f);
  }
  out[i] = __float2half_rn(v);
}

__global__ void latent_out_kernel(const __half* __restrict__ eps, int eps_stride, const float* __restrict__ x,
                                  const float* __restrict__ noise, float a, float s, float* __restrict__ out, int c,
                                  int hw, long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;  // NCHW index
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const long bc = i / hw;
  const int ch = static_cast<int>(bc % c);
  const long b = bc / c;
  const float e = __half2float(eps[(b * hw + p) * eps_stride + ch]);
  const float xt = a * x[i] + (noise ? s * noise[i] : 0.f);
  out[i] = (xt - s * e) / a;
}

// ------------------------------------------------------------------------------------------------ VAE helpers
// In-place row softmax, one CTA (256 threads) per row; the row (<= 8192 fp16) is held in registers between passes.
__global__ void __launch_bounds__(256) softmax_rows_kernel(__half* __restrict__ x, int cols, int row_stride, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  __shared__ float bc;
  __half* xr = x + static_cast<size_t>(blockIdx.x) * row_stride;
  const int nvec = cols >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  constexpr int kMaxVec = 4;                       // 256 threads x 4 vectors x 8 = 8192 columns
  uint4 v[kMaxVec];
  float mx = -INFINITY;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nvec) {
      v[k] = *reinterpret_cast<const uint4*>(xr + (i << 3));
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        mx = fmaxf(mx, fmaxf(f.x, f.y));
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    bc = m;
  }
  __syncthreads();
  const float m = bc * scale_log2;
  float e[kMaxVec][8];
  float sum = 0.f;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nvec) {
      const __half2* h2 = reinterpret_cast<const __half2*>(&v[k]);
#pragma unroll
      for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(h2[j]);
        e[k][2 * j] = fast_exp2(fmaf(f.x, scale_log2, -m));
        e[k][2 * j + 1] = fast_exp2(fmaf(f.y, scale_log2, -m));
        sum += e[k][2 * j] + e[k][2 * j + 1];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    bc = 1.0f / t;
  }
  __syncthreads();
  const float inv = bc;
#pragma unroll
  for (int k = 0; k < kMaxVec; ++k) {
    const int i = threadIdx.x + k * 256;
    if (i < nvec)
      *reinterpret_cast<uint4*>(xr + (i << 3)) =
          make_uint4(pack_half2(e[k][0] * inv, e[k][1] * inv), pack_half2(e[k][2] * inv, e[k][3] * inv),
                     pack_half2(e[k][4] * inv, e[k][5] * inv), pack_half2(e[k][6] * inv, e[k][7] * inv));
  }
}

// Row softmax for rows longer than 8192 columns: three passes over global memory (max, sum, write).
__global__ void __launch_bounds__(256) softmax_rows_long_kernel(__half* __restrict__ x, int cols, int row_stride, float scale_log2) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ float red[8];
  __shared__ float bc;
  __half* xr = x + static_cast<size_t>(blockIdx.x) * row_stride;
  const int nvec = cols >> 3;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  float mx = -INFINITY;
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + (i << 3));
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      mx = fmaxf(mx, fmaxf(f.x, f.y));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, o));
  if (lane == 0) red[warp] = mx;
  __syncthreads();
  if (threadIdx.x == 0) {
    float m = red[0];
    for (int w = 1; w < 8; ++w) m = fmaxf(m, red[w]);
    bc = m;
  }
  __syncthreads();
  const float m = bc * scale_log2;
  float sum = 0.f;
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + (i << 3));
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      sum += fast_exp2(fmaf(f.x, scale_log2, -m)) + fast_exp2(fmaf(f.y, scale_log2, -m));
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sum += __shfl_xor_sync(0xffffffffu, sum, o);
  __syncthreads();
  if (lane == 0) red[warp] = sum;
  __syncthreads();
  if (threadIdx.x == 0) {
    float t = 0.f;
    for (int w = 0; w < 8; ++w) t += red[w];
    bc = 1.0f / t;
  }
  __syncthreads();
  const float inv = bc;
  for (int i = threadIdx.x; i < nvec; i += 256) {
    const uint4 u = *reinterpret_cast<const uint4*>(xr + (i << 3));
    const __half2* h2 = reinterpret_cast<const __half2*>(&u);
    float e[8];
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const float2 f = __half22float2(h2[j]);
      e[2 * j] = fast_exp2(fmaf(f.x, scale_log2, -m)) * inv;
      e[2 * j + 1] = fast_exp2(fmaf(f.y, scale_log2, -m)) * inv;
    }
    *reinterpret_cast<uint4*>(xr + (i << 3)) =
        make_uint4(pack_half2(e[0], e[1]), pack_half2(e[2], e[3]), pack_half2(e[4], e[5]), pack_half2(e[6], e[7]));
  }
}

template <typename T>
__global__ void image_in_kernel(const T* __restrict__ x, __half* __restrict__ out, int c, int hw, int c_pad, long total_px) {
  pdl_launch_dependents();
  pdl_wait();
  // one thread per pixel: reads c planes (coalesced across threads), writes c_pad halfs (16-byte stores)
  const long px = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  if (px >= total_px) return;
  const long b = px / hw, p = px - b * hw;
  __half* o = out + px * c_pad;
  for (int c0 = 0; c0 < c_pad; c0 += 8) {
    __half h[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
      const int ch = c0 + j;
      h[j] = ch < c ? __float2half_rn(static_cast<float>(x[(b * c + ch) * hw + p])) : __float2half_rn(0.f);
    }
    *reinterpret_cast<uint4*>(o + c0) = *reinterpret_cast<const uint4*>(h);
  }
}

// 3x3 patches of a few-channel NCHW image as ONE 64-wide K block per pixel: out[px, (ky*3+kx)*c + ch] = x[b, ch, y+ky-1, x+kx-1]
// (zero outside the image and for k >= 9c). The VAE's conv_in (3 -> 128 channels, reference models/model.py:17 /
// diffusers Encoder.conv_in) then is a K = 64 GEMM instead of a 3x3 convolution over 64 zero-padded channels (K = 576, 27
// useful): the same 27 products per output, a ninth of the tensor work and of the operand traffic of that layer.
// Eight threads per pixel, one 16-byte store each: a warp writes 512 contiguous bytes.
template <typename T>
__global__ void __launch_bounds__(256) image_patches_kernel(const T* __restrict__ x, __half* __restrict__ out, int c, int h, int w, long total_px) {
  pdl_launch_dependents();
  pdl_wait();
  const long t = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;
  const long px = t >> 3;
  if (px >= total_px) return;
  const int part = static_cast<int>(t & 7);
  const int hw = h * w;
  const long b = px / hw;
  const int p = static_cast<int>(px - b * hw);
  const int y = p / w, xx = p - y * w;
  const int kmax = 9 * c;
  __half v[8];
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int k = part * 8 + j;
    float f = 0.f;
    if (k < kmax) {
      const int tap = k / c, ch = k - tap * c;
      const int yy = y + tap / 3 - 1, xc = xx + tap % 3 - 1;
      if (yy >= 0 && yy < h && xc >= 0 && xc < w) f = static_cast<float>(x[(b * c + ch) * hw + yy * w + xc]);
    }
    v[j] = __float2half_rn(f);
  }
  *reinterpret_cast<uint4*>(out + px * 64 + part * 8) = *reinterpret_cast<const uint4*>(v);
}

// The same for C channels known at compile time (the RGB image): one thread per pixel, so the 9C loads of a warp are
// coalesced along the image row; the warp's 32 x 128 B of patches go through shared memory (144-byte rows: conflict-free
// 16-byte accesses) and leave as eight fully coalesced 512-byte stores.
template <typename T, int C>
__global__ void __launch_bounds__(256) image_patches_rows_kernel(const T* __restrict__ x, __half* __restrict__ out, int h, int w, long total_px) {
  pdl_launch_dependents();
  pdl_wait();
  __shared__ __align__(16) uint8_t tile[8][32 * 144];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const long px0 = (blockIdx.x * 8L + warp) * 32;        // first pixel of this warp (total_px % 32 == 0)
  if (px0 >= total_px) return;
  const long px = px0 + lane;
  const int hw = h * w;
  const long b = px / hw;
  const int p = static_cast<int>(px - b * hw);
  const int y = p / w, xx = p - y * w;
  constexpr int K = 9 * C, KP = (K + 7) / 8 * 8;
  __half v[KP];
#pragma unroll
  for (int k = 0; k < KP; ++k) {
    float f = 0.f;
    if (k < K) {
      const int tap = k / C, ch = k % C;
      const int yy = y + tap / 3 - 1, xc = xx + tap % 3 - 1;
      if (yy >= 0 && yy < h && xc >= 0 && xc < w) f = static_cast<float>(x[(b * C + ch) * hw + yy * w + xc]);
    }
    v[k] = __float2half_rn(f);
  }
  uint8_t* mine = tile[warp] + lane * 144;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    *reinterpret_cast<uint4*>(mine + q * 16) = q < KP / 8 ? *reinterpret_cast<const uint4*>(&v[q * 8]) : make_uint4(0, 0, 0, 0);
  __syncwarp();
  uint4* dst = reinterpret_cast<uint4*>(out + px0 * 64);
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    const int idx = j * 32 + lane;
    dst[idx] = *reinterpret_cast<const uint4*>(tile[warp] + (idx >> 3) * 144 + (idx & 7) * 16);
  }
}

template <typename T>
__global__ void image_out_kernel(const __half* __restrict__ y, int stride, float lo, float hi, T* __restrict__ out, int c,
                                 int hw, long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;   // NCHW index
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const long bc = i / hw;
  const int ch = static_cast<int>(bc % c);
  const long b = bc / c;
  const float v = __half2float(y[(b * hw + p) * stride + ch]);
  out[i] = static_cast<T>(fminf(fmaxf(v, lo), hi));
}

__global__ void vae_sample_kernel(const __half* __restrict__ mom, int stride, const float* __restrict__ eps, float scale,
                                  float* __restrict__ out, int c, int hw, long total) {
  pdl_launch_dependents();
  pdl_wait();
  const long i = blockIdx.x * static_cast<long>(blockDim.x) + threadIdx.x;   // NCHW index
  if (i >= total) return;
  const int p = static_cast<int>(i % hw);
  const long bc = i / hw;
  const int ch = static_cast<int>(bc % c);
  const long b = bc / c;
  const __half* m = mom + (b * hw + p) * stride;
  const float mean = __half2float(m[ch]);
  float v = mean;
  if (eps) {
    const float logvar = fminf(fmaxf(__half2float(m[c + ch]), -30.0f), 20.0f);
    v = fmaf(__expf(0.5f * logvar), eps[i], mean);
  }
  out[i] = v * scale;
}

static inline int grid_for(long total, int block, int cap = 148 * 16) {
  long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ir

static int adain_slabs(int batch, int n_ref, int channels, int rows) {
  const long base = static_cast<long>(channels / 64) * (1 + n_ref) * batch;
  long want = (592 + base - 1) / base;               // ~4 CTAs per SM in total
  const long max_slabs = rows / 32 > 0 ? rows / 32 : 1;
  if (want > max_slabs) want = max_slabs;
  if (want > 32) want = 32;
  if (want < 1) want = 1;
  return static_cast<int>(want);
}

extern "C" size_t ir_adain_workspace_bytes(int batch, int n_ref, int channels) {
  return static_cast<size_t>(batch) * (1 + n_ref) * 32 * channels * sizeof(float2);   // up to 32 slabs per chunk
}

extern "C" int ir_adain_coeffs(const ir_adain_coeffs_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->scale || !p->shift) return set_error(IR_ERR_ARG, "ir_adain_coeffs: NULL argument");
  if ((p->own_partial == nullptr) != (p->ref_partial == nullptr))
    return set_error(IR_ERR_ARG, "ir_adain_coeffs: own_partial and ref_partial must both be set or both NULL");
  const bool from_partials = p->own_partial != nullptr;
  if (!from_partials && (!p->v_own || !p->v_ref || !p->workspace)) return set_error(IR_ERR_ARG, "ir_adain_coeffs: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->channels % 64 != 0 || p->n_ref <= 0 || p->batch <= 0 || p->s_own <= 0 || p->s_ref <= 0)
    return set_error(IR_ERR_SHAPE, "ir_adain_coeffs: channels=%d n_ref=%d", p->channels, p->n_ref);
  if (from_partials) {
    // statistics already reduced to 32-row slabs by the GEMM epilogues that produced V: one small merge kernel
    if (p->s_own % 32 != 0 || p->s_ref % 32 != 0 || p->n_ref > 31)
      return set_error(IR_ERR_SHAPE, "ir_adain_coeffs: slab moments need s_own %% 32 == 0, s_ref %% 32 == 0, n_ref <= 31 (s_own=%d s_ref=%d n_ref=%d)", p->s_own, p->s_ref, p->n_ref);
    if ((reinterpret_cast<uintptr_t>(p->own_partial) | reinterpret_cast<uintptr_t>(p->ref_partial)) & 7)
      return set_error(IR_ERR_ALIGN, "ir_adain_coeffs: partial pointers must be 8-byte aligned");
    const int chunks = 1 + p->n_ref;
    const int lanes = chunks <= 8 ? 4 : (chunks <= 16 ? 2 : 1);
    IR_LAUNCH(adain_from_partials_kernel, dim3(p->channels / 32, p->batch), dim3(32, lanes, chunks), 0, static_cast<cudaStream_t>(stream_), 
        static_cast<const float2*>(p->own_partial), static_cast<const float2*>(p->ref_partial), p->n_ref, p->channels, p->s_own, p->s_ref,
        p->eps, p->scale, p->shift);
    IR_CUDA_LAUNCH_CHECK("adain_from_partials launch");
    return 0;
  }
  if ((p->own_row_stride | p->ref_row_stride | p->v_col_off | p->ref_col_off) & 7)
    return set_error(IR_ERR_ALIGN, "ir_adain_coeffs: strides/offsets must be multiples of 8 elements");
  if ((reinterpret_cast<uintptr_t>(p->v_own) | reinterpret_cast<uintptr_t>(p->v_ref)) & 15)
    return set_error(IR_ERR_ALIGN, "ir_adain_coeffs: pointers must be 16-byte aligned");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const int slabs = adain_slabs(p->batch, p->n_ref, p->channels, p->s_own < p->s_ref ? p->s_own : p->s_ref);
  dim3 grid(p->channels / 64, slabs, (1 + p->n_ref) * p->batch);
  IR_LAUNCH(colstats_kernel, grid, 256, 0, stream, static_cast<const __half*>(p->v_own), p->own_row_stride, p->v_col_off, p->s_own,
                                            static_cast<const __half*>(p->v_ref), p->ref_row_stride, p->ref_col_off, p->n_ref,
                                            p->s_ref, p->channels, slabs, static_cast<float2*>(p->workspace));
  IR_CUDA_LAUNCH_CHECK("colstats launch");
  const int total = p->batch * p->n_ref * p->channels;
  IR_LAUNCH(adain_finalize_kernel, (total + 127) / 128, 128, 0, stream, static_cast<const float2*>(p->workspace), p->n_ref, p->channels,
                                                                 slabs, p->s_own, p->s_ref, p->eps, p->scale, p->shift, total);
  IR_CUDA_LAUNCH_CHECK("adain_finalize launch");
  return 0;
}

extern "C" int ir_concat_freeu(const ir_concat_freeu_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->hidden || !p->skip || !p->out) return set_error(IR_ERR_ARG, "ir_concat_freeu: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->c_hidden % 16 != 0 || p->c_skip % 8 != 0 || p->batch <= 0 || p->h <= 0 || p->w <= 0)
    return set_error(IR_ERR_SHAPE, "ir_concat_freeu: c_hidden=%d c_skip=%d", p->c_hidden, p->c_skip);
  const bool filt = p->skip_scale != 1.0f;
  if (filt && p->c_skip % 32 != 0) return set_error(IR_ERR_SHAPE, "ir_concat_freeu: FreeU needs c_skip %% 32 == 0");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const long rows = static_cast<long>(p->batch) * p->h * p->w;
  const long total = rows * ((filt ? p->c_hidden : p->c_hidden + p->c_skip) >> 3);
  const int hw = p->h * p->w;
  if (filt && !p->two_pass && (hw == 64 || hw == 256) && p->h <= 64 && p->w <= 64) {     // the step's two FreeU stages: one launch
    const int n_filter = p->c_skip / 32;
    const long vec = static_cast<long>(hw) * (p->c_hidden >> 3);
    const int n_copy = static_cast<int>(vec / 1024 < 1 ? 1 : (vec / 1024 > 108 ? 108 : vec / 1024));
    dim3 grid(n_filter + n_copy, p->batch);
    if (hw == 64) {
      IR_LAUNCH(freeu_concat_kernel<8>, grid, 256, 0, stream, static_cast<const __half*>(p->hidden), static_cast<const __half*>(p->skip), p->h, p->w,
                p->c_skip, p->c_hidden, p->backbone_scale, p->skip_scale, static_cast<__half*>(p->out), n_filter);
    } else {
      IR_LAUNCH(freeu_concat_kernel<32>, grid, 256, 0, stream, static_cast<const __half*>(p->hidden), static_cast<const __half*>(p->skip), p->h, p->w,
                p->c_skip, p->c_hidden, p->backbone_scale, p->skip_scale, static_cast<__half*>(p->out), n_filter);
    }
    IR_CUDA_LAUNCH_CHECK("freeu_concat launch");
    return 0;
  }
  IR_LAUNCH(concat_kernel, grid_for(total, 256), 256, 0, stream, static_cast<const __half*>(p->hidden), static_cast<const __half*>(p->skip),
                                                          p->c_hidden, p->c_skip, p->backbone_scale, filt ? 0 : 1,
                                                          static_cast<__half*>(p->out), rows);
  IR_CUDA_LAUNCH_CHECK("concat launch");
  if (filt) {
    dim3 grid(p->c_skip / 32, p->batch);
    IR_LAUNCH(freeu_skip_kernel, grid, 256, 0, stream, static_cast<const __half*>(p->skip), p->h, p->w, p->c_skip, p->c_hidden,
                                                p->skip_scale, static_cast<__half*>(p->out));
    IR_CUDA_LAUNCH_CHECK("freeu_skip launch");
  }
  return 0;
}

extern "C" int ir_upsample_nearest2x(const void* x, void* out, int batch, int h, int w, int c, ir_stream_t stream_) {
  using namespace ir;
  if (!x || !out) return set_error(IR_ERR_ARG, "ir_upsample_nearest2x: NULL argument");
  if (int rc = check_arch()) return rc;
  if (c % 8 != 0 || batch <= 0 || h <= 0 || w <= 0) return set_error(IR_ERR_SHAPE, "ir_upsample_nearest2x: c=%d", c);
  const long total_vec = static_cast<long>(batch) * 4 * h * w * (c >> 3);
  IR_LAUNCH(upsample2x_kernel, grid_for(total_vec, 256), 256, 0, static_cast<cudaStream_t>(stream_), 
      static_cast<const __half*>(x), static_cast<__half*>(out), h, w, c, total_vec);
  IR_CUDA_LAUNCH_CHECK("upsample launch");
  return 0;
}

extern "C" int ir_latent_in(const float* x, const float* noise, float a, float s, void* out, int batch, int c, int hw,
                            int c_pad, ir_stream_t stream_) {
  using namespace ir;
  if (!x || !out) return set_error(IR_ERR_ARG, "ir_latent_in: NULL argument");
  if (int rc = check_arch()) return rc;
  if (c_pad < c || c_pad % 8 != 0) return set_error(IR_ERR_SHAPE, "ir_latent_in: c_pad=%d", c_pad);
  const long total = static_cast<long>(batch) * hw * c_pad;
  IR_LAUNCH(latent_in_kernel, static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_), 
      x, noise, a, s, static_cast<__half*>(out), c, hw, c_pad, total);
  IR_CUDA_LAUNCH_CHECK("latent_in launch");
  return 0;
}

extern "C" int ir_latent_out(const void* eps, int eps_row_stride, const float* x, const float* noise, float a, float s,
                             float* out, int batch, int c, int hw, ir_stream_t stream_) {
  using namespace ir;
  if (!eps || !x || !out || a == 0.f) return set_error(IR_ERR_ARG, "ir_latent_out: NULL argument");
  if (int rc = check_arch()) return rc;
  const long total = static_cast<long>(batch) * c * hw;
  IR_LAUNCH(latent_out_kernel, static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_), 
      static_cast<const __half*>(eps), eps_row_stride, x, noise, a, s, out, c, hw, total);
  IR_CUDA_LAUNCH_CHECK("latent_out launch");
  return 0;
}

extern "C" int ir_softmax_rows(void* x, int rows, int cols, int row_stride, float scale, ir_stream_t stream_) {
  using namespace ir;
  if (!x) return set_error(IR_ERR_ARG, "ir_softmax_rows: NULL argument");
  if (int rc = check_arch()) return rc;
  if (rows <= 0 || cols <= 0 || cols % 8 != 0 || row_stride % 8 != 0 || row_stride < cols)
    return set_error(IR_ERR_SHAPE, "ir_softmax_rows: rows=%d cols=%d stride=%d (cols %% 8 == 0)", rows, cols, row_stride);
  if (reinterpret_cast<uintptr_t>(x) & 15) return set_error(IR_ERR_ALIGN, "ir_softmax_rows: pointer not 16-byte aligned");
  if (cols <= 8192)
    IR_LAUNCH(softmax_rows_kernel, rows, 256, 0, static_cast<cudaStream_t>(stream_), static_cast<__half*>(x), cols, row_stride,
                                                                              scale * 1.4426950408889634f);
  else
    IR_LAUNCH(softmax_rows_long_kernel, rows, 256, 0, static_cast<cudaStream_t>(stream_), static_cast<__half*>(x), cols, row_stride,
                                                                                   scale * 1.4426950408889634f);
  IR_CUDA_LAUNCH_CHECK("softmax_rows launch");
  return 0;
}

extern "C" int ir_image_in(const void* x, int x_is_fp32, void* out, int batch, int c, int hw, int c_pad, ir_stream_t stream_) {
  using namespace ir;
  if (!x || !out) return set_error(IR_ERR_ARG, "ir_image_in: NULL argument");
  if (int rc = check_arch()) return rc;
  if (c_pad < c || c_pad % 8 != 0 || batch <= 0 || c <= 0 || hw <= 0) return set_error(IR_ERR_SHAPE, "ir_image_in: c=%d c_pad=%d", c, c_pad);
  const long total_px = static_cast<long>(batch) * hw;
  const int blocks = static_cast<int>((total_px + 255) / 256);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (x_is_fp32) IR_LAUNCH(image_in_kernel<float>, blocks, 256, 0, stream, static_cast<const float*>(x), static_cast<__half*>(out), c, hw, c_pad, total_px);
  else IR_LAUNCH(image_in_kernel<__half>, blocks, 256, 0, stream, static_cast<const __half*>(x), static_cast<__half*>(out), c, hw, c_pad, total_px);
  IR_CUDA_LAUNCH_CHECK("image_in launch");
  return 0;
}

extern "C" int ir_image_in_patches3x3(const void* x, int x_is_fp32, void* out, int batch, int c, int h, int w, ir_stream_t stream_) {
  using namespace ir;
  if (!x || !out) return set_error(IR_ERR_ARG, "ir_image_in_patches3x3: NULL argument");
  if (int rc = check_arch()) return rc;
  if (batch <= 0 || c <= 0 || 9 * c > 64 || h <= 0 || w <= 0) return set_error(IR_ERR_SHAPE, "ir_image_in_patches3x3: c=%d (9c must fit one 64-wide K block)", c);
  if (reinterpret_cast<uintptr_t>(out) & 15) return set_error(IR_ERR_ALIGN, "ir_image_in_patches3x3: out not 16-byte aligned");
  const long total_px = static_cast<long>(batch) * h * w;
  const long blocks = (total_px * 8 + 255) / 256;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (c == 3 && total_px % 32 == 0) {
    const unsigned rblocks = static_cast<unsigned>((total_px / 32 + 7) / 8);
    if (x_is_fp32) IR_LAUNCH((image_patches_rows_kernel<float, 3>), rblocks, 256, 0, stream, static_cast<const float*>(x), static_cast<__half*>(out), h, w, total_px);
    else IR_LAUNCH((image_patches_rows_kernel<__half, 3>), rblocks, 256, 0, stream, static_cast<const __half*>(x), static_cast<__half*>(out), h, w, total_px);
    IR_CUDA_LAUNCH_CHECK("image_patches_rows launch");
    return 0;
  }
  if (x_is_fp32) IR_LAUNCH(image_patches_kernel<float>, static_cast<unsigned>(blocks), 256, 0, stream, static_cast<const float*>(x), static_cast<__half*>(out), c, h, w, total_px);
  else IR_LAUNCH(image_patches_kernel<__half>, static_cast<unsigned>(blocks), 256, 0, stream, static_cast<const __half*>(x), static_cast<__half*>(out), c, h, w, total_px);
  IR_CUDA_LAUNCH_CHECK("image_patches launch");
  return 0;
}

extern "C" int ir_image_out(const void* y, int y_row_stride, float lo, float hi, void* out, int out_is_fp32, int batch, int c,
                            int hw, ir_stream_t stream_) {
  using namespace ir;
  if (!y || !out) return set_error(IR_ERR_ARG, "ir_image_out: NULL argument");
  if (int rc = check_arch()) return rc;
  if (y_row_stride < c || batch <= 0 || c <= 0 || hw <= 0) return set_error(IR_ERR_SHAPE, "ir_image_out: c=%d stride=%d", c, y_row_stride);
  const long total = static_cast<long>(batch) * c * hw;
  const int blocks = static_cast<int>((total + 255) / 256);
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  if (out_is_fp32) IR_LAUNCH(image_out_kernel<float>, blocks, 256, 0, stream, static_cast<const __half*>(y), y_row_stride, lo, hi, static_cast<float*>(out), c, hw, total);
  else IR_LAUNCH(image_out_kernel<__half>, blocks, 256, 0, stream, static_cast<const __half*>(y), y_row_stride, lo, hi, static_cast<__half*>(out), c, hw, total);
  IR_CUDA_LAUNCH_CHECK("image_out launch");
  return 0;
}

extern "C" int ir_vae_sample(const void* moments, int m_row_stride, const float* eps, float scale, float* out, int batch, int c,
                             int hw, ir_stream_t stream_) {
  using namespace ir;
  if (!moments || !out) return set_error(IR_ERR_ARG, "ir_vae_sample: NULL argument");
  if (int rc = check_arch()) return rc;
  if (m_row_stride < 2 * c || batch <= 0 || c <= 0 || hw <= 0) return set_error(IR_ERR_SHAPE, "ir_vae_sample: c=%d stride=%d", c, m_row_stride);
  const long total = static_cast<long>(batch) * c * hw;
  IR_LAUNCH(vae_sample_kernel, static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_), 
      static_cast<const __half*>(moments), m_row_stride, eps, scale, out, c, hw, total);
  IR_CUDA_LAUNCH_CHECK("vae_sample launch");
  return 0;
}
