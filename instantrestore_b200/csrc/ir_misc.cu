// HBM-bound helpers of the hot path: AdaIN statistics -> affine, FreeU (closed form) + concat, nearest 2x
// upsample, scheduler AXPYs fused with the NCHW <-> channel-last layout change.
#include "ir_host.h"
#include "ir_ptx.cuh"

namespace ir {

// ------------------------------------------------------------------------------------------------ AdaIN
// Column statistics (over tokens) of a token-major fp16 matrix slab: 64 channels per CTA, two passes
// (mean, then centred sum of squares; unbiased). grid = (channels/64, 1 + n_ref, batch), block = 256.
// ws[((b * (1 + n_ref) + chunk) * channels + c) * 2 + {0,1}] = {mean, std_unbiased}
__global__ void __launch_bounds__(256) colstats_kernel(const __half* __restrict__ v_own, int own_stride, int own_col_off,
                                                       int s_own, const __half* __restrict__ v_ref, int ref_stride,
                                                       int ref_col_off, int n_ref, int s_ref, int channels,
                                                       float* __restrict__ ws) {
  __shared__ float red[8][64];
  __shared__ float mean_s[64];
  const int cb = blockIdx.x, chunk = blockIdx.y, b = blockIdx.z;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
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
  // pass 1: mean
  float sx = 0.f, sy = 0.f;
  for (int r = warp; r < rows; r += 8) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(base + static_cast<size_t>(r) * stride + 2 * lane));
    sx += f.x;
    sy += f.y;
  }
  red[warp][2 * lane] = sx;
  red[warp][2 * lane + 1] = sy;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    mean_s[threadIdx.x] = t / rows;
  }
  __syncthreads();
  const float mx = mean_s[2 * lane], my = mean_s[2 * lane + 1];
  float qx = 0.f, qy = 0.f;
  for (int r = warp; r < rows; r += 8) {
    const float2 f = __half22float2(*reinterpret_cast<const __half2*>(base + static_cast<size_t>(r) * stride + 2 * lane));
    qx += (f.x - mx) * (f.x - mx);
    qy += (f.y - my) * (f.y - my);
  }
  __syncthreads();
  red[warp][2 * lane] = qx;
  red[warp][2 * lane + 1] = qy;
  __syncthreads();
  if (threadIdx.x < 64) {
    float t = 0.f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const float var = rows > 1 ? t / (rows - 1) : 0.f;
    const size_t o = ((static_cast<size_t>(b) * (1 + n_ref) + chunk) * channels + cb * 64 + threadIdx.x) * 2;
    ws[o] = mean_s[threadIdx.x];
    ws[o + 1] = sqrtf(var);
  }
}

__global__ void adain_finalize_kernel(const float* __restrict__ ws, int n_ref, int channels, float eps,
                                      float* __restrict__ scale, float* __restrict__ shift, int total) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const int c = i % channels;
  const int r = (i / channels) % n_ref;
  const int b = i / (channels * n_ref);
  const size_t so = ((static_cast<size_t>(b) * (1 + n_ref)) * channels + c) * 2;
  const size_t ro = ((static_cast<size_t>(b) * (1 + n_ref) + 1 + r) * channels + c) * 2;
  const float style_mean = ws[so], style_std = ws[so + 1] + eps;
  const float cm = ws[ro], cs = ws[ro + 1] + eps;
  const float a = style_std / cs;
  scale[i] = a;
  shift[i] = style_mean - cm * a;
}

// ------------------------------------------------------------------------------------------------ concat / FreeU
// out[:, :, 0:c_hidden] = hidden * (c < c_hidden/2 ? bscale : 1);  out[:, :, c_hidden:] = skip  (when copy_skip)
__global__ void __launch_bounds__(256) concat_kernel(const __half* __restrict__ hidden, const __half* __restrict__ skip,
                                                     int c_hidden, int c_skip, float bscale, int copy_skip,
                                                     __half* __restrict__ out, long rows) {
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

// ------------------------------------------------------------------------------------------------ upsample
__global__ void __launch_bounds__(256) upsample2x_kernel(const __half* __restrict__ x, __half* __restrict__ out, int h,
                                                         int w, int c, long total_vec) {
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

static inline int grid_for(long total, int block, int cap = 148 * 16) {
  long g = (total + block - 1) / block;
  if (g > cap) g = cap;
  if (g < 1) g = 1;
  return static_cast<int>(g);
}

}  // namespace ir

extern "C" size_t ir_adain_workspace_bytes(int batch, int n_ref, int channels) {
  return static_cast<size_t>(batch) * (1 + n_ref) * channels * 2 * sizeof(float);
}

extern "C" int ir_adain_coeffs(const ir_adain_coeffs_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->v_own || !p->v_ref || !p->scale || !p->shift || !p->workspace) return set_error(IR_ERR_ARG, "ir_adain_coeffs: NULL argument");
  if (int rc = check_arch()) return rc;
  if (p->channels % 64 != 0 || p->n_ref <= 0 || p->batch <= 0 || p->s_own <= 0 || p->s_ref <= 0)
    return set_error(IR_ERR_SHAPE, "ir_adain_coeffs: channels=%d n_ref=%d", p->channels, p->n_ref);
  if ((p->own_row_stride | p->ref_row_stride | p->v_col_off | p->ref_col_off) & 1)
    return set_error(IR_ERR_ALIGN, "ir_adain_coeffs: strides/offsets must be even");
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  dim3 grid(p->channels / 64, 1 + p->n_ref, p->batch);
  colstats_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(p->v_own), p->own_row_stride, p->v_col_off, p->s_own,
                                            static_cast<const __half*>(p->v_ref), p->ref_row_stride, p->ref_col_off, p->n_ref,
                                            p->s_ref, p->channels, static_cast<float*>(p->workspace));
  IR_CUDA_LAUNCH_CHECK("colstats launch");
  const int total = p->batch * p->n_ref * p->channels;
  adain_finalize_kernel<<<(total + 255) / 256, 256, 0, stream>>>(static_cast<const float*>(p->workspace), p->n_ref, p->channels,
                                                                 p->eps, p->scale, p->shift, total);
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
  concat_kernel<<<grid_for(total, 256), 256, 0, stream>>>(static_cast<const __half*>(p->hidden), static_cast<const __half*>(p->skip),
                                                          p->c_hidden, p->c_skip, p->backbone_scale, filt ? 0 : 1,
                                                          static_cast<__half*>(p->out), rows);
  IR_CUDA_LAUNCH_CHECK("concat launch");
  if (filt) {
    dim3 grid(p->c_skip / 32, p->batch);
    freeu_skip_kernel<<<grid, 256, 0, stream>>>(static_cast<const __half*>(p->skip), p->h, p->w, p->c_skip, p->c_hidden,
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
  upsample2x_kernel<<<grid_for(total_vec, 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
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
  latent_in_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
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
  latent_out_kernel<<<static_cast<int>((total + 255) / 256), 256, 0, static_cast<cudaStream_t>(stream_)>>>(
      static_cast<const __half*>(eps), eps_row_stride, x, noise, a, s, out, c, hw, total);
  IR_CUDA_LAUNCH_CHECK("latent_out launch");
  return 0;
}
