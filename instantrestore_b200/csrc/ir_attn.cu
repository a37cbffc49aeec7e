// ir_shared_attn_fwd — fused shared-image attention for sm_100a (head_dim 64).
//
// CTA = one 128-query tile of one (batch, head). KV is streamed in 128-key tiles, chunk by chunk
// ([own tokens] ++ reference 0 ++ reference 1 ...), never concatenated in memory:
//   warp 4 (one lane): TMA producer — Q once, then K/V tiles straight out of the token-major projection outputs
//       (the head split is just the TMA column coordinate), 128B-swizzled, multi-stage ring.
//   warp 5 (one lane): tcgen05.mma issuer — S_j = Q K_j^T into one of two TMEM score buffers, then
//       O_{j-1} = P_{j-1} V_{j-1} (P from shared memory, V consumed MN-major so no transpose is needed);
//       QK of tile j+1 overlaps the softmax of tile j.
//   warps 0-3: online softmax, thread == query row: tcgen05.ld the scores, running max/sum in the exp2 domain,
//       P written to shared memory in the UMMA K-major swizzled layout, per-tile O read back from TMEM and
//       accumulated in registers with the AdaIN affine of the tile's reference:
//           sum_r P_r (a_r*V_r + b_r) = a_r*(P_r V_r) + b_r*rowsum(P_r).
//
// Roofline: tensor-core bound (AI ~ 3300 FLOP/B at S=4096, N=4); algorithmic FLOPs per launch =
// 4 * batch * heads * s_q * s_kv_total * 64.  At d=64 the MUFU exp rate caps tensor utilisation near 50%.
#include "ir_host.h"
#include "ir_ptx.cuh"

namespace ir {

constexpr int kQT = 128;       // query rows per CTA
constexpr int kKT = 128;       // keys per tile
constexpr int kD = 64;         // head dim
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB: Q, K or V tile
constexpr int kKVStages = 3;
constexpr int kMaxRef = 16;

struct AttnKParams {
  CUtensorMap tma_q, tma_k_own, tma_v_own, tma_k_ref, tma_v_ref;
  int q_col_off, k_own_col_off, v_own_col_off, ref_col_off;
  int has_own, own_shared, s_own;
  int n_ref, s_ref;
  int s_q, heads;
  float scale_log2;   // scale * log2(e)
  const float* adain_scale;
  const float* adain_shift;
  __half* out;
  int out_stride;
  float* chunk_mass;
  int n_chunks;
};

// shared-memory carve-up (bytes, from the 1024-aligned base)
constexpr int kOffQ = 0;
constexpr int kOffK = kOffQ + kTileBytes;
constexpr int kOffV = kOffK + kKVStages * kTileBytes;
constexpr int kOffP = kOffV + kKVStages * kTileBytes;       // 2 buffers x 32 KB
constexpr int kOffAdain = kOffP + 2 * 2 * kTileBytes;       // [kMaxRef][2][64] fp32
constexpr int kOffBar = kOffAdain + kMaxRef * 2 * 64 * 4;
constexpr int kAttnSmem = kOffBar + 256 + 1024;

__global__ void __launch_bounds__(192, 1) shared_attn_kernel(const __grid_constant__ AttnKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  uint8_t* sP = smem + kOffP;
  float* sAd = reinterpret_cast<float*>(smem + kOffAdain);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;                 // 1
  uint64_t* kv_full = bars + 1;            // kKVStages
  uint64_t* kv_empty = kv_full + kKVStages;
  uint64_t* s_full = kv_empty + kKVStages; // 2
  uint64_t* s_empty = s_full + 2;
  uint64_t* p_full = s_empty + 2;
  uint64_t* o_full = p_full + 2;
  uint64_t* o_empty = o_full + 2;
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(o_empty + 2);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z;

  const int own_tiles = p.has_own ? (p.s_own + kKT - 1) / kKT : 0;
  const int ref_tiles = p.n_ref > 0 ? (p.s_ref + kKT - 1) / kKT : 0;
  const int total_tiles = own_tiles + p.n_ref * ref_tiles;

  if (warp == 4 && lane == 0) {
    tma_prefetch_desc(&p.tma_q);
    if (p.has_own) { tma_prefetch_desc(&p.tma_k_own); tma_prefetch_desc(&p.tma_v_own); }
    if (p.n_ref) { tma_prefetch_desc(&p.tma_k_ref); tma_prefetch_desc(&p.tma_v_ref); }
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < 2; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&s_empty[i], 128);
      mbar_init(&p_full[i], 128);
      mbar_init(&o_full[i], 1);
      mbar_init(&o_empty[i], 128);
    }
    fence_mbar_init();
  }
  if (warp == 5) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  // stage the AdaIN affine of every reference for this (batch, head): sAd[r][0][d] = scale, sAd[r][1][d] = shift
  if (p.adain_scale != nullptr) {
    const int C = p.heads * kD;
    for (int i = threadIdx.x; i < p.n_ref * kD; i += blockDim.x) {
      const int r = i / kD, d = i % kD;
      const size_t g = (static_cast<size_t>(b) * p.n_ref + r) * C + head * kD + d;
      sAd[(r * 2 + 0) * kD + d] = p.adain_scale[g];
      sAd[(r * 2 + 1) * kD + d] = p.adain_shift[g];
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  const uint32_t tmem_S = tmem_base;         // 2 x 128 columns
  const uint32_t tmem_O = tmem_base + 256;   // 2 x 64 columns

  if (warp == 4) {
    // ------------------------------------------------------------------ TMA producer
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, kTileBytes);
      tma_load_3d(sQ, &p.tma_q, q_full, p.q_col_off + head * kD, qt * kQT, b);
      int j = 0;
      for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
        const bool is_own = p.has_own && chunk == 0;
        const int r = chunk - (p.has_own ? 1 : 0);
        const int tiles = is_own ? own_tiles : ref_tiles;
        for (int t = 0; t < tiles; ++t, ++j) {
          const int s = j % kKVStages;
          const uint32_t ph = (j / kKVStages) & 1;
          mbar_wait(&kv_empty[s], ph ^ 1);
          mbar_arrive_expect_tx(&kv_full[s], 2 * kTileBytes);
          if (is_own) {
            const int bb = p.own_shared ? 0 : b;
            tma_load_3d(sK + s * kTileBytes, &p.tma_k_own, &kv_full[s], p.k_own_col_off + head * kD, t * kKT, bb);
            tma_load_3d(sV + s * kTileBytes, &p.tma_v_own, &kv_full[s], p.v_own_col_off + head * kD, t * kKT, bb);
          } else {
            tma_load_4d(sK + s * kTileBytes, &p.tma_k_ref, &kv_full[s], p.ref_col_off + head * kD, t * kKT, r, b);
            tma_load_4d(sV + s * kTileBytes, &p.tma_v_ref, &kv_full[s], p.ref_col_off + head * kD, t * kKT, r, b);
          }
        }
      }
    }
    __syncwarp();
  } else if (warp == 5) {
    // ------------------------------------------------------------------ MMA issuer
    if (lane == 0) {
      constexpr uint32_t IDESC_QK = umma_idesc_f16(128, kKT, 0, 0);   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t IDESC_PV = umma_idesc_f16(128, kD, 0, 1);    // A = P (K-major), B = V (MN-major)
      const uint32_t q_base = smem_u32(sQ);
      auto issue_pv = [&](int i) {
        const int bi = i & 1, s = i % kKVStages;
        mbar_wait(&p_full[bi], (i >> 1) & 1);
        mbar_wait(&o_empty[bi], ((i >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t p_base = smem_u32(sP + bi * 2 * kTileBytes);
        const uint32_t v_base = smem_u32(sV + s * kTileBytes);
#pragma unroll
        for (int kk = 0; kk < kKT / 16; ++kk) {
          const uint64_t adesc = umma_smem_desc(p_base + (kk >> 2) * kTileBytes + (kk & 3) * 32, 16, 1024);
          const uint64_t bdesc = umma_smem_desc(v_base + kk * 2048, 1024, 1024);
          umma_f16_ss(tmem_O + bi * kD, adesc, bdesc, IDESC_PV, kk != 0 ? 1u : 0u);
        }
        umma_commit(&o_full[bi]);
        umma_commit(&kv_empty[s]);
      };
      mbar_wait(q_full, 0);
      for (int j = 0; j < total_tiles; ++j) {
        const int s = j % kKVStages, bj = j & 1;
        mbar_wait(&kv_full[s], (j / kKVStages) & 1);
        mbar_wait(&s_empty[bj], ((j >> 1) & 1) ^ 1);
        tc_fence_after();
        const uint32_t k_base = smem_u32(sK + s * kTileBytes);
#pragma unroll
        for (int k = 0; k < kD / 16; ++k) {
          const uint64_t adesc = umma_smem_desc(q_base + k * 32, 16, 1024);
          const uint64_t bdesc = umma_smem_desc(k_base + k * 32, 16, 1024);
          umma_f16_ss(tmem_S + bj * kKT, adesc, bdesc, IDESC_QK, k != 0 ? 1u : 0u);
        }
        umma_commit(&s_full[bj]);
        if (j >= 1) issue_pv(j - 1);
      }
      issue_pv(total_tiles - 1);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax / accumulate (thread == row)
    const int row = warp * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const bool use_adain = p.adain_scale != nullptr;
    float m_run = -INFINITY, l_run = 0.f;
    float acc[kD];
#pragma unroll
    for (int d = 0; d < kD; ++d) acc[d] = 0.f;
    float alpha_pend = 1.f, rs_pend = 0.f;
    int ref_pend = -1;

    auto consume = [&](int i, float alpha, float rs, int ref) {
      const int bi = i & 1;
      mbar_wait(&o_full[bi], (i >> 1) & 1);
      tc_fence_after();
      const float* ad = sAd + (ref < 0 ? 0 : ref) * 2 * kD;
      const bool affine = use_adain && ref >= 0;
#pragma unroll
      for (int c = 0; c < 2; ++c) {
        uint32_t r[32];
        tmem_ld32(tmem_O + bi * kD + c * 32 + lane_addr, r);
        tmem_ld_wait();
#pragma unroll
        for (int d = 0; d < 32; ++d) {
          float o = __uint_as_float(r[d]);
          if (affine) o = fmaf(ad[c * 32 + d], o, ad[kD + c * 32 + d] * rs);
          acc[c * 32 + d] = fmaf(acc[c * 32 + d], alpha, o);
        }
      }
      tc_fence_before();
      mbar_arrive(&o_empty[bi]);
    };

    int j = 0;
    for (int chunk = 0; chunk < p.n_chunks; ++chunk) {
      const bool is_own = p.has_own && chunk == 0;
      const int ref = is_own ? -1 : chunk - (p.has_own ? 1 : 0);
      const int tiles = is_own ? own_tiles : ref_tiles;
      const int len = is_own ? p.s_own : p.s_ref;
      for (int t = 0; t < tiles; ++t, ++j) {
        const int bj = j & 1;
        const int valid = min(kKT, len - t * kKT);   // keys of this tile that exist
        mbar_wait(&s_full[bj], (j >> 1) & 1);
        tc_fence_after();
        const uint32_t s_addr = tmem_S + bj * kKT + lane_addr;
        // pass 1: row max
        float mx = -INFINITY;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(s_addr + c * 32, r);
          tmem_ld_wait();
#pragma unroll
          for (int i = 0; i < 32; ++i) {
            const float v = __uint_as_float(r[i]);
            if (c * 32 + i < valid) mx = fmaxf(mx, v);
          }
        }
        const float m_new = fmaxf(m_run, mx * p.scale_log2);
        const float alpha = fast_exp2(m_run - m_new);
        // pass 2: P = exp2(s*scale_log2 - m_new), fp16, into the swizzled K-major A-operand layout
        float rowsum = 0.f;
        uint8_t* prow = sP + bj * 2 * kTileBytes + row * 128;
#pragma unroll 1
        for (int c = 0; c < 4; ++c) {
          uint32_t r[32];
          tmem_ld32(s_addr + c * 32, r);
          tmem_ld_wait();
          uint32_t pk[16];
#pragma unroll
          for (int i = 0; i < 32; i += 2) {
            float p0 = fast_exp2(fmaf(__uint_as_float(r[i]), p.scale_log2, -m_new));
            float p1 = fast_exp2(fmaf(__uint_as_float(r[i + 1]), p.scale_log2, -m_new));
            if (c * 32 + i >= valid) p0 = 0.f;
            if (c * 32 + i + 1 >= valid) p1 = 0.f;
            rowsum += p0 + p1;
            pk[i >> 1] = pack_half2(p0, p1);
          }
          // 32 columns = 4 x 16-byte chunks; chunk index within the 64-column atom: (c & 1) * 4 + q
          uint8_t* atom = prow + (c >> 1) * kTileBytes;
#pragma unroll
          for (int q = 0; q < 4; ++q) {
            const int kc = (c & 1) * 4 + q;
            *reinterpret_cast<uint4*>(atom + ((kc ^ (row & 7)) << 4)) =
                make_uint4(pk[4 * q], pk[4 * q + 1], pk[4 * q + 2], pk[4 * q + 3]);
          }
        }
        tc_fence_before();
        mbar_arrive(&s_empty[bj]);
        fence_proxy_async_smem();
        mbar_arrive(&p_full[bj]);

        l_run = fmaf(l_run, alpha, rowsum);
        m_run = m_new;
        if (j > 0) consume(j - 1, alpha_pend, rs_pend, ref_pend);
        alpha_pend = alpha;
        rs_pend = rowsum;
        ref_pend = ref;
      }
    }
    consume(total_tiles - 1, alpha_pend, rs_pend, ref_pend);

    const int qrow = qt * kQT + row;
    if (qrow < p.s_q) {
      const float inv = 1.0f / l_run;
      __half* op = p.out + (static_cast<size_t>(b) * p.s_q + qrow) * p.out_stride + head * kD;
#pragma unroll
      for (int q = 0; q < 8; ++q) {
        uint4 u = make_uint4(pack_half2(acc[q * 8] * inv, acc[q * 8 + 1] * inv),
                             pack_half2(acc[q * 8 + 2] * inv, acc[q * 8 + 3] * inv),
                             pack_half2(acc[q * 8 + 4] * inv, acc[q * 8 + 5] * inv),
                             pack_half2(acc[q * 8 + 6] * inv, acc[q * 8 + 7] * inv));
        reinterpret_cast<uint4*>(op)[q] = u;
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 5) tmem_dealloc(tmem_base, 512);
}

}  // namespace ir

extern "C" int ir_shared_attn_fwd(const ir_shared_attn_params* p, ir_stream_t stream_) {
  using namespace ir;
  if (!p || !p->q || !p->out) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: NULL argument");
  if (int rc = check_arch()) return rc;
  cudaStream_t stream = static_cast<cudaStream_t>(stream_);
  const bool has_own = p->k_own != nullptr;
  if (has_own != (p->v_own != nullptr)) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: k_own/v_own must both be set or both NULL");
  if (p->n_ref < 0 || p->n_ref > kMaxRef) return set_error(IR_ERR_SHAPE, "ir_shared_attn_fwd: n_ref=%d (max %d)", p->n_ref, kMaxRef);
  if (p->n_ref > 0 && (!p->k_ref || !p->v_ref || p->s_ref <= 0)) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: reference K/V missing");
  if (!has_own && p->n_ref == 0) return set_error(IR_ERR_SHAPE, "ir_shared_attn_fwd: no keys");
  if (has_own && p->s_own <= 0) return set_error(IR_ERR_SHAPE, "ir_shared_attn_fwd: s_own=%d", p->s_own);
  if (p->batch <= 0 || p->heads <= 0 || p->s_q <= 0) return set_error(IR_ERR_SHAPE, "ir_shared_attn_fwd: non-positive dims");
  if ((p->adain_scale == nullptr) != (p->adain_shift == nullptr)) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: adain scale/shift mismatch");
  if (p->chunk_mass) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: chunk_mass output not implemented yet");
  if (p->q_row_stride < p->q_col_off + p->heads * kD || (has_own && p->own_row_stride < p->heads * kD) ||
      (p->n_ref && p->ref_row_stride < p->ref_col_off + p->heads * kD))
    return set_error(IR_ERR_SHAPE, "ir_shared_attn_fwd: row stride smaller than heads*64 columns");
  if (p->q_row_stride % 8 || p->out_row_stride % 8 || (has_own && p->own_row_stride % 8) || (p->n_ref && p->ref_row_stride % 8))
    return set_error(IR_ERR_ALIGN, "ir_shared_attn_fwd: row strides must be multiples of 8 elements");
  if ((p->q_col_off | p->k_own_col_off | p->v_own_col_off | p->ref_col_off) % 8)
    return set_error(IR_ERR_ALIGN, "ir_shared_attn_fwd: column offsets must be multiples of 8");
  if (reinterpret_cast<uintptr_t>(p->out) & 15) return set_error(IR_ERR_ALIGN, "ir_shared_attn_fwd: out not 16-byte aligned");

  AttnKParams kp;
  memset(&kp, 0, sizeof(kp));
  uint32_t box3[3] = {64, 128, 1};
  {
    uint64_t dims[3] = {static_cast<uint64_t>(p->q_col_off + p->heads * kD), static_cast<uint64_t>(p->s_q), static_cast<uint64_t>(p->batch)};
    uint64_t str[2] = {static_cast<uint64_t>(p->q_row_stride) * 2, static_cast<uint64_t>(p->q_row_stride) * 2 * p->s_q};
    if (int rc = make_tmap_f16(&kp.tma_q, p->q, 3, dims, str, box3)) return rc;
  }
  if (has_own) {
    const int nb = p->own_shared ? 1 : p->batch;
    uint64_t dims[3] = {static_cast<uint64_t>(p->k_own_col_off + p->heads * kD), static_cast<uint64_t>(p->s_own), static_cast<uint64_t>(nb)};
    uint64_t str[2] = {static_cast<uint64_t>(p->own_row_stride) * 2, static_cast<uint64_t>(p->own_row_stride) * 2 * p->s_own};
    if (int rc = make_tmap_f16(&kp.tma_k_own, p->k_own, 3, dims, str, box3)) return rc;
    dims[0] = static_cast<uint64_t>(p->v_own_col_off + p->heads * kD);
    if (int rc = make_tmap_f16(&kp.tma_v_own, p->v_own, 3, dims, str, box3)) return rc;
  }
  if (p->n_ref > 0) {
    uint32_t box4[4] = {64, 128, 1, 1};
    const uint64_t rs = static_cast<uint64_t>(p->ref_row_stride) * 2;
    uint64_t dims[4] = {static_cast<uint64_t>(p->ref_col_off + p->heads * kD), static_cast<uint64_t>(p->s_ref), static_cast<uint64_t>(p->n_ref),
                        static_cast<uint64_t>(p->batch)};
    uint64_t str[3] = {rs, rs * p->s_ref, rs * p->s_ref * p->n_ref};
    if (int rc = make_tmap_f16(&kp.tma_k_ref, p->k_ref, 4, dims, str, box4)) return rc;
    if (int rc = make_tmap_f16(&kp.tma_v_ref, p->v_ref, 4, dims, str, box4)) return rc;
  }
  kp.q_col_off = p->q_col_off;
  kp.k_own_col_off = p->k_own_col_off;
  kp.v_own_col_off = p->v_own_col_off;
  kp.ref_col_off = p->ref_col_off;
  kp.has_own = has_own ? 1 : 0;
  kp.own_shared = p->own_shared;
  kp.s_own = p->s_own;
  kp.n_ref = p->n_ref;
  kp.s_ref = p->s_ref;
  kp.s_q = p->s_q;
  kp.heads = p->heads;
  kp.scale_log2 = p->scale * 1.4426950408889634f;
  kp.adain_scale = p->n_ref > 0 ? p->adain_scale : nullptr;
  kp.adain_shift = p->n_ref > 0 ? p->adain_shift : nullptr;
  kp.out = static_cast<__half*>(p->out);
  kp.out_stride = p->out_row_stride;
  kp.chunk_mass = nullptr;
  kp.n_chunks = (has_own ? 1 : 0) + p->n_ref;

  static bool attr_done = false;
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(shared_attn_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaFuncSetAttribute(shared_attn): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid((p->s_q + kQT - 1) / kQT, p->heads, p->batch);
  shared_attn_kernel<<<grid, 192, kAttnSmem, stream>>>(kp);
  IR_CUDA_LAUNCH_CHECK("shared_attn launch");
  return 0;
}
