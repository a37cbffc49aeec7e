// ir_shared_attn_fwd — fused shared-image attention for sm_100a (head_dim 64).
//
// A CTA owns TWO 128-query tiles of one (batch, head) and streams KV in 128-key tiles, chunk by chunk
// ([own tokens] ++ reference 0 ++ reference 1 ...), never concatenated in memory. 10 warps:
//   warps 0-3 / 4-7: softmax warpgroup of query tile 0 / 1, thread == query row. One tcgen05.ld pass brings the 128
//       scores of the row into registers, row max, exp2, and P goes back as packed fp16 INTO THE SAME TMEM COLUMNS
//       (tcgen05.st), from where the PV MMA reads it as its A operand: P never touches shared memory, which keeps the
//       128 B/clk shared-memory port for the K/V operand reads. The two warpgroups run out of phase, so the MUFU pipe
//       (the bound at d=64: 16 exp/clk/SM vs 8192 MMA FLOP/clk/SM) always has work while the tensor pipe runs the
//       other tile's QK^T / PV.
//   Score buffers ROTATE between the two query tiles: slot k = 2 * (KV tile) + (query tile) uses buffer k % NB. The
//       AdaIN variant needs the TMEM for its second accumulator (NB = 2: every query tile owns one buffer); the plain
//       variant has 128 free columns (NB = 3), so QK^T of slot k + 3 is issued right after PV of slot k -- one slot
//       (half a period) EARLIER than with two buffers -- and the round trip softmax -> P -> PV -> QK^T -> scores (about
//       800 clk of issue, MMA and commit latency, 21 % of the stall samples of the two-buffer kernel) is hidden.
//   warp 8 (one lane): TMA producer — both Q tiles once, then K/V tiles straight out of the token-major projection
//       outputs (the head split is just the TMA column coordinate), 128B-swizzled, 4-stage ring.
//   warp 9 (one lane): tcgen05.mma issuer — S = Q_i K_j^T (TMEM, 128 columns per buffer), O_i += P_i V_j (A = P from
//       TMEM, B = V consumed MN-major, no transpose), issued as PV of slot k, QK^T of slot k + NB back to back so a score
//       buffer is refilled as soon as its P has been consumed; the K tiles of the QK^Ts to come are waited for at the top
//       of the iteration, off the chain "P published -> PV -> QK^T -> scores" (every instruction between the P barrier and
//       the QK^T issue is exposed on the two-buffer variant: hoisting that wait was worth 3 %).
//       O accumulates IN TMEM across the KV tiles of a segment; the running max is only
//       raised when it grows by more than 2^8 (lazy rescale, done in TMEM by the softmax warps), so the common tile
//       costs no accumulator traffic at all.
// AdaIN: sum_r P_r (a_r*V_r + b_r) = a_r*(P_r V_r) + b_r*rowsum(P_r) — a segment is one reference chunk; at its end
//   the segment accumulator is folded with the chunk's affine into a second TMEM accumulator.
// Split-KV (few (batch, head, query) units, e.g. one identity): gridDim.x also enumerates KV ranges; CTAs write
//   un-normalised partials (O, m, l) and a small combine kernel merges them in a fixed order.
//
// Roofline: tensor-core bound by FLOPs (AI ~ 3300 FLOP/B at S=4096, N=4) but capped by the exp rate: 2 x 128 x 128
// exps per KV step = 2048 MUFU cycles vs 1024 MMA cycles, i.e. <= 50 % of the tcgen05 peak without exp emulation.
// Algorithmic FLOPs per launch = 4 * batch * heads * s_q * s_kv_total * 64.
#include "ir_host.h"
#include "ir_ptx.cuh"

namespace ir {

constexpr int kQT = 128;                  // query rows per tile (two tiles per CTA)
constexpr int kKT = 128;                  // keys per tile
constexpr int kD = 64;                    // head dim
constexpr int kTileBytes = 128 * 64 * 2;  // 16 KB: a Q, K or V tile, or one 64-key half of a P tile
constexpr int kKVStages = 4;
constexpr int kMaxRef = 16;
constexpr float kRescaleThreshold = 8.0f;  // log2 units: P stays <= 2^8, exact in fp16 / fp32 accumulation
constexpr int kAttnThreads = 320;
#ifndef IR_ATTN_POLY_OF8
#define IR_ATTN_POLY_OF8 2
#endif
constexpr int kPolyOf8 = IR_ATTN_POLY_OF8;  // of every 8 exponentials, this many run on the FMA pipes instead of MUFU

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
  int own_tiles, ref_tiles, total_tiles;
  int n_splits, tiles_per_split;
  float2* mass_ws;    // [batch, heads, q tiles * 128, n_chunks + 1] (row sum, reference max) per chunk, then the row total
  int n_chunks;
  float* part_o;      // [batch, heads, q tiles, n_splits, 128, 64] fp32 un-normalised partial outputs
  float2* part_ml;    // [batch, heads, q tiles, n_splits, 128] (running max in log2 units, row sum)
};

// shared-memory carve-up (bytes, from the 1024-aligned base)
constexpr int kOffQ = 0;                                    // 2 tiles
constexpr int kOffK = kOffQ + 2 * kTileBytes;
constexpr int kOffV = kOffK + kKVStages * kTileBytes;
constexpr int kOffAdain = kOffV + kKVStages * kTileBytes;   // [kMaxRef][2][64] fp32
constexpr int kOffBar = kOffAdain + kMaxRef * 2 * 64 * 4;
constexpr int kAttnSmem = kOffBar + 256 + 1024;

// TMEM columns: AdaIN  S(buf 0) | S(buf 1) | O0 | O1 | ACC0 | ACC1
//               plain  S(buf 0) | S(buf 1) | S(buf 2) | O0 | O1        (P = packed fp16, aliases the first 64 columns of its S)
constexpr uint32_t kTmemS = 0, kTmemAcc = 384;
template <bool ADAIN> constexpr int kNBuf = ADAIN ? 2 : 3;
template <bool ADAIN> constexpr uint32_t kTmemO = kNBuf<ADAIN> * kKT;

struct TileRef {
  int chunk;   // 0 = own (when present), else 1 + reference index (or reference index when no own chunk)
  int ref;     // -1 for the own chunk
  int t;       // tile index inside the chunk
};

__device__ __forceinline__ TileRef locate_tile(const AttnKParams& p, int g) {
  TileRef r;
  if (g < p.own_tiles) {
    r.chunk = 0; r.ref = -1; r.t = g;
  } else {
    const int gg = g - p.own_tiles;
    r.ref = gg / p.ref_tiles;
    r.t = gg - r.ref * p.ref_tiles;
    r.chunk = r.ref + (p.has_own ? 1 : 0);
  }
  return r;
}

__device__ __forceinline__ void tmem_st32(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16, "
      "%17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31, %32};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15]), "r"(r[16]), "r"(r[17]), "r"(r[18]),
      "r"(r[19]), "r"(r[20]), "r"(r[21]), "r"(r[22]), "r"(r[23]), "r"(r[24]), "r"(r[25]), "r"(r[26]), "r"(r[27]),
      "r"(r[28]), "r"(r[29]), "r"(r[30]), "r"(r[31])
      : "memory");
}
__device__ __forceinline__ void tmem_st16(uint32_t taddr, const uint32_t* r) {
  asm volatile(
      "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
      "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
      "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
      "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
      : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// multiplies 64 fp32 TMEM columns of this thread's lane by f
__device__ __forceinline__ void tmem_scale64(uint32_t taddr, float f) {
#pragma unroll
  for (int c = 0; c < 2; ++c) {
    uint32_t r[32];
    tmem_ld32(taddr + c * 32, r);
    tmem_ld_wait();
#pragma unroll
    for (int i = 0; i < 32; ++i) r[i] = __float_as_uint(__uint_as_float(r[i]) * f);
    tmem_st32(taddr + c * 32, r);
  }
}

// Row max over the 128 scores of this thread's row (two 64-column TMEM reads). MASKED: keys >= valid do not exist.
template <bool MASKED>
__device__ __forceinline__ float row_max128(uint32_t t_S, int valid) {
  float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
  for (int h = 0; h < 2; ++h) {
    uint32_t sr[64];
    tmem_ld32(t_S + h * 64, sr);
    tmem_ld32(t_S + h * 64 + 32, sr + 32);
    tmem_ld_wait();
#pragma unroll
    for (int k = 0; k < 64; k += 4) {
      float a = __uint_as_float(sr[k]), b = __uint_as_float(sr[k + 1]);
      float cc = __uint_as_float(sr[k + 2]), d = __uint_as_float(sr[k + 3]);
      if (MASKED) {
        if (h * 64 + k >= valid) a = -INFINITY;
        if (h * 64 + k + 1 >= valid) b = -INFINITY;
        if (h * 64 + k + 2 >= valid) cc = -INFINITY;
        if (h * 64 + k + 3 >= valid) d = -INFINITY;
      }
      mx0 = fmaxf(mx0, a); mx1 = fmaxf(mx1, b); mx2 = fmaxf(mx2, cc); mx3 = fmaxf(mx3, d);
    }
  }
  return fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
}

// P = exp2(s * c - m_ref) as packed fp16 (keys 2k, 2k+1 in TMEM column k), written over the scores: chunk q reads
// score columns [32q, 32q+32) and writes P columns [16q, 16q+16), always behind the read pointer. kPolyOf8 of every
// 8 exponentials run on the FMA pipes. Returns the row sum of the (unrounded) exponentials.
template <bool MASKED>
__device__ __forceinline__ float exp_pass128(uint32_t t_S, float c, float neg_m, int valid) {
  float rs0 = 0.f, rs1 = 0.f, rs2 = 0.f, rs3 = 0.f;
  uint32_t sa[32], sb[32];
  tmem_ld32(t_S, sa);
#pragma unroll
  for (int q = 0; q < 4; ++q) {
    uint32_t* cur = (q & 1) ? sb : sa;
    uint32_t* nxt = (q & 1) ? sa : sb;
    tmem_ld_wait();
    if (q < 3) tmem_ld32(t_S + (q + 1) * 32, nxt);     // prefetch the next chunk while this one is exponentiated
    uint32_t pk[16];
#pragma unroll
    for (int k = 0; k < 32; k += 4) {
      const float x0 = fmaf(__uint_as_float(cur[k]), c, neg_m), x1 = fmaf(__uint_as_float(cur[k + 1]), c, neg_m);
      const float x2 = fmaf(__uint_as_float(cur[k + 2]), c, neg_m), x3 = fmaf(__uint_as_float(cur[k + 3]), c, neg_m);
      // (k & 4): elements 4..7 of every 8; the first kPolyOf8 of those take the polynomial path
      float p0 = ((k & 4) && kPolyOf8 > 0) ? poly_exp2(x0) : fast_exp2(x0);
      float p1 = ((k & 4) && kPolyOf8 > 1) ? poly_exp2(x1) : fast_exp2(x1);
      float p2 = ((k & 4) && kPolyOf8 > 2) ? poly_exp2(x2) : fast_exp2(x2);
      float p3 = ((k & 4) && kPolyOf8 > 3) ? poly_exp2(x3) : fast_exp2(x3);
      if (MASKED) {
        if (q * 32 + k >= valid) p0 = 0.f;
        if (q * 32 + k + 1 >= valid) p1 = 0.f;
        if (q * 32 + k + 2 >= valid) p2 = 0.f;
        if (q * 32 + k + 3 >= valid) p3 = 0.f;
      }
      rs0 += p0; rs1 += p1; rs2 += p2; rs3 += p3;
      pk[k >> 1] = pack_half2(p0, p1);
      pk[(k >> 1) + 1] = pack_half2(p2, p3);
    }
    tmem_st16(t_S + q * 16, pk);
  }
  return (rs0 + rs1) + (rs2 + rs3);
}

template <bool ADAIN>
__global__ void __launch_bounds__(kAttnThreads, 1) shared_attn_kernel(const __grid_constant__ AttnKParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
  uint8_t* sQ = smem + kOffQ;
  uint8_t* sK = smem + kOffK;
  uint8_t* sV = smem + kOffV;
  float* sAd = reinterpret_cast<float*>(smem + kOffAdain);
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + kOffBar);
  uint64_t* q_full = bars;                   // 1
  uint64_t* kv_full = bars + 1;              // kKVStages
  uint64_t* kv_empty = kv_full + kKVStages;  // kKVStages
  constexpr int NB = kNBuf<ADAIN>;
  uint64_t* s_full = kv_empty + kKVStages;   // 3 (per score buffer)
  uint64_t* p_full = s_full + 3;             // 3
  uint64_t* o_done = p_full + 3;             // 2 (per query tile): one phase per PV
  uint64_t* all_done = o_done + 2;           // 1: every MMA of the CTA has completed
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(all_done + 1);

  const int warp = threadIdx.x >> 5;
  const int lane = threadIdx.x & 31;
  const int pair = blockIdx.x / p.n_splits, split = blockIdx.x - pair * p.n_splits;
  const int head = blockIdx.y, b = blockIdx.z;
  const int g_begin = split * p.tiles_per_split;
  const int g_end = min(g_begin + p.tiles_per_split, p.total_tiles);
  const int n_tiles = g_end - g_begin;

  if (warp == 8 && lane == 0) {
    tma_prefetch_desc(&p.tma_q);
    if (p.has_own) { tma_prefetch_desc(&p.tma_k_own); tma_prefetch_desc(&p.tma_v_own); }
    if (p.n_ref) { tma_prefetch_desc(&p.tma_k_ref); tma_prefetch_desc(&p.tma_v_ref); }
    mbar_init(q_full, 1);
    for (int s = 0; s < kKVStages; ++s) { mbar_init(&kv_full[s], 1); mbar_init(&kv_empty[s], 1); }
    for (int i = 0; i < 3; ++i) {
      mbar_init(&s_full[i], 1);
      mbar_init(&p_full[i], 128);
    }
    for (int i = 0; i < 2; ++i) mbar_init(&o_done[i], 1);
    mbar_init(all_done, 1);
    fence_mbar_init();
  }
  if (warp == 9) {
    tmem_alloc(tmem_slot, 512);
    tmem_relinquish();
  }
  pdl_wait();     // barriers / TMEM are set up; everything below reads what the previous kernels wrote
  if (ADAIN) {
    // stage the AdaIN affine of every reference for this (batch, head): sAd[r][0][d] = scale, sAd[r][1][d] = shift
    const int C = p.heads * kD;
    for (int i = threadIdx.x; i < p.n_ref * kD; i += blockDim.x) {
      const int r = i / kD, d = i % kD;
      const size_t g = (static_cast<size_t>(b) * p.n_ref + r) * C + head * kD + d;
      sAd[(r * 2 + 0) * kD + d] = p.adain_scale ? p.adain_scale[g] : 1.0f;   // identity affine: per-chunk bookkeeping only
      sAd[(r * 2 + 1) * kD + d] = p.adain_scale ? p.adain_shift[g] : 0.0f;
    }
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  if (warp == 8) {
    // ------------------------------------------------------------------ TMA producer
    if (elect_one()) {   // elect.sync, not lane == 0: TMA / MMA operands then stay in uniform registers (no R2UR loops)
      mbar_arrive_expect_tx(q_full, 2 * kTileBytes);
      tma_load_3d(sQ, &p.tma_q, q_full, p.q_col_off + head * kD, (2 * pair) * kQT, b);
      tma_load_3d(sQ + kTileBytes, &p.tma_q, q_full, p.q_col_off + head * kD, (2 * pair + 1) * kQT, b);
      TileRef tr = locate_tile(p, g_begin);      // advanced at the end of the loop body (no division per tile)
      for (int jj = 0; jj < n_tiles; ++jj) {
        const int s = jj % kKVStages;
        mbar_wait(&kv_empty[s], ((jj / kKVStages) & 1) ^ 1);
        mbar_arrive_expect_tx(&kv_full[s], 2 * kTileBytes);
        if (tr.ref < 0) {
          const int bb = p.own_shared ? 0 : b;
          tma_load_3d(sK + s * kTileBytes, &p.tma_k_own, &kv_full[s], p.k_own_col_off + head * kD, tr.t * kKT, bb);
          tma_load_3d(sV + s * kTileBytes, &p.tma_v_own, &kv_full[s], p.v_own_col_off + head * kD, tr.t * kKT, bb);
        } else {
          tma_load_4d(sK + s * kTileBytes, &p.tma_k_ref, &kv_full[s], p.ref_col_off + head * kD, tr.t * kKT, tr.ref, b);
          tma_load_4d(sV + s * kTileBytes, &p.tma_v_ref, &kv_full[s], p.ref_col_off + head * kD, tr.t * kKT, tr.ref, b);
        }
        if (++tr.t == (tr.ref < 0 ? p.own_tiles : p.ref_tiles)) { tr.t = 0; ++tr.ref; }
      }
    }
    __syncwarp();
  } else if (warp == 9) {
    // ------------------------------------------------------------------ MMA issuer
    if (elect_one()) {
      constexpr uint32_t IDESC_QK = umma_idesc_f16(128, kKT, 0, 0);   // A = Q (K-major), B = K (K-major)
      constexpr uint32_t IDESC_PV = umma_idesc_f16(128, kD, 0, 1);    // A = P (TMEM), B = V (MN-major)
      // descriptors are (constant high word, low word = address >> 4 | LBO): stepping K or the stage is an integer add;
      // the issue loop has to stay far below the 512 clk of tensor work per (KV tile, query tile)
      const uint32_t q_lo0 = umma_desc_lo(smem_u32(sQ)), k_lo0 = umma_desc_lo(smem_u32(sK)), v_lo0 = umma_desc_lo_mn(smem_u32(sV));
      // slot k = 2 * (KV tile) + (query tile) lives in score buffer k % NB
      auto issue_qk = [&](int k, int buf) {
        const uint32_t k_lo = k_lo0 + ((k >> 1) % kKVStages) * (kTileBytes >> 4);
        const uint32_t q_lo = q_lo0 + (k & 1) * (kTileBytes >> 4);
#pragma unroll
        for (int kk = 0; kk < kD / 16; ++kk)
          umma_f16_ss_lo(tmem_base + kTmemS + buf * kKT, q_lo + 2 * kk, k_lo + 2 * kk, IDESC_QK, kk != 0 ? 1u : 0u);
        umma_commit(&s_full[buf]);
      };
      int kv_ready = -1;          // highest KV tile whose arrival has been waited for
      auto need_kv = [&](int jj) {
        if (jj > kv_ready) {
          mbar_wait(&kv_full[jj % kKVStages], (jj / kKVStages) & 1);
          tc_fence_after();
          kv_ready = jj;
        }
      };
      const int n_slots = 2 * n_tiles;
      mbar_wait(q_full, 0);
      for (int k = 0; k < NB && k < n_slots; ++k) {
        need_kv(k >> 1);
        issue_qk(k, k);
      }
      TileRef mt = locate_tile(p, g_begin);
      int kb_run = 0;             // buffer and barrier phase of the current slot (NB = 2: compile-time, = query tile)
      uint32_t kp_run = 0;
      for (int jj = 0; jj < n_tiles; ++jj) {
        const int s = jj % kKVStages;
        // a segment (AdaIN: one chunk; otherwise the whole range) starts with a fresh accumulator
        const bool fresh = (jj == 0) || (ADAIN && mt.t == 0);
        if (++mt.t == (mt.ref < 0 ? p.own_tiles : p.ref_tiles)) { mt.t = 0; ++mt.ref; }
        const uint32_t v_lo = v_lo0 + s * (kTileBytes >> 4);
        // K tiles of the QK^Ts issued below: waited for here, off the chain P published -> PV -> QK^T -> scores
        if (2 * jj + NB < n_slots) need_kv((2 * jj + NB) >> 1);
        if (2 * jj + 1 + NB < n_slots) need_kv((2 * jj + 1 + NB) >> 1);
#pragma unroll
        for (int i = 0; i < 2; ++i) {
          const int kb = NB == 2 ? i : kb_run;
          const uint32_t kp = NB == 2 ? static_cast<uint32_t>(jj & 1) : kp_run;
          mbar_wait(&p_full[kb], kp);
          tc_fence_after();
#pragma unroll
          for (int kk = 0; kk < kKT / 16; ++kk)
            umma_f16_ts_lo(tmem_base + kTmemO<ADAIN> + i * kD, tmem_base + kTmemS + kb * kKT + kk * 8, v_lo + kk * (2048 >> 4), IDESC_PV,
                           (kk != 0 || !fresh) ? 1u : 0u);
          umma_commit(&o_done[i]);
          // tcgen05.mma executes in issue order: QK^T of slot k + NB refills this buffer only after this PV has read P
          // out of it
          const int k2 = 2 * jj + i + NB;
          if (k2 < n_slots) issue_qk(k2, kb);
          if (i == 1) umma_commit(&kv_empty[s]);   // both query tiles have consumed K (QK^T issued earlier) and V of this tile
          if (NB != 2 && ++kb_run == NB) { kb_run = 0; kp_run ^= 1; }
        }
      }
      umma_commit(all_done);
    }
    __syncwarp();
  } else {
    // ------------------------------------------------------------------ softmax warpgroups (thread == query row)
    const int i = warp >> 2;                         // query tile of this warpgroup
    const int row = (warp & 3) * 32 + lane;
    const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
    const uint32_t t_S0 = tmem_base + kTmemS + lane_addr;
    const uint32_t t_O = tmem_base + kTmemO<ADAIN> + i * kD + lane_addr;
    const uint32_t t_A = tmem_base + kTmemAcc + i * kD + lane_addr;
    const float c = p.scale_log2;

    float m_ref = -INFINITY;     // reference max (log2 units) every stored exponential is relative to
    float l_seg = 0.f;           // row sum of the current segment
    float l_tot = 0.f;           // row sum of the finished segments (AdaIN path)
    bool acc_valid = false;      // the second accumulator holds finished segments
    int buf_run = i;             // score buffer of slot 2 * jj + i and the phase of its barriers (NB = 2: buffer = query tile)
    uint32_t sp = 0;

    // Tile bookkeeping is carried across the loop: ONE division at the start instead of three per tile and thread (the
    // AdaIN variant executed 8100 warp instructions per CTA and KV tile against 7020 for the plain one, all of the
    // difference in these divisions at the head of the dependency chain; removing them: 539 -> 655 TFLOP/s at B=8).
    TileRef tr = locate_tile(p, g_begin);
    for (int jj = 0; jj < n_tiles; ++jj) {
      const int chunk_tiles = tr.ref < 0 ? p.own_tiles : p.ref_tiles;
      const bool seg_first = (jj == 0) || (ADAIN && tr.t == 0);
      const bool seg_last = (jj == n_tiles - 1) || (ADAIN && tr.t == chunk_tiles - 1);
      const int len = tr.ref < 0 ? p.s_own : p.s_ref;
      const int valid = min(kKT, len - tr.t * kKT);   // keys of this tile that exist
      const int buf = NB == 2 ? i : buf_run;
      const uint32_t t_S = t_S0 + buf * kKT;

      mbar_wait(&s_full[buf], sp);
      tc_fence_after();
      // pass 1: row max (TMEM reads are cheap; keeping all 128 scores live across both passes would spill)
      const float m_new = (valid < kKT ? row_max128<true>(t_S, valid) : row_max128<false>(t_S, valid)) * c;

      if (jj == 0) {
        m_ref = m_new;                                // nothing accumulated yet
      } else {
        const bool need = m_new > m_ref + kRescaleThreshold;
        if (__any_sync(0xffffffffu, need)) {          // TMEM ld/st are warp-collective: the warp rescales together
          // PV(jj - 1) of this query tile is the last one issued (the next waits for the P published below). With two
          // score buffers these scores were committed after it; with three it may still be accumulating into O.
          if (NB > 2) {
            mbar_wait(&o_done[i], (jj - 1) & 1);
            tc_fence_after();
          }
          const float f = need ? fast_exp2(m_ref - m_new) : 1.0f;
          if (need) m_ref = m_new;
          if (!seg_first) tmem_scale64(t_O, f);
          if (ADAIN && acc_valid) tmem_scale64(t_A, f);
          tmem_st_wait();
          tc_fence_before();
          l_seg *= f;
          l_tot *= f;
        }
      }

      // pass 2: exponentials -> P (TMEM), row sum
      l_seg += valid < kKT ? exp_pass128<true>(t_S, c, -m_ref, valid) : exp_pass128<false>(t_S, c, -m_ref, valid);
      tmem_st_wait();
      tc_fence_before();
      mbar_arrive(&p_full[buf]);
      if (NB == 2) {
        sp ^= 1;
      } else {
        buf_run += 2;
        if (buf_run >= NB) { buf_run -= NB; sp ^= 1; }
      }

      if (ADAIN && seg_last) {
        // fold the finished segment into the second accumulator: acc += a * O_seg + b * rowsum(P_seg)
        mbar_wait(&o_done[i], jj & 1);
        tc_fence_after();
        const bool affine = tr.ref >= 0;
        const float* ad = sAd + (affine ? tr.ref : 0) * 2 * kD;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[32];
          tmem_ld32(t_O + h * 32, o);
          tmem_ld_wait();
          if (affine) {
#pragma unroll
            for (int d = 0; d < 32; ++d)
              o[d] = __float_as_uint(fmaf(ad[h * 32 + d], __uint_as_float(o[d]), ad[kD + h * 32 + d] * l_seg));
          }
          if (acc_valid) {
            uint32_t a[32];
            tmem_ld32(t_A + h * 32, a);
            tmem_ld_wait();
#pragma unroll
            for (int d = 0; d < 32; ++d) o[d] = __float_as_uint(__uint_as_float(o[d]) + __uint_as_float(a[d]));
          }
          tmem_st32(t_A + h * 32, o);
        }
        tmem_st_wait();
        tc_fence_before();
        acc_valid = true;
        if (p.mass_ws) {
          const int n_qrows = 2 * (gridDim.x / p.n_splits) * kQT;
          p.mass_ws[((static_cast<size_t>(b) * p.heads + head) * n_qrows + (2 * pair + i) * kQT + row) * (p.n_chunks + 1) + tr.chunk] =
              make_float2(l_seg, m_ref);
        }
        l_tot += l_seg;
        l_seg = 0.f;
      }
      // next tile: same chunk, or the first tile of the next one (own -> reference 0 -> reference 1 ...)
      if (++tr.t == chunk_tiles) {
        tr.t = 0;
        ++tr.ref;
        ++tr.chunk;
      }
    }

    // ---------------------------------------------------------------- epilogue
    pdl_launch_dependents();
    const int qrow = (2 * pair + i) * kQT + row;
    if (n_tiles > 0) {
      if (!ADAIN) {
        mbar_wait(all_done, 0);      // not o_done: with three score buffers its phase may be two PVs behind
        tc_fence_after();
      }
      const uint32_t t_src = ADAIN ? t_A : t_O;
      const float l = ADAIN ? l_tot : l_seg;
      if (ADAIN && p.mass_ws) {
        const int n_qrows = 2 * (gridDim.x / p.n_splits) * kQT;
        p.mass_ws[((static_cast<size_t>(b) * p.heads + head) * n_qrows + (2 * pair + i) * kQT + row) * (p.n_chunks + 1) + p.n_chunks] =
            make_float2(l, m_ref);
      }
      if (p.n_splits == 1) {
        const float inv = 1.0f / l;
        __half* op = p.out + (static_cast<size_t>(b) * p.s_q + qrow) * p.out_stride + head * kD;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[32];
          tmem_ld32(t_src + h * 32, o);
          tmem_ld_wait();
          if (qrow < p.s_q) {
#pragma unroll
            for (int q = 0; q < 4; ++q) {
              const uint4 u = make_uint4(
                  pack_half2(__uint_as_float(o[q * 8]) * inv, __uint_as_float(o[q * 8 + 1]) * inv),
                  pack_half2(__uint_as_float(o[q * 8 + 2]) * inv, __uint_as_float(o[q * 8 + 3]) * inv),
                  pack_half2(__uint_as_float(o[q * 8 + 4]) * inv, __uint_as_float(o[q * 8 + 5]) * inv),
                  pack_half2(__uint_as_float(o[q * 8 + 6]) * inv, __uint_as_float(o[q * 8 + 7]) * inv));
              reinterpret_cast<uint4*>(op)[h * 4 + q] = u;
            }
          }
        }
      } else {
        const int n_qt = 2 * (gridDim.x / p.n_splits);
        const size_t unit = ((static_cast<size_t>(b) * p.heads + head) * n_qt + (2 * pair + i)) * p.n_splits + split;
        float* po = p.part_o + (unit * 128 + row) * kD;
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          uint32_t o[32];
          tmem_ld32(t_src + h * 32, o);
          tmem_ld_wait();
#pragma unroll
          for (int q = 0; q < 8; ++q)
            reinterpret_cast<uint4*>(po)[h * 8 + q] = make_uint4(o[4 * q], o[4 * q + 1], o[4 * q + 2], o[4 * q + 3]);
        }
        p.part_ml[unit * 128 + row] = make_float2(m_ref, l);
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 9) tmem_dealloc(tmem_base, 512);
}

// Merges the split-KV partials of one query row in split order: out = sum_s 2^(m_s - M) O_s / sum_s 2^(m_s - M) l_s.
// grid = (q tiles, heads, batch), block = 128 (thread == row).
__global__ void __launch_bounds__(128) attn_combine_kernel(const float* __restrict__ part_o, const float2* __restrict__ part_ml,
                                                           int n_splits, int n_qt, int heads, int s_q,
                                                           __half* __restrict__ out, int out_stride) {
  const int qt = blockIdx.x, head = blockIdx.y, b = blockIdx.z, row = threadIdx.x;
  const int qrow = qt * kQT + row;
  pdl_launch_dependents();
  pdl_wait();
  if (qrow >= s_q) return;
  const size_t unit0 = ((static_cast<size_t>(b) * heads + head) * n_qt + qt) * n_splits;
  float M = -INFINITY;
  for (int s = 0; s < n_splits; ++s) M = fmaxf(M, part_ml[(unit0 + s) * 128 + row].x);
  float acc[kD];
#pragma unroll
  for (int d = 0; d < kD; ++d) acc[d] = 0.f;
  float l = 0.f;
  for (int s = 0; s < n_splits; ++s) {
    const float2 ml = part_ml[(unit0 + s) * 128 + row];
    const float w = fast_exp2(ml.x - M);
    l = fmaf(ml.y, w, l);
    const float4* po = reinterpret_cast<const float4*>(part_o + ((unit0 + s) * 128 + row) * kD);
#pragma unroll
    for (int q = 0; q < kD / 4; ++q) {
      const float4 v = po[q];
      acc[4 * q] = fmaf(v.x, w, acc[4 * q]);
      acc[4 * q + 1] = fmaf(v.y, w, acc[4 * q + 1]);
      acc[4 * q + 2] = fmaf(v.z, w, acc[4 * q + 2]);
      acc[4 * q + 3] = fmaf(v.w, w, acc[4 * q + 3]);
    }
  }
  const float inv = 1.0f / l;
  __half* op = out + (static_cast<size_t>(b) * s_q + qrow) * out_stride + head * kD;
#pragma unroll
  for (int q = 0; q < 8; ++q)
    reinterpret_cast<uint4*>(op)[q] =
        make_uint4(pack_half2(acc[q * 8] * inv, acc[q * 8 + 1] * inv), pack_half2(acc[q * 8 + 2] * inv, acc[q * 8 + 3] * inv),
                   pack_half2(acc[q * 8 + 4] * inv, acc[q * 8 + 5] * inv), pack_half2(acc[q * 8 + 6] * inv, acc[q * 8 + 7] * inv));
}

// chunk_mass[b, h, c] = mean over queries of the softmax mass that lands on KV chunk c (what gradio_demo.py:118-133
// derives from the dense attention_probs). grid = (heads, batch), block = 256; fixed-order block reduction.
__global__ void __launch_bounds__(256) attn_mass_kernel(const float2* __restrict__ ws, int n_qrows, int s_q, int n_chunks,
                                                        int heads, float* __restrict__ mass) {
  __shared__ float red[256];
  const int head = blockIdx.x, b = blockIdx.y;
  const float2* base = ws + (static_cast<size_t>(b) * heads + head) * n_qrows * (n_chunks + 1);
  for (int c = 0; c < n_chunks; ++c) {
    float acc = 0.f;
    for (int q = threadIdx.x; q < s_q; q += 256) {
      const float2 tot = base[static_cast<size_t>(q) * (n_chunks + 1) + n_chunks];
      const float2 seg = base[static_cast<size_t>(q) * (n_chunks + 1) + c];
      acc += seg.x * fast_exp2(seg.y - tot.y) / tot.x;
    }
    red[threadIdx.x] = acc;
    __syncthreads();
    for (int o = 128; o > 0; o >>= 1) {
      if (threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
      __syncthreads();
    }
    if (threadIdx.x == 0) mass[(static_cast<size_t>(b) * heads + head) * n_chunks + c] = red[0] / s_q;
    __syncthreads();
  }
}

// Split-KV plan: the number of KV ranges that minimises (rounds over 148 SMs) x (tiles per range), with a small
// charge per split for the partial write + combine.
static int plan_splits(long units, int total_tiles, int requested) {
  if (requested > 0) return requested < total_tiles ? requested : total_tiles;
  if (units >= 120 || total_tiles < 4) return 1;
  int best = 1;
  double best_cost = 1e30;
  for (int s = 1; s <= 16 && s <= total_tiles / 2; ++s) {
    const int tps = (total_tiles + s - 1) / s;
    const long rounds = (units * s + 147) / 148;
    const double cost = static_cast<double>(rounds) * (tps + 1.5) + (s > 1 ? 1.0 + 0.25 * s : 0.0);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = s; }
  }
  return best;
}

}  // namespace ir

extern "C" size_t ir_shared_attn_workspace_bytes(int batch, int heads, int s_q) {
  const size_t pairs = static_cast<size_t>((s_q + 255) / 256);
  if (pairs * heads * batch >= 120) return 0;   // plan_splits never splits these
  return static_cast<size_t>(batch) * heads * 2 * pairs * 16 * 128 * (ir::kD * sizeof(float) + sizeof(float2));
}

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
  if (p->q_row_stride < p->q_col_off + p->heads * kD || (has_own && p->own_row_stride < p->heads * kD) ||
      (p->n_ref && p->ref_row_stride < p->ref_col_off + p->heads * kD))
    return set_error(IR_ERR_SHAPE, "ir_shared_attn_fwd: row stride smaller than heads*64 columns");
  if (p->q_row_stride % 8 || p->out_row_stride % 8 || (has_own && p->own_row_stride % 8) || (p->n_ref && p->ref_row_stride % 8))
    return set_error(IR_ERR_ALIGN, "ir_shared_attn_fwd: row strides must be multiples of 8 elements");
  if ((p->q_col_off | p->k_own_col_off | p->v_own_col_off | p->ref_col_off) % 8)
    return set_error(IR_ERR_ALIGN, "ir_shared_attn_fwd: column offsets must be multiples of 8");
  if (reinterpret_cast<uintptr_t>(p->out) & 15) return set_error(IR_ERR_ALIGN, "ir_shared_attn_fwd: out not 16-byte aligned");
  if (p->kv_splits < 0 || p->kv_splits > 16) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: kv_splits=%d (0 = auto, 1..16)", p->kv_splits);

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
  const bool want_mass = p->chunk_mass != nullptr;
  const bool adain = p->n_ref > 0 && (p->adain_scale != nullptr || want_mass);   // per-chunk segments (identity affine for mass only)
  kp.adain_scale = (p->n_ref > 0) ? p->adain_scale : nullptr;
  kp.adain_shift = (p->n_ref > 0) ? p->adain_shift : nullptr;
  kp.n_chunks = (has_own ? 1 : 0) + p->n_ref;
  kp.out = static_cast<__half*>(p->out);
  kp.out_stride = p->out_row_stride;
  kp.own_tiles = has_own ? (p->s_own + kKT - 1) / kKT : 0;
  kp.ref_tiles = p->n_ref > 0 ? (p->s_ref + kKT - 1) / kKT : 1;
  kp.total_tiles = kp.own_tiles + p->n_ref * (p->n_ref > 0 ? kp.ref_tiles : 0);

  const int pairs = (p->s_q + 2 * kQT - 1) / (2 * kQT);
  if (want_mass) {
    if (p->n_ref == 0) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: chunk_mass needs reference chunks");
    const size_t need = static_cast<size_t>(p->batch) * p->heads * (2 * pairs * kQT) * (kp.n_chunks + 1) * sizeof(float2);
    if (!p->workspace || p->workspace_bytes < need)
      return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: chunk_mass needs %zu bytes of workspace", need);
    if (p->kv_splits > 1) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: chunk_mass and kv_splits > 1 are exclusive");
    kp.mass_ws = static_cast<float2*>(p->workspace);
  }
  int n_splits = want_mass ? 1 : plan_splits(static_cast<long>(pairs) * p->heads * p->batch, kp.total_tiles, p->kv_splits);
  if (n_splits > 1 && !p->workspace) {
    if (p->kv_splits > 1) return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: kv_splits=%d needs a workspace", p->kv_splits);
    n_splits = 1;
  }
  kp.tiles_per_split = (kp.total_tiles + n_splits - 1) / n_splits;
  n_splits = (kp.total_tiles + kp.tiles_per_split - 1) / kp.tiles_per_split;   // no empty ranges
  kp.n_splits = n_splits;
  if (n_splits > 1) {
    const size_t units = static_cast<size_t>(p->batch) * p->heads * (2 * pairs) * n_splits;
    if (p->workspace_bytes < units * 128 * (kD * sizeof(float) + sizeof(float2)))
      return set_error(IR_ERR_ARG, "ir_shared_attn_fwd: workspace too small (%zu bytes)", p->workspace_bytes);
    kp.part_o = static_cast<float*>(p->workspace);
    kp.part_ml = reinterpret_cast<float2*>(kp.part_o + units * 128 * kD);
  }

  static PerDeviceOnce attr_once;   // function attributes are per device
  bool& attr_done = attr_once.slot();
  if (!attr_done) {
    cudaError_t e = cudaFuncSetAttribute(shared_attn_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e == cudaSuccess) e = cudaFuncSetAttribute(shared_attn_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kAttnSmem);
    if (e != cudaSuccess) return set_error(IR_ERR_CUDA, "cudaFuncSetAttribute(shared_attn): %s", cudaGetErrorString(e));
    attr_done = true;
  }
  dim3 grid(pairs * n_splits, p->heads, p->batch);
  if (adain) IR_LAUNCH(shared_attn_kernel<true>, grid, kAttnThreads, kAttnSmem, stream, kp);
  else IR_LAUNCH(shared_attn_kernel<false>, grid, kAttnThreads, kAttnSmem, stream, kp);
  IR_CUDA_LAUNCH_CHECK("shared_attn launch");
  if (want_mass) {
    attn_mass_kernel<<<dim3(p->heads, p->batch), 256, 0, stream>>>(kp.mass_ws, 2 * pairs * kQT, p->s_q, kp.n_chunks, p->heads, p->chunk_mass);
    IR_CUDA_LAUNCH_CHECK("attn_mass launch");
  }
  if (n_splits > 1) {
    IR_LAUNCH(attn_combine_kernel, dim3(2 * pairs, p->heads, p->batch), 128, 0, stream, kp.part_o, kp.part_ml, n_splits, 2 * pairs, p->heads,
              p->s_q, kp.out, kp.out_stride);
    IR_CUDA_LAUNCH_CHECK("attn_combine launch");
  }
  return 0;
}
