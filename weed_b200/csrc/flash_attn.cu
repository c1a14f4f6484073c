// flash_attn.cu — attention core with the scores kept on chip (tcgen05 + TMEM), head_dim 64.
//
// Replaces steps 2-4 of attention.cu (S = Q K^T to HBM, softmax pass, P V from HBM) for the chain
// of MultiHeadAttention::forward, src/modules/multihead_attention.cpp:319-345: per (b, h) and per
// tile of 128 queries, loop over tiles of 128 keys (only up to the diagonal when causal):
//     S  = Q K_j^T            tcgen05.mma 128x128x64, accumulator in TMEM (never leaves the SM)
//     online softmax          4 warps, thread == query row: running max m and sum l, p = 2^(t - m)
//     P  -> shared memory     bf16, written in the 128-B swizzled K-major layout the MMA reads
//     O_j = P V_j             tcgen05.mma 128x64x128 into TMEM, folded into a register accumulator
//                             acc = acc * 2^(m_old - m_new) + O_j
// and finally O = acc / l. HBM traffic per (b,h): Q, K, V once per query tile and the output —
// the [T, T] score matrix (403 MB per layer at B=8, H=12, T=1024) is never stored.
// Operands come from the bf16 head layout [B*H][hd][T] (T contiguous) that heads_pack writes: one
// tensor-map geometry serves Q (A, MN-major), K (B, MN-major) and V (B, K-major).
// Warp roles: 0-3 softmax + epilogue (TMEM lane quarter == warp), 4 TMA producer, 5 MMA issuer.
// Two CTAs fit per SM (97 KB smem, 256 TMEM columns each): one runs its MMAs while the other is in
// its softmax phase.
#include "tc_common.cuh"

namespace weedcu {
namespace flash {
using namespace tc;

constexpr uint32_t HD = 64, TQ = 128, TK = 128, NTHREADS = 192;
constexpr uint32_t Q_BYTES = TQ * HD * 2, KV_BYTES = TK * HD * 2, P_BYTES = TQ * TK * 2;
// K_j and V_j tiles stream through one ring of 3 slots in the order K0 V0 K1 V1 ...: K_j's slot is
// free as soon as S_j retires, so K_j, V_j and K_{j+1} are resident while tile j is processed
// (96 KB per CTA with Q and P: two CTAs per SM).
constexpr uint32_t RING = 3;
constexpr uint32_t OFF_Q = 0, OFF_KV = OFF_Q + Q_BYTES, OFF_P = OFF_KV + RING * KV_BYTES;
constexpr uint32_t OFF_BAR = OFF_P + P_BYTES, SMEM_BYTES = OFF_BAR + 128 + 1024;
constexpr uint32_t TMEM_COLS = 256, TMEM_S = 0, TMEM_O = 128;

__device__ __forceinline__ float ex2(float x) {
  float y;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
  return y;
}
__device__ __forceinline__ void st_shared_v4(uint32_t addr, uint4 v) {
  asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w) : "memory");
}

__global__ void __launch_bounds__(NTHREADS, 2)
flash_attn_fwd_kernel(const __grid_constant__ CUtensorMap tmQ, const __grid_constant__ CUtensorMap tmK,
                      const __grid_constant__ CUtensorMap tmV, float *__restrict__ oc, uint32_t T, uint32_t q_tiles,
                      float scale_log2, int causal) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (base - raw);
  const uint32_t sQ = base + OFF_Q, sKV = base + OFF_KV, sP = base + OFF_P, bars = base + OFF_BAR;
  const uint32_t q_full = bars, s_full = bars + 8, p_ready = bars + 16, o_full = bars + 24;
  auto kv_full = [&](uint32_t s) { return bars + 32 + 8 * s; };
  auto kv_empty = [&](uint32_t s) { return bars + 56 + 8 * s; };
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen_base + OFF_BAR + 96);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = (b,h), blockIdx.y counts query tiles from the last one: under a causal mask tile qt
  // costs qt + 1 key tiles, and blocks are dispatched in linear order, so the longest tiles of ALL
  // heads start first and the short ones fill the tail (longest-processing-time-first)
  const uint32_t qt = q_tiles - 1u - blockIdx.y;
  const uint32_t bh = blockIdx.x;
  const uint32_t q0 = qt * TQ;
  const uint32_t nk = causal ? (qt + 1u) : ((T + TK - 1) / TK);

  if (warp == 4 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
  }
  if (warp == 5 && lane == 0) {
    mbar_init(q_full, 1);
    mbar_init(s_full, 1);
    mbar_init(p_ready, 128);
    mbar_init(o_full, 1);
    for (uint32_t s = 0; s < RING; ++s) {
      mbar_init(kv_full(s), 1);
      mbar_init(kv_empty(s), 1);
    }
    fence_barrier_init();
  }
  if (warp == 0) tmem_alloc(smem_u32((const void *)tmem_slot), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync(); // set-up above is CTA-local; operands written by the previous kernel are read below

  if (warp == 4) {
    // ================================ TMA producer =====================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (uint32_t i = 0; i < TQ / 64; ++i) tma_load_3d(sQ + i * (64 * HD * 2), &tmQ, q_full, (int)(q0 + 64 * i), 0, (int)bh);
      for (uint32_t it = 0; it < 2 * nk; ++it) { // item 2j = K_j, item 2j+1 = V_j
        const uint32_t slot = it % RING, ph = (it / RING) & 1u;
        mbar_wait(kv_empty(slot), ph ^ 1u);
        mbar_arrive_expect_tx(kv_full(slot), KV_BYTES);
        const int k0 = (int)((it >> 1) * TK);
        const CUtensorMap *tm = (it & 1u) ? &tmV : &tmK;
#pragma unroll
        for (uint32_t i = 0; i < TK / 64; ++i)
          tma_load_3d(sKV + slot * KV_BYTES + i * (64 * HD * 2), tm, kv_full(slot), k0 + (int)(64 * i), 0, (int)bh);
      }
    }
  } else if (warp == 5) {
    // ================================ MMA issuer ========================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(TQ, TK, 1, 1);  // A = Q (MN-major), B = K (MN-major)
      constexpr uint32_t idesc_o = make_idesc(TQ, HD, 0, 0);  // A = P (K-major),  B = V (K-major)
      mbar_wait(q_full, 0);
      for (uint32_t j = 0; j < nk; ++j) {
        const uint32_t ks = (2 * j) % RING, kph = ((2 * j) / RING) & 1u;
        const uint32_t vs = (2 * j + 1) % RING, vph = ((2 * j + 1) / RING) & 1u;
        const uint32_t sK = sKV + ks * KV_BYTES, sV = sKV + vs * KV_BYTES;
        mbar_wait(kv_full(ks), kph);
        tcgen05_fence_after();
        // S = Q K_j^T : 4 k-steps over head_dim; MN-major operands: 16 k-rows of 128 B per step,
        // LBO = next 64-wide MN block (64 k-rows x 128 B)
#pragma unroll
        for (uint32_t k = 0; k < HD / UMMA_K; ++k) {
          const uint64_t adesc = make_smem_desc(sQ + k * (UMMA_K * 128), HD * 128, 1024);
          const uint64_t bdesc = make_smem_desc(sK + k * (UMMA_K * 128), HD * 128, 1024);
          umma_f16(tmem_base + TMEM_S, adesc, bdesc, idesc_s, k ? 1u : 0u);
        }
        umma_commit(s_full);
        umma_commit(kv_empty(ks)); // K_j's slot is free once S_j retires
        // every softmax thread has read S_j and O_{j-1} and written P_j before it arrives here
        mbar_wait(p_ready, j & 1u);
        mbar_wait(kv_full(vs), vph);
        tcgen05_fence_after();
        // O_j = P V_j : 8 k-steps over the 128 keys; K-major operands: two 64-key swizzle atoms,
        // 16 keys = 32 B inside the 128-B row, SBO = 8 rows x 128 B
#pragma unroll
        for (uint32_t k = 0; k < TK / UMMA_K; ++k) {
          const uint32_t atom = k >> 2, kk = k & 3u;
          const uint64_t adesc = make_smem_desc(sP + atom * (TQ * 128) + kk * (UMMA_K * 2), 16, 1024);
          const uint64_t bdesc = make_smem_desc(sV + atom * (HD * 128) + kk * (UMMA_K * 2), 16, 1024);
          umma_f16(tmem_base + TMEM_O, adesc, bdesc, idesc_o, k ? 1u : 0u);
        }
        umma_commit(o_full);
        umma_commit(kv_empty(vs)); // V_j's slot is free once these MMAs retire
      }
    }
  } else {
    // ================================ softmax + epilogue (thread == query row) ===========
    const uint32_t r = warp * 32 + lane;           // row inside the tile == TMEM lane
    const uint32_t qg = q0 + r;                    // global query index
    const uint32_t t_lane = tmem_base + ((warp * 32u) << 16);
    const float sc = scale_log2;
    float m = -INFINITY, l = 0.0f, alpha_prev = 0.0f;
    float acc[HD];
#pragma unroll
    for (uint32_t c = 0; c < HD; ++c) acc[c] = 0.0f;

    for (uint32_t j = 0; j < nk; ++j) {
      const uint32_t k0 = j * TK;
      // Only the diagonal tile (causal) and a ragged last key tile need per-element masking; every
      // other tile runs the mask-free instantiation (no predicates, no branches in the inner loops).
      const bool edge = (causal && j == qt) || (k0 + TK > T);
      // keys k0 + c with c < lim are visible to this row: causal -> kg <= qg, ragged -> kg < T
      uint32_t lim = TK;
      if (edge) {
        const uint32_t by_t = (T > k0) ? (T - k0) : 0u;
        const uint32_t by_q = causal ? ((qg >= k0) ? (qg - k0 + 1u) : 0u) : TK;
        lim = min(min(by_t, by_q), TK);
      }
      mbar_wait(s_full, j & 1u);
      tcgen05_fence_after();
      // pass 1: row maximum of the raw scores over the visible keys (the scale is positive, so it is
      // applied once to the maximum); four independent chains
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < TK; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + TMEM_S + c0, v);
        tmem_ld_wait();
        if (edge) {
#pragma unroll
          for (uint32_t c = 0; c < 32; ++c)
            if (c0 + c >= lim) v[c] = 0xff800000u; // -inf
        }
#pragma unroll
        for (uint32_t c = 0; c < 32; c += 4) {
          mx0 = fmaxf(mx0, __uint_as_float(v[c]));
          mx1 = fmaxf(mx1, __uint_as_float(v[c + 1]));
          mx2 = fmaxf(mx2, __uint_as_float(v[c + 2]));
          mx3 = fmaxf(mx3, __uint_as_float(v[c + 3]));
        }
      }
      const float mx = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3)) * sc;
      float m_new = fmaxf(m, mx);
      if (m_new == -INFINITY) m_new = 0.0f;       // nothing visible yet: every p below is 2^(-inf) = 0
      const float alpha = ex2(m - m_new);         // m = -inf -> 0
      // fold the previous key tile's product into the accumulator (also frees P for rewriting)
      if (j > 0) {
        mbar_wait(o_full, (j - 1u) & 1u);
        tcgen05_fence_after();
#pragma unroll
        for (uint32_t c0 = 0; c0 < HD; c0 += 32) {
          uint32_t v[32];
          tmem_ld_32x32b_x32(t_lane + TMEM_O + c0, v);
          tmem_ld_wait();
#pragma unroll
          for (uint32_t c = 0; c < 32; ++c) acc[c0 + c] = acc[c0 + c] * alpha_prev + __uint_as_float(v[c]);
        }
      }
      // pass 2: p = 2^(s*sc - m_new), row sum, P -> shared memory (bf16, swizzled K-major)
      float rs0 = 0.0f, rs1 = 0.0f, rs2 = 0.0f, rs3 = 0.0f;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < TK; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32b_x32(t_lane + TMEM_S + c0, v);
        tmem_ld_wait();
        if (edge) {
#pragma unroll
          for (uint32_t c = 0; c < 32; ++c)
            if (c0 + c >= lim) v[c] = 0xff800000u; // -inf -> p = 0
        }
        float pv[32];
#pragma unroll
        for (uint32_t c = 0; c < 32; c += 4) {
          pv[c] = ex2(fmaf(__uint_as_float(v[c]), sc, -m_new));
          pv[c + 1] = ex2(fmaf(__uint_as_float(v[c + 1]), sc, -m_new));
          pv[c + 2] = ex2(fmaf(__uint_as_float(v[c + 2]), sc, -m_new));
          pv[c + 3] = ex2(fmaf(__uint_as_float(v[c + 3]), sc, -m_new));
          rs0 += pv[c];
          rs1 += pv[c + 1];
          rs2 += pv[c + 2];
          rs3 += pv[c + 3];
        }
        const uint32_t atom = c0 >> 6;                 // 64-key swizzle atom
        const uint32_t chunk0 = (c0 & 63u) >> 3;       // first 16-byte chunk of this group inside the 128-B row
        const uint32_t row_addr = sP + atom * (TQ * 128) + r * 128;
#pragma unroll
        for (uint32_t g = 0; g < 4; ++g) {
          __nv_bfloat162 h[4];
#pragma unroll
          for (uint32_t e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(pv[g * 8 + 2 * e], pv[g * 8 + 2 * e + 1]);
          st_shared_v4(row_addr + (((chunk0 + g) ^ (r & 7u)) << 4), *reinterpret_cast<const uint4 *>(h));
        }
      }
      l = l * alpha + ((rs0 + rs1) + (rs2 + rs3));
      m = m_new;
      alpha_prev = alpha;
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      mbar_arrive(p_ready);
    }
    // last key tile
    mbar_wait(o_full, (nk - 1u) & 1u);
    tcgen05_fence_after();
#pragma unroll
    for (uint32_t c0 = 0; c0 < HD; c0 += 32) {
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_lane + TMEM_O + c0, v);
      tmem_ld_wait();
#pragma unroll
      for (uint32_t c = 0; c < 32; ++c) acc[c0 + c] = acc[c0 + c] * alpha_prev + __uint_as_float(v[c]);
    }
    tcgen05_fence_before();
    if (qg < T) {
      const float inv = 1.0f / l;
      float *dst = oc + ((uint64_t)bh * HD) * T + qg; // oc[bh][c][t]: a warp stores 32 adjacent t per column
#pragma unroll
      for (uint32_t c = 0; c < HD; ++c) dst[(uint64_t)c * T] = acc[c] * inv;
    }
  }
  __syncthreads();
  if (warp == 0) tmem_dealloc(tmem_base, TMEM_COLS);
}

} // namespace flash

// qh, kh, vh: bf16 [BH][64][T]; oc: fp32 [BH][64][T]
int launch_flash_attn_fwd(const uint16_t *qh, const uint16_t *kh, const uint16_t *vh, float *oc, uint32_t BH, uint32_t T,
                          float divisor, int causal, cudaStream_t st) {
  using namespace flash;
  CUtensorMap tmQ, tmK, tmV;
  const uint64_t head_elems = (uint64_t)HD * T;
  int rc = tc::make_operand_map(&tmQ, qh, 1, T, HD, T, BH, head_elems, 64);
  if (rc == 0) rc = tc::make_operand_map(&tmK, kh, 1, T, HD, T, BH, head_elems, 64);
  if (rc == 0) rc = tc::make_operand_map(&tmV, vh, 1, T, HD, T, BH, head_elems, 64); // same geometry: box {64 keys, 64 hd}
  if (rc) return rc;
  const uint32_t q_tiles = (T + TQ - 1) / TQ;
  ensure_dynamic_smem((const void *)flash_attn_fwd_kernel, (int)SMEM_BYTES);
  // FLOP of the products actually issued (causal: key tiles up to the diagonal only)
  const double tiles = causal ? 0.5 * q_tiles * (q_tiles + 1.0) : (double)q_tiles * ((T + TK - 1) / TK);
  ProfScope prof(WEEDCU_PROF_ATTENTION, st, 2.0 * 2.0 * TQ * TK * HD * tiles * BH);
  launch_k(flash_attn_fwd_kernel, dim3(BH, q_tiles), dim3(NTHREADS), SMEM_BYTES, st, tmQ, tmK, tmV, oc, T, q_tiles,
                                                                         1.4426950408889634f / divisor, causal);
  return after_launch();
}

} // namespace weedcu
