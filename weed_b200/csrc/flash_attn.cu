// flash_attn.cu — attention core with the scores kept on chip (tcgen05 + TMEM), head_dim 64.
//
// Replaces steps 2-4 of attention.cu (S = Q K^T to HBM, softmax pass, P V from HBM) for the chain
// of MultiHeadAttention::forward, src/modules/multihead_attention.cpp:319-345: per (b, h) and per
// tile of 128 queries, loop over tiles of 64 keys (only up to the diagonal when causal):
//     S  = Q K_j^T            tcgen05.mma 128x64x64, accumulator in TMEM (never leaves the SM)
//     online softmax          4 warps, thread == query row: running max m and sum l, p = 2^(t - m)
//     P  -> shared memory     bf16, written in the 128-B swizzled K-major layout the MMA reads
//     O_j = P V_j             tcgen05.mma 128x64x64 into TMEM, folded into a register accumulator
//                             acc = acc * 2^(m_old - m_new) + O_j
// and finally O = acc / l. HBM traffic per (b,h): Q, K, V once per query tile and the output —
// the [T, T] score matrix (403 MB per layer at B=8, H=12, T=1024) is never stored.
// Operands come from the bf16 head layout [B*H][hd][T] (T contiguous) that heads_pack writes: one
// tensor-map geometry serves Q (A, MN-major), K (B, MN-major) and V (B, K-major).
// Warp roles: 0-7 softmax + epilogue, 8 TMA producer, 9 MMA issuer. A query row is shared by TWO threads (warps w and
// w + 4 own the same TMEM lane quarter): each takes one 64-key half of the score tile and 32 of the 64 output columns, so
// the dependent chain per key tile (max, exp2, pack, fold) is half as long and a scheduler has four softmax warps to
// interleave instead of two — with one thread per row the kernel was bound by that chain (ncu r02: issue slots 42 %
// busy, MUFU 33 %, tensor pipe 16 %). The pair agrees on the row maximum through shared memory once per key tile.
// S, P and O are double-buffered (TMEM: S0 S1 O0 O1 = 4 x 64 columns) and S_{j+1} is issued BEFORE the MMA warp waits for
// P_j: the tensor core forms the next scores while the softmax warps work on the current ones, and O_{j-1} is folded after
// P_j has been handed over, so a softmax warp does not wait for a product on its critical path (with single buffers it
// spent ~20 % of its samples waiting for S_{j+1} behind P V_j, ncu r02). 64-key tiles make the four buffers fit 256
// columns: two CTAs per SM (99 KB smem each).
#include "tc_common.cuh"

namespace weedcu {
namespace flash {
using namespace tc;

constexpr uint32_t HD = 64, TQ = 128, TK = 64, NTHREADS = 320, SOFTMAX_WARPS = 8, WARP_TMA = 8, WARP_MMA = 9;
constexpr uint32_t Q_BYTES = TQ * HD * 2, KV_BYTES = TK * HD * 2, P_BYTES = TQ * TK * 2;
// K_j and V_j tiles (8 KB each) stream through one ring of 6 slots in the order K0 V0 K1 V1 ...; the MMA warp consumes them
// as K0 K1 V0 K2 V1 ... (96 KB per CTA with Q and the two P buffers: two CTAs per SM).
constexpr uint32_t RING = 6;
constexpr uint32_t OFF_Q = 0, OFF_KV = OFF_Q + Q_BYTES, OFF_P = OFF_KV + RING * KV_BYTES;
constexpr uint32_t OFF_BAR = OFF_P + 2 * P_BYTES, OFF_X = OFF_BAR + 256, X_BYTES = 2 * 2 * TQ * 4; // row-max exchange: [tile parity][half][row]
constexpr uint32_t SMEM_BYTES = OFF_X + X_BYTES + 1024;
constexpr uint32_t TMEM_COLS = 256, TMEM_S = 0, TMEM_O = 128; // S buffer b at TMEM_S + 64 b, O buffer b at TMEM_O + 64 b

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
  // tile j uses buffer j & 1 of S / P / O and phase (j >> 1) & 1 of that buffer's barriers
  const uint32_t q_full = bars;
  auto s_full = [&](uint32_t b) { return bars + 8 + 8 * b; };
  auto p_ready = [&](uint32_t b) { return bars + 24 + 8 * b; };
  auto o_full = [&](uint32_t b) { return bars + 40 + 8 * b; };
  auto kv_full = [&](uint32_t s) { return bars + 56 + 8 * s; };
  auto kv_empty = [&](uint32_t s) { return bars + 56 + 8 * RING + 8 * s; };
  static_assert(56 + 16 * RING + 4 <= 256, "barrier area");
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen_base + OFF_BAR + 56 + 16 * RING);
  volatile float *xch = reinterpret_cast<volatile float *>(gen_base + OFF_X);

  const uint32_t warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  // blockIdx.x = (b,h), blockIdx.y counts query tiles from the last one: under a causal mask tile qt
  // costs qt + 1 key tiles, and blocks are dispatched in linear order, so the longest tiles of ALL
  // heads start first and the short ones fill the tail (longest-processing-time-first)
  const uint32_t qt = q_tiles - 1u - blockIdx.y;
  const uint32_t bh = blockIdx.x;
  const uint32_t q0 = qt * TQ;
  const uint32_t nk = causal ? min((TQ / TK) * (qt + 1u), (T + TK - 1) / TK) : ((T + TK - 1) / TK);

  if (warp == WARP_TMA && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmQ) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmK) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmV) : "memory");
  }
  if (warp == WARP_MMA && lane == 0) {
    mbar_init(q_full, 1);
    for (uint32_t b = 0; b < 2; ++b) {
      mbar_init(s_full(b), 1);
      mbar_init(p_ready(b), 32 * SOFTMAX_WARPS);
      mbar_init(o_full(b), 1);
    }
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

  if (warp == WARP_TMA) {
    // ================================ TMA producer =====================================
    if (lane == 0) {
      mbar_arrive_expect_tx(q_full, Q_BYTES);
#pragma unroll
      for (uint32_t i = 0; i < TQ / 64; ++i) tma_load_3d(sQ + i * (64 * HD * 2), &tmQ, q_full, (int)(q0 + 64 * i), 0, (int)bh);
      for (uint32_t it = 0; it < 2 * nk; ++it) { // item 2j = K_j, item 2j+1 = V_j
        const uint32_t slot = it % RING, ph = (it / RING) & 1u;
        mbar_wait(kv_empty(slot), ph ^ 1u);
        mbar_arrive_expect_tx(kv_full(slot), KV_BYTES);
        tma_load_3d(sKV + slot * KV_BYTES, (it & 1u) ? &tmV : &tmK, kv_full(slot), (int)((it >> 1) * TK), 0, (int)bh);
      }
    }
  } else if (warp == WARP_MMA) {
    // ================================ MMA issuer ========================================
    if (lane == 0) {
      constexpr uint32_t idesc_s = make_idesc(TQ, TK, 1, 1);  // A = Q (MN-major), B = K (MN-major)
      constexpr uint32_t idesc_o = make_idesc(TQ, HD, 0, 0);  // A = P (K-major),  B = V (K-major)
      // S_j = Q K_j^T into S buffer j & 1: 4 k-steps over head_dim; MN-major operands: 16 k-rows of 128 B per step,
      // LBO = next 64-wide MN block (64 k-rows x 128 B; K_j is a single block)
      auto issue_s = [&](uint32_t j) {
        const uint32_t it = 2 * j, slot = it % RING, ph = (it / RING) & 1u;
        const uint32_t sK = sKV + slot * KV_BYTES;
        mbar_wait(kv_full(slot), ph);
        tcgen05_fence_after();
#pragma unroll
        for (uint32_t k = 0; k < HD / UMMA_K; ++k) {
          const uint64_t adesc = make_smem_desc(sQ + k * (UMMA_K * 128), HD * 128, 1024);
          const uint64_t bdesc = make_smem_desc(sK + k * (UMMA_K * 128), HD * 128, 1024);
          umma_f16(tmem_base + TMEM_S + (j & 1u) * TK, adesc, bdesc, idesc_s, k ? 1u : 0u);
        }
        umma_commit(s_full(j & 1u));
        umma_commit(kv_empty(slot)); // K_j's slot is free once S_j retires
      };
      mbar_wait(q_full, 0);
      issue_s(0);
      for (uint32_t j = 0; j < nk; ++j) {
        // S buffer (j + 1) & 1 was last read by the softmax of tile j - 1, whose p_ready was waited for below one turn ago
        if (j + 1u < nk) issue_s(j + 1u);
        const uint32_t it = 2 * j + 1u, slot = it % RING, ph = (it / RING) & 1u;
        const uint32_t sV = sKV + slot * KV_BYTES, sPj = sP + (j & 1u) * P_BYTES;
        // every softmax thread has written P_j and folded O_{j-2} (the last reader of O buffer j & 1) before it arrives here
        mbar_wait(p_ready(j & 1u), (j >> 1) & 1u);
        mbar_wait(kv_full(slot), ph);
        tcgen05_fence_after();
        // O_j = P_j V_j : 4 k-steps over the 64 keys; K-major operands: one 64-key swizzle atom,
        // 16 keys = 32 B inside the 128-B row, SBO = 8 rows x 128 B
#pragma unroll
        for (uint32_t k = 0; k < TK / UMMA_K; ++k) {
          const uint64_t adesc = make_smem_desc(sPj + k * (UMMA_K * 2), 16, 1024);
          const uint64_t bdesc = make_smem_desc(sV + k * (UMMA_K * 2), 16, 1024);
          umma_f16(tmem_base + TMEM_O + (j & 1u) * HD, adesc, bdesc, idesc_o, k ? 1u : 0u);
        }
        umma_commit(o_full(j & 1u));
        umma_commit(kv_empty(slot)); // V_j's slot is free once these MMAs retire
      }
    }
  } else {
    // ================================ softmax + epilogue (two threads per query row) ======
    const uint32_t quarter = warp & 3u, half = warp >> 2;   // TMEM lane quarter; 32-key half of S / 32-column half of O
    const uint32_t r = quarter * 32 + lane;                  // row inside the tile == TMEM lane
    const uint32_t qg = q0 + r;                              // global query index
    const uint32_t t_lane = tmem_base + ((quarter * 32u) << 16);
    const float sc = scale_log2;
    float m = -INFINITY, l = 0.0f, alpha_prev = 0.0f;
    float acc[32];
#pragma unroll
    for (uint32_t c = 0; c < 32; ++c) acc[c] = 0.0f;
    // acc = acc * alpha_prev + O_jj (this thread's 32 columns), once P V_jj has retired
    auto fold = [&](uint32_t jj) {
      mbar_wait(o_full(jj & 1u), (jj >> 1) & 1u);
      tcgen05_fence_after();
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_lane + TMEM_O + (jj & 1u) * HD + half * 32u, v);
      tmem_ld_wait();
#pragma unroll
      for (uint32_t c = 0; c < 32; ++c) acc[c] = acc[c] * alpha_prev + __uint_as_float(v[c]);
    };

    for (uint32_t j = 0; j < nk; ++j) {
      const uint32_t k0 = j * TK + half * 32u; // first key of this thread's half
      // Only the tiles on the diagonal (causal) and a ragged last key tile need per-element masking; every
      // other tile runs the mask-free path (no predicates in the inner loops).
      const bool edge = (causal && (j + 1u) * TK > q0) || (j * TK + TK > T);
      // keys k0 + c with c < lim are visible to this row: causal -> kg <= qg, ragged -> kg < T
      uint32_t lim = 32u;
      if (edge) {
        const uint32_t by_t = (T > k0) ? (T - k0) : 0u;
        const uint32_t by_q = causal ? ((qg >= k0) ? (qg - k0 + 1u) : 0u) : 32u;
        lim = min(min(by_t, by_q), 32u);
      }
      mbar_wait(s_full(j & 1u), (j >> 1) & 1u);
      tcgen05_fence_after();
      // the raw scores of this half stay in registers for both passes
      uint32_t v[32];
      tmem_ld_32x32b_x32(t_lane + TMEM_S + (j & 1u) * TK + half * 32u, v);
      tmem_ld_wait();
      if (edge) {
#pragma unroll
        for (uint32_t c = 0; c < 32; ++c)
          if (c >= lim) v[c] = 0xff800000u; // -inf -> p = 0
      }
      // pass 1: maximum of the raw scores (the scale is positive, so it is applied once to the maximum)
      float mx0 = -INFINITY, mx1 = -INFINITY, mx2 = -INFINITY, mx3 = -INFINITY;
#pragma unroll
      for (uint32_t c = 0; c < 32; c += 4) {
        mx0 = fmaxf(mx0, __uint_as_float(v[c]));
        mx1 = fmaxf(mx1, __uint_as_float(v[c + 1]));
        mx2 = fmaxf(mx2, __uint_as_float(v[c + 2]));
        mx3 = fmaxf(mx3, __uint_as_float(v[c + 3]));
      }
      // the row's other half: exchange through shared memory (slot by tile parity: the next write to a slot follows the
      // barrier of the tile in between, which both threads pass only after reading this one)
      const float mh = fmaxf(fmaxf(mx0, mx1), fmaxf(mx2, mx3));
      volatile float *xs = xch + (j & 1u) * (2 * TQ);
      xs[half * TQ + r] = mh;
      asm volatile("bar.sync %0, 64;" ::"r"(1u + quarter) : "memory");
      const float mx = fmaxf(mh, xs[(half ^ 1u) * TQ + r]) * sc;
      float m_new = fmaxf(m, mx);
      if (m_new == -INFINITY) m_new = 0.0f;       // nothing visible yet: every p below is 2^(-inf) = 0
      const float alpha = ex2(m - m_new);         // m = -inf -> 0
      // pass 2: p = 2^(s*sc - m_new), row sum, P -> shared memory (bf16, swizzled K-major): chunks 4 half .. 4 half + 3 of
      // the row's 128 B. P buffer j & 1 was last read by P V_{j-2}, whose o_full the fold of the previous turn waited for.
      float rs0 = 0.0f, rs1 = 0.0f, rs2 = 0.0f, rs3 = 0.0f;
      const uint32_t row_addr = sP + (j & 1u) * P_BYTES + r * 128;
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
#pragma unroll
      for (uint32_t g = 0; g < 4; ++g) {
        __nv_bfloat162 h[4];
#pragma unroll
        for (uint32_t e = 0; e < 4; ++e) h[e] = __floats2bfloat162_rn(pv[g * 8 + 2 * e], pv[g * 8 + 2 * e + 1]);
        st_shared_v4(row_addr + (((half * 4u + g) ^ (r & 7u)) << 4), *reinterpret_cast<const uint4 *>(h));
      }
      l = l * alpha + ((rs0 + rs1) + (rs2 + rs3)); // this half's share of the row sum, on the common maximum
      m = m_new;
      fence_proxy_async();      // generic-proxy smem writes -> visible to the tensor core (async proxy)
      tcgen05_fence_before();
      mbar_arrive(p_ready(j & 1u));
      // the previous tile's product, off the path to P_j (alpha_prev still belongs to tile j - 1)
      if (j > 0) fold(j - 1u);
      alpha_prev = alpha;
    }
    fold(nk - 1u); // last key tile
    tcgen05_fence_before();
    // row sum = the two halves' shares (slot of the parity the last tile did not use)
    volatile float *xs = xch + (nk & 1u) * (2 * TQ);
    xs[half * TQ + r] = l;
    asm volatile("bar.sync %0, 64;" ::"r"(1u + quarter) : "memory");
    l += xs[(half ^ 1u) * TQ + r];
    if (qg < T) {
      const float inv = 1.0f / l;
      float *dst = oc + ((uint64_t)bh * HD + half * 32u) * T + qg; // oc[bh][c][t]: a warp stores 32 adjacent t per column
#pragma unroll
      for (uint32_t c = 0; c < 32; ++c) dst[(uint64_t)c * T] = acc[c] * inv;
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
