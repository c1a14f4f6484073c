// gemm_tc.cu — Blackwell tensor-core matmul: tcgen05.mma (kind::f16, bf16 x bf16 -> fp32) with the
// accumulator in TMEM, operands staged by TMA into 128B-swizzled shared memory, mbarrier pipelines,
// warp-specialised persistent CTAs (one per SM).
//
// Serves rows G1-G3 of SURVEY §8(a): Weed::matmul forward (A MN-major, B K-major), dA = dC*B^T
// (A MN-major, B MN-major) and dB = A^T*dC (A K-major, B K-major) — reference
// src/ops/matmul.cpp:242-279 and src/tensors/tensor.cpp:1361-1400. All three operand-majorness
// cases run on the same kernel; majorness is a template parameter that selects the TMA box shape
// and the UMMA shared-memory descriptor (no transposes are materialised). C is written directly in
// Weed's column-major layout: TMEM lane == row m, so each warp-wide store of one column is a
// coalesced 128-byte line; `accumulate` folds the reference's tmp + add_in_place into the epilogue.
//
// Pipelines:  TMA warp --full[s]--> MMA warp --empty[s]--> TMA warp      (STAGES smem slots)
//             MMA warp --tmem_full[a]--> epilogue warps --tmem_empty[a]--> MMA warp (2 TMEM accs)
// so the epilogue of tile i overlaps the MMAs of tile i+1.
#include "tc_common.cuh"
#include <cstdlib>

namespace weedcu {
namespace tc {

// Warps 4..11 = epilogue, two per scheduler. A warp may only read the TMEM lane quarter warp % 4, so warps w and w + 4
// share a quarter and split the tile's columns: group 0 (warps 4..7) takes the left half of the tile, group 1 (warps 8..11)
// the right half. One epilogue warp per scheduler was latency-bound on its own dependent instruction chain (ncu: 4 warps at
// ~0.3 IPC each; the per-chunk cost did not move with the bytes stored) and that chain, not the stores, was the ~6 us
// "epilogue floor" per 256-wide tile that held the K = 768 products at 0.35-0.6 of the tensor peak.
constexpr uint32_t NUM_THREADS = 384;
constexpr uint32_t EPI_WARP0 = 4;
constexpr uint32_t EPI_GROUPS = 2;
// staging slot s of a tile (the order the store warp drains them): group s % 2, chunk (s / 2) + (s % 2) * (chunks / 2)
__device__ __forceinline__ uint32_t slot_chunk(uint32_t s, uint32_t chunks) { return (s >> 1) + (s & 1u) * (chunks >> 1); }

struct Params {
  float *c;
  uint64_t ldc, c_bs;
  uint32_t M, N, K, batch;
  uint32_t tiles_m, tiles_n;
  int accumulate;
  uint32_t splits, kb_per_split; // split-K: work unit = (tile, k-slice); slices meet in C by TMA reduce-add
  const float *col_bias; // optional [N]: C[m,n] = sum_k A B + col_bias[n]  (Linear::forward's bias add)
  // optional [M, N] column-major with leading dimension ldr: C = (A B + bias) + residual — the `x + Linear(...)` of a
  // transformer block (transformer_encoder_layer.cpp:63-125) without writing the Linear output and re-reading it
  const float *residual;
  uint64_t ldr;
  int tma_store; // C goes out through TMA (needs 16-B aligned base / leading dimension); else direct stores
  // grouped launch: `groups` products that share A (Q / K / V projections of one activation): tile
  // column tn belongs to group tn / tiles_n_group; each group has its own B map, C map / pointer, bias
  uint32_t groups, tiles_n_group;
  float *c_grp[3];
  const float *bias_grp[3];
  // extended epilogue (weedcu_gemm_bf16_ex); all zero = the plain fp32 epilogue above.
  // out_mode 0: fp32 C. 1: bf16 only — C is not written, the staging buffers and the C tensor maps carry bf16 (the operand
  // copy the next GEMM / relayout reads). 2: fp32 C through the staging buffers + a bf16 copy stored directly.
  // act 1: the bf16 copy holds gelu(value) (C keeps the pre-activation the GELU backward needs).
  // stats_kind 1: per row, (mean, M2) of this tile's columns (LayerNorm partials, merged by the consumer);
  // stats_kind 2: per row, (max, sum exp(x - max)) of this tile's columns (log-sum-exp partials of the cross-entropy).
  int out_mode, act, stats_kind;
  __nv_bfloat16 *c16;
  uint64_t ldc16;
  float2 *row_stats; // [tiles_n][M]
  // dynamic tile scheduling (CTA-pair kernel): {next unit, clusters done} of this launch; NULL = unit u goes to cluster
  // u % clusters. With it, a cluster that becomes resident late (its SMs lent to a collective's kernel) finds the queue
  // drained instead of holding a static share of the tiles that everyone else then waits for.
  uint32_t *sched;
};

// per-thread running row statistics across the 32-column chunks of one tile
struct RowStats {
  float a, b, n;
};
// The epilogue warps are latency-bound (one warp per scheduler walking a dependent chain per chunk), so every reduction
// below runs on four independent accumulators: a 32-long serial FADD chain per chunk cost ~20 % of the K = 768 products.
__device__ __forceinline__ void row_stats_update(int kind, RowStats &s, const uint32_t (&r)[32], uint32_t valid) {
  if (valid == 0) return;
  if (kind == 1) { // Chan's parallel update with this chunk's (count, mean, M2)
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) a4[j & 3u] += (j < valid) ? __uint_as_float(r[j]) : 0.0f;
    const float cnt = (float)valid, mean_c = ((a4[0] + a4[1]) + (a4[2] + a4[3])) / cnt;
    float q4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) {
      const float d = __uint_as_float(r[j]) - mean_c;
      q4[j & 3u] += (j < valid) ? d * d : 0.0f;
    }
    const float m2 = (q4[0] + q4[1]) + (q4[2] + q4[3]);
    const float n = s.n + cnt, delta = mean_c - s.a;
    s.a += delta * (cnt / n);
    s.b += m2 + delta * delta * (s.n * cnt / n);
    s.n = n;
  } else { // online (max, sum exp)
    float m4[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) m4[j & 3u] = fmaxf(m4[j & 3u], (j < valid) ? __uint_as_float(r[j]) : -INFINITY);
    const float mx = fmaxf(fmaxf(m4[0], m4[1]), fmaxf(m4[2], m4[3]));
    if (mx > s.a) {
      s.b *= __expf(s.a - mx);
      s.a = mx;
    }
    const float base = s.a;
    float e4[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) e4[j & 3u] += (j < valid) ? __expf(__uint_as_float(r[j]) - base) : 0.0f;
    s.b += (e4[0] + e4[1]) + (e4[2] + e4[3]);
    s.n += (float)valid;
  }
}
// Tensor::gelu (reference src/tensors/tensor.cpp:841-851) for a value that is rounded to bf16 right after: MUFU tanh
// (2^-11 relative) instead of the ~40-instruction tanhf — the four epilogue warps have one issue slot each
__device__ __forceinline__ float gelu_for_bf16(float x) {
  const float k1 = 0.044715f, k2 = 0.7978845608028654f;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k2 * (x + k1 * (x * x) * x)));
  return (0.5f * x) * (1.0f + t);
}
// what the extended epilogue does with one finished chunk (bias / residual already added): statistics, then either the
// bf16 staging tile (out_mode 1; returns true: the fp32 staging store is skipped) or the direct bf16 copy (out_mode 2).
// All values are converted before the first store, and adjacent lanes (rows m, m + 1) swap one value per column pair so
// that every store carries two bf16 rows in one 32-bit word: 16 stores per chunk instead of 32 sub-word ones.
__device__ __forceinline__ bool epilogue_ext_chunk(const Params &p, const uint32_t (&r)[32], RowStats &rs, uint32_t m, uint32_t n_chunk0,
                                                   uint32_t z, uint32_t buf, uint32_t row_in_tile) {
  const uint32_t valid = n_chunk0 < p.N ? min(32u, p.N - n_chunk0) : 0u;
  if (p.stats_kind) row_stats_update(p.stats_kind, rs, r, valid);
  if (p.out_mode == 0) return false;
  float v[32];
  if (p.act) {
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) v[j] = gelu_for_bf16(__uint_as_float(r[j]));
  } else {
#pragma unroll
    for (uint32_t j = 0; j < 32; ++j) v[j] = __uint_as_float(r[j]);
  }
  const uint32_t odd = row_in_tile & 1u;
  uint32_t w[16]; // w[i]: column 2i + odd, rows (m & ~1, m | 1) as bf16x2
#pragma unroll
  for (uint32_t i = 0; i < 16; ++i) {
    const float got = __shfl_xor_sync(0xffffffffu, odd ? v[2 * i] : v[2 * i + 1], 1);
    const __nv_bfloat162 h = odd ? __floats2bfloat162_rn(got, v[2 * i + 1]) : __floats2bfloat162_rn(v[2 * i], got);
    w[i] = *reinterpret_cast<const uint32_t *>(&h);
  }
  if (p.out_mode == 1) {
    const uint32_t dst = buf + (row_in_tile & ~1u) * 2 + odd * (BLOCK_M * 2);
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i) st_shared_f32(dst + i * (2 * BLOCK_M * 2), w[i]);
    return true;
  }
  // out_mode 2: rows (m & ~1, m | 1) of one column are 4 contiguous bytes of the dense bf16 copy (M % 8 == 0)
  if ((m | 1u) < p.M) {
    uint32_t *d16 = reinterpret_cast<uint32_t *>(p.c16 + (uint64_t)z * p.c_bs + (m & ~1u) + (uint64_t)(n_chunk0 + odd) * p.ldc16);
#pragma unroll
    for (uint32_t i = 0; i < 16; ++i)
      if (2 * i + odd < valid) d16[(uint64_t)i * p.ldc16] = w[i]; // (2 columns further = ldc16 32-bit words)
  }
  return false;
}


constexpr uint32_t kMaxGroups = 3;
struct alignas(64) TensorMaps {
  CUtensorMap m[kMaxGroups];
};

constexpr uint32_t EPI_COLS = 32;                             // columns per epilogue chunk
constexpr uint32_t EPI_BUF_BYTES = EPI_COLS * BLOCK_M * 4;    // one [32 cols][128 rows] fp32 staging buffer

// EPI_BUFS staging buffers: the TMA store of chunk i must have finished READING its buffer before chunk
// i + EPI_BUFS may refill it, and that read-completion arrives ~2000 cycles after the issue while the
// TMA unit is busy with the operand loads (ncu: the leader's cp.async.bulk.wait_group.read and the
// other epilogue warps' barrier wait behind it were the top stalls, and the MMA warp spun on
// tmem_empty). Two buffers therefore capped the epilogue at one 16 KB chunk per ~1250 cycles —
// 6.5 us per 128 x 256 tile, longer than the 12 k-blocks of a K = 768 product; four keep 64 KB in flight.
template <uint32_t BLOCK_N, uint32_t STAGES, uint32_t EPI_BUFS> struct SmemLayout {
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = BLOCK_N * BLOCK_K * 2;
  static constexpr uint32_t EPI_OFF = STAGES * (A_BYTES + B_BYTES);
  static constexpr uint32_t BAR_OFF = EPI_OFF + EPI_BUFS * EPI_BUF_BYTES;
  static constexpr uint32_t NUM_BARS = 2 * STAGES + 4 + 2 * EPI_BUFS; // full/empty, tmem full/empty, staging full/empty
  static constexpr uint32_t TOTAL = BAR_OFF + NUM_BARS * 8 + 16;
};

// A_MN / B_MN: operand is MN-major (1) or K-major (0).
template <uint32_t BLOCK_N, uint32_t STAGES, uint32_t EPI_BUFS, int A_MN, int B_MN>
__global__ void __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ TensorMaps tmBs,
                 const __grid_constant__ TensorMaps tmCs, Params p) {
  
  using L = SmemLayout<BLOCK_N, STAGES, EPI_BUFS>;
  constexpr uint32_t NUM_ACC = (2 * BLOCK_N <= 512) ? 2 : 1;
  constexpr uint32_t TMEM_COLS = (NUM_ACC * BLOCK_N <= 32)    ? 32
                                 : (NUM_ACC * BLOCK_N <= 64)  ? 64
                                 : (NUM_ACC * BLOCK_N <= 128) ? 128
                                 : (NUM_ACC * BLOCK_N <= 256) ? 256
                                                              : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  // dynamic smem base is only guaranteed 16-B aligned: round up to the 1024 B the swizzle needs
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + STAGES * L::A_BYTES;
  const uint32_t bars = base + L::BAR_OFF;
  auto full_bar = [&](uint32_t s) { return bars + 8 * s; };
  auto empty_bar = [&](uint32_t s) { return bars + 8 * (STAGES + s); };
  auto tfull_bar = [&](uint32_t a) { return bars + 8 * (2 * STAGES + a); };
  auto tempty_bar = [&](uint32_t a) { return bars + 8 * (2 * STAGES + 2 + a); };
  auto efull_bar = [&](uint32_t b) { return bars + 8 * (2 * STAGES + 4 + b); };            // staging buffer b holds a finished chunk
  auto eempty_bar = [&](uint32_t b) { return bars + 8 * (2 * STAGES + 4 + EPI_BUFS + b); }; // its TMA store has finished reading it
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen_base + L::BAR_OFF + L::NUM_BARS * 8);

  // warp index through a shuffle: provably warp-uniform, so the role loops below stay on the uniform datapath
  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const uint32_t tiles_per_batch = p.tiles_m * p.tiles_n;
  const uint32_t num_tiles = tiles_per_batch * p.batch;
  // unit u -> tile u % num_tiles, k-slice u / num_tiles: CTAs running together work on different
  // tiles of the same slice, so their reduce-adds do not collide
  const uint32_t num_units = num_tiles * p.splits;

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBs.m[0]) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 4 * EPI_GROUPS); // one arrive per epilogue warp
    }
    for (uint32_t b = 0; b < EPI_BUFS; ++b) {
      mbar_init(efull_bar(b), 4);
      mbar_init(eempty_bar(b), 1);
    }
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc(smem_u32((const void *)tmem_slot), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync(); // everything above touched only this CTA's shared memory, TMEM and kernel parameters

  // The producer and MMA loops are run by their WHOLE warp with one elected lane issuing: loop state,
  // barrier addresses and descriptors then live in uniform registers. With `if (lane == 0)` around the
  // loop every operand of every UTMALDG / UTCHMMA went through an R2UR.BROADCAST and the single issuing
  // thread needed ~0.3 us per k-block — as long as the four MMAs of a 128 x 192 tile (ncu: the MMA warp
  // never waited on a barrier, its samples were all issue latency).
  if (warp == 0) {
    // ================================ TMA producer =====================================
    uint32_t stage = 0, phase = 0;
    for (uint32_t unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
      const uint32_t tile = unit % num_tiles, ks = unit / num_tiles;
      const uint32_t z = tile / tiles_per_batch, t = tile % tiles_per_batch;
      const uint32_t tn = t / p.tiles_m, grp = tn / p.tiles_n_group;
      const uint32_t m0 = (t % p.tiles_m) * BLOCK_M, n0 = (tn - grp * p.tiles_n_group) * BLOCK_N;
      const CUtensorMap *tmB = &tmBs.m[grp];
      const uint32_t kb0 = ks * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
      for (uint32_t kb = kb0; kb < kb1; ++kb) {
        mbar_wait_fast(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          mbar_arrive_expect_tx(full_bar(stage), L::A_BYTES + L::B_BYTES);
          const int k0 = (int)(kb * BLOCK_K);
          const uint32_t a_dst = sA + stage * L::A_BYTES, b_dst = sB + stage * L::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (uint32_t i = 0; i < BLOCK_M / 64; ++i)
              tma_load_3d(a_dst + i * (64 * BLOCK_K * 2), &tmA, full_bar(stage), (int)(m0 + 64 * i), k0, (int)z);
          } else {
            tma_load_3d(a_dst, &tmA, full_bar(stage), k0, (int)m0, (int)z);
          }
          if (B_MN) {
#pragma unroll
            for (uint32_t i = 0; i < BLOCK_N / 64; ++i)
              tma_load_3d(b_dst + i * (64 * BLOCK_K * 2), tmB, full_bar(stage), (int)(n0 + 64 * i), k0, (int)z);
          } else {
            tma_load_3d(b_dst, tmB, full_bar(stage), k0, (int)n0, (int)z);
          }
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer ========================================
    constexpr uint32_t idesc = make_idesc(BLOCK_M, BLOCK_N, A_MN, B_MN);
    // K-major: 16 k-elements = 32 B inside the 128-B swizzle row; SBO = 8 rows * 128 B.
    // MN-major: 16 k-rows of 128 B = 2048 B; LBO = next 64-wide MN block (BLOCK_K rows).
    // The descriptor's address field counts 16-byte units: stage and k-step advance it by constants.
    const uint64_t adesc0 = A_MN ? make_smem_desc(sA, BLOCK_K * 128, 1024) : make_smem_desc(sA, 16, 1024);
    const uint64_t bdesc0 = B_MN ? make_smem_desc(sB, BLOCK_K * 128, 1024) : make_smem_desc(sB, 16, 1024);
    constexpr uint32_t A_KSTEP = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4, B_KSTEP = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
    uint32_t stage = 0, phase = 0, it = 0;
    for (uint32_t unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t ks = unit / num_tiles;
      const uint32_t kb0 = ks * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
      const uint32_t acc = it % NUM_ACC, acc_phase = (it / NUM_ACC) & 1;
      mbar_wait_fast(tempty_bar(acc), acc_phase ^ 1); // epilogue has drained this accumulator
      tcgen05_fence_after();
      const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
      for (uint32_t kb = kb0; kb < kb1; ++kb) {
        mbar_wait_fast(full_bar(stage), phase); // TMA bytes have landed
        tcgen05_fence_after();
        if (elect_one()) {
          const uint64_t ad = adesc0 + (uint64_t)(stage * (L::A_BYTES >> 4));
          const uint64_t bd = bdesc0 + (uint64_t)(stage * (L::B_BYTES >> 4));
          umma_f16(d_tmem, ad, bd, idesc, kb != kb0 ? 1u : 0u);
#pragma unroll
          for (uint32_t k = 1; k < BLOCK_K / UMMA_K; ++k) umma_f16(d_tmem, ad + k * A_KSTEP, bd + k * B_KSTEP, idesc, 1u);
          umma_commit(empty_bar(stage)); // frees the smem slot once these MMAs retire
          if (kb == kb1 - 1) umma_commit(tfull_bar(acc));
        }
        __syncwarp();
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
    }
  } else if (warp == 3) {
    // ================================ store warp: staging buffers -> C through TMA ======
    if (lane == 0 && p.tma_store) {
      const uint32_t sEpi = base + L::EPI_OFF;
      uint32_t epi_chunk = 0;
      for (uint32_t unit = blockIdx.x; unit < num_units; unit += gridDim.x) {
        const uint32_t tile = unit % num_tiles;
        const uint32_t z = tile / tiles_per_batch, t = tile % tiles_per_batch;
        const uint32_t tn = t / p.tiles_m, grp = tn / p.tiles_n_group;
        const uint32_t m0 = (t % p.tiles_m) * BLOCK_M, n0 = (tn - grp * p.tiles_n_group) * BLOCK_N;
        const CUtensorMap *tmC = &tmCs.m[grp];
        const bool reduce = p.accumulate || p.splits > 1;
        for (uint32_t sl = 0; sl < BLOCK_N / EPI_COLS; ++sl, ++epi_chunk) {
          const uint32_t c0 = slot_chunk(sl, BLOCK_N / EPI_COLS) * EPI_COLS;
          const uint32_t eb = epi_chunk % EPI_BUFS, buf = sEpi + eb * EPI_BUF_BYTES;
          mbar_wait(efull_bar(eb), (epi_chunk / EPI_BUFS) & 1u);
          if (n0 + c0 < p.N) { // TMA clips rows/columns beyond M/N
            if (reduce) tma_reduce_add_3d(tmC, buf, (int)m0, (int)(n0 + c0), (int)z);
            else tma_store_3d(tmC, buf, (int)m0, (int)(n0 + c0), (int)z);
          }
          bulk_commit(); // (possibly empty) group: keeps the group count in step with the buffer rotation
          if (epi_chunk >= EPI_BUFS - 1) {
            bulk_wait_read<EPI_BUFS - 1>(); // chunk epi_chunk - (EPI_BUFS - 1) has been read out of its buffer
            mbar_arrive(eempty_bar((epi_chunk + 1) % EPI_BUFS));
          }
        }
      }
      bulk_wait_all(); // smem must outlive the last store
    }
  } else if (warp >= EPI_WARP0) {
    // ================================ epilogue: TMEM -> registers -> global C ==========
    const uint32_t q = warp & 3; // TMEM lanes [32q, 32q+32)
    const uint32_t eg = (warp - EPI_WARP0) >> 2; // column half of the tile this warp drains
    constexpr uint32_t CHUNKS = BLOCK_N / EPI_COLS;
    const uint32_t sEpi = base + L::EPI_OFF;
    uint32_t it = 0;
    for (uint32_t unit = blockIdx.x; unit < num_units; unit += gridDim.x, ++it) {
      const uint32_t tile = unit % num_tiles, ks = unit / num_tiles;
      const uint32_t z = tile / tiles_per_batch, t = tile % tiles_per_batch;
      const uint32_t tn = t / p.tiles_m, grp = tn / p.tiles_n_group;
      const uint32_t m0 = (t % p.tiles_m) * BLOCK_M, n0 = (tn - grp * p.tiles_n_group) * BLOCK_N;
      const CUtensorMap *tmC = &tmCs.m[grp];
      const float *col_bias = p.groups > 1 ? p.bias_grp[grp] : p.col_bias;
      float *c_base = p.groups > 1 ? p.c_grp[grp] : p.c;
      const uint32_t acc = it % NUM_ACC, acc_phase = (it / NUM_ACC) & 1;
      const bool add_bias = col_bias && ks == 0; // exactly one k-slice contributes the bias
      const uint32_t m = m0 + q * 32 + lane;
      // residual of the chunk being processed (lane = row: every column is one coalesced 128-byte read per warp); the
      // first chunk's loads are issued before the wait for the accumulator, the next chunk's behind the current staging
      const bool has_res = p.residual && ks == 0 && m < p.M;
      float rv[32];
      auto load_res = [&](uint32_t c0) {
        const float *res = p.residual + (uint64_t)z * p.c_bs + m + (uint64_t)(n0 + c0) * p.ldr;
#pragma unroll
        for (uint32_t j = 0; j < 32; ++j) rv[j] = (n0 + c0 + j < p.N) ? res[(uint64_t)j * p.ldr] : 0.0f;
      };
      if (has_res) load_res(slot_chunk(eg, CHUNKS) * EPI_COLS);
      const bool ext = (p.out_mode | p.stats_kind) != 0;
      RowStats rs = {p.stats_kind == 2 ? -INFINITY : 0.0f, 0.0f, 0.0f};
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
      if (p.tma_store) {
        // TMEM -> registers -> [32 cols][128 rows] fp32 staging tile; the store warp (warp 3) turns each
        // finished buffer into one TMA store (or reduce-add). The four epilogue warps never meet: a warp
        // owns rows [32q, 32q+32) of every buffer and hands over through the buffer's mbarriers.
#pragma unroll 1
        for (uint32_t sl = eg; sl < CHUNKS; sl += EPI_GROUPS) {
          const uint32_t c0 = slot_chunk(sl, CHUNKS) * EPI_COLS, epi_chunk = it * CHUNKS + sl;
          const uint32_t eb = epi_chunk % EPI_BUFS, buf = sEpi + eb * EPI_BUF_BYTES;
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + acc * BLOCK_N + c0, r);
          // the 32 column biases: eight warp-uniform 128-bit loads when aligned and in range (a third fewer epilogue
          // instructions than one value per lane + 32 shuffles), else lane j holds the bias of column c0 + j
          const bool bias_vec = add_bias && n0 + c0 + 32u <= p.N && ((((uintptr_t)(col_bias + n0 + c0)) & 15u) == 0);
          float4 b4[8];
          float bias_lane = 0.0f;
          if (bias_vec) {
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4 *>(col_bias + n0 + c0) + j);
          } else if (add_bias && n0 + c0 + lane < p.N) {
            bias_lane = col_bias[n0 + c0 + lane];
          }
          tmem_ld_wait();
          if (bias_vec) {
#pragma unroll
            for (uint32_t j = 0; j < 8; ++j) {
              r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b4[j].x);
              r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b4[j].y);
              r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b4[j].z);
              r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b4[j].w);
            }
          } else if (add_bias) {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bias_lane, j));
          }
          if (has_res) {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + rv[j]);
            if (sl + EPI_GROUPS < CHUNKS) load_res(c0 + EPI_COLS); // next chunk's residual: in flight behind this chunk's staging
          }
          if (sl + EPI_GROUPS >= CHUNKS) { // this warp's share of the accumulator is read: hand it back to the MMA warp early
            tcgen05_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(tempty_bar(acc));
          }
          mbar_wait(eempty_bar(eb), ((epi_chunk / EPI_BUFS) & 1u) ^ 1u); // the store that last used `buf` has read it
          if (!ext || !epilogue_ext_chunk(p, r, rs, m, n0 + c0, z, buf, q * 32 + lane)) {
            const uint32_t dst = buf + (q * 32 + lane) * 4;
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j) st_shared_f32(dst + j * (BLOCK_M * 4), r[j]);
          }
          fence_proxy_async(); // generic-proxy smem writes -> visible to the TMA (async proxy)
          __syncwarp();
          if (lane == 0) mbar_arrive(efull_bar(eb));
        }
        if (p.stats_kind && m < p.M) p.row_stats[(uint64_t)(tn * EPI_GROUPS + eg) * p.M + m] = make_float2(rs.a, rs.b);
      } else {
        float *crow = c_base + (uint64_t)z * p.c_bs + m;
        const bool full_tile = (m0 + BLOCK_M <= p.M) && (n0 + BLOCK_N <= p.N);
#pragma unroll 1
        for (uint32_t c0 = eg * (BLOCK_N / EPI_GROUPS); c0 < (eg + 1u) * (BLOCK_N / EPI_GROUPS); c0 += 32) {
          uint32_t r[32];
          tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + acc * BLOCK_N + c0, r);
          float bias_lane = 0.0f;
          if (add_bias && n0 + c0 + lane < p.N) bias_lane = col_bias[n0 + c0 + lane];
          tmem_ld_wait();
          if (add_bias) {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bias_lane, j));
          }
          float *dst = crow + (uint64_t)(n0 + c0) * p.ldc;
          if (full_tile && !p.accumulate) {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j, dst += p.ldc) *dst = __uint_as_float(r[j]);
          } else if (m < p.M) {
#pragma unroll
            for (uint32_t j = 0; j < 32; ++j, dst += p.ldc) {
              if (n0 + c0 + j < p.N) {
                const float v = __uint_as_float(r[j]);
                *dst = p.accumulate ? (*dst + v) : v;
              }
            }
          }
        }
        tcgen05_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty_bar(acc));
      }
    }
  }
  tcgen05_fence_before();
  __syncthreads();
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ CTA pairs
// gemm_bf16_pair_kernel: the same pipelines on `tcgen05.mma.cta_group::2`. The two CTAs of a
// cluster (the two SMs of one TPC) own a 256 x BLOCK_N tile: CTA r stages rows [128r, 128r+128) of
// A and columns [r*BLOCK_N/2, ...) of B in its own shared memory, the leader (rank 0) issues one
// M = 256 MMA that reads both halves, and each CTA's TMEM receives its 128 rows x BLOCK_N columns.
// Per k-block an SM now pulls 16 KB + BLOCK_N*64 B through L2 for 128 x BLOCK_N x 64 MACs — 128
// FLOP/B at BLOCK_N = 256 against 85 for the single-CTA 128 x 256 tile, which is what lifts the
// K <= 3072 products off the L2 -> SM bandwidth limit (~6300 B/clk chip-wide).
//   full[s]   : leader's barrier, armed with both CTAs' bytes; the peer's TMA credits it remotely
//   empty[s]  : one per CTA, released by the leader's multicast tcgen05.commit
//   tfull[a]  : one per CTA, multicast commit after a tile's last MMA
//   tempty[a] : leader's barrier, 8 arrivals = 4 epilogue warps x 2 CTAs (peer arrives remotely)
template <uint32_t BLOCK_N, uint32_t STAGES, uint32_t EPI_BUFS> struct PairSmemLayout {
  static constexpr uint32_t HALF_N = BLOCK_N / 2;
  static constexpr uint32_t A_BYTES = BLOCK_M * BLOCK_K * 2;
  static constexpr uint32_t B_BYTES = HALF_N * BLOCK_K * 2;
  static constexpr uint32_t EPI_OFF = STAGES * (A_BYTES + B_BYTES);
  static constexpr uint32_t BAR_OFF = EPI_OFF + EPI_BUFS * EPI_BUF_BYTES;
  static constexpr uint32_t SCHED_SLOTS = 8; // unit ring: the producer runs < 5 units ahead of the slowest role (stages, 2 accumulators, 1 prefetch)
  static constexpr uint32_t NUM_BARS = 2 * STAGES + 4 + 2 * EPI_BUFS + SCHED_SLOTS; // full/empty, tmem full/empty, staging full/empty, unit published
  static constexpr uint32_t TOTAL = BAR_OFF + NUM_BARS * 8 + 16 + SCHED_SLOTS * 4;
};

template <uint32_t BLOCK_N, uint32_t STAGES, uint32_t EPI_BUFS, int A_MN, int B_MN>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(NUM_THREADS, 1)
gemm_bf16_pair_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ TensorMaps tmBs,
                      const __grid_constant__ TensorMaps tmCs, Params p) {
  
  using L = PairSmemLayout<BLOCK_N, STAGES, EPI_BUFS>;
  static_assert(BLOCK_N % 32 == 0 && BLOCK_N <= 256, "pair tile width");
  static_assert(!B_MN || (L::HALF_N % 64 == 0), "an MN-major B half must be whole 64-column swizzle blocks");
  constexpr uint32_t PAIR_M = 2 * BLOCK_M;
  constexpr uint32_t NUM_ACC = 2;
  constexpr uint32_t TMEM_COLS = (NUM_ACC * BLOCK_N <= 256) ? 256 : 512;
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  const uint32_t raw = smem_u32(smem_raw);
  const uint32_t base = (raw + 1023u) & ~1023u;
  uint8_t *gen_base = smem_raw + (base - raw);
  const uint32_t sA = base, sB = base + STAGES * L::A_BYTES;
  const uint32_t bars = base + L::BAR_OFF;
  auto full_bar = [&](uint32_t s) { return bars + 8 * s; };
  auto empty_bar = [&](uint32_t s) { return bars + 8 * (STAGES + s); };
  auto tfull_bar = [&](uint32_t a) { return bars + 8 * (2 * STAGES + a); };
  auto tempty_bar = [&](uint32_t a) { return bars + 8 * (2 * STAGES + 2 + a); };
  auto efull_bar = [&](uint32_t b) { return bars + 8 * (2 * STAGES + 4 + b); };            // staging buffer b holds a finished chunk
  auto eempty_bar = [&](uint32_t b) { return bars + 8 * (2 * STAGES + 4 + EPI_BUFS + b); }; // its TMA store has finished reading it
  auto sfull_bar = [&](uint32_t s) { return bars + 8 * (2 * STAGES + 4 + 2 * EPI_BUFS + s); };           // unit ring entry s is published
  volatile uint32_t *tmem_slot = reinterpret_cast<volatile uint32_t *>(gen_base + L::BAR_OFF + L::NUM_BARS * 8);
  const uint32_t sched_ring = base + L::BAR_OFF + L::NUM_BARS * 8 + 16;
  volatile uint32_t *sched_gen = reinterpret_cast<volatile uint32_t *>(gen_base + L::BAR_OFF + L::NUM_BARS * 8 + 16);

  const uint32_t warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0), lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank(), cid = cluster_id_x(), ncl = num_clusters_x();
  const uint32_t num_kb = (p.K + BLOCK_K - 1) / BLOCK_K;
  const uint32_t tiles_per_batch = p.tiles_m * p.tiles_n;
  const uint32_t num_tiles = tiles_per_batch * p.batch;
  const uint32_t num_units = num_tiles * p.splits;
  // Work units: static (unit it of this cluster = cid + it * clusters) or, with p.sched, drawn from a global counter by the
  // leader's producer lane and handed to every role of both CTAs through a ring in shared memory (entry it & 7, one
  // mbarrier per entry; the slowest role is < 5 entries behind the producer, so no "entry consumed" barrier is needed).
  constexpr uint32_t NO_UNIT = 0xffffffffu;
  const bool dyn = p.sched != nullptr;
  uint32_t first_unit = 0;
  if (dyn && rank == 0 && warp == 0 && lane == 0) first_unit = atomicAdd(p.sched, 1u); // this launch's own counter: no dependency on the previous kernel
  auto unit_at = [&](uint32_t it) -> uint32_t {
    if (!dyn) {
      const uint32_t u = cid + it * ncl;
      return u < num_units ? u : NO_UNIT;
    }
    mbar_wait_cluster(sfull_bar(it & 7u), (it >> 3) & 1u);
    return sched_gen[it & 7u];
  };
  auto publish_unit = [&](uint32_t it, uint32_t u) { // leader's producer lane 0
    const uint32_t val = u < num_units ? u : NO_UNIT, slot = it & 7u;
    sched_gen[slot] = val;
    st_shared_cluster_u32(mapa_shared(sched_ring + 4u * slot, 1), val);
    mbar_arrive_cluster(mapa_shared(sfull_bar(slot), 0));
    mbar_arrive_cluster(mapa_shared(sfull_bar(slot), 1));
  };

  if (warp == 0 && lane == 0) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmA) : "memory");
    asm volatile("prefetch.tensormap [%0];" ::"l"(&tmBs.m[0]) : "memory");
  }
  if (warp == 1 && lane == 0) {
    for (uint32_t s = 0; s < STAGES; ++s) {
      mbar_init(full_bar(s), 1);
      mbar_init(empty_bar(s), 1);
    }
    for (uint32_t a = 0; a < 2; ++a) {
      mbar_init(tfull_bar(a), 1);
      mbar_init(tempty_bar(a), 8 * EPI_GROUPS); // the epilogue warps of both CTAs of the pair
    }
    for (uint32_t b = 0; b < EPI_BUFS; ++b) {
      mbar_init(efull_bar(b), 4);
      mbar_init(eempty_bar(b), 1);
    }
    for (uint32_t s = 0; s < L::SCHED_SLOTS; ++s) mbar_init(sfull_bar(s), 1);
    fence_barrier_init();
  }
  if (warp == 2) tmem_alloc_pair(smem_u32((const void *)tmem_slot), TMEM_COLS);
  tcgen05_fence_before();
  __syncthreads();    // implied by the cluster barrier below; stated for compute-sanitizer, which does not model barrier.cluster
  cluster_sync_all(); // both CTAs' barriers and TMEM exist before anything crosses the pair
  tcgen05_fence_after();
  const uint32_t tmem_base = *tmem_slot;
  pdl_grid_sync(); // everything above touched only the pair's shared memory, TMEM and kernel parameters

  if (warp == 0) {
    // ================================ TMA producer (both CTAs) ==========================
    uint32_t stage = 0, phase = 0, next_unit = 0;
    const bool scheduler = dyn && rank == 0;
    if (scheduler && lane == 0) publish_unit(0, first_unit);
    for (uint32_t it = 0;; ++it) {
      __syncwarp();
      const uint32_t unit = unit_at(it);
      if (unit == NO_UNIT) break;
      if (scheduler && lane == 0) next_unit = atomicAdd(p.sched, 1u); // in flight behind this tile's first loads
      bool publish_pending = scheduler;
      const uint32_t tile = unit % num_tiles, ks = unit / num_tiles;
      const uint32_t z = tile / tiles_per_batch, t = tile % tiles_per_batch;
      const uint32_t tn = t / p.tiles_m, grp = tn / p.tiles_n_group;
      const uint32_t m0 = (t % p.tiles_m) * PAIR_M + rank * BLOCK_M;
      const uint32_t n0 = (tn - grp * p.tiles_n_group) * BLOCK_N + rank * L::HALF_N;
      const CUtensorMap *tmB = &tmBs.m[grp];
      const uint32_t kb0 = ks * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
      for (uint32_t kb = kb0; kb < kb1; ++kb) {
        mbar_wait_fast(empty_bar(stage), phase ^ 1);
        if (elect_one()) {
          if (rank == 0) mbar_arrive_expect_tx(full_bar(stage), 2 * (L::A_BYTES + L::B_BYTES));
          const uint32_t lead_full = mapa_shared(full_bar(stage), 0);
          const int k0 = (int)(kb * BLOCK_K);
          const uint32_t a_dst = sA + stage * L::A_BYTES, b_dst = sB + stage * L::B_BYTES;
          if (A_MN) {
#pragma unroll
            for (uint32_t i = 0; i < BLOCK_M / 64; ++i)
              tma_load_3d_pair(a_dst + i * (64 * BLOCK_K * 2), &tmA, lead_full, (int)(m0 + 64 * i), k0, (int)z);
          } else {
            tma_load_3d_pair(a_dst, &tmA, lead_full, k0, (int)m0, (int)z);
          }
          if (B_MN) {
#pragma unroll
            for (uint32_t i = 0; i < L::HALF_N / 64; ++i)
              tma_load_3d_pair(b_dst + i * (64 * BLOCK_K * 2), tmB, lead_full, (int)(n0 + 64 * i), k0, (int)z);
          } else {
            tma_load_3d_pair(b_dst, tmB, lead_full, k0, (int)n0, (int)z);
          }
        }
        __syncwarp();
        if (publish_pending) { // the next unit reaches the peer while this tile's loads are in flight
          if (lane == 0) publish_unit(it + 1u, next_unit);
          publish_pending = false;
        }
        if (++stage == STAGES) { stage = 0; phase ^= 1; }
      }
      if (publish_pending && lane == 0) publish_unit(it + 1u, next_unit); // (a unit without k-blocks)
    }
    if (scheduler && lane == 0) { // the last cluster to run out of work re-arms this launch's counter slot
      if (atomicAdd(p.sched + 1, 1u) == ncl - 1u) {
        p.sched[0] = 0u;
        p.sched[1] = 0u;
      }
    }
  } else if (warp == 1) {
    // ================================ MMA issuer (leader CTA only) ======================
    if (rank == 0) {
      constexpr uint32_t idesc = make_idesc(PAIR_M, BLOCK_N, A_MN, B_MN);
      const uint64_t adesc0 = A_MN ? make_smem_desc(sA, BLOCK_K * 128, 1024) : make_smem_desc(sA, 16, 1024);
      const uint64_t bdesc0 = B_MN ? make_smem_desc(sB, BLOCK_K * 128, 1024) : make_smem_desc(sB, 16, 1024);
      constexpr uint32_t A_KSTEP = (A_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4, B_KSTEP = (B_MN ? UMMA_K * 128 : UMMA_K * 2) >> 4;
      uint32_t stage = 0, phase = 0;
      for (uint32_t it = 0;; ++it) {
        const uint32_t unit = unit_at(it);
        if (unit == NO_UNIT) break;
        const uint32_t ks = unit / num_tiles;
        const uint32_t kb0 = ks * p.kb_per_split, kb1 = min(num_kb, kb0 + p.kb_per_split);
        const uint32_t acc = it % NUM_ACC, acc_phase = (it / NUM_ACC) & 1;
        mbar_wait_fast(tempty_bar(acc), acc_phase ^ 1); // both CTAs' epilogues have drained this accumulator
        tcgen05_fence_after();
        const uint32_t d_tmem = tmem_base + acc * BLOCK_N;
        for (uint32_t kb = kb0; kb < kb1; ++kb) {
          mbar_wait_fast(full_bar(stage), phase); // both halves have landed
          tcgen05_fence_after();
          if (elect_one()) {
            const uint64_t ad = adesc0 + (uint64_t)(stage * (L::A_BYTES >> 4));
            const uint64_t bd = bdesc0 + (uint64_t)(stage * (L::B_BYTES >> 4));
            umma_f16_pair(d_tmem, ad, bd, idesc, kb != kb0 ? 1u : 0u);
#pragma unroll
            for (uint32_t k = 1; k < BLOCK_K / UMMA_K; ++k) umma_f16_pair(d_tmem, ad + k * A_KSTEP, bd + k * B_KSTEP, idesc, 1u);
            umma_commit_pair(empty_bar(stage)); // frees this slot in both CTAs
            if (kb == kb1 - 1) umma_commit_pair(tfull_bar(acc));
          }
          __syncwarp();
          if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
      }
    }
  } else if (warp == 3) {
    // ================================ store warp (both CTAs, own 128 rows) ==============
    if (lane == 0) {
      const uint32_t sEpi = base + L::EPI_OFF;
      uint32_t epi_chunk = 0;
      for (uint32_t it = 0;; ++it) {
        const uint32_t unit = unit_at(it);
        if (unit == NO_UNIT) break;
        const uint32_t tile = unit % num_tiles;
        const uint32_t z = tile / tiles_per_batch, t = tile % tiles_per_batch;
        const uint32_t tn = t / p.tiles_m, grp = tn / p.tiles_n_group;
        const uint32_t m0 = (t % p.tiles_m) * PAIR_M + rank * BLOCK_M, n0 = (tn - grp * p.tiles_n_group) * BLOCK_N;
        const CUtensorMap *tmC = &tmCs.m[grp];
        const bool reduce = p.accumulate || p.splits > 1;
        for (uint32_t sl = 0; sl < BLOCK_N / EPI_COLS; ++sl, ++epi_chunk) {
          const uint32_t c0 = slot_chunk(sl, BLOCK_N / EPI_COLS) * EPI_COLS;
          const uint32_t eb = epi_chunk % EPI_BUFS, buf = sEpi + eb * EPI_BUF_BYTES;
          mbar_wait(efull_bar(eb), (epi_chunk / EPI_BUFS) & 1u);
          if (n0 + c0 < p.N && m0 < p.M) {
            if (reduce) tma_reduce_add_3d(tmC, buf, (int)m0, (int)(n0 + c0), (int)z);
            else tma_store_3d(tmC, buf, (int)m0, (int)(n0 + c0), (int)z);
          }
          bulk_commit();
          if (epi_chunk >= EPI_BUFS - 1) {
            bulk_wait_read<EPI_BUFS - 1>();
            mbar_arrive(eempty_bar((epi_chunk + 1) % EPI_BUFS));
          }
        }
      }
      bulk_wait_all();
    }
  } else if (warp >= EPI_WARP0) {
    // ================================ epilogue (both CTAs, own 128 rows) ================
    const uint32_t q = warp & 3;
    const uint32_t eg = (warp - EPI_WARP0) >> 2; // column half of the tile this warp drains
    constexpr uint32_t CHUNKS = BLOCK_N / EPI_COLS;
    const uint32_t sEpi = base + L::EPI_OFF;
    for (uint32_t it = 0;; ++it) {
      const uint32_t unit = unit_at(it);
      if (unit == NO_UNIT) break;
      const uint32_t tile = unit % num_tiles, ks = unit / num_tiles;
      const uint32_t z = tile / tiles_per_batch, t = tile % tiles_per_batch;
      const uint32_t tn = t / p.tiles_m, grp = tn / p.tiles_n_group;
      const uint32_t n0 = (tn - grp * p.tiles_n_group) * BLOCK_N;
      const uint32_t m = (t % p.tiles_m) * PAIR_M + rank * BLOCK_M + q * 32 + lane; // this thread's row of C
      const float *col_bias = p.groups > 1 ? p.bias_grp[grp] : p.col_bias;
      const uint32_t acc = it % NUM_ACC, acc_phase = (it / NUM_ACC) & 1;
      const bool add_bias = col_bias && ks == 0;
      const bool has_res = p.residual && ks == 0 && m < p.M;
      float rv[32];
      auto load_res = [&](uint32_t c0) {
        const float *res = p.residual + (uint64_t)z * p.c_bs + m + (uint64_t)(n0 + c0) * p.ldr;
#pragma unroll
        for (uint32_t j = 0; j < 32; ++j) rv[j] = (n0 + c0 + j < p.N) ? res[(uint64_t)j * p.ldr] : 0.0f;
      };
      if (has_res) load_res(slot_chunk(eg, CHUNKS) * EPI_COLS);
      const bool ext = (p.out_mode | p.stats_kind) != 0;
      RowStats rs = {p.stats_kind == 2 ? -INFINITY : 0.0f, 0.0f, 0.0f};
      mbar_wait(tfull_bar(acc), acc_phase);
      tcgen05_fence_after();
#pragma unroll 1
      for (uint32_t sl = eg; sl < CHUNKS; sl += EPI_GROUPS) {
        const uint32_t c0 = slot_chunk(sl, CHUNKS) * EPI_COLS, epi_chunk = it * CHUNKS + sl;
        const uint32_t eb = epi_chunk % EPI_BUFS, buf = sEpi + eb * EPI_BUF_BYTES;
        uint32_t r[32];
        tmem_ld_32x32b_x32(tmem_base + ((q * 32) << 16) + acc * BLOCK_N + c0, r);
        const bool bias_vec = add_bias && n0 + c0 + 32u <= p.N && ((((uintptr_t)(col_bias + n0 + c0)) & 15u) == 0);
        float4 b4[8];
        float bias_lane = 0.0f;
        if (bias_vec) {
#pragma unroll
          for (uint32_t j = 0; j < 8; ++j) b4[j] = __ldg(reinterpret_cast<const float4 *>(col_bias + n0 + c0) + j);
        } else if (add_bias && n0 + c0 + lane < p.N) {
          bias_lane = col_bias[n0 + c0 + lane];
        }
        tmem_ld_wait();
        if (bias_vec) {
#pragma unroll
          for (uint32_t j = 0; j < 8; ++j) {
            r[4 * j + 0] = __float_as_uint(__uint_as_float(r[4 * j + 0]) + b4[j].x);
            r[4 * j + 1] = __float_as_uint(__uint_as_float(r[4 * j + 1]) + b4[j].y);
            r[4 * j + 2] = __float_as_uint(__uint_as_float(r[4 * j + 2]) + b4[j].z);
            r[4 * j + 3] = __float_as_uint(__uint_as_float(r[4 * j + 3]) + b4[j].w);
          }
        } else if (add_bias) {
#pragma unroll
          for (uint32_t j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + __shfl_sync(0xffffffffu, bias_lane, j));
        }
        if (has_res) {
#pragma unroll
          for (uint32_t j = 0; j < 32; ++j) r[j] = __float_as_uint(__uint_as_float(r[j]) + rv[j]);
          if (sl + EPI_GROUPS < CHUNKS) load_res(c0 + EPI_COLS);
        }
        if (sl + EPI_GROUPS >= CHUNKS) { // this warp's share of the accumulator is read: hand it back to the leader's MMA warp
          tcgen05_fence_before();
          __syncwarp();
          if (lane == 0) mbar_arrive_cluster(mapa_shared(tempty_bar(acc), 0));
        }
        mbar_wait(eempty_bar(eb), ((epi_chunk / EPI_BUFS) & 1u) ^ 1u);
        if (!ext || !epilogue_ext_chunk(p, r, rs, m, n0 + c0, z, buf, q * 32 + lane)) {
          const uint32_t dst = buf + (q * 32 + lane) * 4;
#pragma unroll
          for (uint32_t j = 0; j < 32; ++j) st_shared_f32(dst + j * (BLOCK_M * 4), r[j]);
        }
        fence_proxy_async();
        __syncwarp();
        if (lane == 0) mbar_arrive(efull_bar(eb));
      }
      if (p.stats_kind && m < p.M) p.row_stats[(uint64_t)(tn * EPI_GROUPS + eg) * p.M + m] = make_float2(rs.a, rs.b);
    }
  }
  tcgen05_fence_before();
  cluster_sync_all(); // the peer's shared memory and TMEM stay alive until the leader's last MMA is consumed
  if (warp == 2) {
    tcgen05_fence_after();
    tmem_dealloc_pair(tmem_base, TMEM_COLS);
  }
}

// ------------------------------------------------------------------------------------ host side
typedef CUresult (*EncodeTiledFn)(CUtensorMap *, CUtensorMapDataType, cuuint32_t, void *, const cuuint64_t *,
                                  const cuuint64_t *, const cuuint32_t *, const cuuint32_t *,
                                  CUtensorMapInterleave, CUtensorMapSwizzle, CUtensorMapL2promotion,
                                  CUtensorMapFloatOOBfill);
static EncodeTiledFn get_encode_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void *ptr = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &ptr, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = (EncodeTiledFn)ptr;
  }
  return fn;
}

// bf16 matrix [mn, k] per batch. major 0: k contiguous (ld = stride of mn index);
// major 1: mn contiguous (ld = stride of k index). box_mn rows of the MN index per TMA box.
int make_operand_map(CUtensorMap *map, const uint16_t *ptr, int major, uint64_t mn, uint64_t k,
                            uint64_t ld, uint64_t batch, uint64_t batch_stride, uint32_t box_mn) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return WEEDCU_ENOSUP;
  if ((((uintptr_t)ptr) & 15u) || (ld % 8) || (batch > 1 && (batch_stride % 8))) return WEEDCU_ENOSUP;
  cuuint64_t dims[3], strides[2];
  cuuint32_t box[3], estr[3] = {1, 1, 1};
  if (major == 0) {
    dims[0] = k; dims[1] = mn;
    box[0] = BLOCK_K; box[1] = box_mn;
  } else {
    dims[0] = mn; dims[1] = k;
    box[0] = 64; box[1] = BLOCK_K;
  }
  dims[2] = batch;
  box[2] = 1;
  strides[0] = ld * 2;
  strides[1] = (batch > 1 ? batch_stride : (uint64_t)ld * dims[1]) * 2; // ld % 8 == 0 => 16-B multiple
  const CUresult r = enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)ptr, dims, strides, box, estr,
                         CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                         CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  return r == CUDA_SUCCESS ? 0 : WEEDCU_ENOSUP;
}

int make_plain_map_2d(CUtensorMap *map, const void *ptr, int elem_bytes, uint64_t inner, uint64_t outer, uint64_t outer_stride_bytes,
                      uint32_t box_inner, uint32_t box_outer) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return WEEDCU_ENOSUP;
  if ((((uintptr_t)ptr) & 15u) || (outer_stride_bytes % 16) || ((box_inner * (uint32_t)elem_bytes) % 16) || box_inner > 256 || box_outer > 256)
    return WEEDCU_ENOSUP;
  const CUtensorMapDataType dt = elem_bytes == 4 ? CU_TENSOR_MAP_DATA_TYPE_FLOAT32 : CU_TENSOR_MAP_DATA_TYPE_BFLOAT16;
  cuuint64_t dims[2] = {inner, outer}, strides[1] = {outer_stride_bytes};
  cuuint32_t box[2] = {box_inner, box_outer}, estr[2] = {1, 1};
  return enc(map, dt, 2, const_cast<void *>(ptr), dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE,
             CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS
             ? 0
             : WEEDCU_ENOSUP;
}

// fp32 C [M, N] (+batch) column-major with leading dimension ldc: TMA box = 128 rows x 32 columns,
// no swizzle (the staging tile in shared memory is plain [col][row]). Needs a 16-B aligned base and
// 16-B multiples for the column / batch strides; otherwise the kernel stores C directly.
static bool make_c_map(CUtensorMap *map, float *c, uint64_t M, uint64_t N, uint64_t ldc, uint64_t batch, uint64_t c_bs) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  if ((((uintptr_t)c) & 15u) || (ldc % 4) || (batch > 1 && (c_bs % 4))) return false;
  cuuint64_t dims[3] = {M, N, batch};
  cuuint64_t strides[2] = {ldc * 4, (batch > 1 ? c_bs : ldc * N) * 4};
  cuuint32_t box[3] = {BLOCK_M, EPI_COLS, 1}, estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 3, (void *)c, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// the same box over a bf16 C (out_mode 1): the staging tile is [32 cols][128 rows] of 2-byte elements
static bool make_c16_map(CUtensorMap *map, uint16_t *c, uint64_t M, uint64_t N, uint64_t ldc, uint64_t batch, uint64_t c_bs) {
  EncodeTiledFn enc = get_encode_fn();
  if (!enc) return false;
  if ((((uintptr_t)c) & 15u) || (ldc % 8) || (batch > 1 && (c_bs % 8))) return false;
  cuuint64_t dims[3] = {M, N, batch};
  cuuint64_t strides[2] = {ldc * 2, (batch > 1 ? c_bs : ldc * N) * 2};
  cuuint32_t box[3] = {BLOCK_M, EPI_COLS, 1}, estr[3] = {1, 1, 1};
  return enc(map, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 3, (void *)c, dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_NONE, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

// SMs the persistent product kernels may occupy (WEEDCU_GEMM_SMS, default all): a data-parallel run can leave a few SMs to
// the collective's kernels, whose CTAs do not fit next to a 224 KB product CTA
static unsigned gemm_sm_limit() {
  static const unsigned lim = [] {
    const char *e = getenv("WEEDCU_GEMM_SMS");
    const int v = e ? atoi(e) : 0;
    return (v >= 2 && v <= kNumSMs) ? (unsigned)(v & ~1) : (unsigned)kNumSMs;
  }();
  return lim;
}
template <uint32_t BLOCK_N, uint32_t STAGES, uint32_t EPI_BUFS>
static int launch_cfg(const CUtensorMap &tmA, const TensorMaps &tmBs, const TensorMaps &tmCs, const Params &p, int a_major,
                      int b_major, cudaStream_t st) {
  using L = SmemLayout<BLOCK_N, STAGES, EPI_BUFS>;
  static_assert(L::TOTAL + 1024 <= 232448, "shared memory budget (227 KB per CTA)");
  const uint32_t smem = L::TOTAL + 1024; // slack for the 1024-B round-up
  const uint32_t num_tiles = p.tiles_m * p.tiles_n * p.batch;
  const uint32_t num_units = num_tiles * p.splits;
  const unsigned grid = num_units < gemm_sm_limit() ? num_units : gemm_sm_limit();
#define WCU_TC_LAUNCH(AM, BM_)                                                                     \
  {                                                                                                \
    auto k = gemm_bf16_kernel<BLOCK_N, STAGES, EPI_BUFS, AM, BM_>;                                           \
    ensure_dynamic_smem((const void *)k, (int)smem);                                               \
    launch_k(k, dim3(grid), dim3(NUM_THREADS), smem, st, tmA, tmBs, tmCs, p);                                        \
  }
  if (a_major && b_major) WCU_TC_LAUNCH(1, 1)
  else if (a_major) WCU_TC_LAUNCH(1, 0)
  else if (b_major) WCU_TC_LAUNCH(0, 1)
  else WCU_TC_LAUNCH(0, 0)
#undef WCU_TC_LAUNCH
  return after_launch();
}

// Counter slots of the dynamic tile scheduler: {next unit, clusters done} per launch, 256 slots per device used in turn (a
// slot is re-armed by the last cluster of the launch that used it, long before its turn comes again). WEEDCU_GEMM_DYNAMIC /
// weedcu_gemm_set_dynamic: 0 = static striding, 1 = dynamic.
static int g_gemm_dynamic = [] {
  const char *e = getenv("WEEDCU_GEMM_DYNAMIC");
  return e ? atoi(e) : 0;
}();
static uint32_t *next_sched_slot() {
  constexpr int kSlots = 256, kMaxDev = 16;
  static uint32_t *slots[kMaxDev] = {nullptr};
  static unsigned turn[kMaxDev] = {0};
  int dev = 0;
  if (cudaGetDevice(&dev) != cudaSuccess || dev < 0 || dev >= kMaxDev) return nullptr;
  if (!slots[dev]) {
    if (cudaMalloc((void **)&slots[dev], sizeof(uint32_t) * 2 * kSlots) != cudaSuccess) return nullptr;
    if (cudaMemset(slots[dev], 0, sizeof(uint32_t) * 2 * kSlots) != cudaSuccess) return nullptr; // (synchronising: once per device)
  }
  return slots[dev] + 2 * (turn[dev]++ % kSlots);
}

template <uint32_t BLOCK_N, uint32_t STAGES, uint32_t EPI_BUFS>
static int launch_pair_cfg(const CUtensorMap &tmA, const TensorMaps &tmBs, const TensorMaps &tmCs, const Params &p, int a_major,
                           int b_major, cudaStream_t st) {
  using L = PairSmemLayout<BLOCK_N, STAGES, EPI_BUFS>;
  static_assert(L::TOTAL + 1024 <= 232448, "shared memory budget (227 KB per CTA)");
  const uint32_t smem = L::TOTAL + 1024;
  const uint32_t num_units = p.tiles_m * p.tiles_n * p.batch * p.splits;
  const unsigned pairs = num_units < gemm_sm_limit() / 2 ? num_units : gemm_sm_limit() / 2;
  Params pd = p;
  pd.sched = (g_gemm_dynamic && num_units > pairs) ? next_sched_slot() : nullptr; // (one unit per cluster: nothing to balance)
#define WCU_TC_LAUNCH(AM, BM_)                                                                     \
  {                                                                                                \
    auto k = gemm_bf16_pair_kernel<BLOCK_N, STAGES, EPI_BUFS, AM, BM_>;                                      \
    ensure_dynamic_smem((const void *)k, (int)smem);                                               \
    launch_k(k, dim3(2 * pairs), dim3(NUM_THREADS), smem, st, tmA, tmBs, tmCs, pd);                                  \
  }
  if constexpr ((BLOCK_N / 2) % 64 == 0) {
    if (a_major && b_major) WCU_TC_LAUNCH(1, 1)
    else if (a_major) WCU_TC_LAUNCH(1, 0)
    else if (b_major) WCU_TC_LAUNCH(0, 1)
    else WCU_TC_LAUNCH(0, 0)
  } else { // 96-column halves exist for a K-major B only
    if (b_major) return WEEDCU_ENOSUP;
    if (a_major) WCU_TC_LAUNCH(1, 0)
    else WCU_TC_LAUNCH(0, 0)
  }
#undef WCU_TC_LAUNCH
  return after_launch();
}

// Tile configuration: single-CTA 128 x {256,192,128} tiles or CTA-pair 256 x {256,192,128} tiles, and
// the split-K factor. mode 0 = cost model over both families, 1 = single-CTA only, 2 = pairs only,
// >= 1000 = forced (pair * 1e6 + BLOCK_N * 1e3 + splits) for the tuning sweep of tools/microbench.py.
static int g_gemm_mode = -1;
static int gemm_mode() {
  if (g_gemm_mode < 0) {
    const char *e = getenv("WEEDCU_GEMM_MODE");
    g_gemm_mode = e ? atoi(e) : 0;
  }
  return g_gemm_mode;
}
void set_gemm_mode(int mode) { g_gemm_mode = mode < 0 ? 0 : mode; }

static int g_last_block_n = 0; // tile width of the most recent launch (reported with the row statistics)
int last_block_n() { return g_last_block_n; }
struct TileCfg {
  int pair;
  uint32_t bn, splits;
  int variant; // tuning sweeps: 1 = direct stores instead of staged TMA stores (single-CTA kernels)
};

// Estimated launch duration in microseconds, calibrated on the tile-configuration sweep of
// tools/microbench.py --group tune (profiles/r01_tune_v12.log). A CTA runs its work units (tile,
// k-slice) back to back with the epilogue of unit i hidden behind the k-loop of unit i+1, so a wave
// costs max(k-loop, epilogue). Measured per-k-block times: the single-CTA kernels are bound by shared-
// memory bandwidth (TMA fills plus the MMA's operand reads: 96 KB per 128 x 256 x 64 block against
// 128 B/clk), the CTA pairs halve the B bytes per SM and run at the tensor pipe's sustained rate for
// BLOCK_N >= 192. The epilogue floor is the ~0.75 us a 16 KB chunk takes through shared memory.
static double tile_cost_us(const TileCfg &c, uint32_t tiles_m128, uint32_t N, uint32_t num_kb, uint32_t groups, uint32_t batch,
                           int accumulate, uint64_t c_elems) {
  const uint32_t tiles_m = c.pair ? (tiles_m128 + 1) / 2 : tiles_m128;
  const uint32_t tiles = tiles_m * ((N + c.bn - 1) / c.bn) * groups * batch;
  const uint32_t kb_per = (num_kb + c.splits - 1) / c.splits, units = tiles * ((num_kb + kb_per - 1) / kb_per);
  const uint32_t slots = c.pair ? kNumSMs / 2 : kNumSMs;
  const uint32_t waves = (units + slots - 1) / slots;
  const double t_kb = c.pair ? (c.bn == 256 ? 0.417 : c.bn == 192 ? 0.334 : 0.267) : (c.bn == 256 ? 0.4425 : c.bn == 192 ? 0.385 : 0.296);
  // `accumulate` is deliberately NOT an input: C = AB into a zero-filled buffer and C += AB must pick the
  // same tiles and k-slices so that the lazy zero-fill of gradients stays a pure scheduling change
  // (tests/test_host_gpu.py::test_operand_cache_and_lazy_zero_change_nothing).
  (void)accumulate;
  // per-tile epilogue floor measured on the LM-head product (N = 50257, K = 768): 6.04 / 4.66 / 3.39 us for CTA-pair
  // tiles of width 256 / 192 / 128, 6.69 / 5.29 / 3.90 us for single-CTA tiles
  const double epi_us = c.pair ? (c.bn == 256 ? 6.04 : c.bn == 192 ? 4.66 : 3.39) : (c.bn == 256 ? 6.69 : c.bn == 192 ? 5.29 : 3.90);
  const double t_epi = epi_us * (c.splits > 1 ? 1.25 : 1.0);
  const double loop = kb_per * t_kb;
  double us = waves * (loop > t_epi ? loop : t_epi) + 4.0 + 0.5 * (loop < t_epi ? loop : t_epi);
  if (c.splits > 1) us += 3.0 + (double)c_elems * 4.0 / 5.0e6; // zero-fill before the reduce-adds
  return us;
}

// `groups` (<= 3) products A x B_g -> C_g (+ bias_g) that share the A operand and every dimension run
// as ONE launch: their tiles join one persistent tile loop, so the ~10 us of per-launch prologue /
// exposed last epilogue is paid once and the tail wave is filled by the other groups' tiles.
// extended epilogue of one launch (see Params): c16[g] = bf16 output of group g
struct EpiExt {
  int out_mode = 0, act = 0, stats_kind = 0;
  uint16_t *c16[kMaxGroups] = {nullptr, nullptr, nullptr};
  uint64_t ldc16 = 0;
  float *row_stats = nullptr;
  uint32_t stats_capacity_tiles = 0, *stats_tiles = nullptr;
};
int launch_gemm_bf16_grouped(const uint16_t *a, int a_major, uint64_t lda, uint64_t a_bs, uint32_t groups,
                             const uint16_t *const *b, int b_major, uint64_t ldb, uint64_t b_bs, float *const *c, uint64_t ldc,
                             uint64_t c_bs, uint32_t M, uint32_t N, uint32_t K, uint32_t batch, int accumulate, cudaStream_t st,
                             const float *const *col_bias, const float *residual, uint64_t ldr, const EpiExt *ext = nullptr) {
  if (!a || !b || !M || !N || !K || !batch || !groups || groups > kMaxGroups) return WEEDCU_EINVAL;
  const bool bf16_only = ext && ext->out_mode == 1;
  if (!c && !bf16_only) return WEEDCU_EINVAL;
  if (residual && (groups != 1 || batch != 1 || accumulate)) return WEEDCU_EINVAL;
  if (ext) {
    if (accumulate || batch != 1) return WEEDCU_EINVAL;
    if ((ext->stats_kind || ext->out_mode == 2) && groups != 1) return WEEDCU_EINVAL;
    if (ext->out_mode && (ext->ldc16 % 8)) return WEEDCU_ENOSUP;
    for (uint32_t g = 0; g < groups; ++g)
      if (ext->out_mode && (!ext->c16[g] || (((uintptr_t)ext->c16[g]) & 15u))) return WEEDCU_EINVAL;
    if (ext->stats_kind && (!ext->row_stats || !ext->stats_tiles)) return WEEDCU_EINVAL;
  }
  for (uint32_t g = 0; g < groups; ++g)
    if (!b[g] || (!bf16_only && !c[g])) return WEEDCU_EINVAL;
  // Tile family, tile width and split-K are chosen together by the cost model above. Few-tile
  // problems (weight gradients: M, N = layer widths, K = batch*seq) get split along K, slices
  // meeting in C by TMA reduce-add; tile counts just above a multiple of the slot count get a
  // narrower tile instead of a nearly empty last wave.
  const uint32_t num_kb = (K + BLOCK_K - 1) / BLOCK_K, tiles_m128 = (M + BLOCK_M - 1) / BLOCK_M;
  bool c_tma = bf16_only || ((ldc % 4) == 0 && (batch == 1 || (c_bs % 4) == 0));
  for (uint32_t g = 0; g < groups && !bf16_only; ++g) c_tma = c_tma && (((uintptr_t)c[g]) & 15u) == 0;
  if (ext && !c_tma) return WEEDCU_ENOSUP; // the extended epilogue lives in the staged (TMA) path only
  const int mode = gemm_mode();
  TileCfg best = {0, 256, 1, 0};
  if (mode >= 1000) {
    best.pair = mode / 1000000;
    best.bn = (uint32_t)(mode / 1000) % 1000u;
    best.variant = (mode / 100) % 10;
    best.splits = (uint32_t)mode % 100u;
    if (best.pair > 1 || (best.bn != 256 && best.bn != 192 && best.bn != 128) || !best.splits) return WEEDCU_EINVAL;
    if (best.pair && (!c_tma || (best.bn == 192 && b_major))) return WEEDCU_ENOSUP;
    if (!c_tma) best.splits = 1;
  } else {
    double best_cost = 1e300;
    const uint32_t cand[3] = {256, 192, 128};
    for (int pair = 0; pair < 2; ++pair) {
      if ((pair == 0 && mode == 2 && c_tma && M > BLOCK_M) || (pair == 1 && (mode == 1 || !c_tma || M <= BLOCK_M))) continue;
      for (uint32_t ci = 0; ci < 3; ++ci) {
        const uint32_t bn = cand[ci];
        if (bn > 128 && N <= bn - 64) continue; // do not pad a narrow N into a wide tile
        if (pair && bn == 192 && b_major) continue;
        for (uint32_t sp = 1; sp <= 16; ++sp) {
          if (sp > 1 && (!c_tma || ext || sp * 4u > num_kb)) break; // (statistics / converted outputs need whole dot products)
          const TileCfg cfg = {pair, bn, sp, 0};
          const double cost = tile_cost_us(cfg, tiles_m128, N, num_kb, groups, batch, accumulate, (uint64_t)M * N * groups * batch);
          if (cost < best_cost) {
            best_cost = cost;
            best = cfg;
          }
        }
      }
    }
  }
  const uint32_t block_n = best.bn;
  g_last_block_n = (int)block_n;
  uint32_t best_s = best.splits;
  const uint32_t tiles_m = best.pair ? (tiles_m128 + 1) / 2 : tiles_m128;
  CUtensorMap tmA;
  TensorMaps tmBs, tmCs;
  int rc = make_operand_map(&tmA, a, a_major, M, K, lda, batch, a_bs, BLOCK_M);
  if (rc) return rc;
  if (ext && ext->stats_kind && EPI_GROUPS * ((N + block_n - 1) / block_n) > ext->stats_capacity_tiles) return WEEDCU_EINVAL;
  if (ext) best_s = 1;
  Params p;
  p.c = bf16_only ? nullptr : c[0];
  p.ldc = ldc;
  p.out_mode = ext ? ext->out_mode : 0;
  p.act = ext ? ext->act : 0;
  p.stats_kind = ext ? ext->stats_kind : 0;
  p.c16 = ext ? (__nv_bfloat16 *)ext->c16[0] : nullptr;
  p.ldc16 = ext ? ext->ldc16 : 0;
  p.row_stats = ext ? (float2 *)ext->row_stats : nullptr;
  p.sched = nullptr;
  p.c_bs = c_bs;
  p.M = M; p.N = N; p.K = K; p.batch = batch;
  p.tiles_m = tiles_m;
  p.tiles_n_group = (N + block_n - 1) / block_n;
  p.tiles_n = p.tiles_n_group * groups;
  p.groups = groups;
  p.accumulate = accumulate;
  p.col_bias = col_bias ? col_bias[0] : nullptr;
  p.residual = residual;
  p.ldr = ldr;
  p.tma_store = 1;
  for (uint32_t g = 0; g < kMaxGroups; ++g) {
    const uint32_t src = g < groups ? g : 0;
    rc = make_operand_map(&tmBs.m[g], b[src], b_major, N, K, ldb, batch, b_bs, best.pair ? block_n / 2 : block_n);
    if (rc) return rc;
    if (bf16_only) {
      if (!make_c16_map(&tmCs.m[g], ext->c16[src], M, N, ext->ldc16, batch, c_bs)) return WEEDCU_ENOSUP;
    } else if (p.tma_store && !make_c_map(&tmCs.m[g], c[src], M, N, ldc, batch, c_bs))
      p.tma_store = 0;
    p.c_grp[g] = bf16_only ? nullptr : c[src];
    p.bias_grp[g] = col_bias ? col_bias[src] : nullptr;
  }
  if (best.variant == 1 && !best.pair && !ext) p.tma_store = 0;
  if (!p.tma_store && (residual || ext)) return WEEDCU_ENOSUP; // the residual add / extended epilogue live in the staged epilogue only
  if (ext && ext->stats_tiles) *ext->stats_tiles = p.tiles_n * EPI_GROUPS; // every column tile leaves one partial per epilogue group
  if (!p.tma_store) {
    if (best.pair) return WEEDCU_ENOSUP; // unreachable: pairs are only chosen when C meets the TMA rules
    for (uint32_t g = 0; g < kMaxGroups; ++g) tmCs.m[g] = tmA; // unused by the kernel, but must be valid descriptors
    best_s = 1;
  }
  p.kb_per_split = (num_kb + best_s - 1) / best_s;
  p.splits = (num_kb + p.kb_per_split - 1) / p.kb_per_split;
  if (p.splits > 1 && !accumulate) { // slices meet by reduce-add: C starts from zero
    for (uint32_t g = 0; g < groups; ++g) {
      if (batch == 1 && ldc == M) {
        // our own fill kernel, not cudaMemsetAsync: it chains with the launches around it (programmatic dependent
        // launch), a memset node would force plain launches on both sides
        const int rc = weedcu_fill_real(c[g], (uint64_t)M * N, 0.0f, (void *)st);
        if (rc) return rc;
      } else {
        note_stream_op();
        for (uint32_t z = 0; z < batch; ++z)
          WCU_CHECK(cudaMemset2DAsync(c[g] + (uint64_t)z * c_bs, ldc * sizeof(float), 0, (size_t)M * sizeof(float), N, st));
      }
    }
  }
  ProfScope prof(WEEDCU_PROF_GEMM_TC, st, 2.0 * (double)M * N * K * batch * groups);
  if (best.pair) {
    if (block_n == 256) return launch_pair_cfg<256, 5, 4>(tmA, tmBs, tmCs, p, a_major, b_major, st);
    if (block_n == 192) return launch_pair_cfg<192, 5, 4>(tmA, tmBs, tmCs, p, a_major, b_major, st);
    return launch_pair_cfg<128, 6, 4>(tmA, tmBs, tmCs, p, a_major, b_major, st);
  }
  if (block_n == 256) return launch_cfg<256, 4, 2>(tmA, tmBs, tmCs, p, a_major, b_major, st); // 3 stages + 4 buffers measured slower
  if (block_n == 192) return launch_cfg<192, 4, 4>(tmA, tmBs, tmCs, p, a_major, b_major, st);
  return launch_cfg<128, 5, 4>(tmA, tmBs, tmCs, p, a_major, b_major, st);
}

int launch_gemm_bf16(const uint16_t *a, int a_major, uint64_t lda, uint64_t a_bs, const uint16_t *b,
                     int b_major, uint64_t ldb, uint64_t b_bs, float *c, uint64_t ldc, uint64_t c_bs,
                     uint32_t M, uint32_t N, uint32_t K, uint32_t batch, int accumulate, cudaStream_t st,
                     const float *col_bias) {
  if (!a || !b || !c) return WEEDCU_EINVAL;
  const uint16_t *bs[1] = {b};
  float *cs[1] = {c};
  const float *biases[1] = {col_bias};
  return launch_gemm_bf16_grouped(a, a_major, lda, a_bs, 1, bs, b_major, ldb, b_bs, cs, ldc, c_bs, M, N, K, batch, accumulate, st,
                                  col_bias ? biases : nullptr, nullptr, 0);
}

} // namespace tc

// ------------------------------------------------------------------------------------ packing
// fp32 strided [rows, cols] (+batch) -> dense bf16 with leading dimension ld.
// dst_major 1: dst[r + c*ld] (rows contiguous); 0: dst[c + r*ld] (cols contiguous).
// 32x32 tiles through shared memory so both the fp32 read and the bf16 write walk their own
// contiguous index with adjacent lanes.
__global__ void __launch_bounds__(256)
pack_bf16_kernel(const float *__restrict__ src, uint64_t s_bs, uint32_t s0, uint32_t s1, uint32_t rows,
                 uint32_t cols, __nv_bfloat16 *__restrict__ dst, uint64_t d_bs, uint64_t ld, int dst_major,
                 int src_rowfast) {
  pdl_grid_sync();
  __shared__ float tile[32][33]; // tile[r][c]
  const uint32_t r0 = blockIdx.x * 32, c0 = blockIdx.y * 32;
  const float *s = src + (uint64_t)blockIdx.z * s_bs;
  __nv_bfloat16 *d = dst + (uint64_t)blockIdx.z * d_bs;
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5; // 32 x 8
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t r, c;
    if (src_rowfast) { r = tx; c = ty + 8 * i; } else { c = tx; r = ty + 8 * i; }
    const uint32_t gr = r0 + r, gc = c0 + c;
    tile[r][c] = (gr < rows && gc < cols) ? s[(uint64_t)gr * s0 + (uint64_t)gc * s1] : 0.0f;
  }
  __syncthreads();
#pragma unroll
  for (int i = 0; i < 4; ++i) {
    uint32_t r, c;
    if (dst_major) { r = tx; c = ty + 8 * i; } else { c = tx; r = ty + 8 * i; }
    const uint32_t gr = r0 + r, gc = c0 + c;
    if (gr < rows && gc < cols) {
      const uint64_t off = dst_major ? ((uint64_t)gr + (uint64_t)gc * ld) : ((uint64_t)gc + (uint64_t)gr * ld);
      d[off] = __float2bfloat16_rn(tile[r][c]);
    }
  }
}

// Same majorness on both sides (the usual case: operands are packed in the majorness they already
// have): a pure streaming conversion, 8 elements per thread (2 x 128-bit loads, 1 x 128-bit store).
__global__ void __launch_bounds__(256)
pack_bf16_stream_kernel(const float *__restrict__ src, uint64_t s_bs, uint64_t ss, uint32_t n_fast8,
                        uint32_t n_slow, __nv_bfloat16 *__restrict__ dst, uint64_t d_bs, uint64_t ld) {
  pdl_grid_sync();
  const float *s = src + (uint64_t)blockIdx.z * s_bs;
  __nv_bfloat16 *d = dst + (uint64_t)blockIdx.z * d_bs;
  const uint64_t total = (uint64_t)n_fast8 * n_slow, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) {
    const uint32_t j = (uint32_t)(i / n_fast8), f = (uint32_t)(i - (uint64_t)j * n_fast8) << 3;
    const float4 a = *reinterpret_cast<const float4 *>(s + (uint64_t)j * ss + f);
    const float4 b = *reinterpret_cast<const float4 *>(s + (uint64_t)j * ss + f + 4);
    __nv_bfloat162 o[4];
    o[0] = __floats2bfloat162_rn(a.x, a.y);
    o[1] = __floats2bfloat162_rn(a.z, a.w);
    o[2] = __floats2bfloat162_rn(b.x, b.y);
    o[3] = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4 *>(d + (uint64_t)j * ld + f) = *reinterpret_cast<const uint4 *>(o);
  }
}

// Streaming conversion that also reduces: one block per slow index j converts the contiguous run
// src[j*ss .. +n_fast) and leaves its fp32 sum in colsum[j] — the bias gradient (column sums of dY)
// falls out of the pass that packs dY for the two backward GEMMs.
__global__ void __launch_bounds__(256)
pack_bf16_colsum_kernel(const float *__restrict__ src, uint64_t ss, uint32_t n_fast8, __nv_bfloat16 *__restrict__ dst,
                        uint64_t ld, float *colsum, int accumulate) {
  pdl_grid_sync();
  __shared__ float red[32];
  const uint32_t j = blockIdx.x;
  const float *s = src + (uint64_t)j * ss;
  __nv_bfloat16 *d = dst + (uint64_t)j * ld;
  float sum = 0.0f;
  for (uint32_t i = threadIdx.x; i < n_fast8; i += 256) {
    const float4 a = *reinterpret_cast<const float4 *>(s + 8 * i);
    const float4 b = *reinterpret_cast<const float4 *>(s + 8 * i + 4);
    sum += ((a.x + a.y) + (a.z + a.w)) + ((b.x + b.y) + (b.z + b.w));
    __nv_bfloat162 o[4];
    o[0] = __floats2bfloat162_rn(a.x, a.y);
    o[1] = __floats2bfloat162_rn(a.z, a.w);
    o[2] = __floats2bfloat162_rn(b.x, b.y);
    o[3] = __floats2bfloat162_rn(b.z, b.w);
    *reinterpret_cast<uint4 *>(d + 8 * i) = *reinterpret_cast<const uint4 *>(o);
  }
  sum = block_sum(sum, red);
  if (threadIdx.x == 0) colsum[j] = accumulate ? (colsum[j] + sum) : sum;
}

int launch_pack_bf16(const float *src, uint64_t s_bs, uint32_t s0, uint32_t s1, uint32_t rows, uint32_t cols,
                     uint16_t *dst, uint64_t d_bs, uint64_t ld, int dst_major, uint32_t batch,
                     cudaStream_t st) {
  const dim3 grid((rows + 31) / 32, (cols + 31) / 32, batch);
  if (grid.y > 65535 || grid.z > 65535) return WEEDCU_EINVAL;
  const int src_rowfast = (s0 <= s1) ? 1 : 0;
  ProfScope prof(WEEDCU_PROF_PACK, st, 6.0 * (double)rows * cols * batch);
  {
    const uint32_t n_fast = dst_major ? rows : cols, n_slow = dst_major ? cols : rows;
    const uint64_t s_fast = dst_major ? s0 : s1, ss = dst_major ? s1 : s0;
    if (s_fast == 1 && (n_fast % 8u) == 0 && (ss % 4u) == 0 && (s_bs % 4u) == 0 && (d_bs % 8u) == 0 &&
        (((uintptr_t)src) & 15u) == 0 && (((uintptr_t)dst) & 15u) == 0) {
      const uint64_t total = (uint64_t)(n_fast / 8u) * n_slow;
      launch_k(pack_bf16_stream_kernel, dim3(grid_for(total, 256, 16), 1, batch), dim3(256), 0, st, 
          src, s_bs, ss, n_fast / 8u, n_slow, (__nv_bfloat16 *)dst, d_bs, ld);
      return after_launch();
    }
  }
  launch_k(pack_bf16_kernel, dim3(grid), dim3(256), 0, st, src, s_bs, s0, s1, rows, cols, (__nv_bfloat16 *)dst, d_bs, ld,
                                         dst_major, src_rowfast);
  return after_launch();
}

int launch_gemm_f32(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c,
                    const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N, uint32_t batch, int accumulate,
                    cudaStream_t st);

static inline uint64_t round8(uint64_t x) { return (x + 7) & ~7ull; }

} // namespace weedcu

using namespace weedcu;

extern "C" {

int weedcu_gemm_bf16(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major,
                     uint64_t ldb, float *c, uint64_t ldc, uint32_t M, uint32_t N, uint32_t K,
                     int accumulate, const float *col_bias, void *stream) {
  return tc::launch_gemm_bf16(a, a_major, lda, 0, b, b_major, ldb, 0, c, ldc, 0, M, N, K, 1, accumulate,
                              resolve_stream(stream), col_bias);
}

int weedcu_gemm_bf16_grouped(const uint16_t *a, int a_major, uint64_t lda, uint32_t groups, const uint16_t *const *b,
                             int b_major, uint64_t ldb, float *const *c, uint64_t ldc, uint32_t M, uint32_t N, uint32_t K,
                             int accumulate, const float *const *col_bias, void *stream) {
  return tc::launch_gemm_bf16_grouped(a, a_major, lda, 0, groups, b, b_major, ldb, 0, c, ldc, 0, M, N, K, 1, accumulate,
                                      resolve_stream(stream), col_bias, nullptr, 0);
}

int weedcu_gemm_bf16_residual(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major, uint64_t ldb, float *c,
                              uint64_t ldc, uint32_t M, uint32_t N, uint32_t K, const float *col_bias, const float *residual,
                              uint64_t ldr, void *stream) {
  if (!a || !b || !c || !residual) return WEEDCU_EINVAL;
  const uint16_t *bs[1] = {b};
  float *cs[1] = {c};
  const float *biases[1] = {col_bias};
  return tc::launch_gemm_bf16_grouped(a, a_major, lda, 0, 1, bs, b_major, ldb, 0, cs, ldc, 0, M, N, K, 1, 0, resolve_stream(stream),
                                      col_bias ? biases : nullptr, residual, ldr);
}

int weedcu_gemm_bf16_ex(const uint16_t *a, int a_major, uint64_t lda, const uint16_t *b, int b_major, uint64_t ldb, float *c,
                        uint64_t ldc, uint16_t *c_bf16, uint64_t ldc_bf16, uint32_t M, uint32_t N, uint32_t K,
                        const weedcu_gemm_epilogue *epi, void *stream) {
  if (!a || !b || (!c && !c_bf16)) return WEEDCU_EINVAL;
  tc::EpiExt ext;
  ext.out_mode = c_bf16 ? (c ? 2 : 1) : 0;
  ext.c16[0] = c_bf16;
  ext.ldc16 = ldc_bf16;
  uint32_t tiles = 0;
  if (epi) {
    ext.act = epi->activation;
    ext.stats_kind = epi->row_stats;
    ext.row_stats = epi->stats;
    ext.stats_capacity_tiles = epi->stats_capacity_tiles;
    ext.stats_tiles = &tiles;
    if (ext.act && !c_bf16) return WEEDCU_EINVAL; // the activation only exists on the bf16 copy
    if (ext.stats_kind < 0 || ext.stats_kind > 2 || ext.act < 0 || ext.act > 1) return WEEDCU_EINVAL;
  }
  const uint16_t *bs[1] = {b};
  float *cs[1] = {c};
  const float *biases[1] = {epi ? epi->col_bias : nullptr};
  const int rc = tc::launch_gemm_bf16_grouped(a, a_major, lda, 0, 1, bs, b_major, ldb, 0, cs, ldc, 0, M, N, K, 1, 0, resolve_stream(stream),
                                              biases[0] ? biases : nullptr, epi ? epi->residual : nullptr, epi ? epi->ldr : 0, &ext);
  if (rc == 0 && epi && epi->row_stats) {
    if (epi->stats_tiles) *epi->stats_tiles = tiles;
    if (epi->stats_tile_cols) *epi->stats_tile_cols = (uint32_t)tc::last_block_n() / tc::EPI_GROUPS; // the half tile one epilogue group drains
  }
  return rc;
}

int weedcu_gemm_bf16_grouped_bf16out(const uint16_t *a, int a_major, uint64_t lda, uint32_t groups, const uint16_t *const *b,
                                     int b_major, uint64_t ldb, uint16_t *const *c_bf16, uint64_t ldc_bf16, uint32_t M, uint32_t N,
                                     uint32_t K, const float *const *col_bias, void *stream) {
  if (!a || !b || !c_bf16 || !groups || groups > tc::kMaxGroups) return WEEDCU_EINVAL;
  tc::EpiExt ext;
  ext.out_mode = 1;
  ext.ldc16 = ldc_bf16;
  for (uint32_t g = 0; g < groups; ++g) ext.c16[g] = c_bf16[g];
  float *cs[3] = {nullptr, nullptr, nullptr};
  return tc::launch_gemm_bf16_grouped(a, a_major, lda, 0, groups, b, b_major, ldb, 0, cs, 0, 0, M, N, K, 1, 0, resolve_stream(stream), col_bias,
                                      nullptr, 0, &ext);
}

int weedcu_gemm_set_dynamic(int on) {
  weedcu::tc::g_gemm_dynamic = on ? 1 : 0;
  return 0;
}
int weedcu_gemm_set_mode(int mode) {
  tc::set_gemm_mode(mode);
  return 0;
}

int weedcu_pack_bf16(const float *src, uint64_t offset, uint32_t s0, uint32_t s1, uint32_t rows,
                     uint32_t cols, uint16_t *dst, int dst_major, void *stream) {
  if (!src || !dst || !rows || !cols) return WEEDCU_EINVAL;
  const uint64_t ld = round8(dst_major ? rows : cols);
  return launch_pack_bf16(src + offset, 0, s0, s1, rows, cols, dst, 0, ld, dst_major, 1, resolve_stream(stream));
}

int weedcu_pack_bf16_colsum(const float *src, uint64_t offset, uint32_t s0, uint32_t s1, uint32_t rows,
                            uint32_t cols, uint16_t *dst, int dst_major, float *colsum, int accumulate,
                            void *stream) {
  if (!src || !dst || !colsum || !rows || !cols) return WEEDCU_EINVAL;
  const uint32_t n_fast = dst_major ? rows : cols, n_slow = dst_major ? cols : rows;
  const uint64_t s_fast = dst_major ? s0 : s1, ss = dst_major ? s1 : s0;
  const float *base = src + offset;
  if (s_fast != 1 || (n_fast % 8u) || (ss % 4u) || (((uintptr_t)base) & 15u) || (((uintptr_t)dst) & 15u)) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  ProfScope prof(WEEDCU_PROF_PACK, st, 6.0 * (double)rows * cols);
  launch_k(pack_bf16_colsum_kernel, dim3(n_slow), dim3(256), 0, st, base, ss, n_fast / 8u, (__nv_bfloat16 *)dst, round8(n_fast), colsum, accumulate);
  return after_launch();
}

int weedcu_gemm_workspace_bytes(uint32_t M, uint32_t K, uint32_t N, uint32_t batch, int precision,
                                uint64_t *bytes) {
  if (!bytes) return WEEDCU_EINVAL;
  if (precision == WEEDCU_GEMM_FP32) { *bytes = 0; return 0; }
  *bytes = 2ull * batch * (round8(M) * round8(K) + round8(K) * round8(N)) + 512;
  return 0;
}

int weedcu_matmul_real(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm,
                       float *c, const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N,
                       uint32_t batch, int accumulate, int precision, void *stream) {
  if (!a || !am || !b || !bm || !c || !cm || !M || !K || !N || !batch) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  // The tensor-core path writes C column-major (M contiguous). Other C layouts, and problems too
  // small to fill one 128-wide tile, take the FFMA kernel (a precision upgrade, never a host path).
  const bool tc_ok = (precision == WEEDCU_GEMM_BF16) && (cm->s0 == 1) && (M >= 64) && (N >= 16) && (K >= 32);
  if (!tc_ok) return launch_gemm_f32(a, am, b, bm, c, cm, M, K, N, batch, accumulate, st);

  // Pack to bf16 in whichever majorness the fp32 source is already contiguous in.
  const int a_major = (am->s0 <= am->s1) ? 1 : 0; // 1: M contiguous
  const int b_major = (bm->s1 < bm->s0) ? 1 : 0;  // 1: N contiguous; B[k,n]: s0 = k stride
  const uint64_t lda = a_major ? round8(M) : round8(K);
  const uint64_t ldb = b_major ? round8(N) : round8(K);
  const uint64_t a_elems = a_major ? lda * K : lda * M, b_elems = b_major ? ldb * K : ldb * N;
  const uint64_t a_bs = round8(a_elems), b_bs = round8(b_elems);
  uint16_t *ws = nullptr;
  WCU_CHECK(pool_alloc((void **)&ws, 2ull * batch * (a_bs + b_bs), st));
  uint16_t *wa = ws, *wb = ws + (uint64_t)batch * a_bs;
  int rc = launch_pack_bf16(a + am->offset, am->batch_stride, am->s0, am->s1, M, K, wa, a_bs, lda, a_major,
                            batch, st);
  // B as an [N, K] operand: rows = n (stride s1), cols = k (stride s0)
  if (rc == 0)
    rc = launch_pack_bf16(b + bm->offset, bm->batch_stride, bm->s1, bm->s0, N, K, wb, b_bs, ldb, b_major,
                          batch, st);
  if (rc == 0)
    rc = tc::launch_gemm_bf16(wa, a_major, lda, a_bs, wb, b_major, ldb, b_bs, c + cm->offset, cm->s1,
                              cm->batch_stride, M, N, K, batch, accumulate, st, nullptr);
  pool_free(ws, st);
  if (rc == WEEDCU_ENOSUP) return launch_gemm_f32(a, am, b, bm, c, cm, M, K, N, batch, accumulate, st);
  return rc;
}

} // extern "C"
