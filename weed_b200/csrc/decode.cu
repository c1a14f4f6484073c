// decode.cu — the KV-cache decode path: cache append + attention over the cache, and the skinny
// (few-row) matmul that every Linear turns into when a handful of tokens is fed per step.
//
// Reference: MultiHeadAttention::forward with use_kv_cache and kv_quant_bits = 0
// (src/modules/multihead_attention.cpp:169-199 cache allocation, :278-287 slot write + slices,
// :313-345 scores / scale / mask / softmax / P V / transposes). The reference issues ~15 tensor ops
// per call and copies the whole cache contiguous twice (Tensor::matmul's reshape of the sliced
// views, tensor.cpp:1242-1243). Here: one append kernel, one split-key attention kernel that reads
// the cache in place, one combine kernel. All HBM-bound: the step reads each cache element once.
//
// Layouts (column-major like every Weed tensor):
//   q, k, v, out : [B, T_new, H*hd]   element (b, t, c) at b + B*(t + T_new*c)      (Linear outputs);
//                  feature c = h + H*j: the column-major reshape to {B, T, H, hd} (:155-157) makes H fast
//   k/v cache    : [B, H, S, hd]      element (b, h, s, j) at bh + BH*(s + S*j), bh = b + B*h
// so for a fixed (s, j) the BH = B*H heads are one contiguous run: a warp = 32 adjacent (b, h) pairs
// reads full 128-byte lines, a thread owns one (b, h) pair and keeps q, the running maximum / sum
// and the hd-wide output accumulator in registers (flash-decoding: keys split over warps and
// blocks, partial (m, l, acc) merged afterwards).
#include "common.cuh"

namespace weedcu {

__global__ void __launch_bounds__(256)
kv_append_kernel(const float *__restrict__ k, const float *__restrict__ v, float *k_cache, float *v_cache, uint32_t B,
                 uint32_t T_new, uint32_t H, uint32_t hd, uint32_t S, uint32_t cache_len) {
  pdl_grid_sync();
  const uint64_t BH = (uint64_t)B * H, total = BH * T_new * hd;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint32_t b = (uint32_t)(i % B);
    uint64_t r = i / B;
    const uint32_t t = (uint32_t)(r % T_new);
    r /= T_new;
    const uint32_t h = (uint32_t)(r % H), j = (uint32_t)(r / H);
    const uint64_t src = b + (uint64_t)B * (t + (uint64_t)T_new * ((uint64_t)h + (uint64_t)H * j));
    const uint64_t dst = (b + (uint64_t)B * h) + BH * ((uint64_t)(cache_len + t) + (uint64_t)S * j);
    // add_in_place into the zero-initialised slot (multihead_attention.cpp:282-283)
    k_cache[dst] = k_cache[dst] + k[src];
    v_cache[dst] = v_cache[dst] + v[src];
  }
}

constexpr int kDecWarps = 8;

// grid (ceil(BH/32), n_chunks, T_new); block 32 x kDecWarps. Chunk c covers keys [c*KC, min(L, (c+1)*KC));
// warp w takes keys c*KC + w, + kDecWarps, ... Partials: part[((c*T_new + t)*(hd + 2) + slot)*BH + bh],
// slot < hd: accumulator, slot hd: running maximum, slot hd + 1: running sum.
template <int HDM>
__global__ void __launch_bounds__(32 * kDecWarps)
attn_decode_partial_kernel(const float *__restrict__ q, const float *__restrict__ k_cache, const float *__restrict__ v_cache,
                           float *__restrict__ part, uint32_t B, uint32_t T_new, uint32_t H, uint32_t hd, uint32_t S,
                           uint32_t L, uint32_t KC, float divisor, float mask_val, int causal) {
  pdl_grid_sync();
  extern __shared__ float sm[]; // [kDecWarps][HDM + 2][32]
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t BH = B * H, bh = blockIdx.x * 32u + lane, t = blockIdx.z, c = blockIdx.y;
  const bool valid = bh < BH;
  const uint32_t b = valid ? bh % B : 0u, h = valid ? bh / B : 0u;
  float qv[HDM], acc[HDM];
#pragma unroll
  for (int j = 0; j < HDM; ++j) {
    qv[j] = (valid && j < (int)hd) ? q[b + (uint64_t)B * (t + (uint64_t)T_new * ((uint64_t)h + (uint64_t)H * j))] : 0.0f;
    acc[j] = 0.0f;
  }
  float m = -INFINITY, l = 0.0f;
  const uint32_t s_end = min(L, (c + 1u) * KC);
  const uint64_t js = (uint64_t)BH * S; // stride of the head-dim index in the cache
  if (valid)
    for (uint32_t s = c * KC + w; s < s_end; s += kDecWarps) {
      const float *kp = k_cache + bh + (uint64_t)BH * s;
      const float *vp = v_cache + bh + (uint64_t)BH * s;
      float kv[HDM]; // (K row, then reused for the V row: q, acc and one row fit the register file at hd = 64)
#pragma unroll
      for (int j = 0; j < HDM; ++j) kv[j] = (j < (int)hd) ? kp[js * j] : 0.0f;
      float dot = 0.0f;
#pragma unroll
      for (int j = 0; j < HDM; ++j) dot += qv[j] * kv[j];
#pragma unroll
      for (int j = 0; j < HDM; ++j) kv[j] = (j < (int)hd) ? vp[js * j] : 0.0f;
      float x = dot / divisor;
      if (causal && t + 1u <= s) x = x + mask_val; // triu_fill(mask, val, diagonal 1) on [T_q, T_k]: key index > query index
      const float m_new = fmaxf(m, x);
      const float sc = expf(m - m_new), p = expf(x - m_new); // m = -inf -> sc = 0
      l = l * sc + p;
#pragma unroll
      for (int j = 0; j < HDM; ++j) acc[j] = acc[j] * sc + p * kv[j];
      m = m_new;
    }
  // merge the warps' partials
  float *mine = sm + (size_t)w * (HDM + 2) * 32;
#pragma unroll
  for (int j = 0; j < HDM; ++j) mine[j * 32 + lane] = acc[j];
  mine[HDM * 32 + lane] = m;
  mine[(HDM + 1) * 32 + lane] = l;
  __syncthreads();
  float M = -INFINITY;
#pragma unroll
  for (int k = 0; k < kDecWarps; ++k) M = fmaxf(M, sm[((size_t)k * (HDM + 2) + HDM) * 32 + lane]);
  float wt[kDecWarps];
#pragma unroll
  for (int k = 0; k < kDecWarps; ++k) {
    const float mk = sm[((size_t)k * (HDM + 2) + HDM) * 32 + lane];
    wt[k] = (mk == -INFINITY) ? 0.0f : expf(mk - M);
  }
  if (!valid) return;
  float *dst = part + ((uint64_t)(c * T_new + t) * (hd + 2u)) * BH + bh;
  for (uint32_t slot = w; slot < hd + 2u; slot += kDecWarps) {
    float r;
    if (slot == hd) {
      r = M;
    } else {
      const uint32_t row = (slot == hd + 1u) ? (uint32_t)(HDM + 1) : slot;
      r = 0.0f;
#pragma unroll
      for (int k = 0; k < kDecWarps; ++k) r += wt[k] * sm[((size_t)k * (HDM + 2) + row) * 32 + lane];
    }
    dst[(uint64_t)slot * BH] = r;
  }
}

// thread per (bh, j, t): out[b, t, h + H*j] = sum_c acc_c[j] e^(m_c - M) / sum_c l_c e^(m_c - M)
__global__ void __launch_bounds__(256)
attn_decode_combine_kernel(const float *__restrict__ part, float *__restrict__ out, uint32_t B, uint32_t T_new, uint32_t H,
                           uint32_t hd, uint32_t n_chunks) {
  pdl_grid_sync();
  const uint32_t BH = B * H;
  const uint64_t total = (uint64_t)BH * hd * T_new, i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= total) return;
  const uint32_t bh = (uint32_t)(i % BH);
  const uint64_t r = i / BH;
  const uint32_t j = (uint32_t)(r % hd), t = (uint32_t)(r / hd);
  // (both loops unrolled by 8: a decode step has 16-32 chunks and the loads of one unrolled group are independent —
  // rolled, every chunk cost a dependent L2 round trip and this tiny merge took 13.6 us)
  float M = -INFINITY;
#pragma unroll 8
  for (uint32_t c = 0; c < n_chunks; ++c) M = fmaxf(M, part[((uint64_t)(c * T_new + t) * (hd + 2u) + hd) * BH + bh]);
  float num = 0.0f, den = 0.0f;
#pragma unroll 8
  for (uint32_t c = 0; c < n_chunks; ++c) {
    const float *p = part + ((uint64_t)(c * T_new + t) * (hd + 2u)) * BH + bh;
    const float mc = p[(uint64_t)hd * BH];
    const float wt = (mc == -INFINITY) ? 0.0f : expf(mc - M);
    num += wt * p[(uint64_t)j * BH];
    den += wt * p[(uint64_t)(hd + 1u) * BH];
  }
  const uint32_t b = bh % B, h = bh / B;
  out[b + (uint64_t)B * (t + (uint64_t)T_new * ((uint64_t)h + (uint64_t)H * j))] = num / den;
}

template <int HDM>
static int launch_decode_partial(const float *q, const float *kc, const float *vc, float *part, uint32_t B, uint32_t T_new,
                                 uint32_t H, uint32_t hd, uint32_t S, uint32_t L, uint32_t KC, uint32_t n_chunks, float divisor,
                                 float mask_val, int causal, cudaStream_t st) {
  auto kern = attn_decode_partial_kernel<HDM>;
  const size_t smem = sizeof(float) * kDecWarps * (HDM + 2) * 32;
  ensure_dynamic_smem((const void *)kern, (int)smem);
  const dim3 grid((B * H + 31u) / 32u, n_chunks, T_new);
  launch_k(kern, dim3(grid), dim3(32 * kDecWarps), smem, st, q, kc, vc, part, B, T_new, H, hd, S, L, KC, divisor, mask_val, causal);
  return after_launch();
}

// ------------------------------------------------------------------------------- skinny matmul
// C[M, N] (+)= A[M, K] B[K, N] (+ bias[n]) for M <= 16 rows (a decode step feeds B tokens): the
// weight matrix is read exactly once, which is all the time there is to spend. A block owns CW = 8
// output columns; its 8 warps split K; inside a warp adjacent lanes take adjacent k (weights are
// K-contiguous: every B load is a full 128-byte line), a lane loads its M values of A once per k
// (128-bit loads when A is M-contiguous) and uses them for all 8 columns: 2 + 8 loads per 64 FMAs.
// The MM x 8 partial dots of a lane meet in a halving butterfly (2*MM*8 - 2 shuffles instead of
// 5 per value), then the 8 warps meet in shared memory. Arbitrary operand strides.
constexpr int kSkCW = 8;
struct SkinnyGroups {
  const float *b[3];
  float *c[3];
  const float *bias[3];
  const float *residual[3]; // optional, laid out like c: c = (a b + bias) + residual
};
template <int MM, bool AVEC>
__global__ void __launch_bounds__(256, (MM <= 8 && AVEC) ? 2 : 1) // the decode variant (8 rows, vector A loads) at <= 128 registers: 2 blocks per SM
skinny_matmul_kernel(const float *__restrict__ a, uint32_t a_s0, uint32_t a_s1, SkinnyGroups grp, uint32_t b_s0,
                     uint32_t b_s1, uint32_t c_s0, uint32_t c_s1, uint32_t M, uint32_t K, uint32_t N, int accumulate) {
  pdl_grid_sync();
  // blockIdx.y = product of a grouped launch (the W_q / W_k / W_v projections of one decode step share A)
  const float *__restrict__ b = grp.b[blockIdx.y];
  float *c = grp.c[blockIdx.y];
  const float *__restrict__ bias = grp.bias[blockIdx.y];
  const float *__restrict__ residual = grp.residual[blockIdx.y];
  constexpr int NV = MM * kSkCW; // partial sums per lane
  __shared__ float red[8][NV];
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const uint32_t n0 = blockIdx.x * kSkCW;
  float acc[NV]; // acc[m * kSkCW + j]
#pragma unroll
  for (int i = 0; i < NV; ++i) acc[i] = 0.0f;
  constexpr int U = (MM <= 8) ? 4 : 2; // k-steps whose loads are issued together (latency, not bandwidth, is the enemy here:
                                       // K = 768 is one round of 3 live steps, K = 3072 three rounds instead of six)
  for (uint32_t k0 = w * 32u + lane; k0 < K; k0 += 256u * U) {
    float bv[U][kSkCW], av[U][MM];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const uint32_t k = k0 + 256u * u;
      const bool kok = k < K;
#pragma unroll
      for (int j = 0; j < kSkCW; ++j) bv[u][j] = (kok && n0 + j < N) ? b[(uint64_t)k * b_s0 + (uint64_t)(n0 + j) * b_s1] : 0.0f;
      if (AVEC) { // a_s0 == 1, 16-byte aligned columns, M == MM
#pragma unroll
        for (int m = 0; m < MM; m += 4)
          *reinterpret_cast<float4 *>(&av[u][m]) = kok ? *reinterpret_cast<const float4 *>(a + (uint64_t)k * a_s1 + m) : make_float4(0.f, 0.f, 0.f, 0.f);
      } else {
#pragma unroll
        for (int m = 0; m < MM; ++m) av[u][m] = (kok && m < (int)M) ? a[(uint64_t)m * a_s0 + (uint64_t)k * a_s1] : 0.0f;
      }
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
#pragma unroll
      for (int m = 0; m < MM; ++m)
#pragma unroll
        for (int j = 0; j < kSkCW; ++j) acc[m * kSkCW + j] += av[u][m] * bv[u][j];
  }
  // halving butterfly: after the step with offset o a lane keeps half of its values, each now the
  // sum over the lane pair; after 5 steps lane L holds NV/32 complete sums, the values with index
  // base(L) + i, base = sum over steps of (lane bit set ? half : 0)
  uint32_t base = 0;
#pragma unroll
  for (int o = 16, cnt = NV; o >= 1; o >>= 1, cnt >>= 1) {
    const int half = cnt >> 1;
    const bool up = (lane & o) != 0;
#pragma unroll
    for (int i = 0; i < half; ++i) {
      const float send = up ? acc[i] : acc[i + half];
      const float keep = up ? acc[i + half] : acc[i];
      acc[i] = keep + __shfl_xor_sync(0xffffffffu, send, o);
    }
    if (up) base += half;
  }
  constexpr int LEFT = NV / 32; // MM >= 4 -> LEFT >= 1
#pragma unroll
  for (int i = 0; i < LEFT; ++i) red[w][base + i] = acc[i];
  __syncthreads();
  if (threadIdx.x < NV) {
    const uint32_t m = threadIdx.x / kSkCW, j = threadIdx.x % kSkCW, nn = n0 + j;
    if (m < M && nn < N) {
      float s = 0.0f;
#pragma unroll
      for (int k = 0; k < 8; ++k) s += red[k][threadIdx.x];
      if (bias) s += bias[nn];
      if (residual) s += residual[(uint64_t)m * c_s0 + (uint64_t)nn * c_s1];
      float *dst = c + (uint64_t)m * c_s0 + (uint64_t)nn * c_s1;
      *dst = accumulate ? (*dst + s) : s;
    }
  }
}

static int launch_skinny_groups(const float *a, uint32_t a_s0, uint32_t a_s1, const SkinnyGroups &grp, uint32_t groups, uint32_t b_s0, uint32_t b_s1,
                                uint32_t c_s0, uint32_t c_s1, uint32_t M, uint32_t K, uint32_t N, int accumulate, cudaStream_t st) {
  if (M > 16u) return WEEDCU_ENOSUP;
  const dim3 grid((N + kSkCW - 1) / kSkCW, groups);
  ProfScope prof(WEEDCU_PROF_GEMM_F32, st, 2.0 * (double)M * N * K * groups);
#define WCU_SK(MM, AV) launch_k(skinny_matmul_kernel<MM, AV>, grid, dim3(256), 0, st, a, a_s0, a_s1, grp, b_s0, b_s1, c_s0, c_s1, M, K, N, accumulate)
  const bool avec_ok = a_s0 == 1u && (a_s1 % 4u) == 0 && aligned16(a);
  if (M <= 4u) {
    if (avec_ok && M == 4u) WCU_SK(4, true);
    else WCU_SK(4, false);
  } else if (M <= 8u) {
    if (avec_ok && M == 8u) WCU_SK(8, true);
    else WCU_SK(8, false);
  } else {
    if (avec_ok && M == 16u) WCU_SK(16, true);
    else WCU_SK(16, false);
  }
#undef WCU_SK
  return after_launch();
}
int launch_skinny_matmul(const float *a, uint32_t a_s0, uint32_t a_s1, const float *b, uint32_t b_s0, uint32_t b_s1, float *c,
                         uint32_t c_s0, uint32_t c_s1, uint32_t M, uint32_t K, uint32_t N, const float *bias, int accumulate,
                         cudaStream_t st) {
  SkinnyGroups grp = {{b, b, b}, {c, c, c}, {bias, bias, bias}, {nullptr, nullptr, nullptr}};
  return launch_skinny_groups(a, a_s0, a_s1, grp, 1, b_s0, b_s1, c_s0, c_s1, M, K, N, accumulate, st);
}

} // namespace weedcu

using namespace weedcu;

extern "C" {

int weedcu_attention_decode(const float *q, const float *k, const float *v, float *k_cache, float *v_cache, float *out,
                            uint32_t B, uint32_t T_new, uint32_t H, uint32_t hd, uint32_t S, uint32_t cache_len,
                            float divisor, float mask_val, int causal, void *stream) {
  if (!q || !k || !v || !k_cache || !v_cache || !out || !B || !T_new || !H || !hd || !S) return WEEDCU_EINVAL;
  if ((uint64_t)cache_len + T_new > S) return WEEDCU_EINVAL;
  if (hd > 64u || T_new > 65535u) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  const uint32_t BH = B * H, L = cache_len + T_new;
  {
    ProfScope prof(WEEDCU_PROF_ELEMENTWISE, st, 2.0 * 12.0 * (double)BH * T_new * hd);
    const uint64_t total = (uint64_t)BH * T_new * hd;
    launch_k(kv_append_kernel, dim3(grid_for(total, 256, 8)), dim3(256), 0, st, k, v, k_cache, v_cache, B, T_new, H, hd, S, cache_len);
    const int rc = after_launch();
    if (rc) return rc;
  }
  // enough (bh group, key chunk, query) blocks to cover the SMs twice; a chunk is a multiple of the
  // kDecWarps keys one block pass consumes
  const uint32_t gx = (BH + 31u) / 32u;
  uint32_t n_chunks = (2u * (uint32_t)kNumSMs + gx * T_new - 1u) / (gx * T_new);
  const uint32_t max_chunks = (L + kDecWarps - 1u) / kDecWarps;
  if (n_chunks > max_chunks) n_chunks = max_chunks;
  if (n_chunks < 1u) n_chunks = 1u;
  uint32_t KC = (L + n_chunks - 1u) / n_chunks;
  KC = (KC + kDecWarps - 1u) / kDecWarps * kDecWarps;
  n_chunks = (L + KC - 1u) / KC;
  if (n_chunks > 65535u) return WEEDCU_ENOSUP;
  float *part = nullptr;
  WCU_CHECK(pool_alloc((void **)&part, sizeof(float) * (size_t)n_chunks * T_new * (hd + 2u) * BH, st));
  int rc;
  {
    // one read of the visible part of both caches (+ q, out)
    ProfScope prof(WEEDCU_PROF_ATTENTION, st, 4.0 * (2.0 * (double)BH * L * hd * T_new + 2.0 * (double)BH * hd * T_new));
    const int do_mask = (causal && T_new > 1u) ? 1 : 0;
    if (hd <= 8u) rc = launch_decode_partial<8>(q, k_cache, v_cache, part, B, T_new, H, hd, S, L, KC, n_chunks, divisor, mask_val, do_mask, st);
    else if (hd <= 16u) rc = launch_decode_partial<16>(q, k_cache, v_cache, part, B, T_new, H, hd, S, L, KC, n_chunks, divisor, mask_val, do_mask, st);
    else if (hd <= 32u) rc = launch_decode_partial<32>(q, k_cache, v_cache, part, B, T_new, H, hd, S, L, KC, n_chunks, divisor, mask_val, do_mask, st);
    else rc = launch_decode_partial<64>(q, k_cache, v_cache, part, B, T_new, H, hd, S, L, KC, n_chunks, divisor, mask_val, do_mask, st);
    if (rc == 0) {
      const uint64_t total = (uint64_t)BH * hd * T_new;
      launch_k(attn_decode_combine_kernel, dim3((unsigned)((total + 255) / 256)), dim3(256), 0, st, part, out, B, T_new, H, hd, n_chunks);
      rc = after_launch();
    }
  }
  pool_free(part, st);
  return rc;
}

int weedcu_matmul_skinny(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c,
                         const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N, const float *bias, int accumulate,
                         void *stream) {
  if (!a || !am || !b || !bm || !c || !cm || !M || !K || !N) return WEEDCU_EINVAL;
  return launch_skinny_matmul(a + am->offset, am->s0, am->s1, b + bm->offset, bm->s0, bm->s1, c + cm->offset, cm->s0, cm->s1, M,
                              K, N, bias, accumulate, resolve_stream(stream));
}

int weedcu_matmul_skinny_residual(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm, float *c, const weedcu_mat *cm,
                                  uint32_t M, uint32_t K, uint32_t N, const float *bias, const float *residual, void *stream) {
  if (!a || !am || !b || !bm || !c || !cm || !residual || !M || !K || !N) return WEEDCU_EINVAL;
  const float *bb = b + bm->offset, *rr = residual + cm->offset;
  float *cc = c + cm->offset;
  SkinnyGroups grp = {{bb, bb, bb}, {cc, cc, cc}, {bias, bias, bias}, {rr, rr, rr}};
  return launch_skinny_groups(a + am->offset, am->s0, am->s1, grp, 1, bm->s0, bm->s1, cm->s0, cm->s1, M, K, N, 0, resolve_stream(stream));
}

int weedcu_matmul_skinny_grouped(const float *a, const weedcu_mat *am, uint32_t groups, const float *const *b, const weedcu_mat *bm,
                                 float *const *c, const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N, const float *const *bias,
                                 void *stream) {
  if (!a || !am || !b || !bm || !c || !cm || !M || !K || !N || !groups || groups > 3u) return WEEDCU_EINVAL;
  SkinnyGroups grp;
  for (uint32_t g = 0; g < 3u; ++g) {
    const uint32_t s = g < groups ? g : 0u;
    if (!b[s] || !c[s]) return WEEDCU_EINVAL;
    grp.b[g] = b[s] + bm->offset;
    grp.c[g] = c[s] + cm->offset;
    grp.bias[g] = bias ? bias[s] : nullptr;
    grp.residual[g] = nullptr;
  }
  return launch_skinny_groups(a + am->offset, am->s0, am->s1, grp, groups, bm->s0, bm->s1, cm->s0, cm->s1, M, K, N, 0, resolve_stream(stream));
}

} // extern "C"
