// attention.cu — fused attention core on bf16 tensor cores (no autograd edge, like the reference).
//
// Reference: MultiHeadAttention::forward, src/modules/multihead_attention.cpp:289-345 — reshape to
// heads, transpose, Q K^T (batched matmul, tensor.cpp:1242-1269), / sqrt(hd), triu mask, softmax,
// P V, transpose back. The reference runs ~10 tensor ops with a full copy for every transpose.
//
// Here, for q, k, v = [B, T, H*hd] column-major (b fastest):
//   1. heads_pack   fp32 [B,T,C] -> bf16 [B*H][hd][T]   one kernel for q, k and v: the head
//                    relayout and the fp32->bf16 operand conversion are the same pass (6 B/elem)
//   2. tcgen05 GEMM  S[bh] = Qh Kh^T                      fp32 [T(k)][T(q)], q contiguous
//   3. softmax       scale + causal mask + softmax        reads fp32 S (only the unmasked part),
//                    writes the probabilities as bf16 — the A operand of the next product — so P
//                    is never stored in fp32 nor re-packed (4+2 B/elem instead of 4+4+4+2)
//   4. tcgen05 GEMM  O[bh] = P Vh                          fp32 [hd][T]
//   5. unheads       fp32 [B*H][hd][T] -> [B,T,C]
// Steps 2-4 are the part a flash-style kernel (S kept in TMEM) replaces next; the
// entry point and its oracle model (oracle/weed_oracle.c: wo_attention_fwd) stay the same.
#include "common.cuh"

#include <cuda_bf16.h>

namespace weedcu {
namespace tc {
int launch_gemm_bf16(const uint16_t *a, int a_major, uint64_t lda, uint64_t a_bs, const uint16_t *b,
                     int b_major, uint64_t ldb, uint64_t b_bs, float *c, uint64_t ldc, uint64_t c_bs,
                     uint32_t M, uint32_t N, uint32_t K, uint32_t batch, int accumulate, cudaStream_t st,
                     const float *col_bias);
}

// ------------------------------------------------------------------------------- relayouts
// Column c of x is one contiguous run of B*T floats (index b + B*t). A block takes `tt` tokens of
// one column: reads tt*B contiguous floats, writes B runs of tt contiguous bf16.
struct HeadsPackArgs {
  const float *src[3];
  __nv_bfloat16 *dst[3];
};
__global__ void __launch_bounds__(256)
heads_pack_kernel(HeadsPackArgs a, uint32_t B, uint32_t T, uint32_t H, uint32_t hd, uint32_t tt) {
  extern __shared__ float tile[]; // [tt][B + 1]
  const uint32_t c = blockIdx.y, h = c / hd, j = c - h * hd;
  const uint32_t t0 = blockIdx.x * tt, nt = min(tt, T - t0), n = nt * B;
  const float *src = a.src[blockIdx.z] + ((uint64_t)c * T + t0) * B;
  __nv_bfloat16 *dst = a.dst[blockIdx.z];
  for (uint32_t i = threadIdx.x; i < n; i += 256) {
    const uint32_t t = i / B, b = i - t * B;
    tile[t * (B + 1) + b] = src[i];
  }
  __syncthreads();
  for (uint32_t i = threadIdx.x; i < n; i += 256) {
    const uint32_t b = i / nt, t = i - b * nt;
    dst[(((uint64_t)b * H + h) * hd + j) * T + t0 + t] = __float2bfloat16_rn(tile[t * (B + 1) + b]);
  }
}
__global__ void __launch_bounds__(256)
unheads_kernel(const float *__restrict__ oc, float *__restrict__ out, uint32_t B, uint32_t T, uint32_t H,
               uint32_t hd, uint32_t tt) {
  extern __shared__ float tile[]; // [tt][B + 1]
  const uint32_t c = blockIdx.y, h = c / hd, j = c - h * hd;
  const uint32_t t0 = blockIdx.x * tt, nt = min(tt, T - t0), n = nt * B;
  for (uint32_t i = threadIdx.x; i < n; i += 256) {
    const uint32_t b = i / nt, t = i - b * nt;
    tile[t * (B + 1) + b] = oc[(((uint64_t)b * H + h) * hd + j) * T + t0 + t];
  }
  __syncthreads();
  float *dst = out + ((uint64_t)c * T + t0) * B;
  for (uint32_t i = threadIdx.x; i < n; i += 256) {
    const uint32_t t = i / B, b = i - t * B;
    dst[i] = tile[t * (B + 1) + b];
  }
}

// ------------------------------------------------------------------------------- softmax -> bf16
// S, P: [BH][T(k)][T(q)], q contiguous. Block = 16 queries x 32 key lanes; a thread keeps its NV
// keys of one query in registers (all loads in flight at once). Arithmetic as the reference chain:
// x / divisor, + mask where q + 1 <= k (triu_fill.cpp:48-56), max, exp, sum, divide.
// With a -2^127-like mask the masked probabilities are exactly 0: key columns beyond the last
// query of the tile are not read, only zero-filled.
constexpr int kSmRT = 16, kSmBY = 32;
template <int NV>
__global__ void __launch_bounds__(kSmRT *kSmBY)
attn_softmax_bf16_kernel(const float *__restrict__ S, __nv_bfloat16 *__restrict__ P, uint32_t T, float divisor,
                         float mask_val, int causal) {
  __shared__ float red[kSmBY][kSmRT + 1];
  const uint32_t tx = threadIdx.x % kSmRT, ty = threadIdx.x / kSmRT;
  const uint32_t q0 = blockIdx.x * kSmRT, q = q0 + tx;
  const bool live = q < T;
  const uint64_t slab = (uint64_t)blockIdx.y * T * T;
  const float *p = S + slab + q;
  __nv_bfloat16 *po = P + slab + q;
  const uint32_t Lc = (causal && mask_val <= -1e30f) ? min(T, q0 + kSmRT) : T;

  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t k = ty + i * kSmBY;
    v[i] = (live && k < Lc) ? p[(uint64_t)k * T] : 0.0f;
  }
  float mx = -INFINITY;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t k = ty + i * kSmBY;
    float x = v[i] / divisor;
    if (causal) x = x + ((q + 1u <= k) ? mask_val : 0.0f);
    v[i] = (live && k < Lc) ? x : -INFINITY;
    mx = fmaxf(mx, v[i]);
  }
  red[ty][tx] = mx;
  __syncthreads();
#pragma unroll
  for (int y = 0; y < kSmBY; ++y) mx = fmaxf(mx, red[y][tx]);
  __syncthreads();
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    v[i] = (v[i] > -INFINITY) ? expf(v[i] - mx) : 0.0f;
    s += v[i];
  }
  red[ty][tx] = s;
  __syncthreads();
  s = 0.0f;
#pragma unroll
  for (int y = 0; y < kSmBY; ++y) s += red[y][tx];
  if (!live) return;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t k = ty + i * kSmBY;
    if (k < Lc) po[(uint64_t)k * T] = __float2bfloat16_rn(v[i] / s);
  }
  for (uint32_t k = Lc + ty; k < T; k += kSmBY) po[(uint64_t)k * T] = __float2bfloat16_rn(0.0f);
}

} // namespace weedcu

using namespace weedcu;

extern "C" int weedcu_attention_fwd(const float *q, const float *k, const float *v, float *out, uint32_t B,
                                    uint32_t T, uint32_t H, uint32_t hd, float divisor, float mask_val,
                                    int causal, void *stream) {
  if (!q || !k || !v || !out || !B || !T || !H || !hd) return WEEDCU_EINVAL;
  // tensor-map constraints of the two products (16-byte row pitch) and the register softmax
  if ((T % 8u) || T < 64u || T > 32u * kSmBY || hd < 16u || (hd % 8u) || H * hd > 65535u) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  const uint64_t BH = (uint64_t)B * H, C = (uint64_t)H * hd;
  if (BH > 65535u) return WEEDCU_ENOSUP;
  const uint64_t head_elems = (uint64_t)hd * T; // per (b,h)
  const uint64_t qkv_bytes = 3 * BH * head_elems * sizeof(uint16_t);
  const uint64_t s_bytes = BH * T * T * sizeof(float), p_bytes = BH * T * T * sizeof(uint16_t);
  const uint64_t o_bytes = BH * head_elems * sizeof(float);
  auto up = [](uint64_t x) { return (x + 255) & ~(uint64_t)255; };
  char *ws = nullptr;
  WCU_CHECK(pool_alloc((void **)&ws, up(qkv_bytes) + up(s_bytes) + up(p_bytes) + up(o_bytes), st));
  uint16_t *qh = (uint16_t *)ws, *kh = qh + BH * head_elems, *vh = kh + BH * head_elems;
  float *S = (float *)(ws + up(qkv_bytes));
  uint16_t *P = (uint16_t *)(ws + up(qkv_bytes) + up(s_bytes));
  float *oc = (float *)(ws + up(qkv_bytes) + up(s_bytes) + up(p_bytes));

  uint32_t tt = 128;
  while (tt > 1 && (size_t)tt * (B + 1) * sizeof(float) > 40 * 1024) tt >>= 1;
  const size_t tile_bytes = (size_t)tt * (B + 1) * sizeof(float);
  if (tile_bytes > 48 * 1024) {
    pool_free(ws, st);
    return WEEDCU_ENOSUP;
  }
  int rc;
  {
    ProfScope prof(WEEDCU_PROF_PACK, st, 3.0 * 6.0 * (double)B * T * C);
    HeadsPackArgs a = {{q, k, v}, {(__nv_bfloat16 *)qh, (__nv_bfloat16 *)kh, (__nv_bfloat16 *)vh}};
    heads_pack_kernel<<<dim3((T + tt - 1) / tt, (unsigned)C, 3), 256, tile_bytes, st>>>(a, B, T, H, hd, tt);
    rc = after_launch();
  }
  if (rc == 0) // S[bh] = Qh Kh^T : A [T, hd] and B [T, hd] both with the token index contiguous
    rc = tc::launch_gemm_bf16(qh, 1, T, head_elems, kh, 1, T, head_elems, S, T, (uint64_t)T * T, T, T, hd, (uint32_t)BH, 0, st, nullptr);
  if (rc == 0) {
    ProfScope prof(WEEDCU_PROF_SOFTMAX, st, (causal ? 4.0 : 6.0) * (double)BH * T * T);
    const dim3 grid((T + kSmRT - 1) / kSmRT, (unsigned)BH);
    const int do_mask = (causal && T > 1) ? 1 : 0;
    const uint32_t nv = (T + kSmBY - 1) / kSmBY;
#define WCU_SM(NV) attn_softmax_bf16_kernel<NV><<<grid, kSmRT * kSmBY, 0, st>>>(S, (__nv_bfloat16 *)P, T, divisor, mask_val, do_mask)
    if (nv <= 4) WCU_SM(4);
    else if (nv <= 8) WCU_SM(8);
    else if (nv <= 16) WCU_SM(16);
    else WCU_SM(32);
#undef WCU_SM
    rc = after_launch();
  }
  if (rc == 0) // O[bh] = P Vh : A = P [T(q), T(k)] q contiguous; B = Vh as [hd, T(k)] with k contiguous
    rc = tc::launch_gemm_bf16(P, 1, T, (uint64_t)T * T, vh, 0, T, head_elems, oc, T, head_elems, T, hd, T, (uint32_t)BH, 0, st, nullptr);
  if (rc == 0) {
    ProfScope prof(WEEDCU_PROF_ELEMENTWISE, st, 8.0 * (double)B * T * C);
    unheads_kernel<<<dim3((T + tt - 1) / tt, (unsigned)C), 256, tile_bytes, st>>>(oc, out, B, T, H, hd, tt);
    rc = after_launch();
  }
  pool_free(ws, st);
  return rc;
}
