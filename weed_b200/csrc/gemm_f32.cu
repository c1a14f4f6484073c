// gemm_f32.cu — FpMath-precision matmul: fp32 in, fp32 FFMA accumulate, arbitrary operand strides,
// optional batch and C += accumulate.  This is the parity path (<= 1e-5 of the reference's CPU
// loop, src/ops/matmul.cpp:34-47); the tensor-core paths live in gemm_tc.cu.
//
// 128x128x16 block tile, 256 threads, 8x8 register tile per thread split into four 4x4 quadrants so
// every shared-memory read is a conflict-free 128-bit load; global->register prefetch of the next
// k-slab overlaps the FFMA loop (double-buffered shared memory, one barrier per slab).
// The thread->element mapping of the global loads is picked per operand so that whichever index is
// contiguous in memory (M/N or K) is the one adjacent lanes walk.
#include "common.cuh"

namespace weedcu {

constexpr int BM = 128, BN = 128, BK = 16, PAD = 4;

struct GemmArgs {
  const float *a, *b;
  float *c;
  uint64_t a_bs, b_bs, c_bs;   // batch strides (elements)
  uint32_t as0, as1, bs0, bs1, cs0, cs1;
  uint32_t M, N, K;
  int accumulate;
};

// A_KFAST: adjacent lanes walk k (A is K-contiguous); else they walk m.  Same for B_KFAST.
template <bool A_KFAST, bool B_KFAST>
__global__ void __launch_bounds__(256)
gemm_f32_kernel(GemmArgs g) {
  pdl_grid_sync();
  __shared__ __align__(16) float As[2][BK][BM + PAD];
  __shared__ __align__(16) float Bs[2][BK][BN + PAD];
  const uint32_t tid = threadIdx.x;
  const uint32_t tx = tid & 15, ty = tid >> 4; // tx -> rows (m), ty -> cols (n)
  const uint32_t m0 = blockIdx.x * BM, n0 = blockIdx.y * BN;
  const float *A = g.a + (uint64_t)blockIdx.z * g.a_bs;
  const float *B = g.b + (uint64_t)blockIdx.z * g.b_bs;
  float *C = g.c + (uint64_t)blockIdx.z * g.c_bs;

  float acc[8][8];
#pragma unroll
  for (int i = 0; i < 8; ++i)
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[i][j] = 0.0f;

  float ra[8], rb[8];
  auto load_global = [&](uint32_t k0) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t m, k;
      if (A_KFAST) { k = tid & 15; m = (tid >> 4) + 16 * i; }
      else         { m = tid & 127; k = (tid >> 7) + 2 * i; }
      const uint32_t gm = m0 + m, gk = k0 + k;
      ra[i] = (gm < g.M && gk < g.K) ? A[(uint64_t)gm * g.as0 + (uint64_t)gk * g.as1] : 0.0f;
      uint32_t n, kb;
      if (B_KFAST) { kb = tid & 15; n = (tid >> 4) + 16 * i; }
      else         { n = tid & 127; kb = (tid >> 7) + 2 * i; }
      const uint32_t gn = n0 + n, gkb = k0 + kb;
      rb[i] = (gn < g.N && gkb < g.K) ? B[(uint64_t)gkb * g.bs0 + (uint64_t)gn * g.bs1] : 0.0f;
    }
  };
  auto store_shared = [&](int buf) {
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t m, k;
      if (A_KFAST) { k = tid & 15; m = (tid >> 4) + 16 * i; }
      else         { m = tid & 127; k = (tid >> 7) + 2 * i; }
      As[buf][k][m] = ra[i];
      uint32_t n, kb;
      if (B_KFAST) { kb = tid & 15; n = (tid >> 4) + 16 * i; }
      else         { n = tid & 127; kb = (tid >> 7) + 2 * i; }
      Bs[buf][kb][n] = rb[i];
    }
  };

  const uint32_t nk = (g.K + BK - 1) / BK;
  load_global(0);
  store_shared(0);
  __syncthreads();
  for (uint32_t kt = 0; kt < nk; ++kt) {
    const int cur = kt & 1;
    if (kt + 1 < nk) load_global((kt + 1) * BK);
#pragma unroll
    for (int kk = 0; kk < BK; ++kk) {
      const float4 a0 = *reinterpret_cast<const float4 *>(&As[cur][kk][tx * 4]);
      const float4 a1 = *reinterpret_cast<const float4 *>(&As[cur][kk][64 + tx * 4]);
      const float4 b0 = *reinterpret_cast<const float4 *>(&Bs[cur][kk][ty * 4]);
      const float4 b1 = *reinterpret_cast<const float4 *>(&Bs[cur][kk][64 + ty * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 8; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    if (kt + 1 < nk) {
      store_shared(cur ^ 1);
      __syncthreads();
    }
  }

  // epilogue: rows {tx*4..+3, 64+tx*4..+3}, cols {ty*4..+3, 64+ty*4..+3}
  const bool vec_ok = (g.cs0 == 1) && ((((uintptr_t)C) & 15u) == 0) && (g.cs1 % 4u == 0);
#pragma unroll
  for (int hj = 0; hj < 2; ++hj)
#pragma unroll
    for (int j = 0; j < 4; ++j) {
      const uint32_t n = n0 + hj * 64 + ty * 4 + j;
      if (n >= g.N) continue;
#pragma unroll
      for (int hi = 0; hi < 2; ++hi) {
        const uint32_t m = m0 + hi * 64 + tx * 4;
        if (m >= g.M) continue;
        float *dst = C + (uint64_t)m * g.cs0 + (uint64_t)n * g.cs1;
        float v[4] = {acc[hi * 4 + 0][hj * 4 + j], acc[hi * 4 + 1][hj * 4 + j],
                      acc[hi * 4 + 2][hj * 4 + j], acc[hi * 4 + 3][hj * 4 + j]};
        if (vec_ok && m + 3 < g.M) {
          float4 o = make_float4(v[0], v[1], v[2], v[3]);
          if (g.accumulate) {
            const float4 old = *reinterpret_cast<const float4 *>(dst);
            o.x += old.x; o.y += old.y; o.z += old.z; o.w += old.w;
          }
          *reinterpret_cast<float4 *>(dst) = o;
        } else {
#pragma unroll
          for (int r = 0; r < 4; ++r)
            if (m + r < g.M) {
              float *d = dst + (uint64_t)r * g.cs0;
              *d = g.accumulate ? (*d + v[r]) : v[r];
            }
        }
      }
    }
}

// Skinny problems (N or M tiny, e.g. the [65536,26]x[26,1] head of the heart_attack MLP, or M=1
// decode GEMV): one thread per output, K serial; adjacent threads walk m so A reads coalesce.
__global__ void __launch_bounds__(256)
gemm_f32_thin_kernel(GemmArgs g) {
  pdl_grid_sync();
  const uint64_t total = (uint64_t)g.M * g.N;
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  const uint32_t m = (uint32_t)(idx % g.M), n = (uint32_t)(idx / g.M);
  const float *A = g.a + (uint64_t)blockIdx.z * g.a_bs + (uint64_t)m * g.as0;
  const float *B = g.b + (uint64_t)blockIdx.z * g.b_bs + (uint64_t)n * g.bs1;
  float s = 0.0f;
  for (uint32_t k = 0; k < g.K; ++k) s = fmaf(A[(uint64_t)k * g.as1], B[(uint64_t)k * g.bs0], s);
  float *d = g.c + (uint64_t)blockIdx.z * g.c_bs + (uint64_t)m * g.cs0 + (uint64_t)n * g.cs1;
  *d = g.accumulate ? (*d + s) : s;
}

// Few outputs, long reduction (the weight gradients of a small layer over a big batch: dW = X^T dY with X [65536, 13] in the
// heart_attack MLP, or the [512, 1] head of the binary-addition transformer over 32768 tokens): the 128 x 128 tile kernel
// puts the whole reduction on one block (1.6 ms for 13 x 26 x 65536), one-thread-per-output likewise. Here the reduction
// is split over blockIdx.z: a block stages 64 x 64-deep slabs of both operands in shared memory (whichever index is
// contiguous in memory is the one adjacent threads walk), every thread carries a 4 x 4 register tile over its slice, and
// the per-slice partial tiles are summed in slice order by a second kernel (deterministic: no atomics).
constexpr int SK_T = 64, SK_K = 64;
__global__ void __launch_bounds__(256)
gemm_f32_splitk_kernel(GemmArgs g, uint32_t k_per_slice, float *__restrict__ partial) {
  pdl_grid_sync();
  __shared__ float As[SK_K][SK_T + 1]; // As[k][m]
  __shared__ float Bs[SK_K][SK_T + 1]; // Bs[k][n]
  const uint32_t tid = threadIdx.x, tx = tid & 15, ty = tid >> 4; // tx -> rows m (4 each), ty -> cols n (4 each)
  const uint32_t m0 = blockIdx.x * SK_T, n0 = blockIdx.y * SK_T;
  const uint32_t k_begin = blockIdx.z * k_per_slice, k_end = min(g.K, k_begin + k_per_slice);
  const bool a_kfast = g.as1 <= g.as0, b_kfast = g.bs0 <= g.bs1;
  float acc[4][4];
#pragma unroll
  for (int i = 0; i < 4; ++i)
#pragma unroll
    for (int j = 0; j < 4; ++j) acc[i][j] = 0.0f;
  for (uint32_t k0 = k_begin; k0 < k_end; k0 += SK_K) {
    for (uint32_t i = tid; i < SK_T * SK_K; i += 256) {
      const uint32_t am_ = a_kfast ? i / SK_K : i % SK_T, ak = a_kfast ? i % SK_K : i / SK_T;
      const uint32_t gm = m0 + am_, gk = k0 + ak;
      As[ak][am_] = (gm < g.M && gk < k_end) ? g.a[(uint64_t)gm * g.as0 + (uint64_t)gk * g.as1] : 0.0f;
      const uint32_t bn = b_kfast ? i / SK_K : i % SK_T, bk = b_kfast ? i % SK_K : i / SK_T;
      const uint32_t gn = n0 + bn, gkb = k0 + bk;
      Bs[bk][bn] = (gn < g.N && gkb < k_end) ? g.b[(uint64_t)gkb * g.bs0 + (uint64_t)gn * g.bs1] : 0.0f;
    }
    __syncthreads();
#pragma unroll 8
    for (int kk = 0; kk < SK_K; ++kk) {
      float av[4], bv[4];
#pragma unroll
      for (int i = 0; i < 4; ++i) {
        av[i] = As[kk][tx * 4 + i];
        bv[i] = Bs[kk][ty * 4 + i];
      }
#pragma unroll
      for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
    }
    __syncthreads();
  }
  float *dst = partial + (uint64_t)blockIdx.z * g.M * g.N;
#pragma unroll
  for (int j = 0; j < 4; ++j)
#pragma unroll
    for (int i = 0; i < 4; ++i) {
      const uint32_t m = m0 + tx * 4 + i, n = n0 + ty * 4 + j;
      if (m < g.M && n < g.N) dst[(uint64_t)n * g.M + m] = acc[i][j];
    }
}
__global__ void __launch_bounds__(256)
gemm_f32_splitk_reduce_kernel(GemmArgs g, uint32_t slices, const float *__restrict__ partial) {
  pdl_grid_sync();
  const uint64_t total = (uint64_t)g.M * g.N, idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= total) return;
  float s = 0.0f;
  for (uint32_t z = 0; z < slices; ++z) s += partial[(uint64_t)z * total + idx];
  const uint32_t m = (uint32_t)(idx % g.M), n = (uint32_t)(idx / g.M);
  float *d = g.c + (uint64_t)m * g.cs0 + (uint64_t)n * g.cs1;
  *d = g.accumulate ? (*d + s) : s;
}

int launch_gemm_f32(const float *a, const weedcu_mat *am, const float *b, const weedcu_mat *bm,
                    float *c, const weedcu_mat *cm, uint32_t M, uint32_t K, uint32_t N,
                    uint32_t batch, int accumulate, cudaStream_t st) {
  GemmArgs g;
  g.a = a + am->offset;
  g.b = b + bm->offset;
  g.c = c + cm->offset;
  g.a_bs = am->batch_stride;
  g.b_bs = bm->batch_stride;
  g.c_bs = cm->batch_stride;
  g.as0 = am->s0; g.as1 = am->s1;
  g.bs0 = bm->s0; g.bs1 = bm->s1;
  g.cs0 = cm->s0; g.cs1 = cm->s1;
  g.M = M; g.N = N; g.K = K;
  g.accumulate = accumulate;
  if (batch > 65535) return WEEDCU_EINVAL;
  ProfScope prof(WEEDCU_PROF_GEMM_F32, st, 2.0 * (double)M * N * K * batch);
  {
    const uint32_t tiles = ((M + SK_T - 1) / SK_T) * ((N + SK_T - 1) / SK_T);
    if (batch == 1 && K >= 2048u && tiles <= 64u && (uint64_t)M * N <= (1ull << 18)) {
      uint32_t slices = (2u * (uint32_t)kNumSMs + tiles - 1) / tiles;
      const uint32_t max_slices = (K + 4 * SK_K - 1) / (4 * SK_K); // at least 4 slabs per slice
      if (slices > max_slices) slices = max_slices;
      if (slices >= 2u) {
        uint32_t kps = (K + slices - 1) / slices;
        kps = (kps + SK_K - 1) / SK_K * SK_K;
        slices = (K + kps - 1) / kps;
        float *partial = nullptr;
        WCU_CHECK(pool_alloc((void **)&partial, sizeof(float) * (size_t)slices * M * N, st));
        launch_k(gemm_f32_splitk_kernel, dim3((M + SK_T - 1) / SK_T, (N + SK_T - 1) / SK_T, slices), dim3(256), 0, st, g, kps, partial);
        int rc = after_launch();
        if (rc == 0) {
          launch_k(gemm_f32_splitk_reduce_kernel, dim3((unsigned)(((uint64_t)M * N + 255) / 256)), dim3(256), 0, st, g, slices, (const float *)partial);
          rc = after_launch();
        }
        pool_free(partial, st);
        return rc;
      }
    }
  }
  if ((uint64_t)M * N * (uint64_t)K <= (1ull << 22) || N <= 4 || M <= 4) {
    if ((uint64_t)M * N <= (1ull << 24) && (N <= 4 || M <= 4 || (uint64_t)M * N * K <= (1ull << 18))) {
      const uint64_t total = (uint64_t)M * N;
      launch_k(gemm_f32_thin_kernel, dim3((unsigned)((total + 255) / 256), 1, batch), dim3(256), 0, st, g);
      return after_launch();
    }
  }
  const dim3 grid((M + BM - 1) / BM, (N + BN - 1) / BN, batch);
  if (grid.y > 65535) return WEEDCU_EINVAL;
  const bool a_k = (am->s1 == 1 && am->s0 != 1), b_k = (bm->s0 == 1 && bm->s1 != 1);
  if (a_k && b_k) launch_k(gemm_f32_kernel<true, true>, dim3(grid), dim3(256), 0, st, g);
  else if (a_k) launch_k(gemm_f32_kernel<true, false>, dim3(grid), dim3(256), 0, st, g);
  else if (b_k) launch_k(gemm_f32_kernel<false, true>, dim3(grid), dim3(256), 0, st, g);
  else launch_k(gemm_f32_kernel<false, false>, dim3(grid), dim3(256), 0, st, g);
  return after_launch();
}

} // namespace weedcu
