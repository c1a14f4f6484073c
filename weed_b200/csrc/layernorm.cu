// layernorm.cu — fused LayerNorm forward/backward, embedding gather / scatter-add, triu fill.
//
// LayerNorm reference: LayerNorm::forward (src/modules/layernorm.cpp:29-42) composes ~12 tensor ops
// (mean, sub, mul, mean, +eps, ^0.5, div, *gamma, +beta) and ~30 backward closures. Here: one
// forward kernel (8 B/elem) and one backward kernel (16 B/elem + parameter partials).
// x is [rows, F] with row stride 1 and feature stride `rows`: the normalised axis is the slowest,
// so RT adjacent rows are tiled against all F features, staged once in shared memory, and the
// reference's two-pass statistics (mean, then mean of centred squares) run from the staged copy.
#include "common.cuh"
#include <cuda_bf16.h>
#include <cstdlib>

namespace weedcu {

// Both directions are split into streaming passes, because a tile kernel that loads, reduces,
// synchronises and only then stores is latency-bound here (ncu: 0.9 TB/s of DRAM traffic with the
// SMs 70 % idle): the normalised axis is the SLOWEST one, so per-row statistics need a cross-warp
// reduction per row tile while everything else is plain elementwise work.
//   pass 1 (row statistics)  block = 32 adjacent rows x BY warps striding the features: every
//                            warp-level load is one full 128-byte line; writes 2-3 floats per row
//   pass 2 (apply)           block = 256 x 4 adjacent rows x FB features: 128-bit loads/stores, the
//                            per-row coefficients stay in registers across the FB features, the
//                            second read of x (and dy) is served by the 126 MB L2
// HBM traffic stays at the algorithmic 8 B/elem forward and 12-16 B/elem backward.
constexpr int kLnBY = 16; // warps per row tile in the statistics kernels
constexpr int kLnFB = 8;  // features per block in the apply kernels

// mean, (var + eps)^0.5 and its reciprocal per row. NV > 0: a thread keeps its <= NV features in
// registers between the reference's two passes (mean, then mean of centred squares); NV == 0:
// any F, the second pass re-reads x (L1/L2 hits).
template <int NV>
__global__ void __launch_bounds__(32 * kLnBY)
layernorm_fwd_stats_kernel(const float *__restrict__ x, uint32_t rows, uint32_t F, float eps,
                           float *__restrict__ mean, float *__restrict__ rstd, float *__restrict__ den_out) {
  pdl_grid_sync();
  __shared__ float red[kLnBY][33];
  const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * 32u + tx;
  const bool live = r < rows;
  const float *p = x + r;
  float v[NV > 0 ? NV : 1];
  float s = 0.0f;
  if (NV > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t f = ty + i * kLnBY;
      v[i] = (live && f < F) ? p[(uint64_t)f * rows] : 0.0f;
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) s += v[i];
  } else if (live) {
    for (uint32_t f0 = ty; f0 < F; f0 += 4 * kLnBY) {
      float t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = (f0 + u * kLnBY < F) ? p[(uint64_t)(f0 + u * kLnBY) * rows] : 0.0f;
#pragma unroll
      for (int u = 0; u < 4; ++u) s += t[u];
    }
  }
  red[ty][tx] = s;
  __syncthreads();
  s = 0.0f;
#pragma unroll
  for (int k = 0; k < kLnBY; ++k) s += red[k][tx];
  const float mu = s / (float)F;
  __syncthreads();
  float q = 0.0f;
  if (NV > 0) {
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const float xc = v[i] - mu;
      if (ty + i * kLnBY < F) q += xc * xc;
    }
  } else if (live) {
    for (uint32_t f0 = ty; f0 < F; f0 += 4 * kLnBY) {
      float t[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) t[u] = (f0 + u * kLnBY < F) ? p[(uint64_t)(f0 + u * kLnBY) * rows] : mu;
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const float xc = t[u] - mu;
        q += xc * xc;
      }
    }
  }
  red[ty][tx] = q;
  __syncthreads();
  if (ty == 0 && live) {
    q = 0.0f;
#pragma unroll
    for (int k = 0; k < kLnBY; ++k) q += red[k][tx];
    const float den = sqrtf(q / (float)F + eps); // (var + eps) ^ 0.5, layernorm.cpp:35
    if (mean) mean[r] = mu;
    if (rstd) rstd[r] = 1.0f / den;
    den_out[r] = den;
    if (!mean) den_out[rows + r] = mu; // the apply pass needs mu even when the caller does not
  }
}

// Few rows (a decode step normalises B tokens): statistics and output in ONE launch, the row tile
// lives in registers exactly as in layernorm_fwd_stats_kernel<NV>.
template <int NV>
__global__ void __launch_bounds__(32 * kLnBY)
layernorm_fwd_small_kernel(const float *__restrict__ x, uint32_t rows, uint32_t F, const float *__restrict__ gamma,
                           const float *__restrict__ beta, float eps, float *__restrict__ y, float *__restrict__ mean,
                           float *__restrict__ rstd) {
  pdl_grid_sync();
  __shared__ float red[kLnBY][33];
  // gamma / beta staged once: with a single resident block, 2 x NV dependent global loads per thread
  // in the output loop would each expose the full L2 latency
  __shared__ float s_gamma[kLnBY * NV], s_beta[kLnBY * NV];
  const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * 32u + tx;
  const bool live = r < rows;
  float v[NV];
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t f = ty + i * kLnBY;
    v[i] = (live && f < F) ? x[r + (uint64_t)f * rows] : 0.0f;
  }
  for (uint32_t f = threadIdx.x; f < F; f += 32u * kLnBY) {
    s_gamma[f] = gamma[f];
    s_beta[f] = beta[f];
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i];
  red[ty][tx] = s;
  __syncthreads();
  s = 0.0f;
#pragma unroll
  for (int k = 0; k < kLnBY; ++k) s += red[k][tx];
  const float mu = s / (float)F;
  __syncthreads();
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float xc = v[i] - mu;
    if (ty + i * kLnBY < F) q += xc * xc;
  }
  red[ty][tx] = q;
  __syncthreads();
  q = 0.0f;
#pragma unroll
  for (int k = 0; k < kLnBY; ++k) q += red[k][tx];
  const float den = sqrtf(q / (float)F + eps); // (var + eps) ^ 0.5, layernorm.cpp:35
  if (!live) return;
  if (ty == 0) {
    if (mean) mean[r] = mu;
    if (rstd) rstd[r] = 1.0f / den;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t f = ty + i * kLnBY;
    if (f < F) y[r + (uint64_t)f * rows] = ((v[i] - mu) / den) * s_gamma[f] + s_beta[f];
  }
}

// A handful of rows (a decode step normalises the B tokens of the batch): one block per row, the row's
// features spread over 256 threads (<= NV each, in registers across the two statistics passes), gamma /
// beta loaded in the same round as x. The 32-rows-per-block kernel above leaves 24 of 32 lanes idle at
// 8 rows and walks 48 features per thread: 11.7 us per call, 25 calls per decode step.
template <int NV>
__global__ void __launch_bounds__(256)
layernorm_fwd_row_kernel(const float *__restrict__ x, uint32_t rows, uint32_t F, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float eps, float *__restrict__ y, float *__restrict__ mean,
                         float *__restrict__ rstd) {
  pdl_grid_sync();
  __shared__ float red[32];
  const uint32_t r = blockIdx.x;
  float v[NV], g[NV], b[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t f = threadIdx.x + i * 256u;
    const bool ok = f < F;
    v[i] = ok ? x[r + (uint64_t)f * rows] : 0.0f;
    g[i] = ok ? gamma[f] : 0.0f;
    b[i] = ok ? beta[f] : 0.0f;
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i];
  const float mu = block_sum(s, red) / (float)F;
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float xc = v[i] - mu;
    if (threadIdx.x + i * 256u < F) q += xc * xc;
  }
  const float den = sqrtf(block_sum(q, red) / (float)F + eps); // (var + eps) ^ 0.5, layernorm.cpp:35
  if (threadIdx.x == 0) {
    if (mean) mean[r] = mu;
    if (rstd) rstd[r] = 1.0f / den;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t f = threadIdx.x + i * 256u;
    if (f < F) y[r + (uint64_t)f * rows] = ((v[i] - mu) / den) * g[i] + b[i];
  }
}

// y = ((x - mu) / den) * gamma + beta for VEC adjacent rows x kLnFB features per thread.
template <int VEC>
__global__ void __launch_bounds__(256)
layernorm_fwd_apply_kernel(const float *__restrict__ x, uint32_t rows, uint32_t F, const float *__restrict__ gamma,
                           const float *__restrict__ beta, const float *__restrict__ mean,
                           const float *__restrict__ den, float *__restrict__ y, __nv_bfloat16 *__restrict__ yb) {
  pdl_grid_sync();
  const uint32_t r = (blockIdx.x * 256u + threadIdx.x) * VEC;
  if (r >= rows) return;
  float mu[VEC], dn[VEC];
  if (VEC == 4) {
    *reinterpret_cast<float4 *>(mu) = *reinterpret_cast<const float4 *>(mean + r);
    *reinterpret_cast<float4 *>(dn) = *reinterpret_cast<const float4 *>(den + r);
  } else {
    mu[0] = mean[r];
    dn[0] = den[r];
  }
  const uint32_t f_begin = blockIdx.y * kLnFB, f_end = min(F, f_begin + kLnFB);
  float xv[kLnFB][VEC];
#pragma unroll
  for (int i = 0; i < kLnFB; ++i) {
    const uint32_t f = f_begin + i;
    if (f < f_end) {
      const float *src = x + (uint64_t)f * rows + r;
      if (VEC == 4) *reinterpret_cast<float4 *>(xv[i]) = *reinterpret_cast<const float4 *>(src);
      else xv[i][0] = *src;
    }
  }
#pragma unroll
  for (int i = 0; i < kLnFB; ++i) {
    const uint32_t f = f_begin + i;
    if (f < f_end) {
      const float g = gamma[f], b = beta[f];
      float o[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = ((xv[i][k] - mu[k]) / dn[k]) * g + b;
      float *dst = y + (uint64_t)f * rows + r;
      if (VEC == 4) *reinterpret_cast<float4 *>(dst) = *reinterpret_cast<const float4 *>(o);
      else *dst = o[0];
      if (VEC == 4 && yb) { // bf16 copy at the same linear index: the next Linear's GEMM operand, no pack pass
        __nv_bfloat162 h[2];
        h[0] = __floats2bfloat162_rn(o[0], o[VEC > 1 ? 1 : 0]);
        h[1] = __floats2bfloat162_rn(o[VEC > 2 ? 2 : 0], o[VEC > 3 ? 3 : 0]);
        *reinterpret_cast<uint2 *>(yb + (uint64_t)f * rows + r) = *reinterpret_cast<const uint2 *>(h);
      }
    }
  }
}

// Row statistics that arrive as per-column-tile partials (mean_t, M2_t) from the epilogue of the GEMM that produced x
// (weedcu_gemm_bf16_ex, row_stats 1): one thread per row merges the partials (Chan's update, counts known from the tile
// width) into mean / den = (var + eps)^0.5 / rstd; layernorm_fwd_apply_kernel follows — the statistics pass over x is gone.
// (Merging inside the apply kernel repeated the ~12-partial merge in each of its F / 8 feature groups: 18.5 us against
// 10.6 us for the plain apply pass at 8192 x 768.)
__global__ void __launch_bounds__(128)
layernorm_stats_merge_kernel(const float2 *__restrict__ stats, uint32_t rows, uint32_t F, uint32_t tiles, uint32_t tile_cols, float eps,
                             float *__restrict__ mean_out, float *__restrict__ den_out, float *__restrict__ rstd_out) {
  pdl_grid_sync();
  const uint32_t r = blockIdx.x * 128u + threadIdx.x;
  if (r >= rows) return;
  float mu = 0.0f, m2 = 0.0f, n = 0.0f;
  // all partials in flight at once (<= 16 by the producer's contract): one memory round trip, not one per partial
  float2 pt[16];
#pragma unroll
  for (uint32_t t = 0; t < 16u; ++t) pt[t] = (t < tiles && t * tile_cols < F) ? stats[(uint64_t)t * rows + r] : make_float2(0.0f, 0.0f);
#pragma unroll
  for (uint32_t t = 0; t < 16u; ++t) {
    if (!(t < tiles && t * tile_cols < F)) break; // (trailing partials beyond F are empty)
    const float cnt = (float)(min(F, (t + 1u) * tile_cols) - t * tile_cols), tot = n + cnt;
    const float2 p = pt[t];
    const float delta = p.x - mu;
    mu += delta * (cnt / tot);
    m2 += p.y + delta * delta * (n * cnt / tot);
    n = tot;
  }
  const float dn = sqrtf(m2 / (float)F + eps); // (var + eps) ^ 0.5, layernorm.cpp:35
  mean_out[r] = mu;
  den_out[r] = dn;
  if (rstd_out) rstd_out[r] = 1.0f / dn;
}

// Backward pass 1: sg = sum_f dy*gamma, sgx = sum_f g*xhat (mode 1) or sum_f xhat (mode 0) per row,
// folded straight into the two per-row coefficients of
//     dx (+)= (rstd * (dy*gamma) + k_x * xhat) + k_0
// mode 1 (analytic):   dx += rs*(g - mean(g) - xh*mean(g*xh))
// mode 0 (reference chain, see weedcu.h): dxc = g*rs + c*xc with c = -rs^3*sum(xc)/F;
//                         dx += dxc - mean(dxc).  In xhat terms xc = xh/rs, sum(xc) = sgx/rs.
__global__ void __launch_bounds__(32 * kLnBY)
layernorm_bwd_rows_kernel(const float *__restrict__ x, const float *__restrict__ dy, uint32_t rows, uint32_t F,
                          const float *__restrict__ gamma, const float *__restrict__ mean,
                          const float *__restrict__ rstd, float *__restrict__ k_x_out, float *__restrict__ k_0_out,
                          int grad_mode) {
  pdl_grid_sync();
  __shared__ float red_a[kLnBY][33];
  __shared__ float red_b[kLnBY][33];
  const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * 32u + tx;
  const bool live = r < rows;
  const float mu = live ? mean[r] : 0.0f, rs = live ? rstd[r] : 0.0f;
  float sg = 0.0f, sgx = 0.0f;
  if (live) {
    // 24 independent 128-byte row loads in flight per thread: with 8 the pass was bound by the 12 dependent round trips of
    // a 768-feature row (17.3 us for 50 MB at 8192 x 768); the order of the per-thread sums does not depend on U
    constexpr int U = 12;
    for (uint32_t f0 = ty; f0 < F; f0 += U * kLnBY) {
      float xv[U], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f0 + u * kLnBY;
        const bool ok = f < F;
        xv[u] = ok ? x[r + (uint64_t)f * rows] : mu;
        dv[u] = ok ? dy[r + (uint64_t)f * rows] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f0 + u * kLnBY;
        if (f < F) {
          const float xh = (xv[u] - mu) * rs;
          const float g = dv[u] * gamma[f];
          sg += g;
          sgx += grad_mode ? g * xh : xh;
        }
      }
    }
  }
  red_a[ty][tx] = sg;
  red_b[ty][tx] = sgx;
  __syncthreads();
  if (ty == 0 && live) {
    sg = sgx = 0.0f;
#pragma unroll
    for (int k = 0; k < kLnBY; ++k) {
      sg += red_a[k][tx];
      sgx += red_b[k][tx];
    }
    float k_x, k_0;
    if (grad_mode) {
      k_x = -rs * (sgx / (float)F);
      k_0 = -rs * (sg / (float)F);
    } else {
      const float sx = (rs != 0.0f) ? sgx / rs : 0.0f;
      const float c = -(rs * rs * rs) * sx / (float)F;
      k_x = (rs != 0.0f) ? c / rs : 0.0f;
      k_0 = -((rs * sg + c * sx) / (float)F);
    }
    k_x_out[r] = k_x;
    k_0_out[r] = k_0;
  }
}

// Backward pass 2: dx for 256 x VEC adjacent rows x kLnFB features per block, and the block's share
// of the column sums dgamma_f = sum_r dy*xhat, dbeta_f = sum_r dy, which leave as one partial row per
// row chunk: part_g/part_b[blockIdx.x][F] (summed in a fixed order by layernorm_param_reduce_kernel).
template <int VEC>
__global__ void __launch_bounds__(256)
layernorm_bwd_apply_kernel(const float *__restrict__ x, const float *__restrict__ dy, uint32_t rows, uint32_t F,
                           const float *__restrict__ gamma, const float *__restrict__ mean,
                           const float *__restrict__ rstd, const float *__restrict__ k_x, const float *__restrict__ k_0,
                           const float *dx_in, float *dx, float *__restrict__ part_g, float *__restrict__ part_b) {
  pdl_grid_sync();
  __shared__ float red[8][2 * kLnFB];
  const bool accumulate = dx_in != nullptr; // dx = dx_in + ...; dx_in may be dx itself (in place) or another buffer
  const uint32_t r = (blockIdx.x * 256u + threadIdx.x) * VEC;
  const bool live = r < rows;
  float mu[VEC], rs[VEC], kx[VEC], k0[VEC];
#pragma unroll
  for (int k = 0; k < VEC; ++k) mu[k] = rs[k] = kx[k] = k0[k] = 0.0f;
  if (live) {
    if (VEC == 4) {
      *reinterpret_cast<float4 *>(mu) = *reinterpret_cast<const float4 *>(mean + r);
      *reinterpret_cast<float4 *>(rs) = *reinterpret_cast<const float4 *>(rstd + r);
      *reinterpret_cast<float4 *>(kx) = *reinterpret_cast<const float4 *>(k_x + r);
      *reinterpret_cast<float4 *>(k0) = *reinterpret_cast<const float4 *>(k_0 + r);
    } else {
      mu[0] = mean[r];
      rs[0] = rstd[r];
      kx[0] = k_x[r];
      k0[0] = k_0[r];
    }
  }
  const uint32_t f_begin = blockIdx.y * kLnFB, f_end = min(F, f_begin + kLnFB);
  float cg[kLnFB], cb[kLnFB];
#pragma unroll
  for (int i = 0; i < kLnFB; ++i) cg[i] = cb[i] = 0.0f;
  if (live) {
    constexpr int U = 4; // features whose loads are in flight together
#pragma unroll
    for (int i0 = 0; i0 < kLnFB; i0 += U) {
      float xv[U][VEC], dv[U][VEC], ov[U][VEC];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f_begin + i0 + u;
        if (f < f_end) {
          const uint64_t off = (uint64_t)f * rows + r;
          if (VEC == 4) {
            *reinterpret_cast<float4 *>(xv[u]) = *reinterpret_cast<const float4 *>(x + off);
            *reinterpret_cast<float4 *>(dv[u]) = *reinterpret_cast<const float4 *>(dy + off);
            if (accumulate) *reinterpret_cast<float4 *>(ov[u]) = *reinterpret_cast<const float4 *>(dx_in + off);
          } else {
            xv[u][0] = x[off];
            dv[u][0] = dy[off];
            if (accumulate) ov[u][0] = dx_in[off];
          }
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f_begin + i0 + u;
        if (f < f_end) {
          const float gm = gamma[f];
          float o[VEC];
#pragma unroll
          for (int k = 0; k < VEC; ++k) {
            const float xh = (xv[u][k] - mu[k]) * rs[k];
            const float val = (rs[k] * (dv[u][k] * gm) + kx[k] * xh) + k0[k];
            o[k] = accumulate ? (ov[u][k] + val) : val;
            cg[i0 + u] += dv[u][k] * xh;
            cb[i0 + u] += dv[u][k];
          }
          float *dst = dx + (uint64_t)f * rows + r;
          if (VEC == 4) *reinterpret_cast<float4 *>(dst) = *reinterpret_cast<const float4 *>(o);
          else *dst = o[0];
        }
      }
    }
  }
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kLnFB; ++i) {
    const float a = warp_sum(cg[i]), b = warp_sum(cb[i]);
    if (lane == 0) {
      red[warp][i] = a;
      red[warp][kLnFB + i] = b;
    }
  }
  __syncthreads();
  if (threadIdx.x < 2 * kLnFB) {
    float t = 0.0f;
#pragma unroll
    for (int w = 0; w < 8; ++w) t += red[w][threadIdx.x];
    const uint32_t i = threadIdx.x % kLnFB, f = f_begin + i;
    if (f < f_end) (threadIdx.x < kLnFB ? part_g : part_b)[(uint64_t)blockIdx.x * F + f] = t;
  }
}

// dgamma[f] += sum_b part_g[b][f], dbeta likewise: 32 columns x 8 partial-row lanes per block.
__global__ void __launch_bounds__(256)
layernorm_param_reduce_kernel(const float *__restrict__ part_g, const float *__restrict__ part_b,
                              uint32_t nblocks, uint32_t F, float *dgamma, float *dbeta) {
  pdl_grid_sync();
  __shared__ float sg[8][33], sb[8][33];
  const uint32_t tx = threadIdx.x & 31, ty = threadIdx.x >> 5;
  const uint32_t f = blockIdx.x * 32 + tx;
  float a = 0.0f, c = 0.0f;
  if (f < F)
    for (uint32_t b = ty; b < nblocks; b += 8) {
      a += part_g[(uint64_t)b * F + f];
      c += part_b[(uint64_t)b * F + f];
    }
  sg[ty][tx] = a;
  sb[ty][tx] = c;
  __syncthreads();
  if (ty == 0 && f < F) {
#pragma unroll
    for (int k = 1; k < 8; ++k) {
      a += sg[k][tx];
      c += sb[k][tx];
    }
    if (dgamma) dgamma[f] += a;
    if (dbeta) dbeta[f] += c;
  }
}

// --------------------------------------------------------------------------- embedding / mask
// Reference cpu_forward / cpu_backward, src/ops/embedding.cpp:56-110. Thread per (token, feature)
// with the token index fastest: output rows are contiguous, weight reads are gathers.
__global__ void __launch_bounds__(256)
embedding_gather_kernel(const int32_t *__restrict__ idx, uint32_t idx_stride, uint32_t n,
                        const float *__restrict__ W, uint32_t w_s0, uint32_t w_s1, uint32_t D,
                        float *__restrict__ out, uint32_t o_s0, uint32_t o_s1) {
  pdl_grid_sync();
  const uint64_t total = (uint64_t)n * D, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const uint32_t i = (uint32_t)(t % n), d = (uint32_t)(t / n);
    const uint64_t tok = (uint32_t)idx[(uint64_t)i * idx_stride];
    out[(uint64_t)i * o_s0 + (uint64_t)d * o_s1] = W[tok * w_s0 + (uint64_t)d * w_s1];
  }
}
__global__ void __launch_bounds__(256)
embedding_scatter_kernel(float *dW, uint32_t w_s0, uint32_t w_s1, const int32_t *__restrict__ idx,
                         uint32_t idx_stride, uint32_t n, uint32_t D, const float *__restrict__ dout,
                         uint32_t o_s0, uint32_t o_s1) {
  pdl_grid_sync();
  const uint64_t total = (uint64_t)n * D, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const uint32_t i = (uint32_t)(t % n), d = (uint32_t)(t / n);
    const uint64_t tok = (uint32_t)idx[(uint64_t)i * idx_stride];
    atomicAdd(&dW[tok * w_s0 + (uint64_t)d * w_s1], dout[(uint64_t)i * o_s0 + (uint64_t)d * o_s1]);
  }
}
// cpu_triu_fill, src/ops/triu_fill.cpp:41-59
__global__ void __launch_bounds__(256)
triu_fill_kernel(float *a, uint32_t t0, uint32_t t1, uint32_t s0, uint32_t s1, float val,
                 uint32_t diagonal) {
  pdl_grid_sync();
  const uint64_t total = (uint64_t)t0 * t1, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const uint32_t i = (uint32_t)(t % t0), j = (uint32_t)(t / t0);
    if ((uint64_t)i + diagonal <= j) a[(uint64_t)i * s0 + (uint64_t)j * s1] = val;
  }
}

} // namespace weedcu

using namespace weedcu;

extern "C" {

int weedcu_layernorm_fwd(const float *x, uint32_t rows, uint32_t F, const float *gamma,
                         const float *beta, float eps, float *y, float *mean, float *rstd,
                         void *stream) {
  return weedcu_layernorm_fwd_bf16(x, rows, F, gamma, beta, eps, y, mean, rstd, nullptr, stream);
}

int weedcu_layernorm_fwd_bf16(const float *x, uint32_t rows, uint32_t F, const float *gamma, const float *beta, float eps,
                              float *y, float *mean, float *rstd, uint16_t *y_bf16, void *stream) {
  if (!x || !gamma || !beta || !y || !rows || !F) return WEEDCU_EINVAL;
  // the bf16 copy rides on the vectorised apply pass only
  if (y_bf16 && ((rows % 8u) || rows <= 256u || !aligned16(x) || !aligned16(y) || !aligned16(y_bf16) || (mean && !aligned16(mean)))) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  // (the single-launch register-tile kernel was also measured at 8192 x 768: 24.1 us against 22.9 us
  // for the two streaming passes, and a TMA-staged one-pass kernel — 32-row x F slab in shared memory,
  // two blocks per SM, in-place normalise, TMA store — at 20.5 us against 21.0 us (36.9 against 25.9 us
  // at F = 1024, one block per SM): every block sits in the same load / compute / store phase at the same
  // time, so the phases do not overlap. Cluster variants that split the features of a row tile over 4 or 8
  // CTAs with the partial sums exchanged through distributed shared memory (values in registers, x read
  // once, 3-4 CTAs per SM) measured 28.9 us (32-row tiles, cluster of 4) and 32.4 us (128-row tiles with
  // 128-bit loads, cluster of 8). Large inputs keep the two passes)
  if (rows <= 16u && F <= 2048u) { // a decode step's few tokens: block per row
    ProfScope prof(WEEDCU_PROF_LAYERNORM, st, 8.0 * (double)rows * F);
#define WCU_LN_ROW(NV) launch_k(layernorm_fwd_row_kernel<NV>, dim3(rows), dim3(256), 0, st, x, rows, F, gamma, beta, eps, y, mean, rstd)
    if (F <= 256u) WCU_LN_ROW(1);
    else if (F <= 512u) WCU_LN_ROW(2);
    else if (F <= 1024u) WCU_LN_ROW(4);
    else WCU_LN_ROW(8);
#undef WCU_LN_ROW
    return after_launch();
  }
  if (rows <= 256u && F <= (uint32_t)kLnBY * 64u) { // decode-sized inputs: one launch
    ProfScope prof(WEEDCU_PROF_LAYERNORM, st, 8.0 * (double)rows * F);
    const unsigned tiles = (rows + 31u) / 32u;
#define WCU_LN_SMALL(NV) launch_k(layernorm_fwd_small_kernel<NV>, dim3(tiles), dim3(32 * kLnBY), 0, st, x, rows, F, gamma, beta, eps, y, mean, rstd)
    if (F <= kLnBY * 16) WCU_LN_SMALL(16);
    else if (F <= kLnBY * 32) WCU_LN_SMALL(32);
    else if (F <= kLnBY * 48) WCU_LN_SMALL(48);
    else WCU_LN_SMALL(64);
#undef WCU_LN_SMALL
    return after_launch();
  }
  float *tmp = nullptr; // den[rows] (+ mu[rows] when the caller does not keep the mean)
  WCU_CHECK(pool_alloc((void **)&tmp, sizeof(float) * 2 * (size_t)rows, st));
  ProfScope prof(WEEDCU_PROF_LAYERNORM, st, (y_bf16 ? 10.0 : 8.0) * (double)rows * F);
  const unsigned tiles = (rows + 31u) / 32u;
#define WCU_LN_STATS(NV) launch_k(layernorm_fwd_stats_kernel<NV>, dim3(tiles), dim3(32 * kLnBY), 0, st, x, rows, F, eps, mean, rstd, tmp)
  if (F <= kLnBY * 16) WCU_LN_STATS(16);
  else if (F <= kLnBY * 32) WCU_LN_STATS(32);
  else if (F <= kLnBY * 48) WCU_LN_STATS(48);
  else if (F <= kLnBY * 64) WCU_LN_STATS(64);
  else WCU_LN_STATS(0);
#undef WCU_LN_STATS
  int rc = after_launch();
  if (rc == 0) {
    const float *mu = mean ? mean : tmp + rows;
    const unsigned fgroups = (F + kLnFB - 1) / kLnFB;
    if ((rows % 4u) == 0 && aligned16(x) && aligned16(y) && aligned16(mu) && fgroups <= 65535u) {
      launch_k(layernorm_fwd_apply_kernel<4>, dim3((rows / 4u + 255u) / 256u, fgroups), dim3(256), 0, st, x, rows, F, gamma, beta, mu, tmp, y, (__nv_bfloat16 *)y_bf16);
    } else if (fgroups <= 65535u) {
      launch_k(layernorm_fwd_apply_kernel<1>, dim3((rows + 255u) / 256u, fgroups), dim3(256), 0, st, x, rows, F, gamma, beta, mu, tmp, y, nullptr);
    } else {
      rc = WEEDCU_EINVAL;
    }
    if (rc == 0) rc = after_launch();
  }
  pool_free(tmp, st);
  return rc;
}

int weedcu_layernorm_fwd_stats(const float *x, uint32_t rows, uint32_t F, const float *stats, uint32_t tiles, uint32_t tile_cols,
                               const float *gamma, const float *beta, float eps, float *y, float *mean, float *rstd, uint16_t *y_bf16,
                               void *stream) {
  if (!x || !stats || !gamma || !beta || !y || !rows || !F || !tiles || !tile_cols) return WEEDCU_EINVAL;
  if ((uint64_t)tiles * tile_cols < F) return WEEDCU_EINVAL; // the tiles must cover [0, F)
  if (tiles > 16u && (uint64_t)16u * tile_cols < F) return WEEDCU_ENOSUP; // the merge kernel holds <= 16 live partials per row
  const unsigned fgroups = (F + kLnFB - 1) / kLnFB;
  if ((rows % 4u) || !aligned16(x) || !aligned16(y) || !aligned16(stats) || (y_bf16 && !aligned16(y_bf16)) || (mean && !aligned16(mean)) ||
      (rstd && !aligned16(rstd)) || fgroups > 65535u)
    return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  ProfScope prof(WEEDCU_PROF_LAYERNORM, st, (y_bf16 ? 10.0 : 8.0) * (double)rows * F);
  float *tmp = nullptr; // den[rows] (+ mu[rows] when the caller does not keep the mean)
  WCU_CHECK(pool_alloc((void **)&tmp, sizeof(float) * 2 * (size_t)rows, st));
  float *mu = mean ? mean : tmp + rows;
  launch_k(layernorm_stats_merge_kernel, dim3((rows + 127u) / 128u), dim3(128), 0, st, (const float2 *)stats, rows, F, tiles, tile_cols, eps, mu, tmp, rstd);
  int rc = after_launch();
  if (rc == 0) {
    launch_k(layernorm_fwd_apply_kernel<4>, dim3((rows / 4u + 255u) / 256u, fgroups), dim3(256), 0, st, x, rows, F, gamma, beta, (const float *)mu,
             (const float *)tmp, y, (__nv_bfloat16 *)y_bf16);
    rc = after_launch();
  }
  pool_free(tmp, st);
  return rc;
}

int weedcu_layernorm_bwd(const float *x, const float *dy, uint32_t rows, uint32_t F,
                         const float *gamma, const float *mean, const float *rstd, float *dx,
                         float *dgamma, float *dbeta, int grad_mode, int accumulate, void *stream) {
  return weedcu_layernorm_bwd_from(x, dy, rows, F, gamma, mean, rstd, accumulate ? dx : nullptr, dx, dgamma, dbeta, grad_mode, stream);
}

int weedcu_layernorm_bwd_from(const float *x, const float *dy, uint32_t rows, uint32_t F, const float *gamma, const float *mean,
                              const float *rstd, const float *dx_in, float *dx, float *dgamma, float *dbeta, int grad_mode,
                              void *stream) {
  if (!x || !dy || !gamma || !mean || !rstd || !dx || !rows || !F) return WEEDCU_EINVAL;
  const int accumulate = dx_in ? 1 : 0;
  const unsigned fgroups = (F + kLnFB - 1) / kLnFB;
  if (fgroups > 65535u) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  const bool vec = (rows % 4u) == 0 && aligned16(x) && aligned16(dy) && aligned16(dx) && (!dx_in || aligned16(dx_in)) && aligned16(mean) && aligned16(rstd);
  const uint32_t rows_per_block = vec ? 1024u : 256u;
  const uint32_t nchunks = (rows + rows_per_block - 1) / rows_per_block;
  // scratch: k_x[rows], k_0[rows] (rounded up to keep 16-byte alignment), part_g / part_b [nchunks][F]
  const size_t rows_up = ((size_t)rows + 3) & ~(size_t)3;
  float *scratch = nullptr;
  WCU_CHECK(pool_alloc((void **)&scratch, sizeof(float) * (2 * rows_up + 2 * (size_t)nchunks * F), st));
  float *kx = scratch, *k0 = scratch + rows_up, *pg = k0 + rows_up, *pb = pg + (size_t)nchunks * F;
  ProfScope prof(WEEDCU_PROF_LAYERNORM, st, (accumulate ? 16.0 : 12.0) * (double)rows * F);
  launch_k(layernorm_bwd_rows_kernel, dim3((rows + 31u) / 32u), dim3(32 * kLnBY), 0, st, x, dy, rows, F, gamma, mean, rstd, kx, k0, grad_mode);
  int rc = after_launch();
  if (rc == 0) {
    if (vec)
      launch_k(layernorm_bwd_apply_kernel<4>, dim3(nchunks, fgroups), dim3(256), 0, st, x, dy, rows, F, gamma, mean, rstd, kx, k0, dx_in, dx, pg, pb);
    else
      launch_k(layernorm_bwd_apply_kernel<1>, dim3(nchunks, fgroups), dim3(256), 0, st, x, dy, rows, F, gamma, mean, rstd, kx, k0, dx_in, dx, pg, pb);
    rc = after_launch();
  }
  if (rc == 0 && (dgamma || dbeta)) {
    launch_k(layernorm_param_reduce_kernel, dim3((F + 31) / 32), dim3(256), 0, st, pg, pb, nchunks, F, dgamma, dbeta);
    rc = after_launch();
  }
  pool_free(scratch, st);
  return rc;
}

int weedcu_embedding_gather(const int32_t *idx, uint64_t idx_off, uint32_t idx_stride, uint32_t n,
                            const float *W, uint64_t w_off, uint32_t w_s0, uint32_t w_s1,
                            uint32_t D, float *out, uint64_t o_off, uint32_t o_s0, uint32_t o_s1,
                            void *stream) {
  if (!idx || !W || !out || !n || !D) return WEEDCU_EINVAL;
  ProfScope prof(WEEDCU_PROF_EMBEDDING, resolve_stream(stream), 8.0 * (double)n * D);
  launch_k(embedding_gather_kernel, dim3(grid_for((uint64_t)n * D, 256, 32)), dim3(256), 0, resolve_stream(stream), 
      idx + idx_off, idx_stride, n, W + w_off, w_s0, w_s1, D, out + o_off, o_s0, o_s1);
  return after_launch();
}

int weedcu_embedding_scatter_add(float *dW, uint64_t w_off, uint32_t w_s0, uint32_t w_s1,
                                 const int32_t *idx, uint64_t idx_off, uint32_t idx_stride,
                                 uint32_t n, uint32_t D, const float *dout, uint64_t o_off,
                                 uint32_t o_s0, uint32_t o_s1, void *stream) {
  if (!dW || !idx || !dout || !n || !D) return WEEDCU_EINVAL;
  ProfScope prof(WEEDCU_PROF_EMBEDDING, resolve_stream(stream), 12.0 * (double)n * D);
  launch_k(embedding_scatter_kernel, dim3(grid_for((uint64_t)n * D, 256, 32)), dim3(256), 0, resolve_stream(stream), 
      dW + w_off, w_s0, w_s1, idx + idx_off, idx_stride, n, D, dout + o_off, o_s0, o_s1);
  return after_launch();
}

int weedcu_triu_fill_real(float *a, const weedcu_view *av, float val, uint32_t diagonal,
                          void *stream) {
  if (!a || !av || av->rank != 2) return WEEDCU_EINVAL;
  const uint64_t total = (uint64_t)av->shape[0] * av->shape[1];
  if (!total) return WEEDCU_EINVAL;
  launch_k(triu_fill_kernel, dim3(grid_for(total, 256, 16)), dim3(256), 0, resolve_stream(stream), 
      a + av->offset, av->shape[0], av->shape[1], av->stride[0], av->stride[1], val, diagonal);
  return after_launch();
}

} // extern "C"
