// layernorm.cu — fused LayerNorm forward/backward, embedding gather / scatter-add, triu fill.
//
// LayerNorm reference: LayerNorm::forward (src/modules/layernorm.cpp:29-42) composes ~12 tensor ops
// (mean, sub, mul, mean, +eps, ^0.5, div, *gamma, +beta) and ~30 backward closures. Here: one
// forward kernel (8 B/elem) and one backward kernel (16 B/elem + parameter partials).
// x is [rows, F] with row stride 1 and feature stride `rows`: the normalised axis is the slowest,
// so RT adjacent rows are tiled against all F features, staged once in shared memory, and the
// reference's two-pass statistics (mean, then mean of centred squares) run from the staged copy.
#include "common.cuh"

namespace weedcu {

constexpr size_t kLnMaxTileBytes = 160 * 1024;

// Loads are issued in batches of U independent requests per thread before anything consumes them:
// with one 4-byte load in flight per thread these kernels are latency-bound, not bandwidth-bound.
template <int RT, int NT, bool STAGED>
__global__ void __launch_bounds__(NT)
layernorm_fwd_kernel(const float *__restrict__ x, uint32_t rows, uint32_t F, const float *__restrict__ gamma,
                     const float *__restrict__ beta, float eps, float *__restrict__ y, float *__restrict__ mean,
                     float *__restrict__ rstd) {
  constexpr int BY = NT / RT, U = 8;
  extern __shared__ float tile[]; // [F][RT]
  __shared__ float red[BY][RT + 1];
  const uint32_t tx = threadIdx.x % RT, ty = threadIdx.x / RT;
  const uint32_t r = blockIdx.x * RT + tx;
  const bool live = r < rows;
  const float *p = x + r;

  float s = 0.0f;
  if (live)
    for (uint32_t f0 = ty; f0 < F; f0 += BY * U) {
      float v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f0 + u * BY;
        v[u] = (f < F) ? p[(uint64_t)f * rows] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f0 + u * BY;
        if (f < F) {
          if (STAGED) tile[f * RT + tx] = v[u];
          s += v[u];
        }
      }
    }
  red[ty][tx] = s;
  __syncthreads();
  s = 0.0f;
#pragma unroll
  for (int k = 0; k < BY; ++k) s += red[k][tx];
  const float mu = s / (float)F;
  __syncthreads();

  float q = 0.0f;
  if (live)
    for (uint32_t f = ty; f < F; f += BY) {
      const float xc = (STAGED ? tile[f * RT + tx] : p[(uint64_t)f * rows]) - mu;
      q += xc * xc;
    }
  red[ty][tx] = q;
  __syncthreads();
  q = 0.0f;
#pragma unroll
  for (int k = 0; k < BY; ++k) q += red[k][tx];
  const float den = sqrtf(q / (float)F + eps); // (var + eps) ^ 0.5, layernorm.cpp:35
  if (live) {
    if (ty == 0) {
      if (mean) mean[r] = mu;
      if (rstd) rstd[r] = 1.0f / den;
    }
    float *py = y + r;
#pragma unroll 4
    for (uint32_t f = ty; f < F; f += BY) {
      const float xc = (STAGED ? tile[f * RT + tx] : p[(uint64_t)f * rows]) - mu;
      py[(uint64_t)f * rows] = (xc / den) * gamma[f] + beta[f];
    }
  }
}

// dx (+)= rstd*(g - mean_f(g) - xhat*mean_f(g*xhat)), g = dy*gamma, xhat = (x-mean)*rstd.
// Persistent over row tiles; column sums of dy*xhat and dy accumulate in shared memory and leave
// as one partial row per block: part_g/part_b[blockIdx][F].
template <int RT, int NT, bool STAGED>
__global__ void __launch_bounds__(NT)
layernorm_bwd_kernel(const float *__restrict__ x, const float *__restrict__ dy, uint32_t rows, uint32_t F,
                     const float *__restrict__ gamma, const float *__restrict__ mean,
                     const float *__restrict__ rstd, float *dx, float *__restrict__ part_g,
                     float *__restrict__ part_b, int grad_mode, int accumulate, uint32_t ntiles) {
  constexpr int BY = NT / RT, U = 4;
  extern __shared__ float tile[]; // colg [F], colb [F], then (STAGED) xhat [F][RT], dy [F][RT]
  __shared__ float red_a[BY][RT + 1];
  __shared__ float red_b[BY][RT + 1];
  float *colg = tile, *colb = tile + F, *t_xh = tile + 2 * (size_t)F, *t_dy = t_xh + (size_t)F * RT;
  const uint32_t tx = threadIdx.x % RT, ty = threadIdx.x / RT;
  for (uint32_t f = threadIdx.x; f < F; f += NT) colg[f] = colb[f] = 0.0f;

  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint32_t r = t * RT + tx;
    const bool live = r < rows;
    const float mu = live ? mean[r] : 0.0f, rs = live ? rstd[r] : 0.0f;
    float sg = 0.0f, sgx = 0.0f;
    for (uint32_t f0 = ty; f0 < F; f0 += BY * U) {
      float xv[U], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f0 + u * BY;
        const bool ok = live && f < F;
        xv[u] = ok ? x[r + (uint64_t)f * rows] : mu;
        dv[u] = ok ? dy[r + (uint64_t)f * rows] : 0.0f;
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t f = f0 + u * BY;
        if (f < F) {
          const float xh = (xv[u] - mu) * rs;
          if (STAGED) {
            t_xh[f * RT + tx] = xh;
            t_dy[f * RT + tx] = dv[u];
          }
          const float g = dv[u] * gamma[f];
          sg += g;
          sgx += grad_mode ? g * xh : xh; // mode 0 needs sum_f(xc) = sum_f(xhat)/rstd instead
        }
      }
    }
    red_a[ty][tx] = sg;
    red_b[ty][tx] = sgx;
    __syncthreads();
    sg = sgx = 0.0f;
#pragma unroll
    for (int k = 0; k < BY; ++k) {
      sg += red_a[k][tx];
      sgx += red_b[k][tx];
    }
    // mode 1 (analytic):   dx += rs*(g - mean(g) - xh*mean(g*xh))
    // mode 0 (reference chain, see weedcu.h): dxc = g*rs + c*xc with c = -rs^3*sum(xc)/F;
    //                         dx += dxc - mean(dxc).  In xhat terms xc = xh/rs, sum(xc) = sgx/rs.
    float k_g, k_x, k_0;
    if (grad_mode) {
      k_g = rs;
      k_x = -rs * (sgx / (float)F);
      k_0 = -rs * (sg / (float)F);
    } else {
      const float sx = (rs != 0.0f) ? sgx / rs : 0.0f;
      const float c = -(rs * rs * rs) * sx / (float)F;
      k_g = rs;
      k_x = (rs != 0.0f) ? c / rs : 0.0f;
      k_0 = -((rs * sg + c * sx) / (float)F);
    }
#pragma unroll 2
    for (uint32_t f0 = 0; f0 < F; f0 += BY) { // uniform trip count: the shuffles below need it
      const uint32_t f = f0 + ty;
      const bool fv = f < F;
      float xh = 0.0f, d = 0.0f;
      if (fv) {
        if (STAGED) {
          xh = t_xh[f * RT + tx];
          d = t_dy[f * RT + tx];
        } else if (live) {
          xh = (x[r + (uint64_t)f * rows] - mu) * rs;
          d = dy[r + (uint64_t)f * rows];
        }
      }
      // column partials over the RT rows of this tile (lanes tx of one ty share f)
      float cg = d * xh, cb = d;
#pragma unroll
      for (int o = RT / 2; o > 0; o >>= 1) {
        cg += __shfl_xor_sync(0xffffffffu, cg, o, RT);
        cb += __shfl_xor_sync(0xffffffffu, cb, o, RT);
      }
      if (tx == 0 && fv) { // (ty, f0) owns column f: no two threads touch the same slot
        colg[f] += cg;
        colb[f] += cb;
      }
      if (live && fv) {
        const float val = (k_g * (d * gamma[f]) + k_x * xh) + k_0;
        float *o = dx + r + (uint64_t)f * rows;
        *o = accumulate ? (*o + val) : val;
      }
    }
    __syncthreads(); // the next tile reuses the staging tile and the reduction scratch
  }
  for (uint32_t f = threadIdx.x; f < F; f += NT) {
    part_g[(uint64_t)blockIdx.x * F + f] = colg[f];
    part_b[(uint64_t)blockIdx.x * F + f] = colb[f];
  }
}

// Register-resident variants for F <= BY*NV: a thread keeps its NV features of one row in registers,
// so every global load of the tile is in flight at once and nothing is staged in shared memory.
template <int RT, int BY, int NV>
__global__ void __launch_bounds__(RT *BY)
layernorm_fwd_reg_kernel(const float *__restrict__ x, uint32_t rows, uint32_t F, const float *__restrict__ gamma,
                         const float *__restrict__ beta, float eps, float *__restrict__ y,
                         float *__restrict__ mean, float *__restrict__ rstd) {
  __shared__ float red[BY][RT + 1];
  const uint32_t tx = threadIdx.x % RT, ty = threadIdx.x / RT;
  const uint32_t r = blockIdx.x * RT + tx;
  const bool live = r < rows;
  float v[NV];
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t f = ty + i * BY;
    v[i] = (live && f < F) ? x[r + (uint64_t)f * rows] : 0.0f;
  }
  float s = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) s += v[i];
  red[ty][tx] = s;
  __syncthreads();
  s = 0.0f;
#pragma unroll
  for (int k = 0; k < BY; ++k) s += red[k][tx];
  const float mu = s / (float)F;
  __syncthreads();
  float q = 0.0f;
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const float xc = v[i] - mu;
    if (ty + i * BY < F) q += xc * xc;
  }
  red[ty][tx] = q;
  __syncthreads();
  q = 0.0f;
#pragma unroll
  for (int k = 0; k < BY; ++k) q += red[k][tx];
  const float den = sqrtf(q / (float)F + eps); // (var + eps) ^ 0.5, layernorm.cpp:35
  if (!live) return;
  if (ty == 0) {
    if (mean) mean[r] = mu;
    if (rstd) rstd[r] = 1.0f / den;
  }
#pragma unroll
  for (int i = 0; i < NV; ++i) {
    const uint32_t f = ty + i * BY;
    if (f < F) y[r + (uint64_t)f * rows] = ((v[i] - mu) / den) * gamma[f] + beta[f];
  }
}

template <int RT, int BY, int NV>
__global__ void __launch_bounds__(RT *BY, (NV <= 24) ? 2 : 1)
layernorm_bwd_reg_kernel(const float *__restrict__ x, const float *__restrict__ dy, uint32_t rows, uint32_t F,
                         const float *__restrict__ gamma, const float *__restrict__ mean,
                         const float *__restrict__ rstd, float *dx, float *__restrict__ part_g,
                         float *__restrict__ part_b, int grad_mode, int accumulate, uint32_t ntiles) {
  extern __shared__ float tile[]; // colg [F], colb [F]
  __shared__ float red_a[BY][RT + 1];
  __shared__ float red_b[BY][RT + 1];
  float *colg = tile, *colb = tile + F;
  const uint32_t tx = threadIdx.x % RT, ty = threadIdx.x / RT;
  for (uint32_t f = threadIdx.x; f < F; f += RT * BY) colg[f] = colb[f] = 0.0f;

  for (uint32_t t = blockIdx.x; t < ntiles; t += gridDim.x) {
    const uint32_t r = t * RT + tx;
    const bool live = r < rows;
    const float mu = live ? mean[r] : 0.0f, rs = live ? rstd[r] : 0.0f;
    float xh[NV], d[NV];
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t f = ty + i * BY;
      const bool ok = live && f < F;
      xh[i] = ok ? x[r + (uint64_t)f * rows] : mu;
      d[i] = ok ? dy[r + (uint64_t)f * rows] : 0.0f;
    }
    float sg = 0.0f, sgx = 0.0f;
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t f = ty + i * BY;
      xh[i] = (xh[i] - mu) * rs;
      const float g = (f < F) ? d[i] * gamma[f] : 0.0f;
      sg += g;
      sgx += grad_mode ? g * xh[i] : xh[i]; // mode 0 needs sum_f(xc) = sum_f(xhat)/rstd instead
    }
    red_a[ty][tx] = sg;
    red_b[ty][tx] = sgx;
    __syncthreads();
    sg = sgx = 0.0f;
#pragma unroll
    for (int k = 0; k < BY; ++k) {
      sg += red_a[k][tx];
      sgx += red_b[k][tx];
    }
    float k_g, k_x, k_0; // see layernorm_bwd_kernel
    if (grad_mode) {
      k_g = rs;
      k_x = -rs * (sgx / (float)F);
      k_0 = -rs * (sg / (float)F);
    } else {
      const float sx = (rs != 0.0f) ? sgx / rs : 0.0f;
      const float c = -(rs * rs * rs) * sx / (float)F;
      k_g = rs;
      k_x = (rs != 0.0f) ? c / rs : 0.0f;
      k_0 = -((rs * sg + c * sx) / (float)F);
    }
#pragma unroll
    for (int i = 0; i < NV; ++i) {
      const uint32_t f = ty + i * BY;
      float cg = d[i] * xh[i], cb = d[i];
#pragma unroll
      for (int w = RT / 2; w > 0; w >>= 1) {
        cg += __shfl_xor_sync(0xffffffffu, cg, w, RT);
        cb += __shfl_xor_sync(0xffffffffu, cb, w, RT);
      }
      if (tx == 0 && f < F) { // (ty, i) owns column f
        colg[f] += cg;
        colb[f] += cb;
      }
      if (live && f < F) {
        const float val = (k_g * (d[i] * gamma[f]) + k_x * xh[i]) + k_0;
        float *o = dx + r + (uint64_t)f * rows;
        *o = accumulate ? (*o + val) : val;
      }
    }
    __syncthreads(); // red_a / red_b are reused by the next tile
  }
  for (uint32_t f = threadIdx.x; f < F; f += RT * BY) {
    part_g[(uint64_t)blockIdx.x * F + f] = colg[f];
    part_b[(uint64_t)blockIdx.x * F + f] = colb[f];
  }
}

// dgamma[f] += sum_b part_g[b][f], dbeta likewise: 32 columns x 8 partial-row lanes per block.
__global__ void __launch_bounds__(256)
layernorm_param_reduce_kernel(const float *__restrict__ part_g, const float *__restrict__ part_b,
                              uint32_t nblocks, uint32_t F, float *dgamma, float *dbeta) {
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
  const uint64_t total = (uint64_t)t0 * t1, stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t t = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; t < total; t += stride) {
    const uint32_t i = (uint32_t)(t % t0), j = (uint32_t)(t / t0);
    if ((uint64_t)i + diagonal <= j) a[(uint64_t)i * s0 + (uint64_t)j * s1] = val;
  }
}

template <int RT, int NT>
static int ln_fwd_launch(const float *x, uint32_t rows, uint32_t F, const float *gamma,
                         const float *beta, float eps, float *y, float *mean, float *rstd,
                         cudaStream_t st) {
  const unsigned grid = (rows + RT - 1) / RT;
  const size_t bytes = (size_t)F * RT * sizeof(float);
  if (bytes <= kLnMaxTileBytes) {
    auto k = layernorm_fwd_kernel<RT, NT, true>;
    ensure_dynamic_smem((const void *)k, (int)kLnMaxTileBytes);
    k<<<grid, NT, bytes, st>>>(x, rows, F, gamma, beta, eps, y, mean, rstd);
  } else {
    layernorm_fwd_kernel<RT, NT, false><<<grid, NT, 0, st>>>(x, rows, F, gamma, beta, eps, y, mean, rstd);
  }
  return after_launch();
}
template <int RT, int NT>
static int ln_bwd_launch(const float *x, const float *dy, uint32_t rows, uint32_t F,
                         const float *gamma, const float *mean, const float *rstd, float *dx,
                         float *pg, float *pb, int grad_mode, int accumulate, uint32_t nblocks,
                         cudaStream_t st) {
  const uint32_t ntiles = (rows + RT - 1) / RT;
  const size_t col_bytes = 2 * (size_t)F * sizeof(float);
  const size_t bytes = col_bytes + 2 * (size_t)F * RT * sizeof(float);
  if (bytes <= kLnMaxTileBytes) {
    auto k = layernorm_bwd_kernel<RT, NT, true>;
    ensure_dynamic_smem((const void *)k, (int)kLnMaxTileBytes);
    k<<<nblocks, NT, bytes, st>>>(x, dy, rows, F, gamma, mean, rstd, dx, pg, pb, grad_mode, accumulate, ntiles);
  } else {
    auto k = layernorm_bwd_kernel<RT, NT, false>;
    ensure_dynamic_smem((const void *)k, (int)kLnMaxTileBytes);
    k<<<nblocks, NT, col_bytes, st>>>(x, dy, rows, F, gamma, mean, rstd, dx, pg, pb, grad_mode, accumulate, ntiles);
  }
  return after_launch();
}
static int pick_rt(uint32_t rows) {
  if (rows / 32 >= 4 * kNumSMs) return 32;
  if (rows / 16 >= 2 * kNumSMs) return 16;
  return 8;
}

} // namespace weedcu

using namespace weedcu;

extern "C" {

int weedcu_layernorm_fwd(const float *x, uint32_t rows, uint32_t F, const float *gamma,
                         const float *beta, float eps, float *y, float *mean, float *rstd,
                         void *stream) {
  if (!x || !gamma || !beta || !y || !rows || !F) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  ProfScope prof(WEEDCU_PROF_LAYERNORM, st, 8.0 * (double)rows * F);
  if (F <= 32 * 32) { // register-resident rows: 8 rows x 32 feature lanes per block
    const unsigned grid = (rows + 7) / 8;
    if (F <= 32 * 8) layernorm_fwd_reg_kernel<8, 32, 8><<<grid, 256, 0, st>>>(x, rows, F, gamma, beta, eps, y, mean, rstd);
    else if (F <= 32 * 16) layernorm_fwd_reg_kernel<8, 32, 16><<<grid, 256, 0, st>>>(x, rows, F, gamma, beta, eps, y, mean, rstd);
    else if (F <= 32 * 24) layernorm_fwd_reg_kernel<8, 32, 24><<<grid, 256, 0, st>>>(x, rows, F, gamma, beta, eps, y, mean, rstd);
    else layernorm_fwd_reg_kernel<8, 32, 32><<<grid, 256, 0, st>>>(x, rows, F, gamma, beta, eps, y, mean, rstd);
    return after_launch();
  }
  switch (pick_rt(rows)) {
  case 32: return ln_fwd_launch<32, 512>(x, rows, F, gamma, beta, eps, y, mean, rstd, st);
  case 16: return ln_fwd_launch<16, 256>(x, rows, F, gamma, beta, eps, y, mean, rstd, st);
  default: return ln_fwd_launch<8, 256>(x, rows, F, gamma, beta, eps, y, mean, rstd, st);
  }
}

int weedcu_layernorm_bwd(const float *x, const float *dy, uint32_t rows, uint32_t F,
                         const float *gamma, const float *mean, const float *rstd, float *dx,
                         float *dgamma, float *dbeta, int grad_mode, int accumulate, void *stream) {
  if (!x || !dy || !gamma || !mean || !rstd || !dx || !rows || !F) return WEEDCU_EINVAL;
  if (2 * (size_t)F * sizeof(float) > kLnMaxTileBytes) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  const bool reg = F <= 32 * 32;
  const int rt = reg ? 8 : pick_rt(rows);
  const uint32_t ntiles = (rows + rt - 1) / rt;
  const uint32_t nblocks = ntiles < 4u * kNumSMs ? ntiles : 4u * kNumSMs;
  float *part = nullptr;
  WCU_CHECK(pool_alloc((void **)&part, sizeof(float) * 2 * (size_t)nblocks * F, st));
  float *pg = part, *pb = part + (size_t)nblocks * F;
  ProfScope prof(WEEDCU_PROF_LAYERNORM, st, (accumulate ? 16.0 : 12.0) * (double)rows * F);
  int rc;
  if (reg) {
    const size_t cb = 2 * (size_t)F * sizeof(float);
#define WCU_LN_BWD_REG(NV)                                                                          \
  layernorm_bwd_reg_kernel<8, 32, NV><<<nblocks, 256, cb, st>>>(x, dy, rows, F, gamma, mean, rstd, dx, pg, pb, grad_mode, \
                                                                 accumulate, ntiles)
    if (F <= 32 * 8) WCU_LN_BWD_REG(8);
    else if (F <= 32 * 16) WCU_LN_BWD_REG(16);
    else if (F <= 32 * 24) WCU_LN_BWD_REG(24);
    else WCU_LN_BWD_REG(32);
#undef WCU_LN_BWD_REG
    rc = after_launch();
  } else
  switch (rt) {
  case 32: rc = ln_bwd_launch<32, 512>(x, dy, rows, F, gamma, mean, rstd, dx, pg, pb, grad_mode, accumulate, nblocks, st); break;
  case 16: rc = ln_bwd_launch<16, 256>(x, dy, rows, F, gamma, mean, rstd, dx, pg, pb, grad_mode, accumulate, nblocks, st); break;
  default: rc = ln_bwd_launch<8, 256>(x, dy, rows, F, gamma, mean, rstd, dx, pg, pb, grad_mode, accumulate, nblocks, st); break;
  }
  if (rc == 0 && (dgamma || dbeta)) {
    layernorm_param_reduce_kernel<<<(F + 31) / 32, 256, 0, st>>>(pg, pb, nblocks, F, dgamma, dbeta);
    rc = after_launch();
  }
  pool_free(part, st);
  return rc;
}

int weedcu_embedding_gather(const int32_t *idx, uint64_t idx_off, uint32_t idx_stride, uint32_t n,
                            const float *W, uint64_t w_off, uint32_t w_s0, uint32_t w_s1,
                            uint32_t D, float *out, uint64_t o_off, uint32_t o_s0, uint32_t o_s1,
                            void *stream) {
  if (!idx || !W || !out || !n || !D) return WEEDCU_EINVAL;
  ProfScope prof(WEEDCU_PROF_EMBEDDING, resolve_stream(stream), 8.0 * (double)n * D);
  embedding_gather_kernel<<<grid_for((uint64_t)n * D, 256, 32), 256, 0, resolve_stream(stream)>>>(
      idx + idx_off, idx_stride, n, W + w_off, w_s0, w_s1, D, out + o_off, o_s0, o_s1);
  return after_launch();
}

int weedcu_embedding_scatter_add(float *dW, uint64_t w_off, uint32_t w_s0, uint32_t w_s1,
                                 const int32_t *idx, uint64_t idx_off, uint32_t idx_stride,
                                 uint32_t n, uint32_t D, const float *dout, uint64_t o_off,
                                 uint32_t o_s0, uint32_t o_s1, void *stream) {
  if (!dW || !idx || !dout || !n || !D) return WEEDCU_EINVAL;
  ProfScope prof(WEEDCU_PROF_EMBEDDING, resolve_stream(stream), 12.0 * (double)n * D);
  embedding_scatter_kernel<<<grid_for((uint64_t)n * D, 256, 32), 256, 0, resolve_stream(stream)>>>(
      dW + w_off, w_s0, w_s1, idx + idx_off, idx_stride, n, D, dout + o_off, o_s0, o_s1);
  return after_launch();
}

int weedcu_triu_fill_real(float *a, const weedcu_view *av, float val, uint32_t diagonal,
                          void *stream) {
  if (!a || !av || av->rank != 2) return WEEDCU_EINVAL;
  const uint64_t total = (uint64_t)av->shape[0] * av->shape[1];
  if (!total) return WEEDCU_EINVAL;
  triu_fill_kernel<<<grid_for(total, 256, 16), 256, 0, resolve_stream(stream)>>>(
      a + av->offset, av->shape[0], av->shape[1], av->stride[0], av->stride[1], val, diagonal);
  return after_launch();
}

} // extern "C"
