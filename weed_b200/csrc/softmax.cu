// softmax.cu — softmax / log-softmax forward + backward, fused attention softmax (scale + causal
// mask + softmax), fused cross-entropy.  HBM-bound: 8 B/elem forward, 16 B/elem backward.
//
// Reference: Weed::softmax / softmax_grad (src/ops/softmax.cpp:85-138), Weed::logsoftmax /
// logsoftmax_grad (src/ops/logsoftmax.cpp:87-149), MultiHeadAttention::forward's scale/mask/softmax
// chain (src/modules/multihead_attention.cpp:319-334), cross_entropy_loss
// (include/autograd/cross_entropy_loss.hpp:21-34).
//
// Layout note (SURVEY §7 hard part 7): in column-major tensors the softmax axis is normally the
// SLOWEST dim, so a "row" is strided and adjacent rows are contiguous. The main kernels therefore
// tile RT adjacent rows x all L columns: a warp reads 32 adjacent rows (one 128-B line) per column,
// BY warps split the columns, the tile is staged once in shared memory (<= 200 KB) and the three
// reference passes (max, sum exp, normalise) run out of that staging copy.
#include "common.cuh"
#include <cuda_bf16.h>

namespace weedcu {

constexpr int kRT = 32;                       // rows per tile (one warp wide)
constexpr size_t kMaxTileBytes = 200 * 1024;  // dynamic shared memory budget for a staged tile

struct PlainLoad {
  __device__ float operator()(float x, uint32_t, uint32_t) const { return x; }
  __device__ uint32_t cols(uint32_t, uint32_t L) const { return L; }
};
// scores / sqrt(hd) + triu mask (multihead_attention.cpp:319-328; triu_fill.cpp:48-56)
struct AttnLoad {
  uint32_t batch;
  float divisor, mask_val;
  int causal;
  __device__ float operator()(float x, uint32_t row, uint32_t col) const {
    float v = x / divisor;
    if (causal) {
      const uint32_t q = row / batch;
      v = v + ((q + 1u <= col) ? mask_val : 0.0f);
    }
    return v;
  }
  // columns some row of a tile ending at row_last can see: beyond them exp(v + mask - max) == 0.0f
  __device__ uint32_t cols(uint32_t row_last, uint32_t L) const {
    return (causal && mask_val <= -1e30f) ? min(L, row_last / batch + 1u) : L;
  }
};

// ----------------------------------------------------------------------------- forward, strided
// a[inner, L, outer] canonical contiguous; rows = inner (per outer slab). Block = RT rows x BY column
// lanes (a warp covers RT adjacent rows x 32/RT columns, so every access is whole 32-B sectors); the
// host picks RT so that the staged tile stays ~32 KB and several CTAs are resident per SM — their
// load / exp / store phases then overlap each other.
// LD::cols() lets the attention variant skip columns that are masked for every row of the tile
// (exp underflows to exactly 0 there): they are neither read nor exponentiated, only zero-filled.
template <bool LOG, int RT, int BY, bool STAGED, class LD>
__global__ void __launch_bounds__(RT * BY)
softmax_strided_fwd(const float *a, float *out, uint32_t inner, uint32_t L,
                    LD ld) {
  pdl_grid_sync();
  extern __shared__ float tile[]; // [L][RT] when STAGED
  __shared__ float red_m[BY][RT + 1];
  __shared__ float red_s[BY][RT + 1];
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint32_t r0 = blockIdx.x * RT;
  const uint32_t r = r0 + tx;
  const bool live = r < inner;
  const uint64_t slab = (uint64_t)blockIdx.y * inner * L;
  const float *p = a + slab + r;
  float *po = out + slab + r;
  const uint32_t grow = r; // row id inside the slab (attention: b + batch*q)
  const uint32_t Lc = ld.cols(min(r0 + RT, inner) - 1u, L);

  float mx = -INFINITY, s = 0.0f;
  if (STAGED) {
    if (live)
      for (uint32_t j = ty; j < Lc; j += BY) {
        const float v = ld(p[(uint64_t)j * inner], grow, j);
        tile[j * RT + tx] = v;
        mx = fmaxf(mx, v);
      }
    red_m[ty][tx] = mx;
    __syncthreads();
#pragma unroll
    for (int y = 0; y < BY; ++y) mx = fmaxf(mx, red_m[y][tx]);
    if (live)
      for (uint32_t j = ty; j < Lc; j += BY) {
        const float e = expf(tile[j * RT + tx] - mx);
        if (!LOG) tile[j * RT + tx] = e;
        s += e;
      }
    red_s[ty][tx] = s;
    __syncthreads();
    s = 0.0f;
#pragma unroll
    for (int y = 0; y < BY; ++y) s += red_s[y][tx];
    if (live) {
      if (LOG) {
        const float log_s = logf(s);
        for (uint32_t j = ty; j < Lc; j += BY) po[(uint64_t)j * inner] = (tile[j * RT + tx] - mx) - log_s;
      } else {
        for (uint32_t j = ty; j < Lc; j += BY) po[(uint64_t)j * inner] = tile[j * RT + tx] / s;
      }
    }
  } else {
    // Row too long to stage: one online (max, sum) pass + one normalise pass (12 B/elem).
    if (live)
      for (uint32_t j = ty; j < Lc; j += BY) {
        const float v = ld(p[(uint64_t)j * inner], grow, j);
        if (v > mx) {
          s = s * expf(mx - v) + 1.0f;
          mx = v;
        } else {
          s += expf(v - mx);
        }
      }
    red_m[ty][tx] = mx;
    red_s[ty][tx] = s;
    __syncthreads();
    float M = -INFINITY;
#pragma unroll
    for (int y = 0; y < BY; ++y) M = fmaxf(M, red_m[y][tx]);
    float S = 0.0f;
#pragma unroll
    for (int y = 0; y < BY; ++y) {
      const float my = red_m[y][tx];
      if (my > -INFINITY) S += red_s[y][tx] * expf(my - M);
    }
    if (live) {
      const float log_s = logf(S);
      for (uint32_t j = ty; j < Lc; j += BY) {
        const float v = ld(p[(uint64_t)j * inner], grow, j);
        po[(uint64_t)j * inner] = LOG ? ((v - M) - log_s) : (expf(v - M) / S);
      }
    }
  }
  if (live)
    for (uint32_t j = Lc + ty; j < L; j += BY) po[(uint64_t)j * inner] = 0.0f;
}

// ----------------------------------------------------------------------------- forward, two streaming passes
// For long rows and big tensors the staged tile kernel above is latency-bound (load, reduce,
// synchronise, only then store: 0.37 of HBM at [8,12,1024,1024], 0.07 once a row no longer fits the
// tile). Same split as LayerNorm: pass 1 leaves the online (max, sum exp) of every row — block = 32
// adjacent rows x 16 warps striding the columns, 8 full-line loads in flight per thread, the vocab /
// key range split over blockIdx.z so the grid fills the GPU whatever the row count; pass 2 is plain
// 128-bit elementwise work. 12 B/elem of traffic (8 when the tensor fits the 126 MB L2).
constexpr int kSmBY = 16, kSmU = 8, kSmCols = 8;
__global__ void __launch_bounds__(32 * kSmBY)
softmax_rows_partial_kernel(const float *__restrict__ a, uint32_t inner, uint32_t L, uint32_t cols_per_split,
                            float *__restrict__ part_m, float *__restrict__ part_s) {
  pdl_grid_sync();
  __shared__ float red_m[kSmBY][33];
  __shared__ float red_s[kSmBY][33];
  const uint32_t tx = threadIdx.x & 31u, ty = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * 32u + tx;
  const bool live = r < inner;
  const uint64_t slab = (uint64_t)blockIdx.y * inner * L;
  const float *p = a + slab + r;
  const uint32_t j_begin = blockIdx.z * cols_per_split, j_end = min(L, j_begin + cols_per_split);
  float mx = -INFINITY, s = 0.0f;
  if (live)
    for (uint32_t j0 = j_begin + ty; j0 < j_end; j0 += kSmBY * kSmU) {
      float v[kSmU];
#pragma unroll
      for (int u = 0; u < kSmU; ++u) {
        const uint32_t j = j0 + u * kSmBY;
        v[u] = (j < j_end) ? p[(uint64_t)j * inner] : -INFINITY;
      }
      float m8 = v[0];
#pragma unroll
      for (int u = 1; u < kSmU; ++u) m8 = fmaxf(m8, v[u]);
      if (m8 > mx) { // rescale the running sum once per batch
        s *= __expf(mx - m8);
        mx = m8;
      }
      if (mx > -INFINITY) { // (__expf = ex2.approx of the scaled argument, 2 ulp: the pass was issue-bound with expf's ~15 instructions)
        float s0 = 0.0f, s1 = 0.0f;
#pragma unroll
        for (int u = 0; u < kSmU; u += 2) {
          s0 += __expf(v[u] - mx);
          s1 += __expf(v[u + 1] - mx);
        }
        s += s0 + s1;
      }
    }
  red_m[ty][tx] = mx;
  red_s[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && live) {
    float M = -INFINITY;
#pragma unroll
    for (int y = 0; y < kSmBY; ++y) M = fmaxf(M, red_m[y][tx]);
    float S = 0.0f;
#pragma unroll
    for (int y = 0; y < kSmBY; ++y) {
      const float my = red_m[y][tx];
      if (my > -INFINITY) S += red_s[y][tx] * expf(my - M);
    }
    const uint64_t row = (uint64_t)blockIdx.y * inner + r, n_rows = (uint64_t)gridDim.y * inner;
    part_m[(uint64_t)blockIdx.z * n_rows + row] = M;
    part_s[(uint64_t)blockIdx.z * n_rows + row] = S;
  }
}
// merge the column splits: M[row], S[row] (or log S for log-softmax)
template <bool LOG>
__global__ void __launch_bounds__(256)
softmax_rows_finish_kernel(const float *__restrict__ part_m, const float *__restrict__ part_s, uint32_t splits, uint64_t n_rows,
                           float *__restrict__ row_m, float *__restrict__ row_s) {
  pdl_grid_sync();
  const uint64_t row = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (row >= n_rows) return;
  float M = -INFINITY;
  for (uint32_t k = 0; k < splits; ++k) M = fmaxf(M, part_m[(uint64_t)k * n_rows + row]);
  float S = 0.0f;
  for (uint32_t k = 0; k < splits; ++k) {
    const float my = part_m[(uint64_t)k * n_rows + row];
    if (my > -INFINITY) S += part_s[(uint64_t)k * n_rows + row] * expf(my - M);
  }
  row_m[row] = M;
  row_s[row] = LOG ? logf(S) : 1.0f / S; // softmax_apply multiplies
}
// out = exp(x - M) / S   or   (x - M) - log S ; VEC adjacent rows x kSmCols columns per thread
template <bool LOG, int VEC>
__global__ void __launch_bounds__(256)
softmax_apply_kernel(const float *__restrict__ a, float *__restrict__ out, uint32_t inner, uint32_t L,
                     const float *__restrict__ row_m, const float *__restrict__ row_s) {
  pdl_grid_sync();
  const uint32_t r = (blockIdx.x * 256u + threadIdx.x) * VEC;
  if (r >= inner) return;
  const uint64_t slab = (uint64_t)blockIdx.z * inner * L, rbase = (uint64_t)blockIdx.z * inner + r;
  float M[VEC], S[VEC];
  if (VEC == 4) {
    *reinterpret_cast<float4 *>(M) = *reinterpret_cast<const float4 *>(row_m + rbase);
    *reinterpret_cast<float4 *>(S) = *reinterpret_cast<const float4 *>(row_s + rbase);
  } else {
    M[0] = row_m[rbase];
    S[0] = row_s[rbase];
  }
  const uint32_t j_begin = blockIdx.y * kSmCols, j_end = min(L, j_begin + kSmCols);
  float xv[kSmCols][VEC];
#pragma unroll
  for (int i = 0; i < kSmCols; ++i) {
    const uint32_t j = j_begin + i;
    if (j < j_end) {
      const float *src = a + slab + (uint64_t)j * inner + r;
      if (VEC == 4) *reinterpret_cast<float4 *>(xv[i]) = *reinterpret_cast<const float4 *>(src);
      else xv[i][0] = *src;
    }
  }
#pragma unroll
  for (int i = 0; i < kSmCols; ++i) {
    const uint32_t j = j_begin + i;
    if (j < j_end) {
      float o[VEC];
#pragma unroll
      for (int k = 0; k < VEC; ++k) o[k] = LOG ? ((xv[i][k] - M[k]) - S[k]) : (__expf(xv[i][k] - M[k]) * S[k]);
      float *dst = out + slab + (uint64_t)j * inner + r;
      if (VEC == 4) *reinterpret_cast<float4 *>(dst) = *reinterpret_cast<const float4 *>(o);
      else *dst = o[0];
    }
  }
}

// ----------------------------------------------------------------------------- forward, contiguous
// inner == 1: each row is L contiguous floats; one block per row.
template <bool LOG>
__global__ void __launch_bounds__(256)
softmax_contig_fwd(const float *__restrict__ a, float *__restrict__ out, uint32_t L) {
  pdl_grid_sync();
  __shared__ float red[32];
  const float *p = a + (uint64_t)blockIdx.x * L;
  float *po = out + (uint64_t)blockIdx.x * L;
  float mx = -INFINITY;
  for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) mx = fmaxf(mx, p[j]);
  mx = block_max(mx, red);
  float s = 0.0f;
  for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) s += expf(p[j] - mx);
  s = block_sum(s, red);
  const float log_s = logf(s);
  for (uint32_t j = threadIdx.x; j < L; j += blockDim.x)
    po[j] = LOG ? ((p[j] - mx) - log_s) : (expf(p[j] - mx) / s);
}

// ----------------------------------------------------------------------------- generic (any view)
struct RowView {
  int rank, axis;
  uint64_t offset;
  uint32_t shape[kMaxRank];
  uint32_t stride[kMaxRank];
};
__device__ __forceinline__ uint64_t row_base(const RowView &v, uint32_t o) {
  uint64_t base = v.offset;
  for (int d = 0; d < v.rank; ++d) {
    if (d == v.axis) continue;
    base += (uint64_t)(o % v.shape[d]) * v.stride[d];
    o /= v.shape[d];
  }
  return base;
}
template <bool LOG>
__global__ void __launch_bounds__(128)
softmax_generic_fwd(const float *__restrict__ a, RowView av, float *__restrict__ out, RowView ov, uint32_t n_rows) {
  pdl_grid_sync();
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_rows) return;
  const uint32_t L = av.shape[av.axis];
  const uint64_t ab = row_base(av, o), ob = row_base(ov, o), as = av.stride[av.axis],
                 os = ov.stride[ov.axis];
  float mx = a[ab];
  for (uint32_t j = 1; j < L; ++j) mx = fmaxf(mx, a[ab + j * as]);
  float s = 0.0f;
  for (uint32_t j = 0; j < L; ++j) s += expf(a[ab + j * as] - mx);
  const float log_s = logf(s);
  for (uint32_t j = 0; j < L; ++j) {
    const float v = a[ab + j * as];
    out[ob + j * os] = LOG ? ((v - mx) - log_s) : (expf(v - mx) / s);
  }
}
template <bool LOG>
__global__ void __launch_bounds__(128)
softmax_generic_bwd(float *din, RowView iv, const float *__restrict__ out, RowView ov,
                    const float *__restrict__ dout, RowView dv, uint32_t n_rows) {
  pdl_grid_sync();
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_rows) return;
  const uint32_t L = iv.shape[iv.axis];
  const uint64_t ib = row_base(iv, o), ob = row_base(ov, o), db = row_base(dv, o);
  const uint64_t is = iv.stride[iv.axis], os = ov.stride[ov.axis], ds = dv.stride[dv.axis];
  float acc = 0.0f;
  for (uint32_t j = 0; j < L; ++j)
    acc += LOG ? dout[db + j * ds] : out[ob + j * os] * dout[db + j * ds];
  for (uint32_t j = 0; j < L; ++j) {
    const float y = out[ob + j * os], g = dout[db + j * ds];
    din[ib + j * is] += LOG ? (g - expf(y) * acc) : (y * (g - acc));
  }
}

// ----------------------------------------------------------------------------- backward, strided
// din += out*(dout - sum(dout*out))            (softmax.cpp:110-127)
// din += dout - exp(out)*sum(dout)             (logsoftmax.cpp:119-137)
template <bool LOG, int BY, bool STAGED>
__global__ void __launch_bounds__(kRT * BY)
softmax_strided_bwd(float *din, const float *__restrict__ out, const float *__restrict__ dout,
                    uint32_t inner, uint32_t L) {
  pdl_grid_sync();
  extern __shared__ float tile[]; // STAGED: [2][L][kRT]  (out, dout)
  __shared__ float red[BY][kRT + 1];
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint32_t r = blockIdx.x * kRT + tx;
  const bool live = r < inner;
  const uint64_t slab = (uint64_t)blockIdx.y * inner * L;
  const float *py = out + slab + r, *pg = dout + slab + r;
  float *pd = din + slab + r;
  float *ty_ = tile, *tg_ = tile + (size_t)L * kRT;
  float acc = 0.0f;
  if (live)
    for (uint32_t j = ty; j < L; j += BY) {
      const float y = py[(uint64_t)j * inner], g = pg[(uint64_t)j * inner];
      if (STAGED) {
        ty_[j * kRT + tx] = y;
        tg_[j * kRT + tx] = g;
      }
      acc += LOG ? g : y * g;
    }
  red[ty][tx] = acc;
  __syncthreads();
  acc = 0.0f;
#pragma unroll
  for (int y = 0; y < BY; ++y) acc += red[y][tx];
  if (live)
    for (uint32_t j = ty; j < L; j += BY) {
      const float y = STAGED ? ty_[j * kRT + tx] : py[(uint64_t)j * inner];
      const float g = STAGED ? tg_[j * kRT + tx] : pg[(uint64_t)j * inner];
      pd[(uint64_t)j * inner] += LOG ? (g - expf(y) * acc) : (y * (g - acc));
    }
}
template <bool LOG>
__global__ void __launch_bounds__(256)
softmax_contig_bwd(float *din, const float *__restrict__ out, const float *__restrict__ dout, uint32_t L) {
  pdl_grid_sync();
  __shared__ float red[32];
  const uint64_t b = (uint64_t)blockIdx.x * L;
  float acc = 0.0f;
  for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) acc += LOG ? dout[b + j] : out[b + j] * dout[b + j];
  acc = block_sum(acc, red);
  for (uint32_t j = threadIdx.x; j < L; j += blockDim.x) {
    const float y = out[b + j], g = dout[b + j];
    din[b + j] += LOG ? (g - expf(y) * acc) : (y * (g - acc));
  }
}

// ----------------------------------------------------------------------------- cross entropy
// logits[rows, V], row stride rs, vocab stride vs. One read of the logits (4 B/elem): the vocab
// range is split over blockIdx.y so that the grid fills the GPU whatever `rows` is; each block keeps
// an online (max, sum exp) per row over its slice, `ce_fwd_finish` merges the slices, gathers the
// target logit and writes lse[r] and nll[r] = x[r,target] - lse[r] (loss = -mean).
constexpr int kCeRT = 32, kCeBY = 8, kCeU = 8; // 8 independent 128-B row loads in flight per thread (64 KB per SM at 8 resident blocks)
__global__ void __launch_bounds__(kCeRT * kCeBY)
ce_fwd_partial(const float *__restrict__ x, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs,
               uint32_t v_per_block, float *__restrict__ part_m, float *__restrict__ part_s) {
  pdl_grid_sync();
  __shared__ float red_m[kCeBY][kCeRT + 1];
  __shared__ float red_s[kCeBY][kCeRT + 1];
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint32_t r = blockIdx.x * kCeRT + tx;
  const bool live = r < rows;
  const uint32_t v0 = blockIdx.y * v_per_block, v1 = min(V, v0 + v_per_block);
  const float *p = x + (uint64_t)r * rs;
  float mx = -INFINITY, s = 0.0f;
  if (live)
    for (uint32_t j0 = v0 + ty; j0 < v1; j0 += kCeBY * kCeU) {
      float v[kCeU];
#pragma unroll
      for (int u = 0; u < kCeU; ++u) {
        const uint32_t j = j0 + u * kCeBY;
        v[u] = (j < v1) ? p[(uint64_t)j * vs] : -INFINITY;
      }
      float m4 = v[0];
#pragma unroll
      for (int u = 1; u < kCeU; ++u) m4 = fmaxf(m4, v[u]);
      if (m4 > mx) { // rescale the running sum once per batch
        s *= expf(mx - m4);
        mx = m4;
      }
      if (mx > -INFINITY) {
#pragma unroll
        for (int u = 0; u < kCeU; ++u) s += expf(v[u] - mx);
      }
    }
  red_m[ty][tx] = mx;
  red_s[ty][tx] = s;
  __syncthreads();
  if (ty == 0 && live) {
    float M = -INFINITY;
#pragma unroll
    for (int y = 0; y < kCeBY; ++y) M = fmaxf(M, red_m[y][tx]);
    float S = 0.0f;
#pragma unroll
    for (int y = 0; y < kCeBY; ++y) {
      const float my = red_m[y][tx];
      if (my > -INFINITY) S += red_s[y][tx] * expf(my - M);
    }
    part_m[(uint64_t)blockIdx.y * rows + r] = M;
    part_s[(uint64_t)blockIdx.y * rows + r] = S;
  }
}
// The same for rs == 1, rows % 4 == 0: a thread owns 4 adjacent rows (one 128-bit load per vocabulary entry), a warp
// reads 512 contiguous bytes and a block covers 128 rows — with 32-row tiles every 128-byte line of a tile sits in a
// different DRAM page and the pass ran at 4.4 TB/s.
__global__ void __launch_bounds__(kCeRT * kCeBY)
ce_fwd_partial_vec4(const float *__restrict__ x, uint32_t rows, uint32_t V, uint32_t vs, uint32_t v_per_block,
                    float *__restrict__ part_m, float *__restrict__ part_s) {
  pdl_grid_sync();
  __shared__ __align__(16) float red_m[kCeBY][4 * kCeRT];
  __shared__ __align__(16) float red_s[kCeBY][4 * kCeRT];
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint32_t r = (blockIdx.x * kCeRT + tx) * 4u;
  const bool live = r < rows;
  const uint32_t v0 = blockIdx.y * v_per_block, v1 = min(V, v0 + v_per_block);
  const float *p = x + r;
  float mx[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY}, s[4] = {0.f, 0.f, 0.f, 0.f};
  constexpr int U = 4; // 4 x 128-bit loads in flight per thread
  if (live)
    for (uint32_t j0 = v0 + ty; j0 < v1; j0 += kCeBY * U) {
      float4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = j0 + u * kCeBY;
        v[u] = (j < v1) ? *reinterpret_cast<const float4 *>(p + (uint64_t)j * vs) : make_float4(-INFINITY, -INFINITY, -INFINITY, -INFINITY);
      }
#pragma unroll
      for (int k = 0; k < 4; ++k) {
        float e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) e[u] = k == 0 ? v[u].x : k == 1 ? v[u].y : k == 2 ? v[u].z : v[u].w;
        float m4 = e[0];
#pragma unroll
        for (int u = 1; u < U; ++u) m4 = fmaxf(m4, e[u]);
        if (m4 > mx[k]) { // rescale the running sum once per batch
          s[k] *= expf(mx[k] - m4);
          mx[k] = m4;
        }
        if (mx[k] > -INFINITY) {
#pragma unroll
          for (int u = 0; u < U; ++u) s[k] += expf(e[u] - mx[k]);
        }
      }
    }
  *reinterpret_cast<float4 *>(&red_m[ty][4 * tx]) = make_float4(mx[0], mx[1], mx[2], mx[3]);
  *reinterpret_cast<float4 *>(&red_s[ty][4 * tx]) = make_float4(s[0], s[1], s[2], s[3]);
  __syncthreads();
  const uint32_t t = ty * kCeRT + tx; // 256 threads, 128 rows: the first 128 merge one row each
  if (t < 4 * kCeRT) {
    const uint32_t rr = blockIdx.x * 4 * kCeRT + t;
    if (rr < rows) {
      float M = -INFINITY;
#pragma unroll
      for (int y = 0; y < kCeBY; ++y) M = fmaxf(M, red_m[y][t]);
      float S = 0.0f;
#pragma unroll
      for (int y = 0; y < kCeBY; ++y) {
        const float my = red_m[y][t];
        if (my > -INFINITY) S += red_s[y][t] * expf(my - M);
      }
      part_m[(uint64_t)blockIdx.y * rows + rr] = M;
      part_s[(uint64_t)blockIdx.y * rows + rr] = S;
    }
  }
}
__global__ void __launch_bounds__(256)
ce_fwd_finish(const float *__restrict__ x, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs,
              const int32_t *__restrict__ targets, const float *__restrict__ part_m,
              const float *__restrict__ part_s, uint32_t splits, float *__restrict__ lse,
              float *__restrict__ nll) {
  pdl_grid_sync();
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float M = -INFINITY;
  for (uint32_t k = 0; k < splits; ++k) M = fmaxf(M, part_m[(uint64_t)k * rows + r]);
  float S = 0.0f;
  for (uint32_t k = 0; k < splits; ++k) {
    const float my = part_m[(uint64_t)k * rows + r];
    if (my > -INFINITY) S += part_s[(uint64_t)k * rows + r] * expf(my - M);
  }
  const float log_s = logf(S);
  const uint32_t t = (uint32_t)targets[r];
  const float xt = (t < V) ? x[(uint64_t)r * rs + (uint64_t)t * vs] : NAN;
  lse[r] = M + log_s;
  nll[r] = (xt - M) - log_s; // = lsm[r, target]
}
__device__ __forceinline__ float exp2f_approx_ce(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
// Online (max, sum exp) partials from the bf16 copy of the logits ([rows, V], rows contiguous, rows % 8 == 0): a thread
// owns 8 adjacent rows (one 128-bit load per vocabulary entry), a block 256 rows x a vocabulary slice. 2 B/elem read
// instead of 4; partials in the (max, sum) pair layout ce_fwd_stats_finish merges.
__global__ void __launch_bounds__(kCeRT * kCeBY)
ce_fwd_partial_bf16(const __nv_bfloat16 *__restrict__ x, uint32_t rows, uint32_t V, uint32_t v_per_block, float2 *__restrict__ part) {
  pdl_grid_sync();
  __shared__ __align__(16) float red_m[kCeBY][8 * kCeRT];
  __shared__ __align__(16) float red_s[kCeBY][8 * kCeRT];
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint32_t r = (blockIdx.x * kCeRT + tx) * 8u;
  const bool live = r < rows;
  const uint32_t v0 = blockIdx.y * v_per_block, v1 = min(V, v0 + v_per_block);
  // running maximum kept twice: mx (the value) and nmx = -mx * log2(e), so that a term costs unpack + FFMA + ex2 + add
  // (the pass was issue-bound at 14 instructions per element: ncu r02, 80 % issue slots busy at 3.9 TB/s)
  constexpr float kLog2e = 1.4426950408889634f;
  float mx[8], nmx[8], s[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    mx[k] = -INFINITY;
    nmx[k] = 0.0f;
    s[k] = 0.0f;
  }
  constexpr int U = 8;
  if (live)
    for (uint32_t j0 = v0 + ty; j0 < v1; j0 += kCeBY * U) {
      uint4 v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = j0 + u * kCeBY;
        v[u] = (j < v1) ? __ldg(reinterpret_cast<const uint4 *>(x + (uint64_t)j * rows + r)) : make_uint4(0xff80ff80u, 0xff80ff80u, 0xff80ff80u, 0xff80ff80u); // -inf
      }
#pragma unroll
      for (int k = 0; k < 8; ++k) {
        float e[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
          const uint32_t w = (k < 2) ? v[u].x : (k < 4) ? v[u].y : (k < 6) ? v[u].z : v[u].w;
          e[u] = __uint_as_float((k & 1) ? (w & 0xffff0000u) : (w << 16));
        }
        float m8 = fmaxf(fmaxf(fmaxf(e[0], e[1]), fmaxf(e[2], e[3])), fmaxf(fmaxf(e[4], e[5]), fmaxf(e[6], e[7])));
        if (m8 > mx[k]) { // (first batch: mx = -inf, s = 0: exp2(-inf) * 0 = 0)
          s[k] *= exp2f_approx_ce((mx[k] - m8) * kLog2e);
          mx[k] = m8;
          nmx[k] = -m8 * kLog2e;
        }
        if (mx[k] > -INFINITY) {
          float a0 = 0.0f, a1 = 0.0f;
#pragma unroll
          for (int u = 0; u < U; u += 2) {
            a0 += exp2f_approx_ce(fmaf(e[u], kLog2e, nmx[k]));
            a1 += exp2f_approx_ce(fmaf(e[u + 1], kLog2e, nmx[k]));
          }
          s[k] += a0 + a1;
        }
      }
    }
#pragma unroll
  for (int k = 0; k < 8; ++k) {
    red_m[ty][8 * tx + k] = mx[k];
    red_s[ty][8 * tx + k] = s[k];
  }
  __syncthreads();
  const uint32_t t = ty * kCeRT + tx; // 256 threads, 256 rows: each merges one row
  const uint32_t rr = blockIdx.x * 8 * kCeRT + t;
  if (rr < rows) {
    float M = -INFINITY;
#pragma unroll
    for (int y = 0; y < kCeBY; ++y) M = fmaxf(M, red_m[y][t]);
    float S = 0.0f;
#pragma unroll
    for (int y = 0; y < kCeBY; ++y) {
      const float my = red_m[y][t];
      if (my > -INFINITY) S += red_s[y][t] * expf(my - M);
    }
    part[(uint64_t)blockIdx.y * rows + rr] = make_float2(M, S);
  }
}

// Forward from per-row log-sum-exp partials (from the LM head's GEMM epilogue, weedcu_gemm_bf16_ex row_stats 2, or from
// ce_fwd_partial_bf16): merge the (max, sum exp) pairs, and recompute the ONE logit the loss needs per row — the target
// column — as the same bf16 x bf16 -> fp32 dot product (+ bias) the tensor cores formed, from the GEMM's own operands.
// Block = 32 adjacent rows x 8 k-slices (a warp reads 32 adjacent rows of one k: coalesced for the MN-major activations).
// The 32 target columns of a K-major weight operand are scattered 2*K-byte runs: read per thread they are 32 different
// sectors per load instruction (251 us at 8192 x 768), so each warp first copies 4 of them into shared memory with 16-byte
// loads (row pitch K + 2 halves: the 32 rows of a k land in 32 different banks).
__global__ void __launch_bounds__(256)
ce_fwd_stats_finish(const float2 *__restrict__ stats, uint32_t tiles, uint32_t rows, uint32_t V, const __nv_bfloat16 *__restrict__ a, int a_major,
                    uint64_t lda, const __nv_bfloat16 *__restrict__ b, int b_major, uint64_t ldb, uint32_t K, const float *__restrict__ bias,
                    const int32_t *__restrict__ targets, float *__restrict__ lse, float *__restrict__ nll, int stage_b) {
  pdl_grid_sync();
  extern __shared__ __align__(16) uint8_t ce_finish_smem[];
  __nv_bfloat16 *bsm = reinterpret_cast<__nv_bfloat16 *>(ce_finish_smem);
  __shared__ float red[8][33];
  const uint32_t rx = threadIdx.x & 31u, kx = threadIdx.x >> 5;
  const uint32_t r = blockIdx.x * 32u + rx;
  const bool live = r < rows;
  const uint32_t t = live ? (uint32_t)targets[r] : 0u;
  const uint32_t pitch = K + 2u;
  if (stage_b) { // K-major b, K % 8 == 0, 16-byte aligned rows: warp kx copies the target rows of block rows 4 kx .. 4 kx + 3
    for (uint32_t i = 0; i < 4u; ++i) {
      const uint32_t lr = 4u * kx + i, gr = blockIdx.x * 32u + lr;
      const uint32_t tt = (gr < rows) ? (uint32_t)targets[gr] : V;
      if (tt >= V) continue; // warp-uniform
      const uint4 *src = reinterpret_cast<const uint4 *>(b + (uint64_t)tt * ldb);
      uint32_t *dst = reinterpret_cast<uint32_t *>(bsm + lr * pitch); // pitch is even: 4-byte aligned
      for (uint32_t c = rx; c < K / 8u; c += 32u) {
        const uint4 v = __ldg(src + c);
        dst[4u * c] = v.x, dst[4u * c + 1u] = v.y, dst[4u * c + 2u] = v.z, dst[4u * c + 3u] = v.w;
      }
    }
    __syncthreads();
  }
  float acc = 0.0f;
  if (live && t < V) {
    const __nv_bfloat16 *ap = a_major ? a + r : a + (uint64_t)r * lda;
    const uint64_t as = a_major ? lda : 1u;
    float a4[4] = {0.f, 0.f, 0.f, 0.f};
    uint32_t k = kx;
    if (stage_b) {
      const __nv_bfloat16 *bp = bsm + rx * pitch;
      for (; k + 24u < K; k += 32u) { // 4 independent products in flight
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) a4[u] += __bfloat162float(ap[(uint64_t)(k + 8u * u) * as]) * __bfloat162float(bp[k + 8u * u]);
      }
      for (; k < K; k += 8u) a4[0] += __bfloat162float(ap[(uint64_t)k * as]) * __bfloat162float(bp[k]);
    } else {
      const __nv_bfloat16 *bp = b_major ? b + t : b + (uint64_t)t * ldb;
      const uint64_t bs = b_major ? ldb : 1u;
      for (; k + 24u < K; k += 32u) {
#pragma unroll
        for (uint32_t u = 0; u < 4; ++u) a4[u] += __bfloat162float(ap[(uint64_t)(k + 8u * u) * as]) * __bfloat162float(bp[(uint64_t)(k + 8u * u) * bs]);
      }
      for (; k < K; k += 8u) a4[0] += __bfloat162float(ap[(uint64_t)k * as]) * __bfloat162float(bp[(uint64_t)k * bs]);
    }
    acc = (a4[0] + a4[1]) + (a4[2] + a4[3]);
  }
  red[kx][rx] = acc;
  __syncthreads();
  if (kx != 0 || !live) return;
  float xt = NAN;
  if (t < V) {
    float dot = 0.0f;
#pragma unroll
    for (int y = 0; y < 8; ++y) dot += red[y][rx];
    xt = dot + (bias ? bias[t] : 0.0f);
  }
  float M = -INFINITY;
  for (uint32_t i = 0; i < tiles; ++i) M = fmaxf(M, stats[(uint64_t)i * rows + r].x);
  float S = 0.0f;
  for (uint32_t i = 0; i < tiles; ++i) {
    const float2 p = stats[(uint64_t)i * rows + r];
    if (p.x > -INFINITY) S += p.y * expf(p.x - M);
  }
  const float log_s = logf(S);
  lse[r] = M + log_s;
  nll[r] = (xt - M) - log_s;
}
// dlogits[r,v] (+)= (exp(x - lse[r]) - onehot) * dloss/rows ; block = 32 rows x 8 vocab lanes.
__global__ void __launch_bounds__(kCeRT * kCeBY)
ce_bwd_kernel(const float *__restrict__ x, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs,
              uint32_t v_per_block, const int32_t *__restrict__ targets, const float *__restrict__ lse,
              const float *__restrict__ dloss, float *dlogits, int accumulate) {
  pdl_grid_sync();
  const uint32_t r = blockIdx.x * kCeRT + threadIdx.x;
  if (r >= rows) return;
  const uint32_t v0 = blockIdx.y * v_per_block, v1 = min(V, v0 + v_per_block);
  const float g = dloss[0] / (float)rows, l = lse[r];
  const uint32_t tgt = (uint32_t)targets[r];
  const uint64_t base = (uint64_t)r * rs;
  for (uint32_t j0 = v0 + threadIdx.y; j0 < v1; j0 += kCeBY * kCeU) {
    float v[kCeU], d[kCeU];
#pragma unroll
    for (int u = 0; u < kCeU; ++u) {
      const uint32_t j = j0 + u * kCeBY;
      v[u] = (j < v1) ? x[base + (uint64_t)j * vs] : 0.0f;
      d[u] = (accumulate && j < v1) ? dlogits[base + (uint64_t)j * vs] : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < kCeU; ++u) {
      const uint32_t j = j0 + u * kCeBY;
      if (j < v1) {
        const float pr = __expf(v[u] - l); // the same ex2-based exp as ce_bwd_pack_kernel: the two backward kernels stay bit-identical
        const float oh = (tgt == j) ? 1.0f : 0.0f;
        dlogits[base + (uint64_t)j * vs] = d[u] + (pr - oh) * g;
      }
    }
  }
}
// Backward fused with what the Linear that produced the logits needs next (its matmul_backward packs
// dY to bf16 for the two tensor-core GEMMs and sums its columns for the bias gradient): ONE pass
// writes dlogits (fp32), its bf16 GEMM operand copy and per-row-chunk column partial sums.
// block = 256 threads x 4 adjacent rows, kCePackCols vocab columns; logits / dlogits are [rows, V]
// with rows contiguous (rs == 1, vs == rows, rows % 8 == 0, 16-byte aligned).
constexpr int kCePackCols = 8;
// four adjacent rows of column j: fp32 logits, or their bf16 copy (the LM head's epilogue wrote only that)
__device__ __forceinline__ float4 ce_load4(const float *x, uint64_t off) { return *reinterpret_cast<const float4 *>(x + off); }
__device__ __forceinline__ float4 ce_load4(const __nv_bfloat16 *x, uint64_t off) {
  const uint2 u = *reinterpret_cast<const uint2 *>(x + off);
  const __nv_bfloat162 lo = *reinterpret_cast<const __nv_bfloat162 *>(&u.x), hi = *reinterpret_cast<const __nv_bfloat162 *>(&u.y);
  return make_float4(__low2float(lo), __high2float(lo), __low2float(hi), __high2float(hi));
}
template <typename XT>
__global__ void __launch_bounds__(256)
ce_bwd_pack_kernel(const XT *__restrict__ x, uint32_t rows, uint32_t V, const int32_t *__restrict__ targets,
                   const float *__restrict__ lse, const float *__restrict__ dloss, float *dlogits, int accumulate,
                   __nv_bfloat16 *__restrict__ shadow, float *__restrict__ part) {
  pdl_grid_sync();
  __shared__ float red[8][kCePackCols];
  const uint32_t r = (blockIdx.x * 256u + threadIdx.x) * 4u;
  const bool live = r < rows;
  const uint32_t j0 = blockIdx.y * kCePackCols;
  const float g = dloss[0] / (float)rows;
  float cs[kCePackCols];
#pragma unroll
  for (int i = 0; i < kCePackCols; ++i) cs[i] = 0.0f;
  if (live) {
    const float4 l4 = *reinterpret_cast<const float4 *>(lse + r);
    const int4 t4 = *reinterpret_cast<const int4 *>(targets + r);
    constexpr int U = 4;
#pragma unroll
    for (int i0 = 0; i0 < kCePackCols; i0 += U) {
      float4 xv[U], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = j0 + i0 + u;
        if (j < V) {
          const uint64_t off = (uint64_t)j * rows + r;
          xv[u] = ce_load4(x, off);
          dv[u] = accumulate ? *reinterpret_cast<const float4 *>(dlogits + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = j0 + i0 + u;
        if (j < V) {
          float4 o;
          // __expf (ex2.approx, 2 ulp): with expf's ~20 instructions per element the pass was issue-bound (452 us for
          // 1.65 GB at 8192 x 50257), and the result is scaled by 1/rows and rounded to bf16 right below
          o.x = dv[u].x + (__expf(xv[u].x - l4.x) - (((uint32_t)t4.x == j) ? 1.0f : 0.0f)) * g;
          o.y = dv[u].y + (__expf(xv[u].y - l4.y) - (((uint32_t)t4.y == j) ? 1.0f : 0.0f)) * g;
          o.z = dv[u].z + (__expf(xv[u].z - l4.z) - (((uint32_t)t4.z == j) ? 1.0f : 0.0f)) * g;
          o.w = dv[u].w + (__expf(xv[u].w - l4.w) - (((uint32_t)t4.w == j) ? 1.0f : 0.0f)) * g;
          const uint64_t off = (uint64_t)j * rows + r;
          if (dlogits) *reinterpret_cast<float4 *>(dlogits + off) = o; // NULL: operand copy + column sums only
          __nv_bfloat162 h[2];
          h[0] = __floats2bfloat162_rn(o.x, o.y);
          h[1] = __floats2bfloat162_rn(o.z, o.w);
          *reinterpret_cast<uint2 *>(shadow + off) = *reinterpret_cast<const uint2 *>(h);
          cs[i0 + u] += (o.x + o.y) + (o.z + o.w);
        }
      }
    }
  }
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
#pragma unroll
  for (int i = 0; i < kCePackCols; ++i) {
    const float t = warp_sum(cs[i]);
    if (lane == 0) red[w][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < kCePackCols && j0 + threadIdx.x < V) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    part[(uint64_t)blockIdx.x * V + j0 + threadIdx.x] = t;
  }
}
// ce_bwd_pack_kernel<bf16> for the LM-head case (no fp32 dlogits, nothing to accumulate into): 2 B read + 2 B written per
// element, so the pass is bound by how few instructions an element costs. A thread owns 8 adjacent rows (one 16-byte
// load / store per column), a block 1024 rows x kCeP16Cols columns in batches of 8 columns (8 loads in flight per
// thread). Per element: unpack, one FFMA into the ex2 argument (-lse * log2(e) is per row), ex2.approx, scale, pack, column
// sum. The one-hot term is not in the loop: a row whose target falls into the block's columns patches that one element
// afterwards (and the block's column sum through shared memory), bit-identical to computing (p - 1) * g in place.
constexpr int kCeP16Cols = 32;
__device__ __forceinline__ float exp2f_approx(float v) {
  float r;
  asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(r) : "f"(v));
  return r;
}
__global__ void __launch_bounds__(128)
ce_bwd_pack16_kernel(const __nv_bfloat16 *__restrict__ x, uint32_t rows, uint32_t V, const int32_t *__restrict__ targets,
                     const float *__restrict__ lse, const float *__restrict__ dloss, __nv_bfloat16 *__restrict__ shadow, float *__restrict__ part) {
  pdl_grid_sync();
  __shared__ float red[4][kCeP16Cols];
  __shared__ float fix[kCeP16Cols];
  const uint32_t r = (blockIdx.x * 128u + threadIdx.x) * 8u;
  const bool live = r < rows;
  const uint32_t j0 = blockIdx.y * kCeP16Cols;
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
  const float g = dloss[0] / (float)rows;
  constexpr float kLog2e = 1.4426950408889634f;
  if (threadIdx.x < kCeP16Cols) fix[threadIdx.x] = 0.0f;
  float nl[8];
#pragma unroll
  for (int k = 0; k < 8; ++k) nl[k] = 0.0f;
  if (live) {
    const float4 la = *reinterpret_cast<const float4 *>(lse + r), lb = *reinterpret_cast<const float4 *>(lse + r + 4);
    nl[0] = -la.x * kLog2e, nl[1] = -la.y * kLog2e, nl[2] = -la.z * kLog2e, nl[3] = -la.w * kLog2e;
    nl[4] = -lb.x * kLog2e, nl[5] = -lb.y * kLog2e, nl[6] = -lb.z * kLog2e, nl[7] = -lb.w * kLog2e;
  }
  __syncthreads(); // fix[] zeroed
#pragma unroll 1
  for (uint32_t c0 = 0; c0 < (uint32_t)kCeP16Cols; c0 += 8u) {
    uint4 xv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t j = j0 + c0 + u;
      xv[u] = (live && j < V) ? __ldg(reinterpret_cast<const uint4 *>(x + (uint64_t)j * rows + r)) : make_uint4(0u, 0u, 0u, 0u);
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t j = j0 + c0 + u;
      const uint32_t wd[4] = {xv[u].x, xv[u].y, xv[u].z, xv[u].w};
      uint32_t out[4];
      float cs = 0.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float e0 = __uint_as_float(wd[q] << 16), e1 = __uint_as_float(wd[q] & 0xffff0000u);
        const float o0 = exp2f_approx(fmaf(e0, kLog2e, nl[2 * q])) * g, o1 = exp2f_approx(fmaf(e1, kLog2e, nl[2 * q + 1])) * g;
        const __nv_bfloat162 h = __floats2bfloat162_rn(o0, o1);
        out[q] = *reinterpret_cast<const uint32_t *>(&h);
        cs += o0 + o1;
      }
      if (live && j < V) *reinterpret_cast<uint4 *>(shadow + (uint64_t)j * rows + r) = make_uint4(out[0], out[1], out[2], out[3]);
      else cs = 0.0f;
      cs = warp_sum(cs);
      if (lane == 0) red[w][c0 + u] = cs;
    }
  }
  if (live) { // the one-hot term of the rows whose target is one of this block's columns
    const int4 ta = *reinterpret_cast<const int4 *>(targets + r), tb = *reinterpret_cast<const int4 *>(targets + r + 4);
    const uint32_t tg[8] = {(uint32_t)ta.x, (uint32_t)ta.y, (uint32_t)ta.z, (uint32_t)ta.w, (uint32_t)tb.x, (uint32_t)tb.y, (uint32_t)tb.z, (uint32_t)tb.w};
#pragma unroll
    for (int k = 0; k < 8; ++k) {
      const uint32_t tj = tg[k] - j0;
      if (tj < (uint32_t)kCeP16Cols && tg[k] < V) {
        const uint64_t off = (uint64_t)tg[k] * rows + r + k;
        const float e = __bfloat162float(x[off]);
        const float p = exp2f_approx(fmaf(e, kLog2e, nl[k]));
        shadow[off] = __float2bfloat16_rn((p - 1.0f) * g);
        atomicAdd(&fix[tj], -g); // every addend is the same value: the sum does not depend on the order
      }
    }
  }
  __syncthreads();
  if (threadIdx.x < kCeP16Cols && j0 + threadIdx.x < V)
    part[(uint64_t)blockIdx.x * V + j0 + threadIdx.x] = ((red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x])) + fix[threadIdx.x];
}
// colsum[j] (+)= sum_c part[c][j] in a fixed order
__global__ void __launch_bounds__(256)
ce_colsum_finish_kernel(const float *__restrict__ part, uint32_t nchunks, uint32_t V, float *__restrict__ colsum) {
  pdl_grid_sync();
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= V) return;
  float t = 0.0f;
  for (uint32_t c = 0; c < nchunks; ++c) t += part[(uint64_t)c * V + j];
  colsum[j] = t;
}

// vocab slices so that (row tiles) x (slices) is ~8 blocks per SM
static void ce_split(uint32_t rows, uint32_t V, uint32_t &splits, uint32_t &v_per_block) {
  const uint32_t row_tiles = (rows + kCeRT - 1) / kCeRT;
  uint32_t want = (8u * kNumSMs + row_tiles - 1) / row_tiles;
  const uint32_t max_splits = (V + kCeBY * kCeU - 1) / (kCeBY * kCeU);
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  if (want > 65535u) want = 65535u;
  v_per_block = (V + want - 1) / want;
  v_per_block = (v_per_block + kCeBY * kCeU - 1) / (kCeBY * kCeU) * (kCeBY * kCeU);
  splits = (V + v_per_block - 1) / v_per_block;
}

// ----------------------------------------------------------------------------- host helpers
static bool canonical(const weedcu_view *v, int axis, uint64_t &inner, uint64_t &outer) {
  uint64_t st = 1;
  inner = outer = 1;
  for (int d = 0; d < v->rank; ++d) {
    const uint32_t ext = v->shape[d];
    if (ext != 1 && v->stride[d] != st) return false;
    if (d < axis) inner *= ext;
    if (d > axis) outer *= ext;
    st *= ext;
  }
  return true;
}
static void to_rowview(const weedcu_view *v, int axis, RowView &r) {
  r.rank = v->rank;
  r.axis = axis;
  r.offset = v->offset;
  for (int d = 0; d < kMaxRank; ++d) {
    r.shape[d] = d < v->rank ? v->shape[d] : 1;
    r.stride[d] = d < v->rank ? v->stride[d] : 0;
  }
}
static bool same_shape(const weedcu_view *a, const weedcu_view *b) {
  if (a->rank != b->rank) return false;
  for (int d = 0; d < a->rank; ++d)
    if (a->shape[d] != b->shape[d]) return false;
  return true;
}

template <bool LOG, int RT, int BY, bool STAGED, class LD>
static void launch_strided_fwd_cfg(const float *a, float *out, uint32_t inner, uint32_t L, uint32_t outer,
                                   LD ld, cudaStream_t st) {
  const dim3 grid((inner + RT - 1) / RT, outer);
  const size_t tile_bytes = STAGED ? (size_t)L * RT * sizeof(float) : 0;
  auto k = softmax_strided_fwd<LOG, RT, BY, STAGED, LD>;
  if (tile_bytes > 48 * 1024)
    ensure_dynamic_smem((const void *)k, (int)kMaxTileBytes);
  launch_k(k, dim3(grid), dim3(RT, BY), tile_bytes, st, a, out, inner, L, ld);
}

template <bool LOG, class LD>
static int launch_strided_fwd(const float *a, float *out, uint32_t inner, uint32_t L, uint32_t outer,
                              LD ld, cudaStream_t st) {
  // rows per tile: as wide as keeps the staged tile within ~32 KB (>= 6 CTAs per SM)
  if ((size_t)L * 32 * sizeof(float) <= 32 * 1024)
    launch_strided_fwd_cfg<LOG, 32, 8, true, LD>(a, out, inner, L, outer, ld, st);
  else if ((size_t)L * 16 * sizeof(float) <= 32 * 1024)
    launch_strided_fwd_cfg<LOG, 16, 16, true, LD>(a, out, inner, L, outer, ld, st);
  else if ((size_t)L * 8 * sizeof(float) <= kMaxTileBytes)
    launch_strided_fwd_cfg<LOG, 8, 32, true, LD>(a, out, inner, L, outer, ld, st);
  else
    launch_strided_fwd_cfg<LOG, 8, 64, false, LD>(a, out, inner, L, outer, ld, st);
  return after_launch();
}

template <bool LOG>
static int softmax_fwd_impl(const float *a, const weedcu_view *av, int axis, float *out,
                            const weedcu_view *ov, cudaStream_t st) {
  uint64_t inner, outer, i2, o2;
  const uint32_t L = av->shape[axis];
  uint64_t total = 1;
  for (int d = 0; d < av->rank; ++d) total *= av->shape[d];
  if (!L || !total || total > 0xffffffffull) return WEEDCU_EINVAL;
  const uint32_t n_rows = (uint32_t)(total / L);
  if (canonical(av, axis, inner, outer) && canonical(ov, axis, i2, o2)) {
    const float *pa = a + av->offset;
    float *po = out + ov->offset;
    if (inner == 1) {
      launch_k(softmax_contig_fwd<LOG>, dim3((unsigned)outer), dim3(256), 0, st, pa, po, L);
      return after_launch();
    }
    // big problems: two streaming passes (row statistics, then elementwise); small ones (a few tiles,
    // short rows) stay on the single staged-tile kernel, where one launch matters more than overlap
    const uint64_t n_rows64 = inner * outer;
    if (inner >= 32 && outer <= 65535 && (uint64_t)total >= (1u << 22) && L >= 64) {
      const uint32_t row_tiles = (uint32_t)((inner + 31) / 32);
      uint32_t splits = (uint32_t)((4ull * kNumSMs + (uint64_t)row_tiles * outer - 1) / ((uint64_t)row_tiles * outer));
      const uint32_t max_splits = (L + kSmBY * kSmU - 1) / (kSmBY * kSmU);
      if (splits > max_splits) splits = max_splits;
      if (splits < 1) splits = 1;
      if (splits > 65535u) splits = 65535u;
      uint32_t cps = (L + splits - 1) / splits;
      cps = (cps + kSmBY * kSmU - 1) / (kSmBY * kSmU) * (kSmBY * kSmU);
      splits = (L + cps - 1) / cps;
      const uint32_t cgroups = (L + kSmCols - 1) / kSmCols;
      if (cgroups <= 65535u && splits <= 65535u) {
        float *ws = nullptr; // part_m, part_s [splits][rows]; row_m, row_s [rows]
        const uint64_t rows_up = (n_rows64 + 3) & ~(uint64_t)3;
        WCU_CHECK(pool_alloc((void **)&ws, sizeof(float) * (2 * (size_t)splits * n_rows64 + 2 * rows_up), st));
        float *pm = ws, *ps = pm + (size_t)splits * n_rows64, *rm = ps + (size_t)splits * n_rows64, *rs = rm + rows_up;
        launch_k(softmax_rows_partial_kernel, dim3(row_tiles, (unsigned)outer, splits), dim3(32 * kSmBY), 0, st, pa, (uint32_t)inner, L, cps, pm, ps);
        int rc = after_launch();
        if (rc == 0) {
          launch_k(softmax_rows_finish_kernel<LOG>, dim3((unsigned)((n_rows64 + 255) / 256)), dim3(256), 0, st, pm, ps, splits, n_rows64, rm, rs);
          rc = after_launch();
        }
        if (rc == 0) {
          if ((inner % 4) == 0 && aligned16(pa) && aligned16(po))
            launch_k(softmax_apply_kernel<LOG, 4>, dim3((unsigned)((inner / 4 + 255) / 256), cgroups, (unsigned)outer), dim3(256), 0, st, pa, po, (uint32_t)inner, L, rm, rs);
          else
            launch_k(softmax_apply_kernel<LOG, 1>, dim3((unsigned)((inner + 255) / 256), cgroups, (unsigned)outer), dim3(256), 0, st, pa, po, (uint32_t)inner, L, rm, rs);
          rc = after_launch();
        }
        pool_free(ws, st);
        return rc;
      }
    }
    if (outer <= 65535) return launch_strided_fwd<LOG>(pa, po, (uint32_t)inner, L, (uint32_t)outer, PlainLoad(), st);
  }
  RowView ra, ro;
  to_rowview(av, axis, ra);
  to_rowview(ov, axis, ro);
  launch_k(softmax_generic_fwd<LOG>, dim3((n_rows + 127) / 128), dim3(128), 0, st, a, ra, out, ro, n_rows);
  return after_launch();
}

template <bool LOG>
static int softmax_bwd_impl(float *din, const weedcu_view *iv, const float *out, const weedcu_view *ov,
                            const float *dout, const weedcu_view *dv, int axis, cudaStream_t st) {
  uint64_t inner, outer, i2, o2, i3, o3;
  const uint32_t L = iv->shape[axis];
  uint64_t total = 1;
  for (int d = 0; d < iv->rank; ++d) total *= iv->shape[d];
  if (!L || !total || total > 0xffffffffull) return WEEDCU_EINVAL;
  const uint32_t n_rows = (uint32_t)(total / L);
  if (canonical(iv, axis, inner, outer) && canonical(ov, axis, i2, o2) && canonical(dv, axis, i3, o3)) {
    float *pd = din + iv->offset;
    const float *py = out + ov->offset, *pg = dout + dv->offset;
    if (inner == 1) {
      launch_k(softmax_contig_bwd<LOG>, dim3((unsigned)outer), dim3(256), 0, st, pd, py, pg, L);
      return after_launch();
    }
    if (outer <= 65535) {
      const dim3 grid((unsigned)((inner + kRT - 1) / kRT), (unsigned)outer);
      const size_t tile_bytes = 2 * (size_t)L * kRT * sizeof(float);
      if (tile_bytes <= kMaxTileBytes) {
        if (L >= 256) {
          auto k = softmax_strided_bwd<LOG, 32, true>;
          ensure_dynamic_smem((const void *)k, (int)kMaxTileBytes);
          launch_k(k, dim3(grid), dim3(kRT, 32), tile_bytes, st, pd, py, pg, (uint32_t)inner, L);
        } else {
          auto k = softmax_strided_bwd<LOG, 8, true>;
          ensure_dynamic_smem((const void *)k, (int)kMaxTileBytes);
          launch_k(k, dim3(grid), dim3(kRT, 8), tile_bytes, st, pd, py, pg, (uint32_t)inner, L);
        }
      } else {
        launch_k(softmax_strided_bwd<LOG, 32, false>, dim3(grid), dim3(kRT, 32), 0, st, pd, py, pg, (uint32_t)inner, L);
      }
      return after_launch();
    }
  }
  RowView ri, ro, rd;
  to_rowview(iv, axis, ri);
  to_rowview(ov, axis, ro);
  to_rowview(dv, axis, rd);
  launch_k(softmax_generic_bwd<LOG>, dim3((n_rows + 127) / 128), dim3(128), 0, st, din, ri, out, ro, dout, rd, n_rows);
  return after_launch();
}

} // namespace weedcu

using namespace weedcu;

extern "C" {

int weedcu_softmax_real(int log_mode, const float *a, const weedcu_view *av, int axis, float *out,
                        const weedcu_view *ov, void *stream) {
  if (!a || !av || !out || !ov || av->rank <= 0 || av->rank > kMaxRank || axis < 0 ||
      axis >= av->rank || !same_shape(av, ov))
    return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  double elems = 1.0;
  for (int d = 0; d < av->rank; ++d) elems *= av->shape[d];
  ProfScope prof(WEEDCU_PROF_SOFTMAX, st, 8.0 * elems);
  return log_mode ? softmax_fwd_impl<true>(a, av, axis, out, ov, st)
                  : softmax_fwd_impl<false>(a, av, axis, out, ov, st);
}

int weedcu_softmax_grad_real(int log_mode, float *din, const weedcu_view *dinv, const float *out,
                             const weedcu_view *ov, const float *dout, const weedcu_view *doutv,
                             int axis, void *stream) {
  if (!din || !dinv || !out || !ov || !dout || !doutv || dinv->rank <= 0 ||
      dinv->rank > kMaxRank || axis < 0 || axis >= dinv->rank || !same_shape(dinv, ov) ||
      !same_shape(dinv, doutv))
    return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  double elems = 1.0;
  for (int d = 0; d < dinv->rank; ++d) elems *= dinv->shape[d];
  ProfScope prof(WEEDCU_PROF_SOFTMAX, st, 16.0 * elems);
  return log_mode ? softmax_bwd_impl<true>(din, dinv, out, ov, dout, doutv, axis, st)
                  : softmax_bwd_impl<false>(din, dinv, out, ov, dout, doutv, axis, st);
}

int weedcu_attn_softmax_real(const float *scores, float *out, uint32_t batch, uint32_t Tq,
                             uint32_t Tk, float divisor, float mask_val, int causal, int batch_fastest,
                             void *stream) {
  if (!scores || !out || !batch || !Tq || !Tk) return WEEDCU_EINVAL;
  const int do_mask = (causal && Tq > 1) ? 1 : 0;
  ProfScope prof(WEEDCU_PROF_SOFTMAX, resolve_stream(stream), 8.0 * (double)batch * Tq * Tk);
  if (batch_fastest) {
    const uint64_t inner = (uint64_t)batch * Tq;
    if (inner > 0xffffffffull) return WEEDCU_EINVAL;
    AttnLoad ld = {batch, divisor, mask_val, do_mask};
    return launch_strided_fwd<false>(scores, out, (uint32_t)inner, Tk, 1, ld, resolve_stream(stream));
  }
  if (batch > 65535) return WEEDCU_EINVAL;
  AttnLoad ld = {1u, divisor, mask_val, do_mask};
  return launch_strided_fwd<false>(scores, out, Tq, Tk, batch, ld, resolve_stream(stream));
}

int weedcu_cross_entropy_fwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                             uint32_t rs, uint32_t vs, const int32_t *targets, float *lse,
                             float *loss, void *stream) {
  if (!logits || !targets || !lse || !loss || !rows || !V) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  uint32_t splits, vpb;
  const bool vec4 = rs == 1u && (rows % 4u) == 0 && rows >= 512u && (vs % 4u) == 0 && aligned16(logits + offset);
  if (vec4) { // 128-row tiles: four times fewer row tiles, so four times more vocabulary slices for the same block count
    const uint32_t row_tiles = (rows + 4 * kCeRT - 1) / (4 * kCeRT);
    uint32_t want = (8u * kNumSMs + row_tiles - 1) / row_tiles;
    const uint32_t max_splits = (V + kCeBY * 4 - 1) / (kCeBY * 4);
    if (want > max_splits) want = max_splits;
    if (want < 1) want = 1;
    vpb = (V + want - 1) / want;
    vpb = (vpb + kCeBY * 4 - 1) / (kCeBY * 4) * (kCeBY * 4);
    splits = (V + vpb - 1) / vpb;
  } else {
    ce_split(rows, V, splits, vpb);
  }
  float *ws = nullptr; // nll[rows], part_m[splits][rows], part_s[splits][rows]
  WCU_CHECK(pool_alloc((void **)&ws, sizeof(float) * (size_t)rows * (1 + 2 * (size_t)splits), st));
  float *nll = ws, *pm = ws + rows, *ps = pm + (size_t)splits * rows;
  ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, st, 4.0 * (double)rows * V);
  if (vec4)
    launch_k(ce_fwd_partial_vec4, dim3((rows + 4 * kCeRT - 1) / (4 * kCeRT), splits), dim3(kCeRT, kCeBY), 0, st, logits + offset, rows, V, vs, vpb, pm, ps);
  else
    launch_k(ce_fwd_partial, dim3((rows + kCeRT - 1) / kCeRT, splits), dim3(kCeRT, kCeBY), 0, st, 
      logits + offset, rows, V, rs, vs, vpb, pm, ps);
  int rc = after_launch();
  if (rc == 0) {
    launch_k(ce_fwd_finish, dim3((rows + 255) / 256), dim3(256), 0, st, logits + offset, rows, V, rs, vs, targets, pm, ps, splits, lse, nll);
    rc = after_launch();
  }
  if (rc == 0) {
    weedcu_view v;
    v.offset = 0;
    v.rank = 1;
    v.shape[0] = rows;
    v.stride[0] = 1;
    rc = weedcu_sum_real(nll, &v, -1.0f / (float)rows, loss, (void *)st);
  }
  pool_free(ws, st);
  return rc;
}

int weedcu_cross_entropy_bwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                             uint32_t rs, uint32_t vs, const int32_t *targets, const float *lse,
                             const float *dloss, float *dlogits, uint64_t d_offset, int accumulate,
                             void *stream) {
  if (!logits || !targets || !lse || !dloss || !dlogits || !rows || !V) return WEEDCU_EINVAL;
  const uint64_t n = (uint64_t)rows * V;
  uint32_t splits, vpb;
  ce_split(rows, V, splits, vpb);
  ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, resolve_stream(stream), (accumulate ? 12.0 : 8.0) * (double)n);
  launch_k(ce_bwd_kernel, dim3((rows + kCeRT - 1) / kCeRT, splits), dim3(kCeRT, kCeBY), 0, resolve_stream(stream), 
      logits + offset, rows, V, rs, vs, vpb, targets, lse, dloss, dlogits + d_offset, accumulate);
  return after_launch();
}

int weedcu_cross_entropy_bwd_pack(const float *logits, uint64_t offset, uint32_t rows, uint32_t V, const int32_t *targets,
                                  const float *lse, const float *dloss, float *dlogits, uint64_t d_offset, int accumulate,
                                  uint16_t *dlogits_bf16, float *colsum, void *stream) {
  if (!logits || !targets || !lse || !dloss || !dlogits_bf16 || !colsum || !rows || !V) return WEEDCU_EINVAL;
  if (!dlogits && accumulate) return WEEDCU_EINVAL; // nothing to accumulate into
  const float *x = logits + offset;
  float *d = dlogits ? dlogits + d_offset : nullptr;
  if ((rows % 8u) || !aligned16(x) || (d && !aligned16(d)) || !aligned16(lse) || !aligned16(targets) || !aligned16(dlogits_bf16)) return WEEDCU_ENOSUP;
  const uint32_t nchunks = (rows + 1023u) / 1024u, cgroups = (V + kCePackCols - 1) / kCePackCols;
  if (cgroups > 65535u) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  float *part = nullptr;
  WCU_CHECK(pool_alloc((void **)&part, sizeof(float) * (size_t)nchunks * V, st));
  ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, st, (accumulate ? 14.0 : (d ? 10.0 : 6.0)) * (double)rows * V);
  launch_k(ce_bwd_pack_kernel<float>, dim3(nchunks, cgroups), dim3(256), 0, st, x, rows, V, targets, lse, dloss, d, accumulate, (__nv_bfloat16 *)dlogits_bf16, part);
  int rc = after_launch();
  if (rc == 0) {
    launch_k(ce_colsum_finish_kernel, dim3((V + 255u) / 256u), dim3(256), 0, st, part, nchunks, V, colsum);
    rc = after_launch();
  }
  pool_free(part, st);
  return rc;
}

int weedcu_cross_entropy_fwd_stats(const float *stats, uint32_t tiles, uint32_t rows, uint32_t V, const uint16_t *a, int a_major, uint64_t lda,
                                   const uint16_t *b, int b_major, uint64_t ldb, uint32_t K, const float *col_bias, const int32_t *targets,
                                   float *lse, float *loss, void *stream) {
  if (!stats || !tiles || !rows || !V || !a || !b || !K || !targets || !lse || !loss) return WEEDCU_EINVAL;
  if (!aligned16(stats)) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  float *nll = nullptr;
  WCU_CHECK(pool_alloc((void **)&nll, sizeof(float) * (size_t)rows, st));
  ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, st, 8.0 * (double)rows * tiles + 2.0 * (double)rows * K * 2.0);
  // K-major weights: stage the 32 target columns of a block in shared memory (dynamic: 32 x (K + 2) halves)
  const size_t stage_bytes = 32u * ((size_t)K + 2u) * 2u;
  const int stage_b = !b_major && (K % 8u) == 0 && (ldb % 8u) == 0 && aligned16(b) && stage_bytes <= 200u * 1024u;
  if (stage_b && stage_bytes > 48u * 1024u) {
    static size_t granted = 0; // per process: the attribute only ever grows
    if (stage_bytes > granted) {
      WCU_CHECK(cudaFuncSetAttribute(ce_fwd_stats_finish, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)stage_bytes));
      granted = stage_bytes;
    }
  }
  launch_k(ce_fwd_stats_finish, dim3((rows + 31u) / 32u), dim3(256), stage_b ? stage_bytes : 0, st, (const float2 *)stats, tiles, rows, V,
           (const __nv_bfloat16 *)a, a_major, lda, (const __nv_bfloat16 *)b, b_major, ldb, K, col_bias, targets, lse, nll, stage_b);
  int rc = after_launch();
  if (rc == 0) {
    weedcu_view v;
    v.offset = 0;
    v.rank = 1;
    v.shape[0] = rows;
    v.stride[0] = 1;
    rc = weedcu_sum_real(nll, &v, -1.0f / (float)rows, loss, (void *)st);
  }
  pool_free(nll, st);
  return rc;
}

int weedcu_cross_entropy_bwd_pack_bf16in(const uint16_t *logits_bf16, uint32_t rows, uint32_t V, const int32_t *targets, const float *lse,
                                         const float *dloss, float *dlogits, uint64_t d_offset, int accumulate, uint16_t *dlogits_bf16,
                                         float *colsum, void *stream) {
  if (!logits_bf16 || !targets || !lse || !dloss || !dlogits_bf16 || !colsum || !rows || !V) return WEEDCU_EINVAL;
  if (!dlogits && accumulate) return WEEDCU_EINVAL;
  float *d = dlogits ? dlogits + d_offset : nullptr;
  if ((rows % 8u) || (((uintptr_t)logits_bf16) & 7u) || (d && !aligned16(d)) || !aligned16(lse) || !aligned16(targets) || !aligned16(dlogits_bf16)) return WEEDCU_ENOSUP;
  const uint32_t nchunks = (rows + 1023u) / 1024u, cgroups = (V + kCePackCols - 1) / kCePackCols;
  if (cgroups > 65535u) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  float *part = nullptr;
  WCU_CHECK(pool_alloc((void **)&part, sizeof(float) * (size_t)nchunks * V, st));
  ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, st, (accumulate ? 12.0 : (d ? 8.0 : 4.0)) * (double)rows * V);
  if (!d && aligned16(logits_bf16) && (V + kCeP16Cols - 1) / kCeP16Cols <= 65535u)
    launch_k(ce_bwd_pack16_kernel, dim3(nchunks, (V + kCeP16Cols - 1) / kCeP16Cols), dim3(128), 0, st, (const __nv_bfloat16 *)logits_bf16, rows, V, targets, lse,
             dloss, (__nv_bfloat16 *)dlogits_bf16, part);
  else
    launch_k(ce_bwd_pack_kernel<__nv_bfloat16>, dim3(nchunks, cgroups), dim3(256), 0, st, (const __nv_bfloat16 *)logits_bf16, rows, V, targets, lse, dloss, d,
             accumulate, (__nv_bfloat16 *)dlogits_bf16, part);
  int rc = after_launch();
  if (rc == 0) {
    launch_k(ce_colsum_finish_kernel, dim3((V + 255u) / 256u), dim3(256), 0, st, part, nchunks, V, colsum);
    rc = after_launch();
  }
  pool_free(part, st);
  return rc;
}

int weedcu_cross_entropy_fwd_bf16in(const uint16_t *logits_bf16, uint32_t rows, uint32_t V, const uint16_t *a, int a_major, uint64_t lda,
                                    const uint16_t *b, int b_major, uint64_t ldb, uint32_t K, const float *col_bias, const int32_t *targets,
                                    float *lse, float *loss, void *stream) {
  if (!logits_bf16 || !rows || !V || !a || !b || !K || !targets || !lse || !loss) return WEEDCU_EINVAL;
  if ((rows % 8u) || !aligned16(logits_bf16)) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  const uint32_t row_tiles = (rows + 8 * kCeRT - 1) / (8 * kCeRT);
  uint32_t want = (8u * kNumSMs + row_tiles - 1) / row_tiles;
  const uint32_t max_splits = (V + kCeBY * 4 - 1) / (kCeBY * 4);
  if (want > max_splits) want = max_splits;
  if (want < 1) want = 1;
  uint32_t vpb = (V + want - 1) / want;
  vpb = (vpb + kCeBY * 4 - 1) / (kCeBY * 4) * (kCeBY * 4);
  const uint32_t splits = (V + vpb - 1) / vpb;
  float2 *part = nullptr;
  WCU_CHECK(pool_alloc((void **)&part, sizeof(float2) * (size_t)rows * splits, st));
  int rc;
  {
    ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, st, 2.0 * (double)rows * V);
    launch_k(ce_fwd_partial_bf16, dim3(row_tiles, splits), dim3(kCeRT, kCeBY), 0, st, (const __nv_bfloat16 *)logits_bf16, rows, V, vpb, part);
    rc = after_launch();
  }
  if (rc == 0)
    rc = weedcu_cross_entropy_fwd_stats((const float *)part, splits, rows, V, a, a_major, lda, b, b_major, ldb, K, col_bias, targets, lse, loss, stream);
  pool_free(part, st);
  return rc;
}

} // extern "C"
