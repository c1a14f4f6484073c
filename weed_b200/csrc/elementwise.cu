// elementwise.cu — fills, N-D broadcast binary/in-place/copy, unary + unary-grad, SGD/Adam.
// HBM-bound: 128-bit vectorised, coalesced along the column-major fastest dim, grid sized in
// multiples of the SM count. No shared memory (no reuse to stage).
//
// Reference semantics: every operand resolves the flat column-major index i through its own
// (shape, stride) view — BaseTensor::get_storage_index (include/tensors/base_tensor.hpp:123-142)
// as used by the CPU lambdas in src/ops/commuting.cpp:27-35, in_place.cpp:27-35,
// copy_broadcast.cpp:27-30, real_unary.cpp:47-83, abs.cpp:70-94, pow.cpp:52-77.
#include "common.cuh"
#include <cuda_bf16.h>

#include <cstring>
#include <mutex>
#include <vector>

namespace weedcu {

template <int NIN> struct EwPtrs {
  const float *in[NIN];
  float *out;
};

// One thread = one element. Handles any rank <= 8 / any strides.
template <int NIN, class F>
__global__ void __launch_bounds__(256) ew_scalar_kernel(IndexSpace<NIN + 1> sp, EwPtrs<NIN> p, F f) {
  pdl_grid_sync();
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < sp.n; i += stride) {
    uint64_t off[NIN + 1];
#pragma unroll
    for (int o = 0; o <= NIN; ++o) off[o] = 0;
    uint32_t rem = i;
    for (int d = 0; d < sp.rank; ++d) {
      const uint32_t ext = sp.shape[d];
      const uint32_t c = rem % ext;
      rem /= ext;
#pragma unroll
      for (int o = 0; o <= NIN; ++o) off[o] += (uint64_t)c * sp.stride[o][d];
    }
    float x[NIN];
#pragma unroll
    for (int o = 0; o < NIN; ++o) x[o] = p.in[o][off[o]];
    p.out[off[NIN]] = f(x);
  }
}

// One thread = four consecutive elements of dim 0. Requires shape[0] % 4 == 0, every operand's
// dim-0 stride in {0,1}, 16-byte aligned bases and (for stride-1 operands) higher strides % 4 == 0.
template <int NIN, class F>
__global__ void __launch_bounds__(256) ew_vec4_kernel(IndexSpace<NIN + 1> sp, EwPtrs<NIN> p, F f) {
  pdl_grid_sync();
  const uint32_t nq = sp.n >> 2, s0q = sp.shape[0] >> 2;
  const uint32_t stride = gridDim.x * blockDim.x;
  for (uint32_t q = blockIdx.x * blockDim.x + threadIdx.x; q < nq; q += stride) {
    uint64_t off[NIN + 1];
    uint32_t rem = q / s0q;
    const uint32_t c0 = (q - rem * s0q) << 2;
#pragma unroll
    for (int o = 0; o <= NIN; ++o) off[o] = (uint64_t)c0 * sp.stride[o][0];
    for (int d = 1; d < sp.rank; ++d) {
      const uint32_t ext = sp.shape[d];
      const uint32_t c = rem % ext;
      rem /= ext;
#pragma unroll
      for (int o = 0; o <= NIN; ++o) off[o] += (uint64_t)c * sp.stride[o][d];
    }
    float4 v[NIN];
#pragma unroll
    for (int o = 0; o < NIN; ++o) {
      if (sp.stride[o][0]) {
        v[o] = *reinterpret_cast<const float4 *>(p.in[o] + off[o]);
      } else {
        const float s = p.in[o][off[o]];
        v[o] = make_float4(s, s, s, s);
      }
    }
    float4 r;
    {
      float x[NIN];
#pragma unroll
      for (int o = 0; o < NIN; ++o) x[o] = v[o].x;
      r.x = f(x);
#pragma unroll
      for (int o = 0; o < NIN; ++o) x[o] = v[o].y;
      r.y = f(x);
#pragma unroll
      for (int o = 0; o < NIN; ++o) x[o] = v[o].z;
      r.z = f(x);
#pragma unroll
      for (int o = 0; o < NIN; ++o) x[o] = v[o].w;
      r.w = f(x);
    }
    *reinterpret_cast<float4 *>(p.out + off[NIN]) = r;
  }
}

template <int NIN, class F>
static int launch_ew(const weedcu_view *const *views, const float *const *ins, float *out, F f,
                     cudaStream_t st) {
  IndexSpace<NIN + 1> sp;
  if (!build_index_space<NIN + 1>(views, sp)) return WEEDCU_EINVAL;
  EwPtrs<NIN> p;
  for (int o = 0; o < NIN; ++o) {
    if (!ins[o]) return WEEDCU_EINVAL;
    p.in[o] = ins[o] + views[o]->offset;
  }
  if (!out) return WEEDCU_EINVAL;
  p.out = out + views[NIN]->offset;
  if (sp.stride[NIN][0] == 0 && sp.n > 1) return WEEDCU_EINVAL; // broadcast output is a race

  bool vec = (sp.shape[0] % 4u) == 0;
  for (int o = 0; o <= NIN && vec; ++o) {
    const float *base = (o < NIN) ? p.in[o] : p.out;
    const uint32_t s0 = sp.stride[o][0];
    if (s0 > 1) vec = false;
    if (s0 == 1) {
      if (!aligned16(base)) vec = false;
      for (int d = 1; d < sp.rank; ++d)
        if (sp.stride[o][d] % 4u) vec = false;
    }
  }
  if (sp.stride[NIN][0] != 1) vec = false;
  double moved = 4.0 * sp.n; // algorithmic bytes: the write + every non-broadcast read
  for (int o = 0; o < NIN; ++o) {
    bool streams = false;
    for (int d = 0; d < sp.rank; ++d)
      if (sp.stride[o][d]) streams = true;
    if (streams) moved += 4.0 * sp.n;
  }
  ProfScope prof(WEEDCU_PROF_ELEMENTWISE, st, moved);
  if (vec) {
    const unsigned grid = grid_for(sp.n >> 2, 256, 16);
    launch_k(ew_vec4_kernel<NIN, F>, dim3(grid), dim3(256), 0, st, sp, p, f);
  } else {
    const unsigned grid = grid_for(sp.n, 256, 32);
    launch_k(ew_scalar_kernel<NIN, F>, dim3(grid), dim3(256), 0, st, sp, p, f);
  }
  return after_launch();
}

// ------------------------------------------------------------------------------- functors
struct AddF { __device__ float operator()(const float *x) const { return x[0] + x[1]; } };
struct MulF { __device__ float operator()(const float *x) const { return x[0] * x[1]; } };
struct SubF { __device__ float operator()(const float *x) const { return x[0] - x[1]; } };
struct DivF { __device__ float operator()(const float *x) const { return x[0] / x[1]; } };
struct CopyF { __device__ float operator()(const float *x) const { return x[0]; } };

__device__ __forceinline__ float gelu_fwd(float x) {
  // Tensor::gelu, reference src/tensors/tensor.cpp:841-851 (tanh approximation)
  const float k1 = 0.044715f, k2 = 0.7978845608028654f;
  const float x3 = (x * x) * x;
  const float t = tanhf(k2 * (x + k1 * x3));
  return (0.5f * x) * (1.0f + t);
}
__device__ __forceinline__ float gelu_dfdx(float x) {
  const float k1 = 0.044715f, k2 = 0.7978845608028654f;
  const float t = tanhf(k2 * (x + k1 * ((x * x) * x)));
  const float dinner = k2 * (1.0f + 3.0f * k1 * (x * x));
  return 0.5f * (1.0f + t) + (0.5f * x) * ((1.0f - t * t) * dinner);
}
// gelu_dfdx for a result that is rounded to bf16 right after: MUFU tanh (2^-11 relative) instead of the ~30-instruction
// tanhf (gelu_grad_pack with no fp32 output was bound by instruction issue, not by its 6-8 B/elem: 52 us at 8192 x 3072)
__device__ __forceinline__ float gelu_dfdx_for_bf16(float x) {
  const float k1 = 0.044715f, k2 = 0.7978845608028654f;
  float t;
  asm("tanh.approx.f32 %0, %1;" : "=f"(t) : "f"(k2 * (x + k1 * ((x * x) * x))));
  const float dinner = k2 * (1.0f + 3.0f * k1 * (x * x));
  return 0.5f * (1.0f + t) + (0.5f * x) * ((1.0f - t * t) * dinner);
}

template <int OP> struct UnaryF {
  float param;
  __device__ float operator()(const float *x) const {
    const float v = x[0];
    if (OP == WEEDCU_RELU) return fmaxf(v, 0.0f);
    if (OP == WEEDCU_SIGMOID) return 1.0f / (1.0f + expf(-v));
    if (OP == WEEDCU_TANH) return tanhf(v);
    if (OP == WEEDCU_ABS) return (v < 0.0f) ? -v : v;
    if (OP == WEEDCU_POW) return powf(v, param);
    if (OP == WEEDCU_EXP) return expf(v * param);
    if (OP == WEEDCU_LOG) return logf(v) * param;
    if (OP == WEEDCU_GELU) return gelu_fwd(v);
    if (OP == WEEDCU_SIN) return sinf(v);
    if (OP == WEEDCU_COS) return cosf(v);
    return v;
  }
};

// clamp (reference src/ops/clamp.cpp:67-73): min(max(x, lo), hi); its gradient passes dout where lo < x < hi (:55-60)
struct ClampF {
  float lo, hi;
  __device__ float operator()(const float *x) const { return fminf(fmaxf(x[0], lo), hi); }
};
struct ClampGradF { // x[0] = din (old), x[1] = in, x[2] = dout
  float lo, hi;
  __device__ float operator()(const float *x) const { return (x[1] > lo && x[1] < hi) ? x[0] + x[2] : x[0]; }
};
// full max / min backward (reference src/ops/real_extremum.cpp:50-57): dout goes to every element equal to the
// extremum; x[0] = din (old), x[1] = in, x[2] = dout, x[3] = the extremum (a scalar tensor, all strides 0)
struct MatchGradF {
  __device__ float operator()(const float *x) const { return (x[1] == x[3]) ? x[0] + x[2] : x[0]; }
};

// x[0] = din (old), x[1] = in (forward input or output), x[2] = dout
template <int OP> struct UnaryGradF {
  __device__ float operator()(const float *x) const {
    const float d = x[0], v = x[1], g = x[2];
    if (OP == WEEDCU_RELU) return (v > 0.0f) ? d + g : d;
    if (OP == WEEDCU_SIGMOID) return d + v * (1.0f - v) * g;
    if (OP == WEEDCU_TANH) return d + g * (1.0f - v * v);
    if (OP == WEEDCU_ABS) return (v != 0.0f) ? d + ((v > 0.0f) ? g : -g) : d;
    if (OP == WEEDCU_GELU) return d + g * gelu_dfdx(v);
    if (OP == WEEDCU_SIN) return d + cosf(v) * g;
    if (OP == WEEDCU_COS) return d + (-sinf(v)) * g;
    return d;
  }
};

// store variant (din known to be zero and never read): x[0] = in, x[1] = dout
template <int OP> struct UnaryGradSetF {
  __device__ float operator()(const float *x) const {
    const float y[3] = {0.0f, x[0], x[1]};
    return UnaryGradF<OP>()(y);
  }
};

// ------------------------------------------------------------------------------- fills
template <typename T> struct alignas(16) Quad { T x, y, z, w; };

template <typename T>
__global__ void __launch_bounds__(256) fill_kernel(T *p, uint64_t n, T v) {
  pdl_grid_sync();
  // body in 128-bit stores; head/tail elements handled by the first warp of block 0
  const uintptr_t addr = (uintptr_t)p;
  uint64_t head = ((16 - (addr & 15)) & 15) / sizeof(T);
  if (head > n) head = n;
  const uint64_t nq = (n - head) >> 2;
  Quad<T> *q = reinterpret_cast<Quad<T> *>(p + head);
  const Quad<T> val = {v, v, v, v};
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < nq; i += stride) q[i] = val;
  if (blockIdx.x == 0 && threadIdx.x < 8) {
    if (threadIdx.x < head) p[threadIdx.x] = v;
    const uint64_t tail0 = head + (nq << 2);
    if (threadIdx.x >= 4 && tail0 + (threadIdx.x - 4) < n) p[tail0 + (threadIdx.x - 4)] = v;
  }
}

// GELU forward that also leaves the bf16 GEMM operand copy of its output (the next Linear's A operand):
// 8 + 2 B/elem instead of 8 + a 6 B/elem pack pass. Dense inputs, 4 elements per thread.
__global__ void __launch_bounds__(256)
gelu_fwd_bf16_kernel(const float *__restrict__ x, float *__restrict__ y, __nv_bfloat16 *__restrict__ yb, uint64_t n4) {
  pdl_grid_sync();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += stride) {
    const float4 v = reinterpret_cast<const float4 *>(x)[i];
    float4 o;
    o.x = gelu_fwd(v.x);
    o.y = gelu_fwd(v.y);
    o.z = gelu_fwd(v.z);
    o.w = gelu_fwd(v.w);
    if (y) reinterpret_cast<float4 *>(y)[i] = o; // NULL: bf16 operand copy only
    __nv_bfloat162 h[2];
    h[0] = __floats2bfloat162_rn(o.x, o.y);
    h[1] = __floats2bfloat162_rn(o.z, o.w);
    reinterpret_cast<uint2 *>(yb)[i] = *reinterpret_cast<const uint2 *>(h);
  }
}

// GELU backward fused with what the Linear behind it needs next: din (+)= dout * gelu'(in) for a dense
// [rows, cols] matrix (rows contiguous), plus the bf16 GEMM operand copy of din and its column sums
// (that Linear's bias gradient). Same block shape as the LayerNorm / cross-entropy apply passes:
// 256 threads x 4 adjacent rows x 8 columns, column partials per row chunk.
constexpr int kGgCols = 8;
// four adjacent rows of one column of dout: fp32, or the bf16 copy the product that formed it wrote instead (the ff2
// Linear's dA when this node is its only reader: 2 B/elem less written by the GEMM and read here)
__device__ __forceinline__ float4 gg_load4(const float *x, uint64_t off) { return *reinterpret_cast<const float4 *>(x + off); }
__device__ __forceinline__ float4 gg_load4(const __nv_bfloat16 *x, uint64_t off) {
  const uint2 u = *reinterpret_cast<const uint2 *>(x + off);
  return make_float4(__uint_as_float(u.x << 16), __uint_as_float(u.x & 0xffff0000u), __uint_as_float(u.y << 16), __uint_as_float(u.y & 0xffff0000u));
}
template <typename DT, bool FAST>
__global__ void __launch_bounds__(256)
gelu_grad_pack_kernel(float *din, const float *__restrict__ in, const DT *__restrict__ dout, uint32_t rows, uint32_t cols,
                      int accumulate, __nv_bfloat16 *__restrict__ shadow, float *__restrict__ part) {
  pdl_grid_sync();
  __shared__ float red[8][kGgCols];
  const uint32_t r = (blockIdx.x * 256u + threadIdx.x) * 4u;
  const bool live = r < rows;
  const uint32_t j0 = blockIdx.y * kGgCols;
  float cs[kGgCols];
#pragma unroll
  for (int i = 0; i < kGgCols; ++i) cs[i] = 0.0f;
  if (live) {
    constexpr int U = 4;
#pragma unroll
    for (int i0 = 0; i0 < kGgCols; i0 += U) {
      float4 xv[U], gv[U], dv[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = j0 + i0 + u;
        if (j < cols) {
          const uint64_t off = (uint64_t)j * rows + r;
          xv[u] = *reinterpret_cast<const float4 *>(in + off);
          gv[u] = gg_load4(dout, off);
          dv[u] = accumulate ? *reinterpret_cast<const float4 *>(din + off) : make_float4(0.f, 0.f, 0.f, 0.f);
        }
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const uint32_t j = j0 + i0 + u;
        if (j < cols) {
          float4 o;
          if (FAST) { // only the bf16 copy and the column sums leave the kernel
            o.x = dv[u].x + gv[u].x * gelu_dfdx_for_bf16(xv[u].x);
            o.y = dv[u].y + gv[u].y * gelu_dfdx_for_bf16(xv[u].y);
            o.z = dv[u].z + gv[u].z * gelu_dfdx_for_bf16(xv[u].z);
            o.w = dv[u].w + gv[u].w * gelu_dfdx_for_bf16(xv[u].w);
          } else {
            o.x = dv[u].x + gv[u].x * gelu_dfdx(xv[u].x);
            o.y = dv[u].y + gv[u].y * gelu_dfdx(xv[u].y);
            o.z = dv[u].z + gv[u].z * gelu_dfdx(xv[u].z);
            o.w = dv[u].w + gv[u].w * gelu_dfdx(xv[u].w);
          }
          const uint64_t off = (uint64_t)j * rows + r;
          if (din) *reinterpret_cast<float4 *>(din + off) = o; // NULL: operand copy + column sums only
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
  for (int i = 0; i < kGgCols; ++i) {
    const float t = warp_sum(cs[i]);
    if (lane == 0) red[w][i] = t;
  }
  __syncthreads();
  if (threadIdx.x < kGgCols && j0 + threadIdx.x < cols) {
    float t = 0.0f;
#pragma unroll
    for (int k = 0; k < 8; ++k) t += red[k][threadIdx.x];
    part[(uint64_t)blockIdx.x * cols + j0 + threadIdx.x] = t;
  }
}
// gelu_grad_pack for the production case (bf16 dout, no fp32 output, nothing to accumulate into): a thread owns 8 adjacent
// rows — one 16-byte load of dout, two of the pre-activation and one 16-byte store per column — a block 1024 rows x 16
// columns in two batches of 8 columns. 8 B/elem.
constexpr int kGg16Cols = 16;
__global__ void __launch_bounds__(128)
gelu_grad_pack16_kernel(const float *__restrict__ in, const __nv_bfloat16 *__restrict__ dout, uint32_t rows, uint32_t cols,
                        __nv_bfloat16 *__restrict__ shadow, float *__restrict__ part) {
  pdl_grid_sync();
  __shared__ float red[4][kGg16Cols];
  const uint32_t r = (blockIdx.x * 128u + threadIdx.x) * 8u;
  const bool live = r < rows;
  const uint32_t j0 = blockIdx.y * kGg16Cols;
  const uint32_t lane = threadIdx.x & 31u, w = threadIdx.x >> 5;
#pragma unroll 1
  for (uint32_t c0 = 0; c0 < (uint32_t)kGg16Cols; c0 += 8u) {
    float4 xa[8], xb[8];
    uint4 gv[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t j = j0 + c0 + u;
      if (live && j < cols) {
        const uint64_t off = (uint64_t)j * rows + r;
        xa[u] = __ldg(reinterpret_cast<const float4 *>(in + off));
        xb[u] = __ldg(reinterpret_cast<const float4 *>(in + off + 4));
        gv[u] = __ldg(reinterpret_cast<const uint4 *>(dout + off));
      } else {
        xa[u] = xb[u] = make_float4(0.f, 0.f, 0.f, 0.f);
        gv[u] = make_uint4(0u, 0u, 0u, 0u);
      }
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t j = j0 + c0 + u;
      const float xs[8] = {xa[u].x, xa[u].y, xa[u].z, xa[u].w, xb[u].x, xb[u].y, xb[u].z, xb[u].w};
      const uint32_t gw[4] = {gv[u].x, gv[u].y, gv[u].z, gv[u].w};
      uint32_t out[4];
      float cs = 0.0f;
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const float o0 = __uint_as_float(gw[q] << 16) * gelu_dfdx_for_bf16(xs[2 * q]);
        const float o1 = __uint_as_float(gw[q] & 0xffff0000u) * gelu_dfdx_for_bf16(xs[2 * q + 1]);
        const __nv_bfloat162 h = __floats2bfloat162_rn(o0, o1);
        out[q] = *reinterpret_cast<const uint32_t *>(&h);
        cs += o0 + o1;
      }
      if (live && j < cols) *reinterpret_cast<uint4 *>(shadow + (uint64_t)j * rows + r) = make_uint4(out[0], out[1], out[2], out[3]);
      cs = warp_sum(cs);
      if (lane == 0) red[w][c0 + u] = cs;
    }
  }
  __syncthreads();
  if (threadIdx.x < kGg16Cols && j0 + threadIdx.x < cols)
    part[(uint64_t)blockIdx.x * cols + j0 + threadIdx.x] = (red[0][threadIdx.x] + red[1][threadIdx.x]) + (red[2][threadIdx.x] + red[3][threadIdx.x]);
}
// up to 64 (source, destination, length) triples per launch: the small gradients of a data-parallel bucket are gathered into
// one staging buffer, all-reduced as one message and scattered back (autograd.cpp: GradientBuckets::flush)
struct MultiCopyArgs {
  const float *src[64];
  float *dst[64];
  uint64_t n[64];
  uint32_t count;
};
__global__ void __launch_bounds__(256)
multi_copy_kernel(const __grid_constant__ MultiCopyArgs a) {
  pdl_grid_sync();
  const uint32_t t = blockIdx.x;
  const float *__restrict__ s = a.src[t];
  float *__restrict__ d = a.dst[t];
  const uint64_t n = a.n[t];
  for (uint64_t i = (uint64_t)blockIdx.y * 256u + threadIdx.x; i < n; i += (uint64_t)gridDim.y * 256u) d[i] = s[i];
}
__global__ void __launch_bounds__(256)
colsum_finish_kernel(const float *__restrict__ part, uint32_t nchunks, uint32_t cols, float *__restrict__ colsum) {
  pdl_grid_sync();
  const uint32_t j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  float t = 0.0f;
  for (uint32_t c = 0; c < nchunks; ++c) t += part[(uint64_t)c * cols + j];
  colsum[j] = t;
}

// ------------------------------------------------------------------------------- optimisers
struct AdamChunk {
  float *p;
  const float *g;
  float *m, *v;
  uint16_t *shadow; // optional bf16 copy of p at the same linear index (the GEMM operand shadow)
  uint32_t n;
  int vec; // bit 0: every pointer is 16-byte aligned (128-bit path); bit 1: write zeros back into g after reading it
};
constexpr uint64_t kAdamChunk = 32768; // elements per block of the multi-tensor launch
// sgd_step (reference include/autograd/sgd.hpp:23-37): p -= lr * g.  12 B/param.
__global__ void __launch_bounds__(256)
sgd_kernel(float *__restrict__ p, const float *__restrict__ g, uint64_t n, float lr, float gscale,
           bool vec) {
  pdl_grid_sync();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const uint64_t nq = n >> 2;
    for (uint64_t i = tid; i < nq; i += stride) {
      float4 pv = reinterpret_cast<float4 *>(p)[i];
      const float4 gv = reinterpret_cast<const float4 *>(g)[i];
      pv.x -= lr * (gscale * gv.x);
      pv.y -= lr * (gscale * gv.y);
      pv.z -= lr * (gscale * gv.z);
      pv.w -= lr * (gscale * gv.w);
      reinterpret_cast<float4 *>(p)[i] = pv;
    }
    for (uint64_t i = (nq << 2) + tid; i < n; i += stride) p[i] -= lr * (gscale * g[i]);
  } else {
    for (uint64_t i = tid; i < n; i += stride) p[i] -= lr * (gscale * g[i]);
  }
}

// adam_step (reference include/autograd/adam.hpp:70-106), one fused pass, 28 B/param:
// m = b1*m + (1-b1)*g ; v = b2*v + ((1-b2)*g)*g ; p -= (lr*m) / (bc1*(sqrt(v/bc2)+eps)).
struct AdamArgs {
  float lr, beta1, beta2, eps, bc1, bc2, gscale, omb1, omb2;
};
__device__ __forceinline__ void adam_one(float &p, float g, float &m, float &v, const AdamArgs &a) {
  g = a.gscale * g;
  m = a.beta1 * m + a.omb1 * g;
  v = a.beta2 * v + (a.omb2 * g) * g;
  p -= (a.lr * m) / (a.bc1 * (sqrtf(v / a.bc2) + a.eps));
}
__global__ void __launch_bounds__(256)
adam_kernel(float *__restrict__ p, const float *__restrict__ g, float *__restrict__ m,
            float *__restrict__ v, uint64_t n, AdamArgs a, bool vec) {
  pdl_grid_sync();
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t tid = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (vec) {
    const uint64_t nq = n >> 2;
    for (uint64_t i = tid; i < nq; i += stride) {
      float4 pv = reinterpret_cast<float4 *>(p)[i];
      const float4 gv = reinterpret_cast<const float4 *>(g)[i];
      float4 mv = reinterpret_cast<float4 *>(m)[i];
      float4 vv = reinterpret_cast<float4 *>(v)[i];
      adam_one(pv.x, gv.x, mv.x, vv.x, a);
      adam_one(pv.y, gv.y, mv.y, vv.y, a);
      adam_one(pv.z, gv.z, mv.z, vv.z, a);
      adam_one(pv.w, gv.w, mv.w, vv.w, a);
      reinterpret_cast<float4 *>(p)[i] = pv;
      reinterpret_cast<float4 *>(m)[i] = mv;
      reinterpret_cast<float4 *>(v)[i] = vv;
    }
    for (uint64_t i = (nq << 2) + tid; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], a);
  } else {
    for (uint64_t i = tid; i < n; i += stride) adam_one(p[i], g[i], m[i], v[i], a);
  }
}

// One block per chunk of one parameter: `count` parameters updated by a single launch.
__global__ void __launch_bounds__(256) adam_multi_kernel(const AdamChunk *__restrict__ table, AdamArgs a) {
  pdl_grid_sync();
  const AdamChunk c = table[blockIdx.x];
  // zero_g: the gradient is read here for the last time this step, so zero_grad's fill rides on this pass (the buffer is
  // then genuinely zero when the next backward accumulates into it: no per-gradient fill launch, no split-K zero fill)
  const bool zero_g = (c.vec & 2) && c.g;
  float *gw = const_cast<float *>(c.g);
  if (c.vec & 1) {
    const uint32_t nq = c.n >> 2;
    for (uint32_t i = threadIdx.x; i < nq; i += 256) {
      float4 pv = reinterpret_cast<float4 *>(c.p)[i];
      const float4 gv = c.g ? reinterpret_cast<const float4 *>(c.g)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
      float4 mv = reinterpret_cast<float4 *>(c.m)[i];
      float4 vv = reinterpret_cast<float4 *>(c.v)[i];
      adam_one(pv.x, gv.x, mv.x, vv.x, a);
      adam_one(pv.y, gv.y, mv.y, vv.y, a);
      adam_one(pv.z, gv.z, mv.z, vv.z, a);
      adam_one(pv.w, gv.w, mv.w, vv.w, a);
      reinterpret_cast<float4 *>(c.p)[i] = pv;
      reinterpret_cast<float4 *>(c.m)[i] = mv;
      reinterpret_cast<float4 *>(c.v)[i] = vv;
      if (zero_g) reinterpret_cast<float4 *>(gw)[i] = make_float4(0.f, 0.f, 0.f, 0.f);
      if (c.shadow) { // the updated weight leaves as fp32 and as the bf16 GEMM operand in the same pass
        __nv_bfloat162 o[2];
        o[0] = __floats2bfloat162_rn(pv.x, pv.y);
        o[1] = __floats2bfloat162_rn(pv.z, pv.w);
        reinterpret_cast<uint2 *>(c.shadow)[i] = *reinterpret_cast<const uint2 *>(o);
      }
    }
    for (uint32_t i = (nq << 2) + threadIdx.x; i < c.n; i += 256) {
      adam_one(c.p[i], c.g ? c.g[i] : 0.0f, c.m[i], c.v[i], a);
      if (zero_g) gw[i] = 0.0f;
      if (c.shadow) reinterpret_cast<__nv_bfloat16 *>(c.shadow)[i] = __float2bfloat16_rn(c.p[i]);
    }
  } else {
    for (uint32_t i = threadIdx.x; i < c.n; i += 256) {
      adam_one(c.p[i], c.g ? c.g[i] : 0.0f, c.m[i], c.v[i], a);
      if (zero_g) gw[i] = 0.0f;
      if (c.shadow) reinterpret_cast<__nv_bfloat16 *>(c.shadow)[i] = __float2bfloat16_rn(c.p[i]);
    }
  }
}

} // namespace weedcu

using namespace weedcu;

extern "C" {

// dst[t][0 .. n[t]) = src[t][0 .. n[t]) for up to 64 tensors per launch (pointers travel as kernel parameters)
int weedcu_multi_copy(uint32_t count, const float *const *src, float *const *dst, const uint64_t *n, void *stream) {
  if (!count) return 0;
  if (!src || !dst || !n) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  for (uint32_t t0 = 0; t0 < count; t0 += 64u) {
    MultiCopyArgs a;
    a.count = min(64u, count - t0);
    uint64_t longest = 0;
    for (uint32_t t = 0; t < a.count; ++t) {
      if (!src[t0 + t] || !dst[t0 + t]) return WEEDCU_EINVAL;
      a.src[t] = src[t0 + t];
      a.dst[t] = dst[t0 + t];
      a.n[t] = n[t0 + t];
      longest = longest > a.n[t] ? longest : a.n[t];
    }
    if (!longest) continue;
    const uint64_t per_block = 256u * 16u;
    const unsigned by = (unsigned)((longest + per_block - 1) / per_block < 1024u ? (longest + per_block - 1) / per_block : 1024u);
    ProfScope prof(WEEDCU_PROF_ELEMENTWISE, st, 0.0);
    launch_k(multi_copy_kernel, dim3(a.count, by), dim3(256), 0, st, a);
    const int rc = after_launch();
    if (rc) return rc;
  }
  return 0;
}
int weedcu_fill_real(float *p, uint64_t n, float value, void *stream) {
  if (!p) return WEEDCU_EINVAL;
  if (!n) return 0;
  ProfScope prof(WEEDCU_PROF_FILL, resolve_stream(stream), 4.0 * n);
  launch_k(fill_kernel<float>, dim3(grid_for((n + 3) / 4, 256, 8)), dim3(256), 0, resolve_stream(stream), p, n, value);
  return after_launch();
}
int weedcu_fill_int(int32_t *p, uint64_t n, int32_t value, void *stream) {
  if (!p) return WEEDCU_EINVAL;
  if (!n) return 0;
  launch_k(fill_kernel<int32_t>, dim3(grid_for((n + 3) / 4, 256, 8)), dim3(256), 0, resolve_stream(stream), p, n, value);
  return after_launch();
}

int weedcu_binary_real(int op, const float *a, const weedcu_view *av, const float *b,
                       const weedcu_view *bv, float *out, const weedcu_view *ov, void *stream) {
  if (!av || !bv || !ov) return WEEDCU_EINVAL;
  const weedcu_view *views[3] = {av, bv, ov};
  const float *ins[2] = {a, b};
  cudaStream_t st = resolve_stream(stream);
  switch (op) {
  case WEEDCU_ADD: return launch_ew<2>(views, ins, out, AddF(), st);
  case WEEDCU_MUL: return launch_ew<2>(views, ins, out, MulF(), st);
  case WEEDCU_SUB: return launch_ew<2>(views, ins, out, SubF(), st);
  case WEEDCU_DIV: return launch_ew<2>(views, ins, out, DivF(), st);
  }
  return WEEDCU_EINVAL;
}

int weedcu_inplace_real(int op, float *a, const weedcu_view *av, const float *b,
                        const weedcu_view *bv, void *stream) {
  if (!av || !bv) return WEEDCU_EINVAL;
  const weedcu_view *views[3] = {av, bv, av};
  const float *ins[2] = {a, b};
  cudaStream_t st = resolve_stream(stream);
  if (op == WEEDCU_ADD) return launch_ew<2>(views, ins, a, AddF(), st);
  if (op == WEEDCU_SUB) return launch_ew<2>(views, ins, a, SubF(), st);
  return WEEDCU_EINVAL;
}

int weedcu_copy_real(float *dst, const weedcu_view *dv, const float *src, const weedcu_view *sv,
                     void *stream) {
  if (!dv || !sv) return WEEDCU_EINVAL;
  const weedcu_view *views[2] = {sv, dv};
  const float *ins[1] = {src};
  return launch_ew<1>(views, ins, dst, CopyF(), resolve_stream(stream));
}

#define WCU_UNARY_CASE(OP)                                                                         \
  case OP: {                                                                                       \
    UnaryF<OP> f;                                                                                  \
    f.param = param;                                                                               \
    return launch_ew<1>(views, ins, out, f, st);                                                   \
  }
int weedcu_unary_real(int op, float param, const float *a, const weedcu_view *av, float *out,
                      const weedcu_view *ov, void *stream) {
  if (!av || !ov) return WEEDCU_EINVAL;
  const weedcu_view *views[2] = {av, ov};
  const float *ins[1] = {a};
  cudaStream_t st = resolve_stream(stream);
  switch (op) {
    WCU_UNARY_CASE(WEEDCU_RELU)
    WCU_UNARY_CASE(WEEDCU_SIGMOID)
    WCU_UNARY_CASE(WEEDCU_TANH)
    WCU_UNARY_CASE(WEEDCU_ABS)
    WCU_UNARY_CASE(WEEDCU_POW)
    WCU_UNARY_CASE(WEEDCU_EXP)
    WCU_UNARY_CASE(WEEDCU_LOG)
    WCU_UNARY_CASE(WEEDCU_GELU)
    WCU_UNARY_CASE(WEEDCU_SIN)
    WCU_UNARY_CASE(WEEDCU_COS)
  }
  return WEEDCU_EINVAL;
}

#define WCU_GRAD_CASE(OP)                                                                          \
  case OP:                                                                                         \
    return accumulate ? launch_ew<3>(views, ins, din, UnaryGradF<OP>(), st)                        \
                      : launch_ew<2>(views + 1, ins + 1, din, UnaryGradSetF<OP>(), st);
int weedcu_unary_grad_real(int op, float *din, const weedcu_view *dinv, const float *in,
                           const weedcu_view *inv, const float *dout, const weedcu_view *doutv,
                           int accumulate, void *stream) {
  if (!dinv || !inv || !doutv) return WEEDCU_EINVAL;
  const weedcu_view *views[4] = {dinv, inv, doutv, dinv};
  const float *ins[3] = {din, in, dout};
  cudaStream_t st = resolve_stream(stream);
  switch (op) {
    WCU_GRAD_CASE(WEEDCU_RELU)
    WCU_GRAD_CASE(WEEDCU_SIGMOID)
    WCU_GRAD_CASE(WEEDCU_TANH)
    WCU_GRAD_CASE(WEEDCU_ABS)
    WCU_GRAD_CASE(WEEDCU_GELU)
    WCU_GRAD_CASE(WEEDCU_SIN)
    WCU_GRAD_CASE(WEEDCU_COS)
  }
  return WEEDCU_EINVAL;
}

int weedcu_clamp_real(const float *a, const weedcu_view *av, float lo, float hi, float *out, const weedcu_view *ov, void *stream) {
  if (!av || !ov) return WEEDCU_EINVAL;
  const weedcu_view *views[2] = {av, ov};
  const float *ins[1] = {a};
  ClampF f;
  f.lo = lo;
  f.hi = hi;
  return launch_ew<1>(views, ins, out, f, resolve_stream(stream));
}
int weedcu_clamp_grad_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout,
                           const weedcu_view *doutv, float lo, float hi, void *stream) {
  if (!dinv || !inv || !doutv) return WEEDCU_EINVAL;
  const weedcu_view *views[4] = {dinv, inv, doutv, dinv};
  const float *ins[3] = {din, in, dout};
  ClampGradF f;
  f.lo = lo;
  f.hi = hi;
  return launch_ew<3>(views, ins, din, f, resolve_stream(stream));
}
int weedcu_match_grad_full_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout,
                                const weedcu_view *doutv, const float *extremum, void *stream) {
  if (!dinv || !inv || !doutv || !extremum) return WEEDCU_EINVAL;
  weedcu_view mv = *dinv; // the scalar, broadcast over din's index space
  mv.offset = 0;
  for (int d = 0; d < WEEDCU_MAX_RANK; ++d) mv.stride[d] = 0;
  const weedcu_view *views[5] = {dinv, inv, doutv, &mv, dinv};
  const float *ins[4] = {din, in, dout, extremum};
  return launch_ew<4>(views, ins, din, MatchGradF(), resolve_stream(stream));
}

extern "C++" {
template <typename DT>
static int gelu_grad_pack_impl(float *din, const float *in, const DT *dout, uint32_t rows, uint32_t cols, int accumulate, uint16_t *din_bf16,
                               float *colsum, void *stream) {
  if (!in || !dout || !din_bf16 || !colsum || !rows || !cols) return WEEDCU_EINVAL;
  if (!din && accumulate) return WEEDCU_EINVAL; // nothing to accumulate into
  if ((rows % 8u) || (din && !aligned16(din)) || !aligned16(in) || !aligned16(dout) || !aligned16(din_bf16)) return WEEDCU_ENOSUP;
  const uint32_t nchunks = (rows + 1023u) / 1024u, cgroups = (cols + kGgCols - 1) / kGgCols;
  if (cgroups > 65535u) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  float *part = nullptr;
  WCU_CHECK(pool_alloc((void **)&part, sizeof(float) * (size_t)nchunks * cols, st));
  ProfScope prof(WEEDCU_PROF_ELEMENTWISE, st, ((accumulate ? 14.0 : (din ? 10.0 : 6.0)) + sizeof(DT)) * (double)rows * cols);
  if (!din && sizeof(DT) == 2 && (cols + kGg16Cols - 1) / kGg16Cols <= 65535u)
    launch_k(gelu_grad_pack16_kernel, dim3(nchunks, (cols + kGg16Cols - 1) / kGg16Cols), dim3(128), 0, st, in, (const __nv_bfloat16 *)dout, rows, cols,
             (__nv_bfloat16 *)din_bf16, part);
  else // (fp32 dout keeps the accurate tanh with or without an fp32 output: switching the operand cache on changes no value)
    launch_k(gelu_grad_pack_kernel<DT, false>, dim3(nchunks, cgroups), dim3(256), 0, st, din, in, dout, rows, cols, accumulate, (__nv_bfloat16 *)din_bf16, part);
  int rc = after_launch();
  if (rc == 0) {
    launch_k(colsum_finish_kernel, dim3((cols + 255u) / 256u), dim3(256), 0, st, part, nchunks, cols, colsum);
    rc = after_launch();
  }
  pool_free(part, st);
  return rc;
}
} // extern "C++"
int weedcu_gelu_grad_pack(float *din, const float *in, const float *dout, uint32_t rows, uint32_t cols, int accumulate,
                          uint16_t *din_bf16, float *colsum, void *stream) {
  return gelu_grad_pack_impl<float>(din, in, dout, rows, cols, accumulate, din_bf16, colsum, stream);
}
int weedcu_gelu_grad_pack_bf16dy(float *din, const float *in, const uint16_t *dout_bf16, uint32_t rows, uint32_t cols, int accumulate,
                                 uint16_t *din_bf16, float *colsum, void *stream) {
  return gelu_grad_pack_impl<__nv_bfloat16>(din, in, (const __nv_bfloat16 *)dout_bf16, rows, cols, accumulate, din_bf16, colsum, stream);
}

int weedcu_gelu_fwd_bf16(const float *x, float *y, uint16_t *y_bf16, uint64_t n, void *stream) {
  if (!x || !y_bf16 || !n) return WEEDCU_EINVAL;
  if ((n % 4u) || !aligned16(x) || (y && !aligned16(y)) || (((uintptr_t)y_bf16) & 7u)) return WEEDCU_ENOSUP;
  cudaStream_t st = resolve_stream(stream);
  ProfScope prof(WEEDCU_PROF_ELEMENTWISE, st, (y ? 10.0 : 6.0) * (double)n);
  launch_k(gelu_fwd_bf16_kernel, dim3(grid_for(n / 4u, 256, 16)), dim3(256), 0, st, x, y, (__nv_bfloat16 *)y_bf16, n / 4u);
  return after_launch();
}

int weedcu_adam_step_multi(uint32_t count, float *const *p, const float *const *g, float *const *m,
                           float *const *v, const uint64_t *n, float lr, float beta1, float beta2,
                           float eps, float bc1, float bc2, float gscale, void *stream) {
  return weedcu_adam_step_multi_shadow(count, p, g, m, v, n, nullptr, lr, beta1, beta2, eps, bc1, bc2, gscale, stream);
}

int weedcu_adam_step_multi_shadow(uint32_t count, float *const *p, const float *const *g, float *const *m,
                                  float *const *v, const uint64_t *n, uint16_t *const *shadow, float lr, float beta1,
                                  float beta2, float eps, float bc1, float bc2, float gscale, void *stream) {
  return weedcu_adam_step_multi_zero(count, p, g, m, v, n, shadow, nullptr, lr, beta1, beta2, eps, bc1, bc2, gscale, stream);
}

int weedcu_adam_step_multi_zero(uint32_t count, float *const *p, const float *const *g, float *const *m, float *const *v,
                                const uint64_t *n, uint16_t *const *shadow, const uint8_t *zero_grad, float lr, float beta1, float beta2,
                                float eps, float bc1, float bc2, float gscale, void *stream) {
  if (!count) return 0;
  if (!p || !g || !m || !v || !n) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  // Chunk tables stay on the device between calls, keyed by (stream, contents): a training step that updates its
  // parameters bucket by bucket (data parallel: one launch per reduced gradient bucket, on the communication stream)
  // re-uses one cached table per bucket, and two optimisers / devices / streams never share one. Only a table seen
  // for the first time is uploaded (one blocking copy from pageable memory).
  struct CachedTable {
    cudaStream_t stream;
    std::vector<AdamChunk> host;
    AdamChunk *dev;
  };
  static std::vector<CachedTable> cache; // most recently used last
  static std::mutex table_mutex;
  std::lock_guard<std::mutex> lock(table_mutex);
  std::vector<AdamChunk> table;
  double total = 0.0, shadowed = 0.0, zeroed = 0.0;
  for (uint32_t t = 0; t < count; ++t) {
    if (!p[t] || !m[t] || !v[t]) return WEEDCU_EINVAL; // g[t] == NULL: an all-zero gradient
    if (zero_grad && zero_grad[t] && g[t]) zeroed += (double)n[t];
    uint16_t *sh = shadow ? shadow[t] : nullptr;
    const int vec = ((aligned16(p[t]) && (!g[t] || aligned16(g[t])) && aligned16(m[t]) && aligned16(v[t]) && (!sh || aligned16(sh))) ? 1 : 0) |
                    ((zero_grad && zero_grad[t]) ? 2 : 0);
    for (uint64_t o = 0; o < n[t]; o += kAdamChunk) {
      const uint64_t len = (n[t] - o < kAdamChunk) ? n[t] - o : kAdamChunk;
      table.push_back(AdamChunk{p[t] + o, g[t] ? g[t] + o : nullptr, m[t] + o, v[t] + o, sh ? sh + o : nullptr, (uint32_t)len, vec});
    }
    total += (double)n[t];
    if (sh) shadowed += (double)n[t];
  }
  if (table.empty()) return 0;
  size_t hit = cache.size();
  for (size_t i = 0; i < cache.size(); ++i)
    if (cache[i].stream == st && cache[i].host.size() == table.size() &&
        memcmp(cache[i].host.data(), table.data(), table.size() * sizeof(AdamChunk)) == 0) {
      hit = i;
      break;
    }
  if (hit == cache.size()) {
    if (cache.size() >= 64) { // oldest entry goes back to the pool on the stream it was used on
      pool_free(cache.front().dev, cache.front().stream);
      cache.erase(cache.begin());
    }
    CachedTable e{st, std::move(table), nullptr};
    WCU_CHECK(pool_alloc((void **)&e.dev, e.host.size() * sizeof(AdamChunk), st));
    note_stream_op();
    WCU_CHECK(cudaMemcpyAsync(e.dev, e.host.data(), e.host.size() * sizeof(AdamChunk), cudaMemcpyHostToDevice, st));
    WCU_CHECK(cudaStreamSynchronize(st)); // pageable source: make the staging copy complete
    cache.push_back(std::move(e));
  } else if (hit + 1 != cache.size()) {
    CachedTable e = std::move(cache[hit]);
    cache.erase(cache.begin() + (long)hit);
    cache.push_back(std::move(e));
  }
  const CachedTable &use = cache.back();
  AdamArgs a = {lr, beta1, beta2, eps, bc1, bc2, gscale, 1.0f - beta1, 1.0f - beta2};
  ProfScope prof(WEEDCU_PROF_OPTIMIZER, st, 28.0 * total + 2.0 * shadowed + 4.0 * zeroed);
  launch_k(adam_multi_kernel, dim3((unsigned)use.host.size()), dim3(256), 0, st, use.dev, a);
  return after_launch();
}

int weedcu_sgd_step(float *p, const float *g, uint64_t n, float lr, float gscale, void *stream) {
  if (!p || !g) return WEEDCU_EINVAL;
  if (!n) return 0;
  const bool vec = aligned16(p) && aligned16(g);
  ProfScope prof(WEEDCU_PROF_OPTIMIZER, resolve_stream(stream), 12.0 * n);
  launch_k(sgd_kernel, dim3(grid_for((n + 3) / 4, 256, 8)), dim3(256), 0, resolve_stream(stream), p, g, n, lr, gscale,
                                                                               vec);
  return after_launch();
}

int weedcu_adam_step(float *p, const float *g, float *m, float *v, uint64_t n, float lr,
                     float beta1, float beta2, float eps, float bc1, float bc2, float gscale,
                     void *stream) {
  if (!p || !g || !m || !v) return WEEDCU_EINVAL;
  if (!n) return 0;
  AdamArgs a = {lr, beta1, beta2, eps, bc1, bc2, gscale, 1.0f - beta1, 1.0f - beta2};
  const bool vec = aligned16(p) && aligned16(g) && aligned16(m) && aligned16(v);
  ProfScope prof(WEEDCU_PROF_OPTIMIZER, resolve_stream(stream), 28.0 * n);
  launch_k(adam_kernel, dim3(grid_for((n + 3) / 4, 256, 8)), dim3(256), 0, resolve_stream(stream), p, g, m, v, n, a,
                                                                                vec);
  return after_launch();
}

} // extern "C"
