// reduce.cu — axis sum + its gradient, full sum/mean, row argmax.  HBM-bound (4 B / input elem).
//
// Reference: Weed::reduce / reduce_grad (src/ops/reduce.cpp:17-38,60-66,84-113) and Weed::sum /
// mean (src/ops/sum.cpp:74-98). The reference's OpenCL reduce kernels never receive the rank
// (SURVEY §2.3 defect 2) and full sum copies the buffer to the host (sum.cpp:52-67); here both are
// real device reductions: shared-memory staging across column slices + warp-shuffle trees.
#include "common.cuh"

namespace weedcu {

struct ReduceView {
  int rank, axis;
  uint32_t shape[kMaxRank];
  uint32_t stride[kMaxRank];
};

// Generic / reference-order kernel: one thread per output, serial over the axis. Used for
// index_order = 1 (reproduces REDUCE_HEAD's last-dim-fastest decomposition, reduce.cpp:17-31) and
// for layouts the tiled kernels do not cover.
__global__ void __launch_bounds__(256)
reduce_generic_kernel(const float *__restrict__ a, ReduceView v, uint32_t n_out, float *__restrict__ out,
                      int index_order) {
  pdl_grid_sync();
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  uint64_t base = 0;
  uint32_t tmp = o;
  if (index_order) {
    for (int d = v.rank - 1; d >= 0; --d) {
      if (d == v.axis) continue;
      base += (uint64_t)(tmp % v.shape[d]) * v.stride[d];
      tmp /= v.shape[d];
    }
  } else {
    for (int d = 0; d < v.rank; ++d) {
      if (d == v.axis) continue;
      base += (uint64_t)(tmp % v.shape[d]) * v.stride[d];
      tmp /= v.shape[d];
    }
  }
  float s = 0.0f;
  const uint64_t as = v.stride[v.axis];
  for (uint32_t j = 0; j < v.shape[v.axis]; ++j) s += a[base + j * as];
  out[o] = s;
}

// Canonical contiguous form a[inner, L, outer] (strides 1, inner, inner*L), output [inner, outer].
// Case "strided axis" (inner >= 32): a warp spans 32 adjacent outputs (coalesced 128-B rows), the
// BY warps of a block split the axis, partial sums meet in shared memory.
template <int BY>
__global__ void __launch_bounds__(32 * BY)
reduce_strided_kernel(const float *__restrict__ a, uint32_t inner, uint32_t L, float *__restrict__ out) {
  pdl_grid_sync();
  __shared__ float part[BY][33];
  const uint32_t ii = blockIdx.x * 32 + threadIdx.x;
  const uint64_t slab = (uint64_t)blockIdx.y * inner * L;
  float s0 = 0.0f, s1 = 0.0f, s2 = 0.0f, s3 = 0.0f;
  if (ii < inner) {
    const float *p = a + slab + ii;
    uint32_t j = threadIdx.y;
    for (; j + 3 * BY < L; j += 4 * BY) { // four independent loads in flight per thread
      s0 += p[(uint64_t)j * inner];
      s1 += p[(uint64_t)(j + BY) * inner];
      s2 += p[(uint64_t)(j + 2 * BY) * inner];
      s3 += p[(uint64_t)(j + 3 * BY) * inner];
    }
    for (; j < L; j += BY) s0 += p[(uint64_t)j * inner];
  }
  part[threadIdx.y][threadIdx.x] = (s0 + s1) + (s2 + s3);
  __syncthreads();
  if (threadIdx.y == 0 && ii < inner) {
    float t = 0.0f;
#pragma unroll
    for (int y = 0; y < BY; ++y) t += part[y][threadIdx.x];
    out[(uint64_t)blockIdx.y * inner + ii] = t;
  }
}

// Case "contiguous axis" (inner == 1): one block per output, 128-bit loads along the axis.
__global__ void __launch_bounds__(256)
reduce_contig_kernel(const float *__restrict__ a, uint32_t L, float *__restrict__ out) {
  pdl_grid_sync();
  __shared__ float red[32];
  const float *p = a + (uint64_t)blockIdx.x * L;
  float s = 0.0f;
  if ((((uintptr_t)p) & 15u) == 0) {
    const uint32_t nq = L >> 2;
    const float4 *q = reinterpret_cast<const float4 *>(p);
    for (uint32_t i = threadIdx.x; i < nq; i += blockDim.x) {
      const float4 v = q[i];
      s += (v.x + v.y) + (v.z + v.w);
    }
    for (uint32_t i = (nq << 2) + threadIdx.x; i < L; i += blockDim.x) s += p[i];
  } else {
    for (uint32_t i = threadIdx.x; i < L; i += blockDim.x) s += p[i];
  }
  s = block_sum(s, red);
  if (threadIdx.x == 0) out[blockIdx.x] = s;
}

// Short contiguous axis (L <= 32, e.g. the batch sum behind a broadcast positional-encoding
// gradient: [B, T*d] over B = 8): one thread per output walks its L consecutive elements in the
// reference's serial order; adjacent threads read adjacent 4L-byte runs, so the warp's loads cover
// one dense span (128-bit loads when L % 4 == 0).
template <bool VEC>
__global__ void __launch_bounds__(256)
reduce_contig_short_kernel(const float *__restrict__ a, uint32_t L, uint32_t n_out, float *__restrict__ out) {
  pdl_grid_sync();
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  const float *p = a + (uint64_t)o * L;
  float s = 0.0f;
  if (VEC) {
    const float4 *q = reinterpret_cast<const float4 *>(p);
    for (uint32_t i = 0; i < (L >> 2); ++i) {
      const float4 v = q[i];
      s += v.x;
      s += v.y;
      s += v.z;
      s += v.w;
    }
  } else {
    for (uint32_t i = 0; i < L; ++i) s += p[i];
  }
  out[o] = s;
}
// Medium contiguous axis (32 < L < 2048): one warp per output, 8 outputs per block.
__global__ void __launch_bounds__(256)
reduce_contig_warp_kernel(const float *__restrict__ a, uint32_t L, uint32_t n_out, float *__restrict__ out) {
  pdl_grid_sync();
  const uint32_t o = blockIdx.x * 8 + (threadIdx.x >> 5), lane = threadIdx.x & 31;
  if (o >= n_out) return;
  const float *p = a + (uint64_t)o * L;
  float s = 0.0f;
  if ((((uintptr_t)p) & 15u) == 0) {
    const uint32_t nq = L >> 2;
    const float4 *q = reinterpret_cast<const float4 *>(p);
    for (uint32_t i = lane; i < nq; i += 32) {
      const float4 v = q[i];
      s += (v.x + v.y) + (v.z + v.w);
    }
    for (uint32_t i = (nq << 2) + lane; i < L; i += 32) s += p[i];
  } else {
    for (uint32_t i = lane; i < L; i += 32) s += p[i];
  }
  s = warp_sum(s);
  if (lane == 0) out[o] = s;
}

// reduce_grad, reference order (REDUCE_GRAD_HEAD, reduce.cpp:84-101): i is decomposed over the
// NON-axis dims only, last dim fastest.
template <int NOPS>
__global__ void __launch_bounds__(256)
reduce_grad_reforder_kernel(float *din, IndexSpace<NOPS> sp_din, ReduceView dims,
                            const float *__restrict__ dout, ReduceView dv) {
  pdl_grid_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= sp_din.n) return;
  uint64_t o = 0;
  uint32_t tmp = i;
  for (int d = dims.rank - 1; d >= 0; --d) {
    if (d == dims.axis) continue;
    o += (uint64_t)(tmp % dims.shape[d]) * dv.stride[d];
    tmp /= dims.shape[d];
  }
  uint64_t off = 0;
  uint32_t rem = i;
  for (int d = 0; d < sp_din.rank; ++d) {
    off += (uint64_t)(rem % sp_din.shape[d]) * sp_din.stride[0][d];
    rem /= sp_din.shape[d];
  }
  din[off] += dout[o];
}

// ------------------------------------------------------------------------ full sum (two pass)
template <int NOPS>
__global__ void __launch_bounds__(256)
sum_pass1_kernel(const float *__restrict__ a, IndexSpace<NOPS> sp, bool linear_vec, float *__restrict__ partial) {
  pdl_grid_sync();
  __shared__ float red[32];
  float s0 = 0.0f, s1 = 0.0f;
  const uint32_t stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (linear_vec) {
    const uint32_t nq = sp.n >> 2;
    const float4 *q = reinterpret_cast<const float4 *>(a);
    uint32_t i = tid;
    for (; i + stride < nq; i += 2 * stride) {
      const float4 v = q[i], w = q[i + stride];
      s0 += (v.x + v.y) + (v.z + v.w);
      s1 += (w.x + w.y) + (w.z + w.w);
    }
    if (i < nq) {
      const float4 v = q[i];
      s0 += (v.x + v.y) + (v.z + v.w);
    }
    for (uint32_t k = (nq << 2) + tid; k < sp.n; k += stride) s0 += a[k];
  } else {
    for (uint32_t i = tid; i < sp.n; i += stride) {
      uint64_t off = 0;
      uint32_t rem = i;
      for (int d = 0; d < sp.rank; ++d) {
        off += (uint64_t)(rem % sp.shape[d]) * sp.stride[0][d];
        rem /= sp.shape[d];
      }
      s0 += a[off];
    }
  }
  const float t = block_sum(s0 + s1, red);
  if (threadIdx.x == 0) partial[blockIdx.x] = t;
}
__global__ void __launch_bounds__(1024)
sum_pass2_kernel(const float *__restrict__ partial, uint32_t n, float scale, float *__restrict__ out) {
  pdl_grid_sync();
  __shared__ float red[32];
  float s = 0.0f;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) s += partial[i];
  s = block_sum(s, red);
  if (threadIdx.x == 0) *out = s * scale;
}

// ------------------------------------------------------------------------ max / min
// Full max / min (reference src/ops/real_extremum.cpp:88-106: per-thread running extremum, then the extremum of those)
// as the same two passes as the full sum; the axis forms (src/ops/reduce.cpp:40-58,239-248) share reduce_generic's
// index walk, and match_grad (:84-101, MATCH_GRAD_OUT) routes dout to the elements equal to their reduced value.
template <bool IS_MIN> __device__ __forceinline__ float pick(float a, float b) { return IS_MIN ? fminf(a, b) : fmaxf(a, b); }
template <bool IS_MIN>
__global__ void __launch_bounds__(256)
extremum_pass1_kernel(const float *__restrict__ a, IndexSpace<1> sp, bool linear_vec, float *__restrict__ partial) {
  pdl_grid_sync();
  __shared__ float red[32];
  const float init = IS_MIN ? INFINITY : -INFINITY;
  float m = init;
  const uint32_t stride = gridDim.x * blockDim.x, tid = blockIdx.x * blockDim.x + threadIdx.x;
  if (linear_vec) {
    const uint32_t nq = sp.n >> 2;
    const float4 *q = reinterpret_cast<const float4 *>(a);
    for (uint32_t i = tid; i < nq; i += stride) {
      const float4 v = q[i];
      m = pick<IS_MIN>(m, pick<IS_MIN>(pick<IS_MIN>(v.x, v.y), pick<IS_MIN>(v.z, v.w)));
    }
    for (uint32_t k = (nq << 2) + tid; k < sp.n; k += stride) m = pick<IS_MIN>(m, a[k]);
  } else {
    for (uint32_t i = tid; i < sp.n; i += stride) {
      uint64_t off = 0;
      uint32_t rem = i;
      for (int d = 0; d < sp.rank; ++d) {
        off += (uint64_t)(rem % sp.shape[d]) * sp.stride[0][d];
        rem /= sp.shape[d];
      }
      m = pick<IS_MIN>(m, a[off]);
    }
  }
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = pick<IS_MIN>(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[w] = m;
  __syncthreads();
  if (w == 0) {
    m = lane < (blockDim.x >> 5) ? red[lane] : init;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = pick<IS_MIN>(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) partial[blockIdx.x] = m;
  }
}
template <bool IS_MIN>
__global__ void __launch_bounds__(1024) extremum_pass2_kernel(const float *__restrict__ partial, uint32_t n, float *__restrict__ out) {
  pdl_grid_sync();
  __shared__ float red[32];
  const float init = IS_MIN ? INFINITY : -INFINITY;
  float m = init;
  for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) m = pick<IS_MIN>(m, partial[i]);
  const uint32_t lane = threadIdx.x & 31, w = threadIdx.x >> 5;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) m = pick<IS_MIN>(m, __shfl_xor_sync(0xffffffffu, m, o));
  if (lane == 0) red[w] = m;
  __syncthreads();
  if (w == 0) {
    m = red[lane];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = pick<IS_MIN>(m, __shfl_xor_sync(0xffffffffu, m, o));
    if (lane == 0) *out = m;
  }
}
// axis form: a warp covers 32 adjacent outputs (coalesced when the first non-axis dim is the contiguous one)
template <bool IS_MIN>
__global__ void __launch_bounds__(256)
extremum_axis_kernel(const float *__restrict__ a, ReduceView v, uint32_t n_out, float *__restrict__ out, int index_order) {
  pdl_grid_sync();
  const uint32_t o = blockIdx.x * blockDim.x + threadIdx.x;
  if (o >= n_out) return;
  uint64_t base = 0;
  uint32_t tmp = o;
  if (index_order) {
    for (int d = v.rank - 1; d >= 0; --d) {
      if (d == v.axis) continue;
      base += (uint64_t)(tmp % v.shape[d]) * v.stride[d];
      tmp /= v.shape[d];
    }
  } else {
    for (int d = 0; d < v.rank; ++d) {
      if (d == v.axis) continue;
      base += (uint64_t)(tmp % v.shape[d]) * v.stride[d];
      tmp /= v.shape[d];
    }
  }
  const uint64_t as = v.stride[v.axis];
  float m = a[base];
  for (uint32_t j = 1; j < v.shape[v.axis]; ++j) {
    const float x = a[base + j * as];
    if (IS_MIN ? (x < m) : (x > m)) m = x; // the reference's strict comparison (MAX_LOOP / MIN_LOOP): NaNs never win
  }
  out[o] = m;
}
// din[i] += dout[o(i)] where in[i] == red[o(i)]; o(i) drops the axis coordinate. index_order 1 decomposes i over the
// non-axis dims only, last dim fastest, exactly as REDUCE_GRAD_HEAD does (correct only where the reference is, D2);
// index_order 0 decomposes i over all dims, first dim fastest (the intended column-major walk).
__global__ void __launch_bounds__(256)
match_grad_kernel(float *din, ReduceView dinv, const float *__restrict__ in, ReduceView inv, const float *__restrict__ dout,
                  const float *__restrict__ red, ReduceView ov, uint32_t n, int index_order) {
  pdl_grid_sync();
  const uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint64_t o = 0, off_d = 0, off_i = 0;
  uint32_t rem = i;
  for (int d = 0; d < dinv.rank; ++d) {
    const uint32_t c = rem % dinv.shape[d];
    rem /= dinv.shape[d];
    off_d += (uint64_t)c * dinv.stride[d];
    off_i += (uint64_t)c * inv.stride[d];
    if (!index_order && d != dinv.axis) o += (uint64_t)c * ov.stride[d];
  }
  if (index_order) {
    uint32_t tmp = i;
    for (int d = dinv.rank - 1; d >= 0; --d) {
      if (d == dinv.axis) continue;
      o += (uint64_t)(tmp % dinv.shape[d]) * ov.stride[d];
      tmp /= dinv.shape[d];
    }
  }
  if (in[off_i] == red[o]) din[off_d] += dout[o];
}

// ------------------------------------------------------------------------ argmax over rows
// logits[rows, V] with row stride rs (normally 1) and vocab stride vs: 32 rows x BY slices per
// block, lowest index wins ties (matches a serial first-max scan).
// The vocabulary is split over gridDim.y slices so that a decode step's 8 x 50257 scan fills the chip
// (one block took 347 us, a fifth of the step); a slice leaves (value, index) per row, the merge kernel
// keeps the serial-scan winner: larger value, lowest index on ties, slices in ascending order.
template <int BY>
__global__ void __launch_bounds__(32 * BY)
argmax_rows_kernel(const float *__restrict__ x, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs, uint32_t v_per_slice,
                   float *__restrict__ part_v, uint32_t *__restrict__ part_i) {
  pdl_grid_sync();
  __shared__ float mv[BY][33];
  __shared__ uint32_t mi[BY][33];
  const uint32_t r = blockIdx.x * 32 + threadIdx.x;
  const uint32_t v_begin = blockIdx.y * v_per_slice, v_end = min(V, v_begin + v_per_slice);
  float best = -INFINITY;
  uint32_t bi = 0xffffffffu;
  if (r < rows) {
    const float *p = x + (uint64_t)r * rs;
    for (uint32_t v = v_begin + threadIdx.y; v < v_end; v += BY) {
      const float y = p[(uint64_t)v * vs];
      if (y > best || bi == 0xffffffffu) {
        best = y;
        bi = v;
      }
    }
  }
  mv[threadIdx.y][threadIdx.x] = best;
  mi[threadIdx.y][threadIdx.x] = bi;
  __syncthreads();
  if (threadIdx.y == 0 && r < rows) {
    for (int y = 1; y < BY; ++y) {
      const float c = mv[y][threadIdx.x];
      const uint32_t ci = mi[y][threadIdx.x];
      if (ci != 0xffffffffu && (bi == 0xffffffffu || c > best || (c == best && ci < bi))) {
        best = c;
        bi = ci;
      }
    }
    part_v[(uint64_t)blockIdx.y * rows + r] = best;
    part_i[(uint64_t)blockIdx.y * rows + r] = bi;
  }
}
__global__ void __launch_bounds__(256)
argmax_merge_kernel(const float *__restrict__ part_v, const uint32_t *__restrict__ part_i, uint32_t rows, uint32_t slices,
                    int32_t *__restrict__ out) {
  pdl_grid_sync();
  const uint32_t r = blockIdx.x * blockDim.x + threadIdx.x;
  if (r >= rows) return;
  float best = part_v[r];
  uint32_t bi = part_i[r];
  for (uint32_t s = 1; s < slices; ++s) {
    const float c = part_v[(uint64_t)s * rows + r];
    const uint32_t ci = part_i[(uint64_t)s * rows + r];
    if (ci != 0xffffffffu && (bi == 0xffffffffu || c > best || (c == best && ci < bi))) {
      best = c;
      bi = ci;
    }
  }
  out[r] = (int32_t)bi;
}

static bool canonical_contiguous(const weedcu_view *v, int axis, uint64_t &inner, uint64_t &outer) {
  uint64_t st = 1;
  inner = 1;
  outer = 1;
  for (int d = 0; d < v->rank; ++d) {
    const uint32_t ext = v->shape[d];
    if (ext != 1 && v->stride[d] != st) return false;
    if (d < axis) inner *= ext;
    if (d > axis) outer *= ext;
    st *= ext;
  }
  return true;
}

} // namespace weedcu

using namespace weedcu;

extern "C" {

int weedcu_reduce_real(const float *a, const weedcu_view *av, int axis, float *out, int index_order,
                       void *stream) {
  if (!a || !av || !out || av->rank <= 0 || av->rank > kMaxRank || axis < 0 || axis >= av->rank)
    return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  uint64_t total = 1;
  for (int d = 0; d < av->rank; ++d) total *= av->shape[d];
  const uint32_t L = av->shape[axis];
  if (!L || !total) return WEEDCU_EINVAL;
  const uint64_t n_out = total / L;
  const float *base = a + av->offset;
  ProfScope prof(WEEDCU_PROF_REDUCE, st, 4.0 * total);
  uint64_t inner, outer;
  // Row-major and column-major enumeration of the non-axis coordinates coincide when at most one
  // non-axis dim has extent > 1; then the fast kernels serve index_order 1 as well.
  int big = 0;
  for (int d = 0; d < av->rank; ++d)
    if (d != axis && av->shape[d] > 1) ++big;
  const bool order_free = (index_order == 0) || (big <= 1);
  if (order_free && canonical_contiguous(av, axis, inner, outer)) {
    if (inner == 1) {
      if (L <= 32) {
        const unsigned blocks = (unsigned)((outer + 255) / 256);
        if ((L % 4u) == 0 && aligned16(base))
          launch_k(reduce_contig_short_kernel<true>, dim3(blocks), dim3(256), 0, st, base, L, (uint32_t)outer, out);
        else
          launch_k(reduce_contig_short_kernel<false>, dim3(blocks), dim3(256), 0, st, base, L, (uint32_t)outer, out);
      } else if (L < 2048 && outer >= 2 * (uint64_t)kNumSMs) {
        launch_k(reduce_contig_warp_kernel, dim3((unsigned)((outer + 7) / 8)), dim3(256), 0, st, base, L, (uint32_t)outer, out);
      } else {
        launch_k(reduce_contig_kernel, dim3((unsigned)outer), dim3(256), 0, st, base, L, out);
      }
      return after_launch();
    }
    if (inner >= 32 && outer <= 65535) {
      dim3 grid((unsigned)((inner + 31) / 32), (unsigned)outer);
      const uint64_t blocks = (uint64_t)grid.x * grid.y;
      if (L >= 64 && blocks < 4 * kNumSMs) {
        launch_k(reduce_strided_kernel<16>, dim3(grid), dim3(32, 16), 0, st, base, (uint32_t)inner, L, out);
      } else {
        launch_k(reduce_strided_kernel<4>, dim3(grid), dim3(32, 4), 0, st, base, (uint32_t)inner, L, out);
      }
      return after_launch();
    }
  }
  ReduceView v;
  v.rank = av->rank;
  v.axis = axis;
  for (int d = 0; d < kMaxRank; ++d) {
    v.shape[d] = d < av->rank ? av->shape[d] : 1;
    v.stride[d] = d < av->rank ? av->stride[d] : 0;
  }
  launch_k(reduce_generic_kernel, dim3((unsigned)((n_out + 255) / 256)), dim3(256), 0, st, base, v, (uint32_t)n_out, out,
                                                                       index_order);
  return after_launch();
}

int weedcu_reduce_grad_real(float *din, const weedcu_view *dinv, const float *dout,
                            const weedcu_view *doutv, int axis, int index_order, void *stream) {
  if (!din || !dinv || !dout || !doutv || dinv->rank != doutv->rank || axis < 0 ||
      axis >= dinv->rank)
    return WEEDCU_EINVAL;
  int big = 0;
  for (int d = 0; d < dinv->rank; ++d)
    if (d != axis && dinv->shape[d] > 1) ++big;
  // The reference's decomposition skips the axis coordinate entirely, so it agrees with the
  // intended broadcast only when the axis is the last dim with extent > 1 and <= 1 other dim is.
  bool axis_last = true;
  for (int d = axis + 1; d < dinv->rank; ++d)
    if (dinv->shape[d] > 1) axis_last = false;
  if (index_order == 0 || (big <= 1 && axis_last)) {
    // intended semantics: din[i] += dout[i] with dout broadcast (stride 0) along `axis`
    weedcu_view dv = *doutv;
    dv.stride[axis] = 0;
    return weedcu_inplace_real(WEEDCU_ADD, din, dinv, dout, &dv, stream);
  }
  const weedcu_view *views[1] = {dinv};
  IndexSpace<1> sp;
  // no collapsing: the reference formula needs the original dims
  sp.rank = dinv->rank;
  uint64_t n = 1;
  for (int d = 0; d < kMaxRank; ++d) {
    sp.shape[d] = d < dinv->rank ? dinv->shape[d] : 1;
    sp.stride[0][d] = d < dinv->rank ? dinv->stride[d] : 0;
    n *= sp.shape[d];
  }
  (void)views;
  if (n > 0xffffffffull) return WEEDCU_EINVAL;
  sp.n = (uint32_t)n;
  ReduceView dims, dv;
  dims.rank = dv.rank = dinv->rank;
  dims.axis = dv.axis = axis;
  for (int d = 0; d < kMaxRank; ++d) {
    dims.shape[d] = sp.shape[d];
    dims.stride[d] = sp.stride[0][d];
    dv.shape[d] = sp.shape[d];
    dv.stride[d] = d < doutv->rank ? doutv->stride[d] : 0;
  }
  launch_k(reduce_grad_reforder_kernel<1>, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, resolve_stream(stream), 
      din + dinv->offset, sp, dims, dout + doutv->offset, dv);
  return after_launch();
}

int weedcu_sum_real(const float *a, const weedcu_view *av, float scale, float *out, void *stream) {
  if (!a || !av || !out) return WEEDCU_EINVAL;
  const weedcu_view *views[1] = {av};
  IndexSpace<1> sp;
  if (!build_index_space<1>(views, sp)) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  const float *base = a + av->offset;
  const bool linear = (sp.rank == 1 && sp.stride[0][0] == 1);
  const bool vec = linear && aligned16(base);
  if (sp.rank == 1 && sp.stride[0][0] == 0) { // broadcast scalar: n * value
    sp.stride[0][0] = 0;
  }
  unsigned blocks = grid_for(vec ? (sp.n >> 2) : sp.n, 256, 4);
  if (blocks > 1024) blocks = 1024;
  float *partial = nullptr;
  WCU_CHECK(pool_alloc((void **)&partial, sizeof(float) * blocks, st));
  ProfScope prof(WEEDCU_PROF_REDUCE, st, 4.0 * sp.n);
  launch_k(sum_pass1_kernel<1>, dim3(blocks), dim3(256), 0, st, base, sp, vec, partial);
  int rc = after_launch();
  if (rc == 0) {
    launch_k(sum_pass2_kernel, dim3(1), dim3(1024), 0, st, partial, blocks, scale, out);
    rc = after_launch();
  }
  pool_free(partial, st);
  return rc;
}

int weedcu_argmax_rows(const float *x, uint64_t offset, uint32_t rows, uint32_t V, uint32_t rs,
                       uint32_t vs, int32_t *out, void *stream) {
  if (!x || !out || !rows || !V) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  const uint32_t row_tiles = (rows + 31) / 32;
  uint32_t slices = (2u * (uint32_t)kNumSMs + row_tiles - 1) / row_tiles; // ~2 blocks per SM in total
  const uint32_t max_slices = (V + 255u) / 256u;                          // at least 16 values per thread of a slice
  if (slices > max_slices) slices = max_slices;
  if (slices < 1u) slices = 1u;
  const uint32_t v_per_slice = (V + slices - 1) / slices;
  slices = (V + v_per_slice - 1) / v_per_slice;
  float *ws = nullptr;
  WCU_CHECK(pool_alloc((void **)&ws, 8ull * slices * rows, st));
  uint32_t *wi = reinterpret_cast<uint32_t *>(ws + (size_t)slices * rows);
  launch_k(argmax_rows_kernel<16>, dim3(row_tiles, slices), dim3(32, 16), 0, st, x + offset, rows, V, rs, vs, v_per_slice, ws, wi);
  int rc = after_launch();
  if (rc == 0) {
    launch_k(argmax_merge_kernel, dim3((rows + 255u) / 256u), dim3(256), 0, st, (const float *)ws, (const uint32_t *)wi, rows, slices, out);
    rc = after_launch();
  }
  pool_free(ws, st);
  return rc;
}

int weedcu_extremum_real(int is_min, const float *a, const weedcu_view *av, float *out, void *stream) {
  if (!a || !av || !out) return WEEDCU_EINVAL;
  const weedcu_view *views[1] = {av};
  IndexSpace<1> sp;
  if (!build_index_space<1>(views, sp)) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  const float *base = a + av->offset;
  const bool vec = (sp.rank == 1 && sp.stride[0][0] == 1) && aligned16(base);
  unsigned blocks = grid_for(vec ? (sp.n >> 2) : sp.n, 256, 4);
  if (blocks > 1024) blocks = 1024;
  float *partial = nullptr;
  WCU_CHECK(pool_alloc((void **)&partial, sizeof(float) * blocks, st));
  ProfScope prof(WEEDCU_PROF_REDUCE, st, 4.0 * sp.n);
  if (is_min) launch_k(extremum_pass1_kernel<true>, dim3(blocks), dim3(256), 0, st, base, sp, vec, partial);
  else launch_k(extremum_pass1_kernel<false>, dim3(blocks), dim3(256), 0, st, base, sp, vec, partial);
  int rc = after_launch();
  if (rc == 0) {
    if (is_min) launch_k(extremum_pass2_kernel<true>, dim3(1), dim3(1024), 0, st, (const float *)partial, blocks, out);
    else launch_k(extremum_pass2_kernel<false>, dim3(1), dim3(1024), 0, st, (const float *)partial, blocks, out);
    rc = after_launch();
  }
  pool_free(partial, st);
  return rc;
}

static void fill_reduce_view(ReduceView &v, const weedcu_view *src, int rank, int axis) {
  v.rank = rank;
  v.axis = axis;
  for (int d = 0; d < kMaxRank; ++d) {
    v.shape[d] = d < rank ? src->shape[d] : 1;
    v.stride[d] = d < rank ? src->stride[d] : 0;
  }
}

int weedcu_extremum_axis_real(int is_min, const float *a, const weedcu_view *av, int axis, float *out, int index_order, void *stream) {
  if (!a || !av || !out || av->rank <= 0 || av->rank > kMaxRank || axis < 0 || axis >= av->rank) return WEEDCU_EINVAL;
  uint64_t total = 1;
  for (int d = 0; d < av->rank; ++d) total *= av->shape[d];
  const uint32_t L = av->shape[axis];
  if (!L || !total || total / L > 0xffffffffull) return WEEDCU_EINVAL;
  const uint32_t n_out = (uint32_t)(total / L);
  cudaStream_t st = resolve_stream(stream);
  ReduceView v;
  fill_reduce_view(v, av, av->rank, axis);
  ProfScope prof(WEEDCU_PROF_REDUCE, st, 4.0 * total);
  if (is_min) launch_k(extremum_axis_kernel<true>, dim3((n_out + 255u) / 256u), dim3(256), 0, st, a + av->offset, v, n_out, out, index_order);
  else launch_k(extremum_axis_kernel<false>, dim3((n_out + 255u) / 256u), dim3(256), 0, st, a + av->offset, v, n_out, out, index_order);
  return after_launch();
}

int weedcu_match_grad_real(float *din, const weedcu_view *dinv, const float *in, const weedcu_view *inv, const float *dout,
                           const weedcu_view *doutv, const float *reduced, int axis, int index_order, void *stream) {
  if (!din || !dinv || !in || !inv || !dout || !doutv || !reduced || dinv->rank != doutv->rank || dinv->rank != inv->rank || axis < 0 ||
      axis >= dinv->rank || dinv->rank > kMaxRank)
    return WEEDCU_EINVAL;
  uint64_t n = 1;
  for (int d = 0; d < dinv->rank; ++d) {
    if (inv->shape[d] != dinv->shape[d]) return WEEDCU_EINVAL;
    n *= dinv->shape[d];
  }
  if (!n || n > 0xffffffffull) return WEEDCU_EINVAL;
  cudaStream_t st = resolve_stream(stream);
  ReduceView dv, iv, ov;
  fill_reduce_view(dv, dinv, dinv->rank, axis);
  fill_reduce_view(iv, inv, dinv->rank, axis);
  fill_reduce_view(ov, doutv, dinv->rank, axis);
  ProfScope prof(WEEDCU_PROF_REDUCE, st, 16.0 * n);
  // dout and the reduced values are read through the same (dout) view: they are the gradient and the value of one tensor
  launch_k(match_grad_kernel, dim3((unsigned)((n + 255) / 256)), dim3(256), 0, st, din + dinv->offset, dv, in + inv->offset, iv, dout + doutv->offset,
           reduced + doutv->offset, ov, (uint32_t)n, index_order);
  return after_launch();
}

} // extern "C"
