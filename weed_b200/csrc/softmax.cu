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

namespace weedcu {

constexpr int kRT = 32;                       // rows per tile (one warp wide)
constexpr size_t kMaxTileBytes = 200 * 1024;  // dynamic shared memory budget for a staged tile

struct PlainLoad {
  __device__ float operator()(float x, uint32_t, uint32_t) const { return x; }
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
};

// ----------------------------------------------------------------------------- forward, strided
// a[inner, L, outer] canonical contiguous; rows = inner (per outer slab).
template <bool LOG, int BY, bool STAGED, class LD>
__global__ void __launch_bounds__(kRT * BY)
softmax_strided_fwd(const float *a, float *out, uint32_t inner, uint32_t L,
                    LD ld) {
  extern __shared__ float tile[]; // [L][kRT] when STAGED
  __shared__ float red_m[BY][kRT + 1];
  __shared__ float red_s[BY][kRT + 1];
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint32_t r = blockIdx.x * kRT + tx;
  const bool live = r < inner;
  const uint64_t slab = (uint64_t)blockIdx.y * inner * L;
  const float *p = a + slab + r;
  float *po = out + slab + r;
  const uint32_t grow = r; // row id inside the slab (attention: b + batch*q)

  float mx = -INFINITY, s = 0.0f;
  if (STAGED) {
    if (live)
      for (uint32_t j = ty; j < L; j += BY) {
        const float v = ld(p[(uint64_t)j * inner], grow, j);
        tile[j * kRT + tx] = v;
        mx = fmaxf(mx, v);
      }
    red_m[ty][tx] = mx;
    __syncthreads();
#pragma unroll
    for (int y = 0; y < BY; ++y) mx = fmaxf(mx, red_m[y][tx]);
    if (live)
      for (uint32_t j = ty; j < L; j += BY) {
        const float e = expf(tile[j * kRT + tx] - mx);
        if (!LOG) tile[j * kRT + tx] = e;
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
        for (uint32_t j = ty; j < L; j += BY) po[(uint64_t)j * inner] = (tile[j * kRT + tx] - mx) - log_s;
      } else {
        for (uint32_t j = ty; j < L; j += BY) po[(uint64_t)j * inner] = tile[j * kRT + tx] / s;
      }
    }
  } else {
    // Row too long to stage: one online (max, sum) pass + one normalise pass (12 B/elem).
    if (live)
      for (uint32_t j = ty; j < L; j += BY) {
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
      for (uint32_t j = ty; j < L; j += BY) {
        const float v = ld(p[(uint64_t)j * inner], grow, j);
        po[(uint64_t)j * inner] = LOG ? ((v - M) - log_s) : (expf(v - M) / S);
      }
    }
  }
}

// ----------------------------------------------------------------------------- forward, contiguous
// inner == 1: each row is L contiguous floats; one block per row.
template <bool LOG>
__global__ void __launch_bounds__(256)
softmax_contig_fwd(const float *__restrict__ a, float *__restrict__ out, uint32_t L) {
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
// logits[rows, V], row stride rs, vocab stride vs. One read of the logits (4 B/elem): online
// (max, sum exp) per row, gather of the target logit, lse[r] and nll[r] = lse[r] - x[r,target].
template <int BY>
__global__ void __launch_bounds__(kRT * BY)
ce_fwd_kernel(const float *__restrict__ x, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs,
              const int32_t *__restrict__ targets, float *__restrict__ lse, float *__restrict__ nll) {
  __shared__ float red_m[BY][kRT + 1];
  __shared__ float red_s[BY][kRT + 1];
  const uint32_t tx = threadIdx.x, ty = threadIdx.y;
  const uint32_t r = blockIdx.x * kRT + tx;
  const bool live = r < rows;
  const float *p = x + (uint64_t)r * rs;
  float mx = -INFINITY, s = 0.0f;
  if (live)
    for (uint32_t j = ty; j < V; j += BY) {
      const float v = p[(uint64_t)j * vs];
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
  if (ty == 0 && live) {
    float M = -INFINITY;
#pragma unroll
    for (int y = 0; y < BY; ++y) M = fmaxf(M, red_m[y][tx]);
    float S = 0.0f;
#pragma unroll
    for (int y = 0; y < BY; ++y) {
      const float my = red_m[y][tx];
      if (my > -INFINITY) S += red_s[y][tx] * expf(my - M);
    }
    const float log_s = logf(S);
    const uint32_t t = (uint32_t)targets[r];
    const float xt = (t < V) ? p[(uint64_t)t * vs] : NAN;
    lse[r] = M + log_s;
    nll[r] = (xt - M) - log_s; // = lsm[r, target]; loss = -mean
  }
}
// dlogits[r,v] += (exp(x - lse[r]) - onehot) * dloss/rows ; thread per element, r fastest.
__global__ void __launch_bounds__(256)
ce_bwd_kernel(const float *__restrict__ x, uint32_t rows, uint32_t V, uint32_t rs, uint32_t vs,
              const int32_t *__restrict__ targets, const float *__restrict__ lse,
              const float *__restrict__ dloss, float *dlogits) {
  const uint64_t n = (uint64_t)rows * V;
  const float g = dloss[0] / (float)rows;
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
    const uint32_t r = (uint32_t)(i % rows), v = (uint32_t)(i / rows);
    const uint64_t off = (uint64_t)r * rs + (uint64_t)v * vs;
    const float pr = expf(x[off] - lse[r]);
    const float oh = ((uint32_t)targets[r] == v) ? 1.0f : 0.0f;
    dlogits[off] += (pr - oh) * g;
  }
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

template <bool LOG, class LD>
static int launch_strided_fwd(const float *a, float *out, uint32_t inner, uint32_t L, uint32_t outer,
                              LD ld, cudaStream_t st) {
  const dim3 grid((inner + kRT - 1) / kRT, outer);
  const size_t tile_bytes = (size_t)L * kRT * sizeof(float);
  if (tile_bytes <= kMaxTileBytes) {
    if (L >= 512) {
      auto k = softmax_strided_fwd<LOG, 32, true, LD>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileBytes);
      k<<<grid, dim3(kRT, 32), tile_bytes, st>>>(a, out, inner, L, ld);
    } else {
      auto k = softmax_strided_fwd<LOG, 8, true, LD>;
      cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileBytes);
      k<<<grid, dim3(kRT, 8), tile_bytes, st>>>(a, out, inner, L, ld);
    }
  } else {
    softmax_strided_fwd<LOG, 32, false, LD><<<grid, dim3(kRT, 32), 0, st>>>(a, out, inner, L, ld);
  }
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
      softmax_contig_fwd<LOG><<<(unsigned)outer, 256, 0, st>>>(pa, po, L);
      return after_launch();
    }
    if (outer <= 65535) return launch_strided_fwd<LOG>(pa, po, (uint32_t)inner, L, (uint32_t)outer, PlainLoad(), st);
  }
  RowView ra, ro;
  to_rowview(av, axis, ra);
  to_rowview(ov, axis, ro);
  softmax_generic_fwd<LOG><<<(n_rows + 127) / 128, 128, 0, st>>>(a, ra, out, ro, n_rows);
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
      softmax_contig_bwd<LOG><<<(unsigned)outer, 256, 0, st>>>(pd, py, pg, L);
      return after_launch();
    }
    if (outer <= 65535) {
      const dim3 grid((unsigned)((inner + kRT - 1) / kRT), (unsigned)outer);
      const size_t tile_bytes = 2 * (size_t)L * kRT * sizeof(float);
      if (tile_bytes <= kMaxTileBytes) {
        if (L >= 256) {
          auto k = softmax_strided_bwd<LOG, 32, true>;
          cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileBytes);
          k<<<grid, dim3(kRT, 32), tile_bytes, st>>>(pd, py, pg, (uint32_t)inner, L);
        } else {
          auto k = softmax_strided_bwd<LOG, 8, true>;
          cudaFuncSetAttribute(k, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMaxTileBytes);
          k<<<grid, dim3(kRT, 8), tile_bytes, st>>>(pd, py, pg, (uint32_t)inner, L);
        }
      } else {
        softmax_strided_bwd<LOG, 32, false><<<grid, dim3(kRT, 32), 0, st>>>(pd, py, pg, (uint32_t)inner, L);
      }
      return after_launch();
    }
  }
  RowView ri, ro, rd;
  to_rowview(iv, axis, ri);
  to_rowview(ov, axis, ro);
  to_rowview(dv, axis, rd);
  softmax_generic_bwd<LOG><<<(n_rows + 127) / 128, 128, 0, st>>>(din, ri, out, ro, dout, rd, n_rows);
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
  float *nll = nullptr;
  WCU_CHECK(cudaMallocAsync((void **)&nll, sizeof(float) * rows, st));
  ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, st, 4.0 * (double)rows * V);
  ce_fwd_kernel<32><<<(rows + kRT - 1) / kRT, dim3(kRT, 32), 0, st>>>(logits + offset, rows, V, rs, vs,
                                                                    targets, lse, nll);
  int rc = after_launch();
  if (rc == 0) {
    weedcu_view v;
    v.offset = 0;
    v.rank = 1;
    v.shape[0] = rows;
    v.stride[0] = 1;
    rc = weedcu_sum_real(nll, &v, -1.0f / (float)rows, loss, (void *)st);
  }
  cudaFreeAsync(nll, st);
  return rc;
}

int weedcu_cross_entropy_bwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                             uint32_t rs, uint32_t vs, const int32_t *targets, const float *lse,
                             const float *dloss, float *dlogits, uint64_t d_offset, void *stream) {
  if (!logits || !targets || !lse || !dloss || !dlogits || !rows || !V) return WEEDCU_EINVAL;
  const uint64_t n = (uint64_t)rows * V;
  ProfScope prof(WEEDCU_PROF_CROSS_ENTROPY, resolve_stream(stream), 12.0 * (double)n);
  ce_bwd_kernel<<<grid_for(n, 256, 32), 256, 0, resolve_stream(stream)>>>(
      logits + offset, rows, V, rs, vs, targets, lse, dloss, dlogits + d_offset);
  return after_launch();
}

} // extern "C"
