/*
 * weed_oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE. See weed_oracle.h.
 *
 * Every function cites the reference (vm6502q/weed v0.7.3, /root/reference) lines it restates.
 * Serial float loops in the reference's own operation order.
 */
#include "weed_oracle.h"
#include <math.h>
#include <stdlib.h>
#include <string.h>

/* BaseTensor::get_storage_index, include/tensors/base_tensor.hpp:123-142
 * (is_scalar shortcut :101-113 is subsumed: all (shape-1)*stride == 0 gives `offset`). */
uint64_t wo_storage_index(const wo_view *v, uint64_t i) {
  uint64_t curr = i, stor = v->offset;
  for (int d = 0; d < v->rank && curr; ++d) {
    const uint64_t l = v->shape[d];
    stor += (curr % l) * (uint64_t)v->stride[d];
    curr /= l;
  }
  return stor;
}

/* BaseTensor::get_broadcast_size, base_tensor.hpp:82-94 */
uint64_t wo_broadcast_size(const wo_view *v) {
  if (v->rank <= 0) return 0;
  uint64_t n = 1;
  for (int d = 0; d < v->rank; ++d) n *= v->shape[d];
  return n;
}

static int same_shape(const wo_view *a, const wo_view *b) {
  if (a->rank != b->rank) return 0;
  for (int d = 0; d < a->rank; ++d)
    if (a->shape[d] != b->shape[d]) return 0;
  return 1;
}

/* TypedStorage::FillValue, include/storage/typed_storage.hpp:57-67 */
int wo_fill_real(float *p, uint64_t n, float value) {
  for (uint64_t i = 0; i < n; ++i) p[i] = value;
  return 0;
}

/* ADD_KERNEL / MUL_KERNEL src/ops/commuting.cpp:27-35; sub src/ops/sub.cpp:37-46;
 * div src/ops/div.cpp:37-46.  out.write(i, a[i] op b[i]) over the flat index. */
int wo_binary_real(int op, const float *a, const wo_view *av, const float *b, const wo_view *bv,
                   float *out, const wo_view *ov) {
  if (!same_shape(av, ov) || !same_shape(bv, ov)) return -1;
  const uint64_t n = wo_broadcast_size(ov);
  for (uint64_t i = 0; i < n; ++i) {
    const float x = a[wo_storage_index(av, i)], y = b[wo_storage_index(bv, i)];
    float r;
    switch (op) {
    case 0: r = x + y; break;
    case 1: r = x * y; break;
    case 2: r = x - y; break;
    case 3: r = x / y; break;
    default: return -1;
    }
    out[wo_storage_index(ov, i)] = r;
  }
  return 0;
}

/* ADD_KERNEL / SUB_KERNEL src/ops/in_place.cpp:27-35: a.add(i, +-b[i]), n = a.get_broadcast_size() */
int wo_inplace_real(int op, float *a, const wo_view *av, const float *b, const wo_view *bv) {
  if (!same_shape(av, bv)) return -1;
  const uint64_t n = wo_broadcast_size(av);
  for (uint64_t i = 0; i < n; ++i) {
    const float y = b[wo_storage_index(bv, i)];
    float *p = &a[wo_storage_index(av, i)];
    if (op == 0) *p = *p + y;
    else if (op == 2) *p = *p + (-y);
    else return -1;
  }
  return 0;
}

/* COPY_KERNEL src/ops/copy_broadcast.cpp:27-30 */
int wo_copy_real(float *dst, const wo_view *dv, const float *src, const wo_view *sv) {
  if (!same_shape(dv, sv)) return -1;
  const uint64_t n = wo_broadcast_size(dv);
  for (uint64_t i = 0; i < n; ++i) dst[wo_storage_index(dv, i)] = src[wo_storage_index(sv, i)];
  return 0;
}

static float gelu_f(float x) {
  /* Tensor::gelu, src/tensors/tensor.cpp:841-851:
   * x3 = x*x*x; inner = k2*(x + k1*x3); t = tanh(inner); k0*x*(1 + t) */
  const float k0 = 0.5f, k1 = 0.044715f, k2 = 0.7978845608028654f;
  const float x3 = (x * x) * x;
  const float inner = k2 * (x + k1 * x3);
  const float t = tanhf(inner);
  return (k0 * x) * (1.0f + t);
}

/* relu/sigmoid/tanh src/ops/real_unary.cpp:77-83,132-138,188-194; sin/cos :243-249,296-302;
 * abs src/ops/abs.cpp:88-94; pow/exp/log src/ops/pow.cpp:52-77 (param = p | log b | 1/log b). */
int wo_unary_real(int op, float param, const float *a, const wo_view *av, float *out,
                  const wo_view *ov) {
  if (!same_shape(av, ov)) return -1;
  const uint64_t n = wo_broadcast_size(ov);
  for (uint64_t i = 0; i < n; ++i) {
    const float x = a[wo_storage_index(av, i)];
    float r;
    switch (op) {
    case 0: r = (x > 0.0f) ? x : 0.0f; break;
    case 1: r = 1.0f / (1.0f + expf(-x)); break;
    case 2: r = tanhf(x); break;
    case 3: r = (x < 0.0f) ? -x : x; break;
    case 4: r = powf(x, param); break;
    case 5: r = expf(x * param); break;
    case 6: r = logf(x) * param; break;
    case 7: r = gelu_f(x); break;
    case 8: r = sinf(x); break;
    case 9: r = cosf(x); break;
    default: return -1;
    }
    out[wo_storage_index(ov, i)] = r;
  }
  return 0;
}

/* CPU_RELU_GRAD / CPU_SIGMOID_GRAD / CPU_TANH_GRAD / CPU_SIN_GRAD / CPU_COS_GRAD
 * src/ops/real_unary.cpp:47-76; REAL_ABS_GRAD_KERNEL src/ops/abs.cpp:70-77.
 * gelu: analytic derivative of the tanh-approximation (the reference back-propagates through the
 * nine composite ops of tensor.cpp:841-851, which is the same function). */
int wo_unary_grad_real(int op, float *din, const wo_view *dinv, const float *in, const wo_view *inv,
                       const float *dout, const wo_view *doutv, int accumulate) {
  if (!same_shape(dinv, inv) || !same_shape(dinv, doutv)) return -1;
  const uint64_t n = wo_broadcast_size(dinv);
  for (uint64_t i = 0; i < n; ++i) {
    const float v = in[wo_storage_index(inv, i)];
    const float g = dout[wo_storage_index(doutv, i)];
    float *p = &din[wo_storage_index(dinv, i)];
    if (!accumulate) *p = 0.0f; /* store variant: din is not read */
    switch (op) {
    case 0: if (v > 0.0f) *p += g; break;
    case 1: *p += v * (1.0f - v) * g; break;
    case 2: *p += g * (1.0f - v * v); break;
    case 3: if (v != 0.0f) *p += (v > 0.0f) ? g : -g; break;
    case 7: {
      const float k1 = 0.044715f, k2 = 0.7978845608028654f;
      const float inner = k2 * (v + k1 * ((v * v) * v));
      const float t = tanhf(inner);
      const float dinner = k2 * (1.0f + 3.0f * k1 * (v * v));
      *p += g * (0.5f * (1.0f + t) + (0.5f * v) * ((1.0f - t * t) * dinner));
      break;
    }
    case 8: *p += cosf(v) * g; break;
    case 9: *p += -sinf(v) * g; break;
    default: return -1;
    }
  }
  return 0;
}

/* REDUCE_HEAD + SUM_LOOP, src/ops/reduce.cpp:17-38,60-66.
 * index_order 1 = verbatim (o decomposed last-dim-fastest); 0 = column-major (intended). */
int wo_reduce_real(const float *a, const wo_view *av, int axis, float *out, int index_order) {
  if (axis < 0 || axis >= av->rank) return -1;
  const uint64_t n = wo_broadcast_size(av) / av->shape[axis];
  for (uint64_t o = 0; o < n; ++o) {
    uint64_t base = 0, tmp = o;
    if (index_order) {
      for (int d = av->rank - 1; d >= 0; --d) {
        if (d == axis) continue;
        const uint64_t dim = av->shape[d];
        base += (tmp % dim) * av->stride[d];
        tmp /= dim;
      }
    } else {
      for (int d = 0; d < av->rank; ++d) {
        if (d == axis) continue;
        const uint64_t dim = av->shape[d];
        base += (tmp % dim) * av->stride[d];
        tmp /= dim;
      }
    }
    float sum = 0.0f;
    for (uint32_t j = 0; j < av->shape[axis]; ++j)
      sum += a[av->offset + base + (uint64_t)j * av->stride[axis]];
    out[o] = sum;
  }
  return 0;
}

/* REDUCE_GRAD_HEAD + SUM_GRAD_OUT, src/ops/reduce.cpp:84-101: din.add(i, dout[o]).
 * index_order 1 = verbatim: i decomposed last-dim-fastest over the non-axis dims only.
 * index_order 0 = intended: din[i] += dout[i] with dout broadcast (stride 0) along axis. */
int wo_reduce_grad_real(float *din, const wo_view *dinv, const float *dout, const wo_view *doutv,
                        int axis, int index_order) {
  if (axis < 0 || axis >= dinv->rank || !same_shape(dinv, doutv)) return -1;
  const uint64_t n = wo_broadcast_size(dinv);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t o = 0, tmp = i;
    if (index_order) {
      for (int d = dinv->rank - 1; d >= 0; --d) {
        if (d == axis) continue;
        const uint64_t dim = dinv->shape[d];
        o += (tmp % dim) * doutv->stride[d];
        tmp /= dim;
      }
    } else {
      for (int d = 0; d < dinv->rank; ++d) {
        const uint64_t dim = dinv->shape[d];
        if (d != axis) o += (tmp % dim) * doutv->stride[d];
        tmp /= dim;
      }
    }
    din[wo_storage_index(dinv, i)] += dout[doutv->offset + o];
  }
  return 0;
}

/* cpu() src/ops/clamp.cpp:67-73: out.write(i, min(max(a[i], l), h)) */
int wo_clamp_real(const float *a, const wo_view *av, float lo, float hi, float *out, const wo_view *ov) {
  if (!same_shape(av, ov)) return -1;
  const uint64_t n = wo_broadcast_size(ov);
  for (uint64_t i = 0; i < n; ++i) {
    float x = a[wo_storage_index(av, i)];
    x = (x < lo) ? lo : x; /* std::max(a, l) */
    x = (hi < x) ? hi : x; /* std::min(., h) */
    out[wo_storage_index(ov, i)] = x;
  }
  return 0;
}
/* CPU_GRAD_KERNEL src/ops/clamp.cpp:55-60: if (l < in[i] && in[i] < h) din.add(i, dout[i]) */
int wo_clamp_grad_real(float *din, const wo_view *dinv, const float *in, const wo_view *inv, const float *dout, const wo_view *doutv,
                       float lo, float hi) {
  if (!same_shape(dinv, inv) || !same_shape(dinv, doutv)) return -1;
  const uint64_t n = wo_broadcast_size(dinv);
  for (uint64_t i = 0; i < n; ++i) {
    const float x = in[wo_storage_index(inv, i)];
    if (x > lo && x < hi) din[wo_storage_index(dinv, i)] += dout[wo_storage_index(doutv, i)];
  }
  return 0;
}
/* CPU_MAX / CPU_MIN src/ops/real_extremum.cpp:62-86 (serial par_for: one running extremum seeded with a[0]) */
int wo_extremum_real(int is_min, const float *a, const wo_view *av, float *out) {
  const uint64_t n = wo_broadcast_size(av);
  if (!n) return -1;
  float m = a[wo_storage_index(av, 0)];
  for (uint64_t i = 1; i < n; ++i) {
    const float v = a[wo_storage_index(av, i)];
    if (is_min ? (v < m) : (v > m)) m = v;
  }
  *out = m;
  return 0;
}
/* CPU_GRAD src/ops/real_extremum.cpp:50-57: if (in[i] == m) din.add(i, dout[i]) */
int wo_match_grad_full_real(float *din, const wo_view *dinv, const float *in, const wo_view *inv, const float *dout, const wo_view *doutv,
                            const float *extremum) {
  if (!same_shape(dinv, inv) || !same_shape(dinv, doutv)) return -1;
  const uint64_t n = wo_broadcast_size(dinv);
  const float m = *extremum;
  for (uint64_t i = 0; i < n; ++i)
    if (in[wo_storage_index(inv, i)] == m) din[wo_storage_index(dinv, i)] += dout[wo_storage_index(doutv, i)];
  return 0;
}
/* REDUCE_HEAD + MAX_LOOP / MIN_LOOP src/ops/reduce.cpp:17-31,40-58 (index_order as wo_reduce_real) */
int wo_extremum_axis_real(int is_min, const float *a, const wo_view *av, int axis, float *out, int index_order) {
  if (axis < 0 || axis >= av->rank) return -1;
  const uint64_t n = wo_broadcast_size(av) / av->shape[axis];
  for (uint64_t o = 0; o < n; ++o) {
    uint64_t base = 0, tmp = o;
    if (index_order) {
      for (int d = av->rank - 1; d >= 0; --d) {
        if (d == axis) continue;
        base += (tmp % av->shape[d]) * av->stride[d];
        tmp /= av->shape[d];
      }
    } else {
      for (int d = 0; d < av->rank; ++d) {
        if (d == axis) continue;
        base += (tmp % av->shape[d]) * av->stride[d];
        tmp /= av->shape[d];
      }
    }
    float m = a[av->offset + base];
    for (uint32_t j = 1; j < av->shape[axis]; ++j) {
      const float v = a[av->offset + base + (uint64_t)j * av->stride[axis]];
      if (is_min ? (v < m) : (v > m)) m = v;
    }
    out[o] = m;
  }
  return 0;
}
/* REDUCE_GRAD_HEAD + MATCH_GRAD_OUT src/ops/reduce.cpp:84-113: if (in[i] == out[o]) din.add(i, dout[o]) with the index
 * decompositions of wo_reduce_grad_real; out and dout are read through the same view (value and gradient of one tensor). */
int wo_match_grad_real(float *din, const wo_view *dinv, const float *in, const wo_view *inv, const float *dout, const wo_view *doutv,
                       const float *reduced, int axis, int index_order) {
  if (axis < 0 || axis >= dinv->rank || !same_shape(dinv, doutv) || !same_shape(dinv, inv)) return -1;
  const uint64_t n = wo_broadcast_size(dinv);
  for (uint64_t i = 0; i < n; ++i) {
    uint64_t o = 0, tmp = i;
    if (index_order) {
      for (int d = dinv->rank - 1; d >= 0; --d) {
        if (d == axis) continue;
        o += (tmp % dinv->shape[d]) * doutv->stride[d];
        tmp /= dinv->shape[d];
      }
    } else {
      for (int d = 0; d < dinv->rank; ++d) {
        if (d != axis) o += (tmp % dinv->shape[d]) * doutv->stride[d];
        tmp /= dinv->shape[d];
      }
    }
    if (in[wo_storage_index(inv, i)] == reduced[doutv->offset + o]) din[wo_storage_index(dinv, i)] += dout[doutv->offset + o];
  }
  return 0;
}

/* CPU_KERNEL src/ops/sum.cpp:27-38 (serial branch of par_for, src/common/parallel_for.cpp:94-106);
 * mean = sum / n (sum.cpp:82-86) is expressed by scale = 1/n applied as a division-equivalent
 * multiply only in the test tolerance; callers wanting the exact quotient pass scale = 1. */
int wo_sum_real(const float *a, const wo_view *av, float scale, float *out) {
  const uint64_t n = wo_broadcast_size(av);
  float t = 0.0f;
  for (uint64_t i = 0; i < n; ++i) t += a[wo_storage_index(av, i)];
  *out = t * scale;
  return 0;
}

static uint64_t row_base(const wo_view *v, int axis, uint64_t o) {
  /* SOFTMAX_HEAD src/ops/softmax.cpp:20-43 — row enumeration order does not affect results */
  uint64_t base = v->offset, tmp = o;
  for (int d = v->rank - 1; d >= 0; --d) {
    if (d == axis) continue;
    const uint64_t dim = v->shape[d];
    base += (tmp % dim) * v->stride[d];
    tmp /= dim;
  }
  return base;
}

/* SOFTMAX_FWD_LOOP src/ops/softmax.cpp:85-108; LOGSOFTMAX_FWD_LOOP src/ops/logsoftmax.cpp:87-111 */
int wo_softmax_real(int log_mode, const float *a, const wo_view *av, int axis, float *out,
                    const wo_view *ov) {
  if (axis < 0 || axis >= av->rank || !same_shape(av, ov)) return -1;
  const uint32_t L = av->shape[axis];
  const uint64_t n = wo_broadcast_size(av) / L, as = av->stride[axis], os = ov->stride[axis];
  for (uint64_t o = 0; o < n; ++o) {
    const uint64_t base = row_base(av, axis, o), obase = row_base(ov, axis, o);
    float mx = a[base];
    for (uint32_t j = 1; j < L; ++j) {
      const float v = a[base + j * as];
      if (v > mx) mx = v;
    }
    float s = 0.0f;
    for (uint32_t j = 0; j < L; ++j) s += expf(a[base + j * as] - mx);
    if (log_mode) {
      const float log_s = logf(s);
      for (uint32_t j = 0; j < L; ++j) out[obase + j * os] = (a[base + j * as] - mx) - log_s;
    } else {
      for (uint32_t j = 0; j < L; ++j) out[obase + j * os] = expf(a[base + j * as] - mx) / s;
    }
  }
  return 0;
}

/* SOFTMAX_BWD_LOOP src/ops/softmax.cpp:110-127; LOGSOFTMAX_BWD_LOOP src/ops/logsoftmax.cpp:119-137 */
int wo_softmax_grad_real(int log_mode, float *din, const wo_view *dinv, const float *out,
                         const wo_view *ov, const float *dout, const wo_view *doutv, int axis) {
  if (axis < 0 || axis >= dinv->rank || !same_shape(dinv, ov) || !same_shape(dinv, doutv))
    return -1;
  const uint32_t L = dinv->shape[axis];
  const uint64_t n = wo_broadcast_size(dinv) / L;
  const uint64_t is = dinv->stride[axis], os = ov->stride[axis], ds = doutv->stride[axis];
  for (uint64_t o = 0; o < n; ++o) {
    const uint64_t ib = row_base(dinv, axis, o), ob = row_base(ov, axis, o),
                   db = row_base(doutv, axis, o);
    if (log_mode) {
      float sum_dout = 0.0f;
      for (uint32_t j = 0; j < L; ++j) sum_dout += dout[db + j * ds];
      for (uint32_t j = 0; j < L; ++j)
        din[ib + j * is] += dout[db + j * ds] - expf(out[ob + j * os]) * sum_dout;
    } else {
      float dot = 0.0f;
      for (uint32_t j = 0; j < L; ++j) dot += out[ob + j * os] * dout[db + j * ds];
      for (uint32_t j = 0; j < L; ++j)
        din[ib + j * is] += out[ob + j * os] * (dout[db + j * ds] - dot);
    }
  }
  return 0;
}

/* MultiHeadAttention::forward, src/modules/multihead_attention.cpp:319-334:
 * scores / sqrt(hd); + mask where triu_fill(mask_val, diagonal=1) (src/ops/triu_fill.cpp:41-59:
 * filled where i + 1 <= j); softmax over the last axis. scores[batch,Tq,Tk], batch fastest. */
int wo_attn_softmax_real(const float *scores, float *out, uint32_t batch, uint32_t Tq, uint32_t Tk,
                         float divisor, float mask_val, int causal, int batch_fastest) {
  float *row = (float *)malloc(sizeof(float) * Tk);
  if (!row) return -1;
  for (uint32_t q = 0; q < Tq; ++q)
    for (uint32_t b = 0; b < batch; ++b) {
      /* element (b, q, k): batch-fastest b + batch*(q + Tq*k); per-batch q + Tq*k + Tq*Tk*b */
      const uint64_t base = batch_fastest ? ((uint64_t)b + (uint64_t)batch * q)
                                          : ((uint64_t)q + (uint64_t)Tq * Tk * b);
      const uint64_t st = batch_fastest ? (uint64_t)batch * Tq : (uint64_t)Tq;
      for (uint32_t k = 0; k < Tk; ++k) {
        float v = scores[base + k * st] / divisor;
        if (causal && Tq > 1) v = v + (((uint64_t)q + 1 <= k) ? mask_val : 0.0f);
        row[k] = v;
      }
      float mx = row[0];
      for (uint32_t k = 1; k < Tk; ++k)
        if (row[k] > mx) mx = row[k];
      float s = 0.0f;
      for (uint32_t k = 0; k < Tk; ++k) s += expf(row[k] - mx);
      for (uint32_t k = 0; k < Tk; ++k) out[base + k * st] = expf(row[k] - mx) / s;
    }
  free(row);
  return 0;
}

/* cross_entropy_loss, include/autograd/cross_entropy_loss.hpp:21-34:
 * lsm = logsoftmax(logits,-1); gathered[r] = sum_v lsm[r,v]*onehot[r,v]; loss = mean(gathered)*-1.
 * lse[r] = max + log(sum exp) is what the fused device kernel saves for its backward. */
int wo_cross_entropy_fwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                         uint32_t rs, uint32_t vs, const int32_t *targets, float *lse,
                         float *loss) {
  float total = 0.0f;
  for (uint32_t r = 0; r < rows; ++r) {
    const uint64_t base = offset + (uint64_t)r * rs;
    float mx = logits[base];
    for (uint32_t v = 1; v < V; ++v) {
      const float x = logits[base + (uint64_t)v * vs];
      if (x > mx) mx = x;
    }
    float s = 0.0f;
    for (uint32_t v = 0; v < V; ++v) s += expf(logits[base + (uint64_t)v * vs] - mx);
    const float log_s = logf(s);
    if (lse) lse[r] = mx + log_s;
    const int32_t t = targets[r];
    if (t < 0 || (uint32_t)t >= V) return -1;
    total += (logits[base + (uint64_t)t * vs] - mx) - log_s;
  }
  *loss = (total / (float)rows) * -1.0f;
  return 0;
}

/* Backward of the same chain: mul-by(-1) node, mean node (tensor.cpp:596-612: dout/N), axis-sum
 * node, mul-by-onehot node, logsoftmax_grad (logsoftmax.cpp:119-137):
 * dlogits[r,v] += (exp(lsm[r,v]) - onehot[r,v]) * dloss / rows. */
int wo_cross_entropy_bwd(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                         uint32_t rs, uint32_t vs, const int32_t *targets, const float *lse,
                         const float *dloss, float *dlogits, uint64_t d_offset, int accumulate) {
  const float g = dloss[0] / (float)rows;
  for (uint32_t r = 0; r < rows; ++r) {
    const uint64_t base = offset + (uint64_t)r * rs, dbase = d_offset + (uint64_t)r * rs;
    for (uint32_t v = 0; v < V; ++v) {
      const float p = expf(logits[base + (uint64_t)v * vs] - lse[r]);
      const float oh = ((uint32_t)targets[r] == v) ? 1.0f : 0.0f;
      float *o = &dlogits[dbase + (uint64_t)v * vs];
      *o = (accumulate ? *o : 0.0f) + (p - oh) * g;
    }
  }
  return 0;
}

/* LayerNorm::forward, src/modules/layernorm.cpp:29-42, over x[rows,F] (row stride 1, feature
 * stride rows): mean = sum/F (Tensor::mean axis, tensor.cpp:682-693); xc = x - mean;
 * var = sum(xc*xc)/F; y = xc / pow(var+eps, 0.5) (pow.cpp:52-57); y*gamma + beta. */
int wo_layernorm_fwd(const float *x, uint32_t rows, uint32_t F, const float *gamma,
                     const float *beta, float eps, float *y, float *mean, float *rstd) {
  for (uint32_t r = 0; r < rows; ++r) {
    float s = 0.0f;
    for (uint32_t f = 0; f < F; ++f) s += x[r + (uint64_t)f * rows];
    const float mu = s / (float)F;
    float q = 0.0f;
    for (uint32_t f = 0; f < F; ++f) {
      const float xc = x[r + (uint64_t)f * rows] - mu;
      q += xc * xc;
    }
    const float var = q / (float)F;
    const float den = powf(var + eps, 0.5f);
    if (mean) mean[r] = mu;
    if (rstd) rstd[r] = 1.0f / den;
    for (uint32_t f = 0; f < F; ++f) {
      const float xc = x[r + (uint64_t)f * rows] - mu;
      y[r + (uint64_t)f * rows] = (xc / den) * gamma[f] + beta[f];
    }
  }
  return 0;
}

/* Backward of LayerNorm::forward as the reference's autograd chain computes it (grad_mode 0), node
 * by node over layernorm.cpp:29-42:
 *   y = y0*gamma + beta        mul/add nodes (tensor.cpp:1159-1202,1105-1136): g = dy*gamma,
 *                              dgamma += sum_rows dy*y0, dbeta += sum_rows dy
 *   y0 = xc / d                div node (tensor.cpp:1479-1524): dxc += g/d ;  dd -= sum_f xc/d^2
 *                              -- NOTE the denominator branch does not multiply by dout
 *   d = (v+eps)^0.5            pow node (tensor.cpp:1540-1573): dv = 0.5*dd*d/(v+eps) = 0.5*dd/d
 *   v = mean_f(xc*xc)          mean(axis) + mul nodes: dxc += 2*xc*dv/F
 *   xc = x - mean_f(x)         sub + mean(axis) nodes: dx += dxc - mean_f(dxc)
 * grad_mode 1: the analytic gradient rstd*(g - mean_f(g) - xhat*mean_f(g*xhat)).
 * Accumulation in double: the reference's float chain is pinned against this within 1e-5 by
 * tests/test_oracle_cpu.py (live comparison with the compiled reference). */
int wo_layernorm_bwd(const float *x, const float *dy, uint32_t rows, uint32_t F, const float *gamma,
                     const float *mean, const float *rstd, float *dx, float *dgamma, float *dbeta,
                     int grad_mode, int accumulate) {
  if (!accumulate) /* store variant: dx is not read */
    for (uint64_t i = 0; i < (uint64_t)rows * F; ++i) dx[i] = 0.0f;
  for (uint32_t r = 0; r < rows; ++r) {
    const double rs = rstd[r], mu = mean[r];
    double sg = 0.0, sgx = 0.0, sx = 0.0;
    for (uint32_t f = 0; f < F; ++f) {
      const uint64_t i = r + (uint64_t)f * rows;
      const double xc = (double)x[i] - mu;
      const double g = (double)dy[i] * gamma[f];
      sg += g;
      sgx += g * (xc * rs);
      sx += xc;
    }
    if (grad_mode) {
      const double mg = sg / F, mgx = sgx / F;
      for (uint32_t f = 0; f < F; ++f) {
        const uint64_t i = r + (uint64_t)f * rows;
        const double xh = ((double)x[i] - mu) * rs;
        const double g = (double)dy[i] * gamma[f];
        dx[i] += (float)(rs * (g - mg - xh * mgx));
      }
    } else {
      const double dd = -(rs * rs) * sx;      /* -sum_f xc/d^2 */
      const double c = dd * rs / F;           /* 2 * (0.5*dd/d) / F */
      const double mean_dxc = (rs * sg + c * sx) / F;
      for (uint32_t f = 0; f < F; ++f) {
        const uint64_t i = r + (uint64_t)f * rows;
        const double xc = (double)x[i] - mu;
        const double g = (double)dy[i] * gamma[f];
        dx[i] += (float)(g * rs + c * xc - mean_dxc);
      }
    }
  }
  for (uint32_t f = 0; f < F; ++f) {
    double a = 0.0, b = 0.0;
    for (uint32_t r = 0; r < rows; ++r) {
      const uint64_t i = r + (uint64_t)f * rows;
      a += (double)dy[i] * (((double)x[i] - mean[r]) * rstd[r]);
      b += dy[i];
    }
    if (dgamma) dgamma[f] += (float)a;
    if (dbeta) dbeta[f] += (float)b;
  }
  return 0;
}

/* cpu_forward src/ops/embedding.cpp:56-82 */
int wo_embedding_gather(const int32_t *idx, uint64_t idx_off, uint32_t idx_stride, uint32_t n,
                        const float *W, uint64_t w_off, uint32_t w_s0, uint32_t w_s1, uint32_t D,
                        float *out, uint64_t o_off, uint32_t o_s0, uint32_t o_s1) {
  for (uint32_t i = 0; i < n; ++i) {
    const uint64_t token = (uint32_t)idx[idx_off + (uint64_t)i * idx_stride];
    const uint64_t w_base = w_off + token * w_s0, o_base = o_off + (uint64_t)i * o_s0;
    for (uint32_t d = 0; d < D; ++d) out[o_base + (uint64_t)d * o_s1] = W[w_base + (uint64_t)d * w_s1];
  }
  return 0;
}

/* cpu_backward src/ops/embedding.cpp:84-110 */
int wo_embedding_scatter_add(float *dW, uint64_t w_off, uint32_t w_s0, uint32_t w_s1,
                             const int32_t *idx, uint64_t idx_off, uint32_t idx_stride, uint32_t n,
                             uint32_t D, const float *dout, uint64_t o_off, uint32_t o_s0,
                             uint32_t o_s1) {
  for (uint32_t i = 0; i < n; ++i) {
    const uint64_t token = (uint32_t)idx[idx_off + (uint64_t)i * idx_stride];
    const uint64_t w_base = w_off + token * w_s0, o_base = o_off + (uint64_t)i * o_s0;
    for (uint32_t d = 0; d < D; ++d)
      dW[w_base + (uint64_t)d * w_s1] += dout[o_base + (uint64_t)d * o_s1];
  }
  return 0;
}

/* cpu_triu_fill src/ops/triu_fill.cpp:41-59 */
int wo_triu_fill_real(float *a, const wo_view *av, float val, uint32_t diagonal) {
  if (av->rank != 2) return -1;
  const uint64_t t0 = av->shape[0], t1 = av->shape[1];
  for (uint64_t idx = 0; idx < t0 * t1; ++idx) {
    const uint64_t i = idx % t0, j = idx / t0;
    if (i + diagonal <= j) a[wo_storage_index(av, idx)] = val;
  }
  return 0;
}

int wo_argmax_rows(const float *x, uint64_t offset, uint32_t rows, uint32_t V, uint32_t rs,
                   uint32_t vs, int32_t *out) {
  for (uint32_t r = 0; r < rows; ++r) {
    const uint64_t base = offset + (uint64_t)r * rs;
    float m = x[base];
    int32_t bi = 0;
    for (uint32_t v = 1; v < V; ++v) {
      const float y = x[base + (uint64_t)v * vs];
      if (y > m) { m = y; bi = (int32_t)v; }
    }
    out[r] = bi;
  }
  return 0;
}

/* sgd_step include/autograd/sgd.hpp:23-37: tmp = lr*g; p -= tmp */
int wo_sgd_step(float *p, const float *g, uint64_t n, float lr, float gscale) {
  for (uint64_t i = 0; i < n; ++i) {
    const float gi = (gscale == 1.0f) ? g[i] : gscale * g[i];
    p[i] = p[i] + (-(lr * gi));
  }
  return 0;
}

/* adam_step include/autograd/adam.hpp:70-106, in its operation order:
 * m = b1*m + (1-b1)*g; v = b2*v + ((1-b2)*g)*g;
 * tmp = (lr*m) / (bc1 * (pow(v/bc2, 0.5) + eps)); p -= tmp */
int wo_adam_step(float *p, const float *g, float *m, float *v, uint64_t n, float lr, float beta1,
                 float beta2, float eps, float bc1, float bc2, float gscale) {
  const float omb1 = 1.0f - beta1, omb2 = 1.0f - beta2;
  for (uint64_t i = 0; i < n; ++i) {
    const float gi = (gscale == 1.0f) ? g[i] : gscale * g[i];
    const float mi = beta1 * m[i] + omb1 * gi;
    const float vi = beta2 * v[i] + (omb2 * gi) * gi;
    m[i] = mi;
    v[i] = vi;
    const float tmp = (lr * mi) / (bc1 * (powf(vi / bc2, 0.5f) + eps));
    p[i] = p[i] + (-tmp);
  }
  return 0;
}

/* CPU_BY_TYPE src/ops/matmul.cpp:34-47: for each (i,j): sum over k of a[i,k]*b[k,j], serial float.
 * batch > 1 = the host loop of Tensor::matmul, src/tensors/tensor.cpp:1259-1269.
 * accumulate = tmp + add_in_place of make_matmul_node (tensor.cpp:1372,1383): c += sum. */
int wo_matmul_real(const float *a, const wo_mat *am, const float *b, const wo_mat *bm, float *c,
                   const wo_mat *cm, uint32_t M, uint32_t K, uint32_t N, uint32_t batch,
                   int accumulate) {
  for (uint32_t z = 0; z < batch; ++z) {
    const uint64_t ao = am->offset + z * am->batch_stride, bo = bm->offset + z * bm->batch_stride,
                   co = cm->offset + z * cm->batch_stride;
    for (uint32_t i = 0; i < M; ++i)
      for (uint32_t j = 0; j < N; ++j) {
        float sum = 0.0f;
        for (uint32_t k = 0; k < K; ++k)
          sum += a[ao + (uint64_t)i * am->s0 + (uint64_t)k * am->s1] *
                 b[bo + (uint64_t)k * bm->s0 + (uint64_t)j * bm->s1];
        float *o = &c[co + (uint64_t)i * cm->s0 + (uint64_t)j * cm->s1];
        *o = accumulate ? (*o + sum) : sum;
      }
  }
  return 0;
}

uint16_t wo_f32_to_bf16(float f) {
  uint32_t u;
  memcpy(&u, &f, 4);
  if ((u & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((u >> 16) | 0x40u); /* NaN */
  u += 0x7fffu + ((u >> 16) & 1u);
  return (uint16_t)(u >> 16);
}

static float bf16_round(float f) {
  uint32_t u = (uint32_t)wo_f32_to_bf16(f) << 16;
  float r;
  memcpy(&r, &u, 4);
  return r;
}

int wo_matmul_bf16_model(const float *a, const wo_mat *am, const float *b, const wo_mat *bm,
                         float *c, const wo_mat *cm, uint32_t M, uint32_t K, uint32_t N,
                         uint32_t batch, int accumulate) {
  for (uint32_t z = 0; z < batch; ++z) {
    const uint64_t ao = am->offset + z * am->batch_stride, bo = bm->offset + z * bm->batch_stride,
                   co = cm->offset + z * cm->batch_stride;
    for (uint32_t i = 0; i < M; ++i)
      for (uint32_t j = 0; j < N; ++j) {
        double sum = 0.0;
        for (uint32_t k = 0; k < K; ++k)
          sum += (double)bf16_round(a[ao + (uint64_t)i * am->s0 + (uint64_t)k * am->s1]) *
                 (double)bf16_round(b[bo + (uint64_t)k * bm->s0 + (uint64_t)j * bm->s1]);
        float *o = &c[co + (uint64_t)i * cm->s0 + (uint64_t)j * cm->s1];
        *o = accumulate ? (float)(*o + sum) : (float)sum;
      }
  }
  return 0;
}

/* Model of the fused bf16 attention core (weedcu_attention_fwd): the same chain as
 * MultiHeadAttention::forward, src/modules/multihead_attention.cpp:289-345 — per (b, h):
 * S = Q K^T, S/divisor (+ triu mask), softmax over keys, O = P V — with the rounding points of the
 * tensor-core path: Q, K, V and the probabilities P are rounded to bf16 (RNE), products are exact
 * and sums are carried in double (the device accumulates in fp32; tolerance covers the difference).
 * q, k, v, out: [B, T, H*hd] column-major (b fastest); feature c = h + H*j, because the reference
 * reshapes to {B, T, H, hd} (:155-157) and a column-major reshape makes the first new extent fast. */
int wo_attention_fwd(const float *q, const float *k, const float *v, float *out, uint32_t B,
                           uint32_t T, uint32_t H, uint32_t hd, float divisor, float mask_val,
                           int causal) {
  float *S = (float *)malloc(sizeof(float) * (size_t)T * T);
  float *P = (float *)malloc(sizeof(float) * (size_t)T * T);
  if (!S || !P) {
    free(S);
    free(P);
    return -1;
  }
  const uint64_t sT = B, sC = (uint64_t)B * T; /* strides of the token and the feature index */
  for (uint32_t b = 0; b < B; ++b)
    for (uint32_t h = 0; h < H; ++h) {
      for (uint32_t i = 0; i < T; ++i)
        for (uint32_t j = 0; j < T; ++j) {
          double sum = 0.0;
          for (uint32_t c = 0; c < hd; ++c)
            sum += (double)bf16_round(q[b + i * sT + (uint64_t)(h + H * c) * sC]) *
                   (double)bf16_round(k[b + j * sT + (uint64_t)(h + H * c) * sC]);
          S[i + (size_t)T * j] = (float)sum;
        }
      if (wo_attn_softmax_real(S, P, 1, T, T, divisor, mask_val, causal, 0) != 0) {
        free(S);
        free(P);
        return -1;
      }
      for (uint32_t i = 0; i < T; ++i)
        for (uint32_t c = 0; c < hd; ++c) {
          double sum = 0.0;
          for (uint32_t j = 0; j < T; ++j)
            sum += (double)bf16_round(P[i + (size_t)T * j]) *
                   (double)bf16_round(v[b + j * sT + (uint64_t)(h + H * c) * sC]);
          out[b + i * sT + (uint64_t)(h + H * c) * sC] = (float)sum;
        }
    }
  free(S);
  free(P);
  return 0;
}

/* MultiHeadAttention::forward with use_kv_cache, kv_quant_bits = 0 — the float cache path,
 * src/modules/multihead_attention.cpp:278-287 (slot add_in_place + slices of the cache), :313-345
 * (scores = Q K^T, / sqrt(hd), [T_q, T_k] triu mask when T > 1, softmax, P V, transpose back).
 * Serial float loops in the reference's order: dot over head_dim (matmul.cpp:34-47), x / divisor,
 * + mask (triu_fill.cpp:48-56, diagonal 1: key index > query index), max / exp / sum / divide
 * (softmax.cpp:23-138), then sum over keys of p * v.
 * q, k, v, out: [B, T_new, H*hd] column-major; caches [B, H, S, hd] column-major. */
int wo_attention_decode(const float *q, const float *k, const float *v, float *k_cache, float *v_cache,
                        float *out, uint32_t B, uint32_t T_new, uint32_t H, uint32_t hd, uint32_t S,
                        uint32_t cache_len, float divisor, float mask_val, int causal) {
  if ((uint64_t)cache_len + T_new > S) return -1;
  const uint64_t BH = (uint64_t)B * H;
  const uint32_t L = cache_len + T_new;
  for (uint32_t b = 0; b < B; ++b)
    for (uint32_t h = 0; h < H; ++h)
      for (uint32_t t = 0; t < T_new; ++t)
        for (uint32_t j = 0; j < hd; ++j) {
          const uint64_t src = b + (uint64_t)B * (t + (uint64_t)T_new * ((uint64_t)h + (uint64_t)H * j));
          const uint64_t dst = (b + (uint64_t)B * h) + BH * ((uint64_t)(cache_len + t) + (uint64_t)S * j);
          k_cache[dst] = k_cache[dst] + k[src];
          v_cache[dst] = v_cache[dst] + v[src];
        }
  float *x = (float *)malloc(sizeof(float) * (size_t)L);
  if (!x) return -1;
  const int do_mask = causal && T_new > 1;
  for (uint32_t b = 0; b < B; ++b)
    for (uint32_t h = 0; h < H; ++h) {
      const uint64_t bh = b + (uint64_t)B * h;
      for (uint32_t t = 0; t < T_new; ++t) {
        float mx = -INFINITY;
        for (uint32_t s = 0; s < L; ++s) {
          float sum = 0.0f;
          for (uint32_t j = 0; j < hd; ++j)
            sum += q[b + (uint64_t)B * (t + (uint64_t)T_new * ((uint64_t)h + (uint64_t)H * j))] * k_cache[bh + BH * ((uint64_t)s + (uint64_t)S * j)];
          float y = sum / divisor;
          if (do_mask && t + 1u <= s) y = y + mask_val;
          x[s] = y;
          if (y > mx) mx = y;
        }
        float den = 0.0f;
        for (uint32_t s = 0; s < L; ++s) {
          x[s] = expf(x[s] - mx);
          den += x[s];
        }
        for (uint32_t j = 0; j < hd; ++j) {
          float sum = 0.0f;
          for (uint32_t s = 0; s < L; ++s) sum += (x[s] / den) * v_cache[bh + BH * ((uint64_t)s + (uint64_t)S * j)];
          out[b + (uint64_t)B * (t + (uint64_t)T_new * ((uint64_t)h + (uint64_t)H * j))] = sum;
        }
      }
    }
  free(x);
  return 0;
}

/* Linear::forward on a few rows (src/modules/linear.cpp:86-100: y = x >> W; y = y + bias) = the
 * serial-float matmul of wo_matmul_real followed by the broadcast add. */
int wo_matmul_skinny(const float *a, const wo_mat *am, const float *b, const wo_mat *bm, float *c,
                     const wo_mat *cm, uint32_t M, uint32_t K, uint32_t N, const float *bias, int accumulate) {
  for (uint32_t i = 0; i < M; ++i)
    for (uint32_t j = 0; j < N; ++j) {
      float sum = 0.0f;
      for (uint32_t k = 0; k < K; ++k)
        sum += a[am->offset + (uint64_t)i * am->s0 + (uint64_t)k * am->s1] * b[bm->offset + (uint64_t)k * bm->s0 + (uint64_t)j * bm->s1];
      if (bias) sum = sum + bias[j];
      float *o = &c[cm->offset + (uint64_t)i * cm->s0 + (uint64_t)j * cm->s1];
      *o = accumulate ? (*o + sum) : sum;
    }
  return 0;
}

/* wo_cross_entropy_bwd followed by what Tensor::matmul_backward's bf16 path and the bias node do with
 * dlogits: the RNE bf16 copy (wo_f32_to_bf16) and the column sums (reduce over rows, reduce.cpp:17-38). */
int wo_cross_entropy_bwd_pack(const float *logits, uint64_t offset, uint32_t rows, uint32_t V,
                              const int32_t *targets, const float *lse, const float *dloss, float *dlogits,
                              uint64_t d_offset, int accumulate, uint16_t *dlogits_bf16, float *colsum) {
  const int rc = wo_cross_entropy_bwd(logits, offset, rows, V, 1, rows, targets, lse, dloss, dlogits, d_offset, accumulate);
  if (rc) return rc;
  for (uint32_t v = 0; v < V; ++v) {
    float s = 0.0f;
    for (uint32_t r = 0; r < rows; ++r) {
      const float d = dlogits[d_offset + r + (uint64_t)v * rows];
      dlogits_bf16[r + (uint64_t)v * rows] = wo_f32_to_bf16(d);
      s += d;
    }
    colsum[v] = s;
  }
  return 0;
}

/* The bf16-emitting forms of LayerNorm::forward and Tensor::gelu: the fp32 result of the plain
 * restatement plus its RNE bf16 copy at the same linear index (wo_f32_to_bf16). */
int wo_layernorm_fwd_bf16(const float *x, uint32_t rows, uint32_t F, const float *gamma, const float *beta,
                          float eps, float *y, float *mean, float *rstd, uint16_t *y_bf16) {
  const int rc = wo_layernorm_fwd(x, rows, F, gamma, beta, eps, y, mean, rstd);
  if (rc == 0 && y_bf16)
    for (uint64_t i = 0; i < (uint64_t)rows * F; ++i) y_bf16[i] = wo_f32_to_bf16(y[i]);
  return rc;
}
int wo_gelu_fwd_bf16(const float *x, float *y, uint16_t *y_bf16, uint64_t n) {
  wo_view v;
  memset(&v, 0, sizeof(v));
  v.rank = 1;
  v.shape[0] = (uint32_t)n;
  v.stride[0] = 1;
  const int rc = wo_unary_real(7 /* GELU */, 0.0f, x, &v, y, &v);
  if (rc == 0)
    for (uint64_t i = 0; i < n; ++i) y_bf16[i] = wo_f32_to_bf16(y[i]);
  return rc;
}

/* gelu_grad followed by the bf16 copy and the column sums the next Linear backward takes of din. */
int wo_gelu_grad_pack(float *din, const float *in, const float *dout, uint32_t rows, uint32_t cols, int accumulate,
                      uint16_t *din_bf16, float *colsum) {
  wo_view v;
  memset(&v, 0, sizeof(v));
  v.rank = 1;
  v.shape[0] = rows * cols;
  v.stride[0] = 1;
  const int rc = wo_unary_grad_real(7 /* GELU */, din, &v, in, &v, dout, &v, accumulate);
  if (rc) return rc;
  for (uint32_t c = 0; c < cols; ++c) {
    float s = 0.0f;
    for (uint32_t r = 0; r < rows; ++r) {
      const float d = din[r + (uint64_t)c * rows];
      din_bf16[r + (uint64_t)c * rows] = wo_f32_to_bf16(d);
      s += d;
    }
    colsum[c] = s;
  }
  return 0;
}
