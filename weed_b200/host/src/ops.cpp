// ops.cpp — Weed:: op entry points on the CUDA device. Validation mirrors the reference (same
// exception types at the same checks: src/ops/commuting.cpp:121-140, in_place.cpp:110-125,
// copy_broadcast.cpp:66-80, reduce.cpp:234-278, sum.cpp:127-135, matmul.cpp:95-122,242-256,
// util.cpp:18-34); the body of each op is one call into include/weedcu.h.
#include "weed_b200/ops.hpp"

#include <cmath>

namespace Weed {

void validate_all_same_device(const std::vector<const BaseTensor *> &t, const std::string cls) {
  if (t.size() < 2U) return;
  const DeviceTag dtag = t[0U]->storage->device;
  for (const auto &x : t)
    if (dtag != x->storage->device) throw std::domain_error(std::string("In ") + cls + std::string(", tensor storage devices do not match!"));
}

namespace {
struct Dev {
  real1 *ptr;
  void *stream;
};
inline GpuRealStorage *gpu_storage(const BaseTensor &t, const char *op) {
  if (!t.storage) throw std::invalid_argument(std::string(op) + ": tensor has no storage");
  if (t.storage->dtype != DType::REAL) throw std::invalid_argument(std::string(op) + ": only real-valued tensors are supported on the CUDA device");
  if (t.storage->device != DeviceTag::GPU)
    throw std::domain_error(std::string(op) + ": this backend implements DeviceTag::GPU only; there is no CPU compute path "
                                              "(move the tensor with cast(DeviceTag::GPU))");
  return static_cast<GpuRealStorage *>(t.storage.get());
}
// operand that is only read (keeps the storage version, so bf16 shadows stay valid)
inline Dev dev_of(const BaseTensor &t, const char *op) {
  GpuRealStorage *s = gpu_storage(t, op);
  s->dev->Bind();
  return Dev{const_cast<real1 *>(s->device_ptr_ro()), s->dev->stream};
}
// does the view address every element of its storage exactly once?
inline bool covers_storage(const BaseTensor &t) {
  if (t.offset) return false;
  tcapint expect = 1U;
  for (size_t i = 0U; i < t.shape.size(); ++i) {
    if (t.shape[i] == 1U) continue;
    if (t.stride[i] != expect) return false;
    expect *= t.shape[i];
  }
  return expect == t.storage->size;
}
// operand that is written; `overwrites_all`: the kernel stores every element of the view without
// reading it, so a pending lazy zero-fill of a fully covered storage can be dropped
inline Dev dev_out(const BaseTensor &t, const char *op, bool overwrites_all = false) {
  GpuRealStorage *s = gpu_storage(t, op);
  s->dev->Bind();
  return Dev{(overwrites_all && covers_storage(t)) ? s->device_ptr_overwrite() : s->device_ptr(), s->dev->stream};
}
inline void require_real(const Tensor &t, const char *msg) {
  if (t.storage->dtype != DType::REAL) throw std::invalid_argument(msg);
}

// The reference's CPU loops resolve ONE flat index through every operand's own (shape, stride)
// (BaseTensor::get_storage_index), so operands only need equal element counts, not equal shapes:
// autograd routinely pairs e.g. a [1,T] gradient with a [1,T,1] contribution after
// squeeze()/unsqueeze(). The device kernels want one common shape, so bring the views to it:
// drop extent-1 dims, broadcast true scalars, and re-express a dense operand in the other's shape.
void strip_unit_dims(weedcu_view &v) {
  int r = 0;
  for (int d = 0; d < v.rank; ++d) {
    if (v.shape[d] == 1U) continue;
    v.shape[r] = v.shape[d];
    v.stride[r] = v.stride[d];
    ++r;
  }
  if (r == 0) {
    v.shape[0] = 1U;
    v.stride[0] = 0U;
    r = 1;
  }
  for (int d = r; d < WEEDCU_MAX_RANK; ++d) {
    v.shape[d] = 1U;
    v.stride[d] = 0U;
  }
  v.rank = r;
}
bool same_dims(const weedcu_view &a, const weedcu_view &b) {
  if (a.rank != b.rank) return false;
  for (int d = 0; d < a.rank; ++d)
    if (a.shape[d] != b.shape[d]) return false;
  return true;
}
bool all_broadcast(const weedcu_view &v) {
  for (int d = 0; d < v.rank; ++d)
    if (v.shape[d] > 1U && v.stride[d]) return false;
  return true;
}
bool dense_run(const weedcu_view &v) { // contiguous column-major run: flat index == storage offset
  uint64_t expect = 1;
  for (int d = 0; d < v.rank; ++d) {
    if (v.shape[d] == 1U) continue;
    if (v.stride[d] != expect) return false;
    expect *= v.shape[d];
  }
  return true;
}
void conform_views(std::vector<weedcu_view *> views, const char *name) {
  for (weedcu_view *v : views) strip_unit_dims(*v);
  // reference shape: the first operand that is neither a pure broadcast nor a plain dense run
  // (those two kinds can take any shape); else the first non-broadcast one; else the first
  const weedcu_view *ref = nullptr;
  for (weedcu_view *v : views)
    if (!all_broadcast(*v) && !dense_run(*v)) { ref = v; break; }
  if (!ref)
    for (weedcu_view *v : views)
      if (!all_broadcast(*v)) { ref = v; break; }
  if (!ref) ref = views[0];
  const weedcu_view r = *ref;
  for (weedcu_view *v : views) {
    if (same_dims(*v, r)) continue;
    const uint64_t off = v->offset;
    if (all_broadcast(*v)) {
      *v = r;
      v->offset = off;
      for (int d = 0; d < WEEDCU_MAX_RANK; ++d) v->stride[d] = 0U;
    } else if (dense_run(*v)) {
      *v = r;
      v->offset = off;
      uint32_t acc = 1U;
      for (int d = 0; d < v->rank; ++d) {
        v->stride[d] = (v->shape[d] == 1U) ? 0U : acc;
        acc *= v->shape[d];
      }
    } else {
      throw std::invalid_argument(std::string(name) + ": operands with equal element counts but incompatible strided shapes");
    }
  }
}

void binary(int op, const Tensor &a, const Tensor &b, Tensor &out, const char *name) {
  validate_all_same_device({&a, &b, &out}, name);
  const tcapint aSize = a.get_broadcast_size(), bSize = b.get_broadcast_size(), oSize = out.get_broadcast_size();
  if (aSize != bSize) throw std::invalid_argument(std::string("In ") + name + "(a, b, out), 'a' size does not match 'b' size!");
  if (aSize != oSize) throw std::invalid_argument(std::string("In ") + name + "(a, b, out), out size does not match input size!");
  const Dev da = dev_of(a, name), db = dev_of(b, name), dout = dev_out(out, name, true);
  weedcu_view av = a.view(), bv = b.view(), ov = out.view();
  conform_views({&ov, &av, &bv}, name);
  throw_on_error(weedcu_binary_real(op, da.ptr, &av, db.ptr, &bv, dout.ptr, &ov, dout.stream), name);
}

void in_place(int op, Tensor &a, const Tensor &b, const char *name) {
  validate_all_same_device({&a, &b}, name);
  if (a.get_broadcast_size() != b.get_broadcast_size())
    throw std::invalid_argument(std::string("In ") + name + "(a, b), 'a' size does not match 'b' size!");
  weedcu_view av = a.view(), bv = b.view();
  // Destination with broadcast (stride-0) dims: the reference's serial loop visits every flat index,
  // so each stored element is updated once per broadcast index (this is how sgd_step ends up
  // applying a bias update B times after match_shape mutated the Parameter, sgd.hpp:29-35).
  // On the device that would be a write race; when b is broadcast along the same dims the effect is
  // `times` identical updates, issued as `times` in-order launches over the collapsed views.
  uint64_t times = 1;
  if (same_dims(av, bv)) {
    bool needs_reduce = false;
    for (int d = 0; d < av.rank; ++d)
      if (av.shape[d] > 1U && av.stride[d] == 0U && bv.stride[d] != 0U) needs_reduce = true;
    if (needs_reduce) {
      // a[j] (+/-)= sum over the broadcast indices of b: what the serial loop accumulates when a
      // stale, un-reduced gradient meets a match_shape-mutated Parameter (e.g. adam_step on a bias
      // whose add node never ran). Sum b over those dims (intended index order), then recurse.
      struct QuirkOff {
        bool prev;
        QuirkOff() : prev(backend_config().ref_index_quirks) { backend_config().ref_index_quirks = false; }
        ~QuirkOff() { backend_config().ref_index_quirks = prev; }
      } guard;
      TensorPtr bt = std::make_shared<Tensor>(b);
      bt->requires_grad = false;
      bt->grad = nullptr;
      bt->grad_node = nullptr;
      Tensor a2(a);
      for (int d = (int)a.shape.size() - 1; d >= 0; --d) {
        if (a.shape[(size_t)d] > 1U && a.stride[(size_t)d] == 0U && bt->stride[(size_t)d] != 0U) {
          bt = Tensor::sum(bt, (symint)d);
          a2.shape[(size_t)d] = 1U;
        }
      }
      in_place(op, a2, *bt, name);
      return;
    }
    for (int d = 0; d < av.rank; ++d) {
      if (av.shape[d] > 1U && av.stride[d] == 0U) {
        if (bv.stride[d] != 0U)
          throw std::domain_error(std::string(name) + ": accumulating a non-broadcast source into a broadcast destination is not supported");
        times *= av.shape[d];
        av.shape[d] = 1U;
        bv.shape[d] = 1U;
      }
    }
  }
  if (times > 4096) throw std::domain_error(std::string(name) + ": broadcast destination repeated too many times");
  conform_views({&av, &bv}, name);
  for (int d = 0; d < av.rank; ++d)
    if (av.shape[d] > 1U && av.stride[d] == 0U)
      throw std::domain_error(std::string(name) + ": accumulating into a broadcast destination is not supported for these shapes");
  const Dev db = dev_of(b, name);
  GpuRealStorage *as = gpu_storage(a, name);
  if (op == WEEDCU_ADD && times == 1 && as->zero_pending && covers_storage(a)) {
    // first accumulation into a lazily zeroed gradient: 0 + b is a plain copy (8 B/elem, no fill) — or no copy at all
    // when b is a whole dense storage of the same layout: a then shares b's buffer copy-on-write
    if (backend_config().cow_grads && backend_config().fused && b.storage.get() != a.storage.get() && a.shape == b.shape && a.stride == b.stride &&
        covers_storage(b) && a.storage->size == b.storage->size && b.storage->device == DeviceTag::GPU &&
        a.storage->get_device_id() == b.storage->get_device_id()) {
      as->share_buffer_from(*gpu_storage(b, name));
      return;
    }
    const Dev da = dev_out(a, name, true);
    throw_on_error(weedcu_copy_real(da.ptr, &av, db.ptr, &bv, da.stream), name);
    return;
  }
  const Dev da = dev_out(a, name);
  for (uint64_t t = 0; t < times; ++t)
    throw_on_error(weedcu_inplace_real(op, da.ptr, &av, db.ptr, &bv, da.stream), name);
}

void unary(int op, real1 param, const Tensor &a, Tensor &out, const char *name) {
  validate_all_same_device({&a, &out}, name);
  if (a.get_broadcast_size() != out.get_broadcast_size())
    throw std::invalid_argument(std::string("In Weed::") + name + "(a, out), out size does not match input size!");
  const Dev da = dev_of(a, name), dout = dev_out(out, name, true);
  weedcu_view av = a.view(), ov = out.view();
  conform_views({&ov, &av}, name);
  throw_on_error(weedcu_unary_real(op, param, da.ptr, &av, dout.ptr, &ov, dout.stream), name);
}

void unary_grad(int op, Tensor &din, const Tensor &in, const Tensor &dout, const char *name) {
  validate_all_same_device({&din, &in, &dout}, name);
  const tcapint n = din.get_broadcast_size();
  if ((n != in.get_broadcast_size()) || (n != dout.get_broadcast_size()))
    throw std::invalid_argument(std::string("In Weed::") + name + "(din, in, dout), sizes do not match!");
  // a lazily zeroed, fully covered din is stored to, not accumulated into (no fill, no read)
  const int accumulate = (gpu_storage(din, name)->zero_pending && covers_storage(din)) ? 0 : 1;
  const Dev dd = dev_out(din, name, !accumulate), di = dev_of(in, name), dg = dev_of(dout, name);
  weedcu_view dv = din.view(), iv = in.view(), gv = dout.view();
  conform_views({&dv, &iv, &gv}, name);
  throw_on_error(weedcu_unary_grad_real(op, dd.ptr, &dv, di.ptr, &iv, dg.ptr, &gv, accumulate, dd.stream), name);
}


weedcu_mat mat_of(const Tensor &t, tcapint extra_offset = 0U, uint64_t batch_stride = 0U) {
  weedcu_mat m;
  m.offset = (uint64_t)t.offset + extra_offset;
  m.s0 = t.stride[t.stride.size() - 2U];
  m.s1 = t.stride[t.stride.size() - 1U];
  m.batch_stride = batch_stride;
  return m;
}

// bf16 copy of a matrix view for the tensor-core GEMM, cached on the storage it was packed from.
// The packed layout only depends on the memory the view addresses — [n_slow][round8(n_fast)] with
// the smaller-stride index contiguous — so a view and its transpose share one shadow: X serves the
// forward product and dW = X^T dY, W serves the forward product and dX = dY W^T, dY serves both
// backward products. A shadow is stale once its storage's version moved (any potential write).
struct Bf16Operand {
  BufferPtr buf; // keeps the shadow alive for as long as the operand is used (a later pack may evict it from its storage)
  const uint16_t *ptr;
  int major; // 1: the M (resp. N) index is contiguous, 0: the K index is
  uint64_t ld;
};
inline uint64_t round8(uint64_t x) { return (x + 7U) & ~(uint64_t)7U; }
// `colsum` (optional, dense [n_slow]): the pack pass also accumulates the fp32 sum over the contiguous
// index into it; false (nothing done) if the shadow already exists or the layout cannot stream.
bool bf16_operand(const Tensor &t, tcapint s_mn, tcapint s_k, tcapint n_mn, tcapint n_k, bool is_a, Bf16Operand &op,
                  real1 *colsum = nullptr, int colsum_accumulate = 1) {
  if (!s_mn || !s_k) return false; // broadcast operands take the generic path
  GpuRealStorage *s = gpu_storage(t, "matmul");
  const bool mn_fast = is_a ? (s_mn <= s_k) : (s_mn < s_k);
  const tcapint n_fast = mn_fast ? n_mn : n_k, n_slow = mn_fast ? n_k : n_mn;
  const tcapint s_fast = mn_fast ? s_mn : s_k, s_slow = mn_fast ? s_k : s_mn;
  op.major = mn_fast ? 1 : 0;
  op.ld = round8(n_fast);
  GpuRealStorage::Bf16Shadow *hit = nullptr;
  for (GpuRealStorage::Bf16Shadow &sh : s->shadows)
    if (sh.offset == t.offset && sh.n_fast == n_fast && sh.n_slow == n_slow && sh.s_fast == s_fast && sh.s_slow == s_slow) hit = &sh;
  if (hit && hit->version == s->version) { // (a current shadow is served without touching the fp32 buffer: deferred values stay deferred)
    if (colsum) return false;
    op.buf = hit->buf;
    op.ptr = (const uint16_t *)hit->buf->ptr;
    return true;
  }
  const real1 *src = s->device_ptr_ro(); // (materialises a pending zero fill / deferred values; keeps the version)
  if (!hit) {
    if (s->shadows.size() >= 4U) s->shadows.erase(s->shadows.begin());
    s->shadows.push_back(GpuRealStorage::Bf16Shadow{s->dev->MakeBuffer(2U * (size_t)(op.ld * n_slow + 8U)), 0U, t.offset, n_fast, n_slow, s_fast, s_slow});
    hit = &s->shadows.back();
  }
  s->dev->Bind();
  // weedcu_pack_bf16(rows = mn index, cols = k index): dst_major 1 keeps rows contiguous
  if (colsum) {
    const int rc = weedcu_pack_bf16_colsum(src, t.offset, s_mn, s_k, n_mn, n_k, (uint16_t *)hit->buf->ptr, op.major, colsum, colsum_accumulate,
                                           s->dev->stream);
    if (rc == WEEDCU_ENOSUP) return false; // (the shadow entry stays stale and is packed on first use)
    throw_on_error(rc, "pack_bf16_colsum");
  } else
    throw_on_error(weedcu_pack_bf16(src, t.offset, s_mn, s_k, n_mn, n_k, (uint16_t *)hit->buf->ptr, op.major, s->dev->stream), "pack_bf16");
  hit->version = s->version;
  op.buf = hit->buf;
  op.ptr = (const uint16_t *)hit->buf->ptr;
  return true;
}

// `bias` (optional): a dense [N] vector added to every row in the GEMM epilogue; when given and the
// tensor-core path does not apply, nothing is computed and false is returned.
// `cs` received only the bf16 copy of the product a b: produce the fp32 values on demand with the plain product of the same
// operand copies, as long as neither operand has changed since
void defer_product_output(GpuRealStorage *cs, const Bf16Operand &pa, const Bf16Operand &pb, tcapint M, tcapint N, tcapint K, const Tensor &a,
                          const Tensor &b) {
  const StoragePtr a_s = a.storage, b_s = b.storage;
  const uint64_t a_v = static_cast<GpuRealStorage *>(a_s.get())->version, b_v = static_cast<GpuRealStorage *>(b_s.get())->version;
  const BufferPtr abuf = pa.buf, bbuf = pb.buf;
  const int a_major = pa.major, b_major = pb.major;
  const uint64_t lda = pa.ld, ldb = pb.ld;
  cs->deferred_values = [cs, a_s, b_s, a_v, b_v, abuf, bbuf, a_major, b_major, lda, ldb, M, N, K]() {
    if (static_cast<GpuRealStorage *>(a_s.get())->version != a_v || static_cast<GpuRealStorage *>(b_s.get())->version != b_v)
      throw std::runtime_error("matmul: an operand was modified before the deferred fp32 product was read");
    cs->dev->Bind();
    throw_on_error(weedcu_gemm_bf16((const uint16_t *)abuf->ptr, a_major, lda, (const uint16_t *)bbuf->ptr, b_major, ldb, (real1 *)cs->buffer->ptr, M, M, N, K,
                                    0, nullptr, cs->dev->stream),
                   "matmul (deferred fp32 product)");
  };
}
bool matmul_impl(const Tensor &a, const Tensor &b, Tensor &out, int accumulate, const Tensor *bias = nullptr, const Tensor *residual = nullptr) {
  validate_all_same_device({&a, &b, &out}, "MatMulKernel::matmul");
  if ((a.shape.size() != 2U) || (b.shape.size() != 2U) || (out.shape.size() != 2U))
    throw std::invalid_argument("MatMul is only for matrices with 2 indices!");
  const tcapint K = a.shape[1U];
  if (K != b.shape[0U]) throw std::invalid_argument("MatMul operand dimensions aren't compatible!");
  const tcapint M = a.shape[0U], N = b.shape[1U];
  if ((M != out.shape[0U]) || (N != out.shape[1U])) throw std::invalid_argument("MatMul output dimensions don't match inputs!");
  GpuRealStorage *cs = gpu_storage(out, "matmul");
  if (accumulate && cs->zero_pending && covers_storage(out)) accumulate = 0; // 0 + A*B: plain store, no fill
  const BackendConfig &cfg = backend_config();
  if (cfg.matmul_precision == WEEDCU_GEMM_BF16 && cfg.fused && cfg.operand_cache && out.stride[0U] == 1U && M >= 64U && N >= 16U && K >= 32U) {
    // bf16 operands come from per-storage shadows: a weight is packed once per optimiser step and an
    // activation / output gradient once per graph, not once per GEMM that reads it
    Bf16Operand pa, pb;
    if (bf16_operand(a, a.stride[0U], a.stride[1U], M, K, true, pa) && bf16_operand(b, b.stride[1U], b.stride[0U], N, K, false, pb)) {
      const Dev dc = dev_out(out, "matmul", !accumulate);
      const real1 *bias_ptr = bias ? dev_of(*bias, "matmul").ptr + bias->offset : nullptr;
      int rc;
      if (residual && cfg.epilogue_stats && !out.offset && covers_storage(out) && N <= 8U * 128U) {
        // the residual sum is what the next LayerNorm normalises: its epilogue leaves the per-row (mean, M2) partials, and
        // LayerNorm::forward skips its statistics pass over the tensor (layernorm_forward below)
        BufferPtr stats = cs->dev->MakeBuffer(sizeof(real1) * 2U * (size_t)M * 16U); // <= 2 * ceil(N / 128) partials per row
        uint32_t tiles = 0U, tile_cols = 0U;
        weedcu_gemm_epilogue epi;
        memset(&epi, 0, sizeof(epi));
        epi.col_bias = bias_ptr;
        epi.residual = dev_of(*residual, "matmul").ptr + residual->offset;
        epi.ldr = out.stride[1U];
        epi.row_stats = 1;
        epi.stats = (float *)stats->ptr;
        epi.stats_capacity_tiles = 16U;
        epi.stats_tiles = &tiles;
        epi.stats_tile_cols = &tile_cols;
        rc = weedcu_gemm_bf16_ex(pa.ptr, pa.major, pa.ld, pb.ptr, pb.major, pb.ld, dc.ptr, out.stride[1U], nullptr, 0U, M, N, K, &epi, dc.stream);
        if (rc == 0) {
          cs->row_stats = stats;
          cs->row_stats_kind = 1;
          cs->row_stats_tiles = tiles;
          cs->row_stats_tile_cols = tile_cols;
          cs->row_stats_rows = M;
          cs->row_stats_cols = N;
          cs->row_stats_version = cs->version;
        }
      } else if (!accumulate && !bias && !residual && cs->accept_bf16_values && cfg.defer_grads && cfg.epilogue_stats && !out.offset &&
                 covers_storage(out) && out.stride[1U] == M && (M % 8U) == 0U && N >= 32U) {
        // the whole storage is overwritten and its one reader takes bf16 (core.hpp: accept_bf16_values): write only the
        // bf16 copy; the fp32 values are formed on demand by the plain product of the same operand copies
        const OutputShadow os = begin_output_shadow(out, N);
        rc = WEEDCU_ENOSUP;
        if (os.ptr) {
          cs->dev->Bind();
          cs->device_ptr_overwrite();
          weedcu_gemm_epilogue epi;
          memset(&epi, 0, sizeof(epi));
          rc = weedcu_gemm_bf16_ex(pa.ptr, pa.major, pa.ld, pb.ptr, pb.major, pb.ld, nullptr, 0U, os.ptr, M, M, N, K, &epi, dc.stream);
          if (rc == 0) {
            end_output_shadow(os);
            defer_product_output(cs, pa, pb, M, N, K, a, b);
          }
        }
        if (rc == WEEDCU_ENOSUP)
          rc = weedcu_gemm_bf16(pa.ptr, pa.major, pa.ld, pb.ptr, pb.major, pb.ld, dc.ptr + out.offset, out.stride[1U], M, N, K, accumulate, bias_ptr,
                                dc.stream);
      } else if (residual) // [M, N] with the layout of `out`, added after the bias in the epilogue
        rc = weedcu_gemm_bf16_residual(pa.ptr, pa.major, pa.ld, pb.ptr, pb.major, pb.ld, dc.ptr + out.offset, out.stride[1U], M, N, K, bias_ptr,
                                       dev_of(*residual, "matmul").ptr + residual->offset, out.stride[1U], dc.stream);
      else
        rc = weedcu_gemm_bf16(pa.ptr, pa.major, pa.ld, pb.ptr, pb.major, pb.ld, dc.ptr + out.offset, out.stride[1U], M, N, K, accumulate, bias_ptr,
                              dc.stream);
      if (rc != WEEDCU_ENOSUP) {
        throw_on_error(rc, "matmul");
        return true;
      }
    }
  }
  if (cfg.fused && M <= 16U && residual) { // decode step: the residual rides in the skinny kernel's final add
    // (Tensor::linear has checked that the residual is the dense column-major tensor `out` becomes after its reshape)
    if (out.stride[0U] != 1U || out.stride[1U] != M || residual->get_broadcast_size() != (tcapint)M * N) return false;
    const Dev da = dev_of(a, "matmul"), db = dev_of(b, "matmul"), dr = dev_of(*residual, "matmul"), dc = dev_out(out, "matmul", true);
    const weedcu_mat am = mat_of(a), bm = mat_of(b);
    weedcu_mat cm = mat_of(out);
    // residual and C share one mat descriptor: fold their (possibly different) offsets into the pointers
    const uint64_t c_off = cm.offset, r_off = residual->offset;
    cm.offset = 0U;
    const real1 *bias_ptr = bias ? dev_of(*bias, "matmul").ptr + bias->offset : nullptr;
    throw_on_error(weedcu_matmul_skinny_residual(da.ptr, &am, db.ptr, &bm, dc.ptr + c_off, &cm, M, K, N, bias_ptr, dr.ptr + r_off, dc.stream), "matmul");
    return true;
  }
  if (residual) return false; // no other path adds a residual: the caller composes Linear + add
  if (cfg.fused && M <= 16U) {
    // a handful of rows (a decode step, or a tiny training batch): the product is one pass over the
    // weight matrix — skinny FFMA kernel at fp32 in either precision mode, bias added in the same pass
    const Dev da = dev_of(a, "matmul"), db = dev_of(b, "matmul"), dc = dev_out(out, "matmul", !accumulate);
    const weedcu_mat am = mat_of(a), bm = mat_of(b), cm = mat_of(out);
    const real1 *bias_ptr = bias ? dev_of(*bias, "matmul").ptr + bias->offset : nullptr;
    throw_on_error(weedcu_matmul_skinny(da.ptr, &am, db.ptr, &bm, dc.ptr, &cm, M, K, N, bias_ptr, accumulate, dc.stream), "matmul");
    return true;
  }
  if (bias) return false;
  const Dev da = dev_of(a, "matmul"), db = dev_of(b, "matmul"), dc = dev_out(out, "matmul", !accumulate);
  const weedcu_mat am = mat_of(a), bm = mat_of(b), cm = mat_of(out);
  throw_on_error(weedcu_matmul_real(da.ptr, &am, db.ptr, &bm, dc.ptr, &cm, M, K, N, 1U, accumulate,
                                    backend_config().matmul_precision, dc.stream),
                 "matmul");
  return true;
}
} // namespace

void add(const Tensor &a, const Tensor &b, Tensor &out) { binary(WEEDCU_ADD, a, b, out, "CommutingKernel::commuting"); }
void mul(const Tensor &a, const Tensor &b, Tensor &out) { binary(WEEDCU_MUL, a, b, out, "CommutingKernel::commuting"); }
void sub(const Tensor &a, const Tensor &b, Tensor &out) { binary(WEEDCU_SUB, a, b, out, "SubKernel::sub"); }
void div(const Tensor &a, const Tensor &b, Tensor &out) { binary(WEEDCU_DIV, a, b, out, "DivKernel::div"); }
void add_in_place(Tensor &a, const Tensor &b) { in_place(WEEDCU_ADD, a, b, "InPlaceKernel::in_place"); }
void sub_in_place(Tensor &a, const Tensor &b) { in_place(WEEDCU_SUB, a, b, "InPlaceKernel::in_place"); }

void copy_broadcast(Tensor &a, const Tensor &b) {
  validate_all_same_device({&a, &b}, "CopyKernel::copy_broadcast");
  if (a.get_size() != b.get_broadcast_size())
    throw std::invalid_argument("In CopyKernel::copy_broadcast(a, b), 'a' size does not match 'b' size!");
  const Dev da = dev_out(a, "copy_broadcast", true), db = dev_of(b, "copy_broadcast");
  const weedcu_view av = a.view(), bv = b.view();
  throw_on_error(weedcu_copy_real(da.ptr, &av, db.ptr, &bv, da.stream), "copy_broadcast");
}

void relu(const Tensor &a, Tensor &out) { unary(WEEDCU_RELU, 0, a, out, "relu"); }
void sigmoid(const Tensor &a, Tensor &out) { unary(WEEDCU_SIGMOID, 0, a, out, "sigmoid"); }
void tanh(const Tensor &a, Tensor &out) { unary(WEEDCU_TANH, 0, a, out, "tanh"); }
void sin(const Tensor &a, Tensor &out) { unary(WEEDCU_SIN, 0, a, out, "sin"); }
void cos(const Tensor &a, Tensor &out) { unary(WEEDCU_COS, 0, a, out, "cos"); }
void abs(const Tensor &a, Tensor &out) { unary(WEEDCU_ABS, 0, a, out, "abs"); }
OutputShadow begin_output_shadow(const Tensor &out, tcapint cols) {
  OutputShadow os;
  const BackendConfig &cfg = backend_config();
  if (cfg.matmul_precision != WEEDCU_GEMM_BF16 || !cfg.fused || !cfg.operand_cache || !cols) return os;
  if (out.storage->device != DeviceTag::GPU || out.offset || !covers_storage(out)) return os;
  const tcapint n = out.storage->size;
  if (n % cols) return os;
  const tcapint rows = n / cols;
  // what matmul_impl / bf16_operand will ask for: A = [rows, cols], rows contiguous, tensor-core eligible
  if ((rows % 8U) || rows < 64U || cols < 32U) return os;
  GpuRealStorage *s = gpu_storage(out, "output shadow");
  for (size_t k = 0U; k < s->shadows.size(); ++k) {
    const GpuRealStorage::Bf16Shadow &c = s->shadows[k];
    if (c.offset == 0U && c.n_fast == rows && c.n_slow == cols && c.s_fast == 1U && c.s_slow == rows) {
      os.storage = s;
      os.index = k;
      os.ptr = (uint16_t *)c.buf->ptr;
      return os;
    }
  }
  if (s->shadows.size() >= 4U) s->shadows.erase(s->shadows.begin());
  s->shadows.push_back(GpuRealStorage::Bf16Shadow{s->dev->MakeBuffer(2U * ((size_t)rows * cols + 8U)), 0U, 0U, rows, cols, 1U, rows});
  os.storage = s;
  os.index = s->shadows.size() - 1U;
  os.ptr = (uint16_t *)s->shadows.back().buf->ptr;
  return os;
}
void end_output_shadow(const OutputShadow &os) {
  if (os.storage) os.storage->shadows[os.index].version = os.storage->version;
}

namespace {
void defer_gelu_values(GpuRealStorage *ys, const StoragePtr &a_s, tcapint n);
}
void gelu(const Tensor &a, Tensor &out) {
  // dense input and output of the same layout: the forward can leave the bf16 operand of the Linear
  // that follows (ff2 reads [B*T, d_ff]) in the same pass
  if (a.shape.size() >= 2U && a.shape == out.shape && a.stride == out.stride && !a.offset && covers_storage(a) && a.storage->device == DeviceTag::GPU) {
    const OutputShadow os = begin_output_shadow(out, out.shape.back());
    if (os.ptr) {
      const Dev da = dev_of(a, "gelu"), dout = dev_out(out, "gelu", true);
      const bool defer = backend_config().defer_grads && covers_storage(out);
      const int rc = weedcu_gelu_fwd_bf16(da.ptr, defer ? nullptr : dout.ptr, os.ptr, out.storage->size, dout.stream);
      if (rc == 0) {
        end_output_shadow(os);
        // in a transformer layer y = gelu(h) is read by ff2's forward product and by ff2's weight-gradient product, both
        // through the bf16 shadow; the GELU backward needs h only. The fp32 y is computed if something else reads it
        if (defer) defer_gelu_values(gpu_storage(out, "gelu"), a.storage, out.storage->size);
        return;
      }
      if (rc != WEEDCU_ENOSUP) throw_on_error(rc, "gelu");
    }
  }
  unary(WEEDCU_GELU, 0, a, out, "gelu");
}
void layernorm_forward(const Tensor &x, tcapint rows, tcapint features, const Tensor &gamma, const Tensor &beta, real1 eps, Tensor &y, Tensor &mean,
                       Tensor &rstd) {
  const Dev dx = dev_of(x, "LayerNorm::forward"), dg = dev_of(gamma, "LayerNorm::forward"), db = dev_of(beta, "LayerNorm::forward");
  const OutputShadow os = begin_output_shadow(y, features);
  const Dev dy = dev_out(y, "LayerNorm::forward", true), dm = dev_out(mean, "LayerNorm::forward", true), dr = dev_out(rstd, "LayerNorm::forward", true);
  {
    // x straight out of a residual GEMM whose epilogue left the row partials: one pass instead of two
    GpuRealStorage *xs = gpu_storage(x, "LayerNorm::forward");
    if (xs->row_stats && xs->row_stats_kind == 1 && xs->row_stats_version == xs->version && !x.offset && xs->row_stats_rows == rows &&
        xs->row_stats_cols == features && (uint64_t)rows * features == xs->size) {
      const int rc = weedcu_layernorm_fwd_stats(dx.ptr, rows, features, (const float *)xs->row_stats->ptr, xs->row_stats_tiles, xs->row_stats_tile_cols,
                                                dg.ptr + gamma.offset, db.ptr + beta.offset, eps, dy.ptr, dm.ptr, dr.ptr, os.ptr, dy.stream);
      if (rc == 0) {
        if (os.ptr) end_output_shadow(os);
        return;
      }
      if (rc != WEEDCU_ENOSUP) throw_on_error(rc, "LayerNorm::forward");
    }
  }
  if (os.ptr) {
    const int rc = weedcu_layernorm_fwd_bf16(dx.ptr + x.offset, rows, features, dg.ptr + gamma.offset, db.ptr + beta.offset, eps, dy.ptr, dm.ptr, dr.ptr,
                                             os.ptr, dy.stream);
    if (rc == 0) {
      end_output_shadow(os);
      return;
    }
    if (rc != WEEDCU_ENOSUP) throw_on_error(rc, "LayerNorm::forward");
  }
  throw_on_error(weedcu_layernorm_fwd(dx.ptr + x.offset, rows, features, dg.ptr + gamma.offset, db.ptr + beta.offset, eps, dy.ptr, dm.ptr, dr.ptr, dy.stream),
                 "LayerNorm::forward");
}
namespace {
// y = gelu(a) was left as its bf16 GEMM operand copy only: the fp32 values are computed from `a` if something reads them
// (and `a` must not have been written in between: checked, not assumed)
void defer_gelu_values(GpuRealStorage *ys, const StoragePtr &a_s, tcapint n) {
  const uint64_t a_version = static_cast<GpuRealStorage *>(a_s.get())->version;
  ys->deferred_values = [ys, a_s, a_version, n]() {
    GpuRealStorage *as = static_cast<GpuRealStorage *>(a_s.get());
    if (as->version != a_version) throw std::runtime_error("gelu: the input was modified before the deferred fp32 output was read");
    weedcu_view v;
    memset(&v, 0, sizeof(v));
    v.rank = 1;
    v.shape[0] = n;
    v.stride[0] = 1U;
    ys->dev->Bind();
    throw_on_error(weedcu_unary_real(WEEDCU_GELU, 0, as->device_ptr_ro(), &v, (real1 *)ys->buffer->ptr, &v, ys->dev->stream), "gelu (deferred)");
  };
}
// `cs` received only the bf16 copy of a w + bias: remember the operands, and produce the fp32 values on demand with the
// plain product of the same operands (bit-identical to what Linear::forward writes without this path), as long as the
// weights have not been updated since
void defer_linear_output(GpuRealStorage *cs, const Bf16Operand &pa, const Bf16Operand &pb, tcapint M, tcapint N, tcapint K, const Tensor &w,
                         const Tensor &bias) {
  std::shared_ptr<GpuRealStorage::GemmSource> src = std::make_shared<GpuRealStorage::GemmSource>();
  src->a = pa.buf;
  src->b = pb.buf;
  src->a_major = pa.major;
  src->b_major = pb.major;
  src->lda = pa.ld;
  src->ldb = pb.ld;
  src->M = M;
  src->N = N;
  src->K = K;
  src->w_storage = w.storage;
  src->bias_storage = bias.storage;
  src->w_version = static_cast<GpuRealStorage *>(w.storage.get())->version;
  src->bias_version = static_cast<GpuRealStorage *>(bias.storage.get())->version;
  src->bias_offset = bias.offset;
  cs->gemm_source = src;
  cs->deferred_values = [cs, src]() {
    GpuRealStorage *ws = static_cast<GpuRealStorage *>(src->w_storage.get()), *bs = static_cast<GpuRealStorage *>(src->bias_storage.get());
    if (ws->version != src->w_version || bs->version != src->bias_version)
      throw std::runtime_error("Linear: the weights were modified before the deferred fp32 output was read");
    cs->dev->Bind();
    throw_on_error(weedcu_gemm_bf16((const uint16_t *)src->a->ptr, src->a_major, src->lda, (const uint16_t *)src->b->ptr, src->b_major, src->ldb,
                                    (real1 *)cs->buffer->ptr, src->M, src->M, src->N, src->K, 0, bs->device_ptr_ro() + src->bias_offset, cs->dev->stream),
                   "Linear (deferred fp32 output)");
  };
}
bool tensor_core_linear_ok(const Tensor &a, const Tensor &w, const Tensor &bias, const Tensor &out) {
  const BackendConfig &cfg = backend_config();
  if (cfg.matmul_precision != WEEDCU_GEMM_BF16 || !cfg.fused || !cfg.operand_cache || !cfg.epilogue_stats || !cfg.defer_grads) return false;
  if (a.shape.size() != 2U || w.shape.size() != 2U || out.shape.size() != 2U || a.shape[1U] != w.shape[0U]) return false;
  const tcapint M = a.shape[0U], K = a.shape[1U], N = w.shape[1U];
  if (out.shape[0U] != M || out.shape[1U] != N || out.stride[0U] != 1U || out.stride[1U] != M || out.offset || !covers_storage(out)) return false;
  if ((M % 8U) || M < 64U || N < 32U || K < 32U) return false;
  if (bias.storage->size != N || bias.storage->device != DeviceTag::GPU) return false;
  return true;
}
} // namespace

bool matmul_bias_gelu(const Tensor &a, const Tensor &w, const Tensor &bias, Tensor &h, Tensor &y) {
  if (!tensor_core_linear_ok(a, w, bias, h)) return false;
  if (y.shape != h.shape || y.stride != h.stride || y.offset || !covers_storage(y)) return false;
  validate_all_same_device({&a, &w, &h, &y}, "matmul_bias_gelu");
  const tcapint M = a.shape[0U], K = a.shape[1U], N = w.shape[1U];
  Bf16Operand pa, pb;
  if (!bf16_operand(a, a.stride[0U], a.stride[1U], M, K, true, pa) || !bf16_operand(w, w.stride[1U], w.stride[0U], N, K, false, pb)) return false;
  const OutputShadow os = begin_output_shadow(y, N);
  if (!os.ptr) return false;
  const Dev dh = dev_out(h, "matmul_bias_gelu", true);
  GpuRealStorage *ys = gpu_storage(y, "matmul_bias_gelu");
  ys->device_ptr_overwrite(); // y's fp32 values are not written here: deferred below
  weedcu_gemm_epilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.col_bias = dev_of(bias, "matmul_bias_gelu").ptr + bias.offset;
  epi.activation = 1;
  const int rc = weedcu_gemm_bf16_ex(pa.ptr, pa.major, pa.ld, pb.ptr, pb.major, pb.ld, dh.ptr, h.stride[1U], os.ptr, M, M, N, K, &epi, dh.stream);
  if (rc == WEEDCU_ENOSUP) return false; // (nothing was launched; both outputs are about to be overwritten by the caller's fallback)
  throw_on_error(rc, "matmul_bias_gelu");
  end_output_shadow(os);
  defer_gelu_values(ys, h.storage, h.storage->size);
  return true;
}

bool matmul_bias_lse(const Tensor &a, const Tensor &w, const Tensor &bias, Tensor &out) {
  if (!tensor_core_linear_ok(a, w, bias, out)) return false;
  validate_all_same_device({&a, &w, &out}, "matmul_bias_lse");
  const tcapint M = a.shape[0U], K = a.shape[1U], N = w.shape[1U];
  Bf16Operand pa, pb;
  if (!bf16_operand(a, a.stride[0U], a.stride[1U], M, K, true, pa) || !bf16_operand(w, w.stride[1U], w.stride[0U], N, K, false, pb)) return false;
  const OutputShadow os = begin_output_shadow(out, N);
  if (!os.ptr) return false;
  GpuRealStorage *cs = gpu_storage(out, "matmul_bias_lse");
  cs->dev->Bind();
  cs->device_ptr_overwrite();
  // (log-sum-exp partials from this epilogue were measured too: 412 M exps on the eight epilogue warps cost +320 us on
  //  the 8192 x 50257 product, more than the 2 B/elem pass over the bf16 logits that replaces them)
  weedcu_gemm_epilogue epi;
  memset(&epi, 0, sizeof(epi));
  epi.col_bias = dev_of(bias, "matmul_bias_lse").ptr + bias.offset;
  const int rc = weedcu_gemm_bf16_ex(pa.ptr, pa.major, pa.ld, pb.ptr, pb.major, pb.ld, nullptr, 0U, os.ptr, M, M, N, K, &epi, cs->dev->stream);
  if (rc == WEEDCU_ENOSUP) return false;
  throw_on_error(rc, "matmul_bias_lse");
  end_output_shadow(os);
  defer_linear_output(cs, pa, pb, M, N, K, w, bias);
  return true;
}

static const symint *sym_ptr(const SymbolTensor &s, const char *op);
bool cross_entropy_fwd_from_stats(const Tensor &logits, const SymbolTensor &targets, Tensor &lse, Tensor &loss, tcapint rows, tcapint V) {
  GpuRealStorage *ls = gpu_storage(logits, "cross_entropy_loss");
  if (!ls->deferred_values || !ls->gemm_source || logits.offset) return false;
  const GpuRealStorage::GemmSource &src = *ls->gemm_source;
  if (src.M != rows || src.N != V) return false;
  GpuRealStorage *ws = static_cast<GpuRealStorage *>(src.w_storage.get()), *bs = static_cast<GpuRealStorage *>(src.bias_storage.get());
  if (ws->version != src.w_version || bs->version != src.bias_version) return false;
  const uint16_t *logits_bf16 = nullptr;
  for (const GpuRealStorage::Bf16Shadow &sh : ls->shadows)
    if (sh.offset == 0U && sh.n_fast == rows && sh.n_slow == V && sh.s_fast == 1U && sh.s_slow == rows && sh.version == ls->version)
      logits_bf16 = (const uint16_t *)sh.buf->ptr;
  if (!logits_bf16) return false;
  ls->dev->Bind();
  const Dev dl = dev_out(lse, "cross_entropy_loss", true), dloss = dev_out(loss, "cross_entropy_loss", true);
  const int rc = weedcu_cross_entropy_fwd_bf16in(logits_bf16, rows, V, (const uint16_t *)src.a->ptr, src.a_major, src.lda, (const uint16_t *)src.b->ptr, src.b_major,
                                                 src.ldb, src.K, bs->device_ptr_ro() + src.bias_offset, sym_ptr(targets, "cross_entropy_loss") + targets.offset,
                                                 dl.ptr + lse.offset, dloss.ptr + loss.offset, ls->dev->stream);
  if (rc == WEEDCU_ENOSUP) return false;
  throw_on_error(rc, "cross_entropy_loss");
  return true;
}

void pow(const Tensor &a, const real1 &p, Tensor &out) { unary(WEEDCU_POW, p, a, out, "pow"); }
void exp(const Tensor &a, const real1 &b, Tensor &out) { unary(WEEDCU_EXP, (real1)std::log((real1_s)b), a, out, "exp"); }
void log(const Tensor &a, const real1 &b, Tensor &out) {
  if (b <= ZERO_R1) throw std::invalid_argument("Log base must be positive!");
  unary(WEEDCU_LOG, (real1)(ONE_R1 / std::log((real1_s)b)), a, out, "log");
}
void relu_grad(Tensor &din, const Tensor &in, const Tensor &dout) { unary_grad(WEEDCU_RELU, din, in, dout, "relu_grad"); }
void sigmoid_grad(Tensor &din, const Tensor &in, const Tensor &dout) { unary_grad(WEEDCU_SIGMOID, din, in, dout, "sigmoid_grad"); }
void tanh_grad(Tensor &din, const Tensor &in, const Tensor &dout) { unary_grad(WEEDCU_TANH, din, in, dout, "tanh_grad"); }
void sin_grad(Tensor &din, const Tensor &in, const Tensor &dout) { unary_grad(WEEDCU_SIN, din, in, dout, "sin_grad"); }
void cos_grad(Tensor &din, const Tensor &in, const Tensor &dout) { unary_grad(WEEDCU_COS, din, in, dout, "cos_grad"); }
void abs_grad(Tensor &din, const Tensor &in, const Tensor &dout) { unary_grad(WEEDCU_ABS, din, in, dout, "abs_grad"); }
void gelu_grad(Tensor &din, const Tensor &in, const Tensor &dout) {
  // din = gradient of the pre-activation [.., d_ff] that ff1's Linear node consumes next: leave its bf16
  // GEMM operand copy and its column sums (ff1's bias gradient) on the storage in the same pass
  // (include/weedcu.h: weedcu_gelu_grad_pack; picked up by pack_with_column_sums / bf16_operand)
  const BackendConfig &cfg = backend_config();
  if (cfg.matmul_precision == WEEDCU_GEMM_BF16 && cfg.fused && cfg.operand_cache && din.shape.size() >= 2U && din.shape == in.shape &&
      din.shape == dout.shape && din.stride == in.stride && din.stride == dout.stride && !din.offset && !in.offset && !dout.offset &&
      covers_storage(din) && covers_storage(in) && covers_storage(dout) && din.storage->device == DeviceTag::GPU) {
    const tcapint cols = din.shape.back(), rows = din.storage->size / cols;
    if ((rows % 8U) == 0U && rows >= 64U && cols >= 16U) {
      validate_all_same_device({&din, &in, &dout}, "gelu_grad");
      GpuRealStorage *ds = gpu_storage(din, "gelu_grad");
      const int accumulate = ds->zero_pending ? 0 : 1;
      GpuRealStorage::Bf16Shadow *hit = nullptr;
      for (GpuRealStorage::Bf16Shadow &sh : ds->shadows)
        if (sh.offset == 0U && sh.n_fast == rows && sh.n_slow == cols && sh.s_fast == 1U && sh.s_slow == rows) hit = &sh;
      if (!hit) {
        if (ds->shadows.size() >= 4U) ds->shadows.erase(ds->shadows.begin());
        ds->shadows.push_back(GpuRealStorage::Bf16Shadow{ds->dev->MakeBuffer(2U * ((size_t)rows * cols + 8U)), 0U, 0U, rows, cols, 1U, rows});
        hit = &ds->shadows.back();
      }
      if (!ds->colsum || ds->colsum_n != cols) {
        ds->colsum = ds->dev->MakeBuffer(sizeof(real1) * (size_t)cols);
        ds->colsum_n = cols;
      }
      // dout whose fp32 values are still deferred behind a current bf16 copy (matmul_impl: accept_bf16_values): read that
      GpuRealStorage *gs = gpu_storage(dout, "gelu_grad");
      const uint16_t *dout_bf16 = nullptr;
      if (gs->deferred_values)
        for (const GpuRealStorage::Bf16Shadow &sh : gs->shadows)
          if (sh.offset == 0U && sh.n_fast == rows && sh.n_slow == cols && sh.s_fast == 1U && sh.s_slow == rows && sh.version == gs->version)
            dout_bf16 = (const uint16_t *)sh.buf->ptr;
      const Dev di = dev_of(in, "gelu_grad");
      const bool defer = !accumulate && cfg.defer_grads;
      real1 *out = accumulate ? ds->device_ptr() : ds->device_ptr_overwrite();
      const int rc = dout_bf16 ? weedcu_gelu_grad_pack_bf16dy(defer ? nullptr : out, di.ptr, dout_bf16, rows, cols, accumulate, (uint16_t *)hit->buf->ptr,
                                                              (real1 *)ds->colsum->ptr, ds->dev->stream)
                               : weedcu_gelu_grad_pack(defer ? nullptr : out, di.ptr, dev_of(dout, "gelu_grad").ptr, rows, cols, accumulate,
                                                       (uint16_t *)hit->buf->ptr, (real1 *)ds->colsum->ptr, ds->dev->stream);
      if (rc == 0) {
        hit->version = ds->version;
        ds->colsum_version = ds->version;
        if (defer) {
          // every reader of this gradient in a training step (ff1's dA / dB products, its bias gradient) is served by
          // the shadow and the column sums; the fp32 values are produced only if something else asks for them
          StoragePtr in_s = in.storage, dout_s = dout.storage;
          const tcapint n = rows * cols;
          const uint64_t in_v = static_cast<GpuRealStorage *>(in_s.get())->version, dout_v = static_cast<GpuRealStorage *>(dout_s.get())->version;
          ds->deferred_values = [ds, in_s, dout_s, n, in_v, dout_v]() {
            if (static_cast<GpuRealStorage *>(in_s.get())->version != in_v || static_cast<GpuRealStorage *>(dout_s.get())->version != dout_v)
              throw std::runtime_error("gelu_grad: an input was modified before the deferred fp32 gradient was read");
            weedcu_view v;
            memset(&v, 0, sizeof(v));
            v.rank = 1;
            v.shape[0] = n;
            v.stride[0] = 1U;
            ds->dev->Bind();
            throw_on_error(weedcu_unary_grad_real(WEEDCU_GELU, (real1 *)ds->buffer->ptr, &v, static_cast<GpuRealStorage *>(in_s.get())->device_ptr_ro(), &v,
                                                  static_cast<GpuRealStorage *>(dout_s.get())->device_ptr_ro(), &v, 0, ds->dev->stream),
                           "gelu_grad (deferred)");
          };
        }
        return;
      }
      if (rc != WEEDCU_ENOSUP) throw_on_error(rc, "gelu_grad");
      if (!accumulate) ds->FillZeros(); // nothing was launched: restore the pending zero fill the plain kernel relies on
    }
  }
  unary_grad(WEEDCU_GELU, din, in, dout, "gelu_grad");
}

static void full_reduce(const Tensor &a, Tensor &out, bool is_mean) {
  validate_all_same_device({&a, &out}, "SumKernel::sum");
  if (out.get_broadcast_size() != 1U)
    throw std::invalid_argument("In Weed::sum(a, out) or Weed::mean(a, out), out parameter is not a scalar!");
  const Dev da = dev_of(a, "sum"), dout = dev_out(out, "sum");
  const weedcu_view av = a.view();
  const real1 scale = is_mean ? (ONE_R1 / (real1)a.get_broadcast_size()) : ONE_R1;
  throw_on_error(weedcu_sum_real(da.ptr, &av, scale, dout.ptr + out.offset, dout.stream), "sum");
}
void sum(const Tensor &a, Tensor &out) { full_reduce(a, out, false); }
void mean(const Tensor &a, Tensor &out) { full_reduce(a, out, true); }

void reduce(const tcapint &index, const Tensor &a, Tensor &out) {
  validate_all_same_device({&a, &out}, "ReduceKernel::reduce");
  if (a.storage->dtype != out.storage->dtype) throw std::invalid_argument("Output tensor dtype mismatch in reduce!");
  if (index >= a.shape.size()) throw std::invalid_argument("reduce: axis out of range");
  const Dev da = dev_of(a, "reduce"), dout = dev_out(out, "reduce");
  const weedcu_view av = a.view();
  // the output is written as a dense buffer starting at out.offset, which is how the tensor built
  // by Tensor::sum(axis) (contiguous, axis extent 1) is read
  throw_on_error(weedcu_reduce_real(da.ptr, &av, (int)index, dout.ptr + out.offset, backend_config().ref_index_quirks ? 1 : 0,
                                    dout.stream),
                 "reduce");
}
void reduce_grad(const tcapint &index, Tensor &din, const Tensor &in, const Tensor &dout) {
  validate_all_same_device({&din, &dout}, "ReduceKernel::reduce_grad");
  const tcapint n = din.get_broadcast_size();
  if ((n != in.get_broadcast_size()) || (n != dout.get_broadcast_size()))
    throw std::invalid_argument("In Weed::reduce_grad(din, in, dout), sizes do not match!");
  const Dev dd = dev_out(din, "reduce_grad"), dg = dev_of(dout, "reduce_grad");
  const weedcu_view dv = din.view(), gv = dout.view();
  throw_on_error(weedcu_reduce_grad_real(dd.ptr, &dv, dg.ptr, &gv, (int)index, backend_config().ref_index_quirks ? 1 : 0, dd.stream),
                 "reduce_grad");
}

// ---- clamp / max / min (reference src/ops/clamp.cpp:116-170, real_extremum.cpp:165-262, reduce.cpp:362-449)
void clamp(const Tensor &a, const real1 &l, const real1 &h, Tensor &out) {
  validate_all_same_device({&a, &out}, "ClampKernel::clamp");
  if ((a.storage->dtype != DType::REAL) || (out.storage->dtype != DType::REAL))
    throw std::invalid_argument("In Weed::clamp(a, l, h, out), arguments must all be real-number!");
  if (a.get_broadcast_size() != out.get_broadcast_size()) throw std::invalid_argument("In Weed::clamp(a, l, h, out), out size does not match input size!");
  const Dev da = dev_of(a, "clamp"), dout = dev_out(out, "clamp", true);
  weedcu_view av = a.view(), ov = out.view();
  conform_views({&ov, &av}, "clamp");
  throw_on_error(weedcu_clamp_real(da.ptr, &av, l, h, dout.ptr, &ov, dout.stream), "clamp");
}
void clamp_grad(Tensor &din, const Tensor &in, const Tensor &dout, const real1 &l, const real1 &h) {
  validate_all_same_device({&din, &in, &dout}, "ClampKernel::clamp_grad");
  const tcapint n = din.get_broadcast_size();
  if ((n != in.get_broadcast_size()) || (n != dout.get_broadcast_size())) throw std::invalid_argument("In Weed::clamp_grad(din, in, dout, l, h), sizes do not match!");
  if (in.storage->dtype != DType::REAL) throw std::invalid_argument("In Weed::clamp_grad(din, in, dout, l, h), 'in' dtype must be real-number!");
  const Dev dd = dev_out(din, "clamp_grad"), di = dev_of(in, "clamp_grad"), dg = dev_of(dout, "clamp_grad");
  weedcu_view dv = din.view(), iv = in.view(), gv = dout.view();
  conform_views({&dv, &iv, &gv}, "clamp_grad");
  throw_on_error(weedcu_clamp_grad_real(dd.ptr, &dv, di.ptr, &iv, dg.ptr, &gv, l, h, dd.stream), "clamp_grad");
}
static void extremum_full(int is_min, const Tensor &a, Tensor &out) {
  validate_all_same_device({&a, &out}, "RealExtremumKernel::extremum");
  if ((a.storage->dtype == DType::COMPLEX) || (out.storage->dtype == DType::COMPLEX)) throw std::invalid_argument("Cannot apply extremum reduction on complex tensors!");
  const Dev da = dev_of(a, "extremum"), dout = dev_out(out, "extremum");
  const weedcu_view av = a.view();
  throw_on_error(weedcu_extremum_real(is_min, da.ptr, &av, dout.ptr + out.offset, dout.stream), "extremum");
}
static void extremum_full_grad(Tensor &din, const Tensor &in, const Tensor &dout, const Tensor &out) {
  validate_all_same_device({&din, &in, &dout, &out}, "RealExtremumKernel::extremum_grad");
  if ((in.storage->dtype != DType::REAL) || (out.storage->dtype != DType::REAL))
    throw std::invalid_argument("In RealExtremumKernel::extremum_grad(din, in, dout), in and out dtype must be real-number!");
  const tcapint n = din.get_broadcast_size();
  if ((n != in.get_broadcast_size()) || (n != dout.get_broadcast_size())) throw std::invalid_argument("In Weed::extremum_grad(din, in, dout), sizes do not match!");
  const Dev dd = dev_out(din, "extremum_grad"), di = dev_of(in, "extremum_grad"), dg = dev_of(dout, "extremum_grad"), dm = dev_of(out, "extremum_grad");
  weedcu_view dv = din.view(), iv = in.view(), gv = dout.view();
  conform_views({&dv, &iv, &gv}, "extremum_grad");
  throw_on_error(weedcu_match_grad_full_real(dd.ptr, &dv, di.ptr, &iv, dg.ptr, &gv, dm.ptr + out.offset, dd.stream), "extremum_grad");
}
void max(const Tensor &a, Tensor &out) { extremum_full(0, a, out); }
void min(const Tensor &a, Tensor &out) { extremum_full(1, a, out); }
void max_grad(Tensor &din, const Tensor &in, const Tensor &dout, const Tensor &out) { extremum_full_grad(din, in, dout, out); }
void min_grad(Tensor &din, const Tensor &in, const Tensor &dout, const Tensor &out) { extremum_full_grad(din, in, dout, out); }
static void extremum_axis(int is_min, const tcapint &index, const Tensor &a, Tensor &out, const char *name) {
  validate_all_same_device({&a, &out}, name);
  if ((a.storage->dtype != DType::REAL) || (out.storage->dtype != DType::REAL)) throw std::invalid_argument("Tensor dtype mismatch in max / min!");
  if (index >= a.shape.size()) throw std::invalid_argument("max / min: axis out of range");
  const Dev da = dev_of(a, name), dout = dev_out(out, name);
  const weedcu_view av = a.view();
  // the output is written as a dense buffer starting at out.offset (the tensor Tensor::max(axis) builds, like Tensor::sum)
  throw_on_error(weedcu_extremum_axis_real(is_min, da.ptr, &av, (int)index, dout.ptr + out.offset, backend_config().ref_index_quirks ? 1 : 0, dout.stream), name);
}
void max(const tcapint &index, const Tensor &a, Tensor &out) { extremum_axis(0, index, a, out, "ReduceKernel::max"); }
void min(const tcapint &index, const Tensor &a, Tensor &out) { extremum_axis(1, index, a, out, "ReduceKernel::min"); }
void match_grad(const tcapint &index, Tensor &din, const Tensor &in, const Tensor &dout, const Tensor &out) {
  validate_all_same_device({&din, &dout}, "ReduceKernel::match_grad");
  const tcapint n = din.get_broadcast_size();
  if ((n != in.get_broadcast_size()) || (n != dout.get_broadcast_size())) throw std::invalid_argument("In Weed::match_grad(din, in, dout, out), sizes do not match!");
  const Dev dd = dev_out(din, "match_grad"), di = dev_of(in, "match_grad"), dg = dev_of(dout, "match_grad"), dm = dev_of(out, "match_grad");
  const weedcu_view dv = din.view(), iv = in.view();
  weedcu_view gv = dout.view();
  // the reduced values live at out.offset in the dense layout dout's view describes (the caller match_shape'd dout to din);
  // both are read through dout's strides, each from its own buffer
  const uint64_t g_off = gv.offset;
  gv.offset = 0U;
  throw_on_error(weedcu_match_grad_real(dd.ptr, &dv, di.ptr, &iv, dg.ptr + g_off, &gv, dm.ptr + out.offset, (int)index, backend_config().ref_index_quirks ? 1 : 0,
                                        dd.stream),
                 "match_grad");
}

static void softmax_fwd(int log_mode, const tcapint &index, const Tensor &a, Tensor &out, const char *name) {
  validate_all_same_device({&a, &out}, name);
  require_real(a, "Tensor dtype mismatch in softmax_forward!");
  const Dev da = dev_of(a, name), dout = dev_out(out, name, true);
  const weedcu_view av = a.view(), ov = out.view();
  throw_on_error(weedcu_softmax_real(log_mode, da.ptr, &av, (int)index, dout.ptr, &ov, dout.stream), name);
}
static void softmax_bwd(int log_mode, const tcapint &index, Tensor &din, const Tensor &out, const Tensor &dout, const char *name) {
  validate_all_same_device({&din, &out, &dout}, name);
  const Dev dd = dev_out(din, name), dy = dev_of(out, name), dg = dev_of(dout, name);
  const weedcu_view dv = din.view(), yv = out.view(), gv = dout.view();
  throw_on_error(weedcu_softmax_grad_real(log_mode, dd.ptr, &dv, dy.ptr, &yv, dg.ptr, &gv, (int)index, dd.stream), name);
}
void softmax(const tcapint &index, const Tensor &a, Tensor &out) { softmax_fwd(0, index, a, out, "SoftmaxKernel::softmax_forward"); }
void logsoftmax(const tcapint &index, const Tensor &a, Tensor &out) { softmax_fwd(1, index, a, out, "LogSoftmaxKernel::logsoftmax_forward"); }
void softmax_grad(const tcapint &index, Tensor &din, const Tensor &out, const Tensor &dout) {
  softmax_bwd(0, index, din, out, dout, "SoftmaxKernel::softmax_backward");
}
void logsoftmax_grad(const tcapint &index, Tensor &din, const Tensor &out, const Tensor &dout) {
  softmax_bwd(1, index, din, out, dout, "LogSoftmaxKernel::logsoftmax_backward");
}

void matmul(const Tensor &a, const Tensor &b, Tensor &out) { matmul_impl(a, b, out, 0); }
void matmul_accumulate(const Tensor &a, const Tensor &b, Tensor &out) { matmul_impl(a, b, out, 1); }
bool matmul_bias(const Tensor &a, const Tensor &b, const Tensor &bias, Tensor &out, const Tensor *residual) {
  return matmul_impl(a, b, out, 0, &bias, residual);
}
bool matmul_bias_grouped(const Tensor &a, const std::vector<const Tensor *> &ws, const std::vector<const Tensor *> &biases,
                         const std::vector<Tensor *> &outs, bool bf16_only) {
  const BackendConfig &cfg = backend_config();
  const size_t G = ws.size();
  if (cfg.matmul_precision != WEEDCU_GEMM_BF16 || !cfg.fused || !cfg.operand_cache || G < 2U || G > 3U || biases.size() != G || outs.size() != G) return false;
  if (a.shape.size() != 2U) return false;
  const tcapint M = a.shape[0U], K = a.shape[1U], N = ws[0]->shape[1U];
  if (M < 64U || N < 16U || K < 32U) return false;
  for (size_t g = 0U; g < G; ++g) {
    const Tensor &w = *ws[g], &o = *outs[g], &bi = *biases[g];
    if (w.shape.size() != 2U || w.shape[0U] != K || w.shape[1U] != N || w.stride != ws[0]->stride) return false;
    if (o.shape.size() != 2U || o.shape[0U] != M || o.shape[1U] != N || o.stride[0U] != 1U || o.stride[1U] != outs[0]->stride[1U] || !covers_storage(o)) return false;
    if (bi.storage->size != N || bi.storage->device != DeviceTag::GPU) return false;
  }
  validate_all_same_device({&a, ws[0], outs[0]}, "matmul_bias_grouped");
  Bf16Operand pa, pb[3];
  if (!bf16_operand(a, a.stride[0U], a.stride[1U], M, K, true, pa)) return false;
  for (size_t g = 0U; g < G; ++g) {
    if (!bf16_operand(*ws[g], ws[g]->stride[1U], ws[g]->stride[0U], N, K, false, pb[g])) return false;
    if (pb[g].major != pb[0].major || pb[g].ld != pb[0].ld) return false;
  }
  const uint16_t *bptr[3];
  real1 *cptr[3];
  const real1 *biasptr[3];
  void *stream = nullptr;
  if (bf16_only && cfg.epilogue_stats && cfg.defer_grads && (M % 8U) == 0U && N >= 32U) {
    // the outputs' only reader is the attention core's head relayout, which takes the bf16 operand copies: the products
    // write those and nothing else (fp32 values deferred, like the LM head's logits)
    OutputShadow os[3];
    uint16_t *c16[3];
    bool ok = true;
    for (size_t g = 0U; g < G && ok; ++g) {
      os[g] = begin_output_shadow(*outs[g], N);
      ok = os[g].ptr != nullptr;
      c16[g] = os[g].ptr;
      bptr[g] = pb[g].ptr;
      biasptr[g] = dev_of(*biases[g], "matmul_bias_grouped").ptr + biases[g]->offset;
    }
    if (ok) {
      GpuRealStorage *first = gpu_storage(*outs[0], "matmul_bias_grouped");
      first->dev->Bind();
      for (size_t g = 0U; g < G; ++g) gpu_storage(*outs[g], "matmul_bias_grouped")->device_ptr_overwrite();
      const int rc = weedcu_gemm_bf16_grouped_bf16out(pa.ptr, pa.major, pa.ld, (uint32_t)G, bptr, pb[0].major, pb[0].ld, c16, M, M, N, K, biasptr, first->dev->stream);
      if (rc == 0) {
        for (size_t g = 0U; g < G; ++g) {
          end_output_shadow(os[g]);
          defer_linear_output(gpu_storage(*outs[g], "matmul_bias_grouped"), pa, pb[g], M, N, K, *ws[g], *biases[g]);
        }
        return true;
      }
      if (rc != WEEDCU_ENOSUP) throw_on_error(rc, "matmul_bias_grouped");
    }
  }
  for (size_t g = 0U; g < G; ++g) {
    const Dev dc = dev_out(*outs[g], "matmul_bias_grouped", true);
    stream = dc.stream;
    bptr[g] = pb[g].ptr;
    cptr[g] = dc.ptr + outs[g]->offset;
    biasptr[g] = dev_of(*biases[g], "matmul_bias_grouped").ptr + biases[g]->offset;
  }
  const int rc = weedcu_gemm_bf16_grouped(pa.ptr, pa.major, pa.ld, (uint32_t)G, bptr, pb[0].major, pb[0].ld, cptr, outs[0]->stride[1U], M, N, K, 0, biasptr, stream);
  if (rc == WEEDCU_ENOSUP) return false;
  throw_on_error(rc, "matmul_bias_grouped");
  return true;
}
// The same for a handful of rows (a decode step): one skinny launch for the group, fp32 in either precision mode.
bool matmul_skinny_grouped(const Tensor &a, const std::vector<const Tensor *> &ws, const std::vector<const Tensor *> &biases,
                           const std::vector<Tensor *> &outs) {
  const BackendConfig &cfg = backend_config();
  const size_t G = ws.size();
  if (!cfg.fused || G < 2U || G > 3U || biases.size() != G || outs.size() != G || a.shape.size() != 2U) return false;
  const tcapint M = a.shape[0U], K = a.shape[1U], N = ws[0]->shape[1U];
  if (M > 16U) return false;
  for (size_t g = 0U; g < G; ++g) {
    const Tensor &w = *ws[g], &o = *outs[g], &bi = *biases[g];
    if (w.shape.size() != 2U || w.shape[0U] != K || w.shape[1U] != N || w.stride != ws[0]->stride || w.offset != ws[0]->offset) return false;
    if (o.shape.size() != 2U || o.shape[0U] != M || o.shape[1U] != N || o.stride != outs[0]->stride || o.offset != outs[0]->offset) return false;
    if (bi.storage->size != N || bi.storage->device != DeviceTag::GPU) return false;
  }
  validate_all_same_device({&a, ws[0], outs[0]}, "matmul_skinny_grouped");
  const Dev da = dev_of(a, "matmul_skinny_grouped");
  const weedcu_mat am = mat_of(a), bm = mat_of(*ws[0]), cm = mat_of(*outs[0]);
  const real1 *bptr[3], *biasptr[3];
  real1 *cptr[3];
  void *stream = nullptr;
  for (size_t g = 0U; g < G; ++g) {
    bptr[g] = dev_of(*ws[g], "matmul_skinny_grouped").ptr;
    const Dev dc = dev_out(*outs[g], "matmul_skinny_grouped", true);
    stream = dc.stream;
    cptr[g] = dc.ptr;
    biasptr[g] = dev_of(*biases[g], "matmul_skinny_grouped").ptr + biases[g]->offset;
  }
  throw_on_error(weedcu_matmul_skinny_grouped(da.ptr, &am, (uint32_t)G, bptr, &bm, cptr, &cm, M, K, N, biasptr, stream), "matmul_skinny_grouped");
  return true;
}
bool pack_with_column_sums(const Tensor &dy, Tensor &sums) {
  const BackendConfig &cfg = backend_config();
  if (cfg.matmul_precision != WEEDCU_GEMM_BF16 || !cfg.fused || !cfg.operand_cache || dy.shape.size() != 2U) return false;
  const tcapint M = dy.shape[0U], N = dy.shape[1U];
  // only when the packed layout keeps the row index contiguous (colsum then runs over rows) and the
  // GEMMs that follow will use this very shadow (same eligibility as matmul_impl)
  if (dy.stride[0U] != 1U || M < 64U || N < 16U || sums.storage->size != N || !covers_storage(sums)) return false;
  GpuRealStorage *ss = gpu_storage(sums, "pack_with_column_sums");
  const int accumulate = ss->zero_pending ? 0 : 1;
  Bf16Operand op;
  // probe first: device_ptr_overwrite() below drops the pending zero fill, so the pack must happen
  GpuRealStorage *ds = gpu_storage(dy, "pack_with_column_sums");
  for (const GpuRealStorage::Bf16Shadow &sh : ds->shadows)
    if (sh.offset == dy.offset && sh.n_fast == M && sh.n_slow == N && sh.version == ds->version) {
      // already packed; if its producer also left the column sums, they only need adding in
      if (!ds->colsum || ds->colsum_version != ds->version || ds->colsum_n != N) return false;
      ds->dev->Bind();
      if (accumulate) {
        weedcu_view v;
        v.offset = 0U;
        v.rank = 1;
        v.shape[0] = N;
        v.stride[0] = 1U;
        weedcu_view sv = v;
        sv.offset = sums.offset;
        throw_on_error(weedcu_inplace_real(WEEDCU_ADD, ss->device_ptr(), &sv, (const real1 *)ds->colsum->ptr, &v, ds->dev->stream), "pack_with_column_sums");
      } else {
        throw_on_error(weedcu_memcpy_d2d(ss->device_ptr_overwrite() + sums.offset, ds->colsum->ptr, sizeof(real1) * (size_t)N, ds->dev->stream),
                       "pack_with_column_sums");
      }
      return true;
    }
  if ((M % 8U) || (dy.stride[1U] % 4U) || (dy.offset % 4U)) return false;
  real1 *sp = accumulate ? ss->device_ptr() : ss->device_ptr_overwrite();
  if (!bf16_operand(dy, dy.stride[0U], dy.stride[1U], M, N, true, op, sp + sums.offset, accumulate))
    throw std::runtime_error("pack_with_column_sums: streaming pack refused after the eligibility check");
  return true;
}

bool cross_entropy_bwd_pack(const Tensor &logits, const SymbolTensor &targets, const Tensor &lse, const Tensor &dloss, Tensor &dlogits,
                            tcapint rows, tcapint V) {
  const BackendConfig &cfg = backend_config();
  if (cfg.matmul_precision != WEEDCU_GEMM_BF16 || !cfg.fused || !cfg.operand_cache) return false;
  // same eligibility as the tensor-core GEMMs that will read the shadow (matmul_impl) + the kernel's alignment rules
  if ((rows % 8U) || rows < 64U || V < 16U || (dlogits.offset % 4U) || (logits.offset % 4U)) return false;
  GpuRealStorage *ds = gpu_storage(dlogits, "cross_entropy_bwd_pack");
  const int accumulate = (ds->zero_pending && covers_storage(dlogits)) ? 0 : 1;
  GpuRealStorage::Bf16Shadow *hit = nullptr;
  for (GpuRealStorage::Bf16Shadow &sh : ds->shadows)
    if (sh.offset == dlogits.offset && sh.n_fast == rows && sh.n_slow == V && sh.s_fast == 1U && sh.s_slow == rows) hit = &sh;
  if (!hit) {
    if (ds->shadows.size() >= 4U) ds->shadows.erase(ds->shadows.begin());
    ds->shadows.push_back(GpuRealStorage::Bf16Shadow{ds->dev->MakeBuffer(2U * ((size_t)rows * V + 8U)), 0U, dlogits.offset, rows, V, 1U, rows});
    hit = &ds->shadows.back();
  }
  if (!ds->colsum || ds->colsum_n != V) {
    ds->colsum = ds->dev->MakeBuffer(sizeof(real1) * (size_t)V);
    ds->colsum_n = V;
  }
  // logits whose fp32 values are still deferred (the LM head's epilogue wrote the bf16 copy only): read that copy
  GpuRealStorage *lgs = gpu_storage(logits, "cross_entropy_bwd_pack");
  const uint16_t *logits_bf16 = nullptr;
  if (lgs->deferred_values && !logits.offset)
    for (const GpuRealStorage::Bf16Shadow &sh : lgs->shadows)
      if (sh.offset == 0U && sh.n_fast == rows && sh.n_slow == V && sh.s_fast == 1U && sh.s_slow == rows && sh.version == lgs->version)
        logits_bf16 = (const uint16_t *)sh.buf->ptr;
  const Dev dlse = dev_of(lse, "cross_entropy_bwd_pack"), dg = dev_of(dloss, "cross_entropy_bwd_pack");
  const bool defer = !accumulate && cfg.defer_grads && covers_storage(dlogits);
  real1 *out = accumulate ? ds->device_ptr() : ds->device_ptr_overwrite();
  int rc;
  if (logits_bf16)
    rc = weedcu_cross_entropy_bwd_pack_bf16in(logits_bf16, rows, V, sym_ptr(targets, "cross_entropy_bwd_pack") + targets.offset, dlse.ptr + lse.offset,
                                              dg.ptr + dloss.offset, defer ? nullptr : out, dlogits.offset, accumulate, (uint16_t *)hit->buf->ptr,
                                              (real1 *)ds->colsum->ptr, ds->dev->stream);
  else
    rc = weedcu_cross_entropy_bwd_pack(dev_of(logits, "cross_entropy_bwd_pack").ptr, logits.offset, rows, V, sym_ptr(targets, "cross_entropy_bwd_pack") + targets.offset,
                                       dlse.ptr + lse.offset, dg.ptr + dloss.offset, defer ? nullptr : out, dlogits.offset, accumulate,
                                       (uint16_t *)hit->buf->ptr, (real1 *)ds->colsum->ptr, ds->dev->stream);
  if (rc == WEEDCU_ENOSUP) {
    // nothing was launched; the caller's plain kernel must see the pending zero fill again if we dropped it
    if (!accumulate) ds->FillZeros();
    return false;
  }
  throw_on_error(rc, "cross_entropy_bwd_pack");
  hit->version = ds->version;
  ds->colsum_version = ds->version;
  if (defer) {
    // the LM head's backward reads dlogits only through the bf16 shadow and the column sums: 1.65 GB of fp32 per
    // step at the GPT-2 shape are written only if something else reads this gradient
    StoragePtr l_s = logits.storage, t_s = targets.storage, lse_s = lse.storage, g_s = dloss.storage;
    const tcapint l_off = logits.offset, t_off = targets.offset, lse_off = lse.offset, g_off = dloss.offset, d_off = dlogits.offset;
    const uint64_t l_v = static_cast<GpuRealStorage *>(l_s.get())->version, g_v = static_cast<GpuRealStorage *>(g_s.get())->version;
    ds->deferred_values = [ds, l_s, t_s, lse_s, g_s, l_off, t_off, lse_off, g_off, d_off, rows, V, l_v, g_v]() {
      if (static_cast<GpuRealStorage *>(l_s.get())->version != l_v || static_cast<GpuRealStorage *>(g_s.get())->version != g_v)
        throw std::runtime_error("cross_entropy_loss: logits or the loss gradient were modified before the deferred fp32 gradient was read");
      ds->dev->Bind();
      throw_on_error(weedcu_cross_entropy_bwd(static_cast<GpuRealStorage *>(l_s.get())->device_ptr_ro(), l_off, rows, V, 1U, rows,
                                              static_cast<GpuIntStorage *>(t_s.get())->device_ptr_ro() + t_off,
                                              static_cast<GpuRealStorage *>(lse_s.get())->device_ptr_ro() + lse_off,
                                              static_cast<GpuRealStorage *>(g_s.get())->device_ptr_ro() + g_off, (real1 *)ds->buffer->ptr, d_off, 0,
                                              ds->dev->stream),
                     "cross_entropy_loss backward (deferred)");
    };
  }
  return true;
}

// the bf16 attention entry; q / k / v whose fp32 values are deferred behind a current bf16 copy are read through that copy
int attention_forward_bf16(const Tensor &q, const Tensor &k, const Tensor &v, Tensor &out, uint16_t *out_bf16, tcapint B, tcapint T, tcapint H, tcapint hd,
                           real1 divisor, real1 mask_val, int causal) {
  auto current_copy = [&](const Tensor &t) -> const uint16_t * {
    GpuRealStorage *s = gpu_storage(t, "attention");
    const tcapint rows = B * T, cols = H * hd;
    if (!s->deferred_values || t.offset) return nullptr;
    for (const GpuRealStorage::Bf16Shadow &sh : s->shadows)
      if (sh.offset == 0U && sh.n_fast == rows && sh.n_slow == cols && sh.s_fast == 1U && sh.s_slow == rows && sh.version == s->version)
        return (const uint16_t *)sh.buf->ptr;
    return nullptr;
  };
  const uint16_t *q16 = current_copy(q), *k16 = current_copy(k), *v16 = current_copy(v);
  const Dev dout = dev_out(out, "attention", true);
  if (q16 && k16 && v16) {
    const int rc = weedcu_attention_fwd_bf16in(q16, k16, v16, dout.ptr, out_bf16, B, T, H, hd, divisor, mask_val, causal, dout.stream);
    if (rc != WEEDCU_ENOSUP) return rc;
  }
  return weedcu_attention_fwd_bf16out(dev_of(q, "attention").ptr + q.offset, dev_of(k, "attention").ptr + k.offset, dev_of(v, "attention").ptr + v.offset,
                                      dout.ptr, out_bf16, B, T, H, hd, divisor, mask_val, causal, dout.stream);
}

void matmul_batched(const Tensor &a3, const Tensor &b3, Tensor &out3) {
  validate_all_same_device({&a3, &b3, &out3}, "MatMulKernel::matmul_batched");
  if ((a3.shape.size() != 3U) || (b3.shape.size() != 3U) || (out3.shape.size() != 3U))
    throw std::invalid_argument("matmul_batched is for [batch, M, K] x [batch, K, N] views");
  const tcapint batch = a3.shape[0U], M = a3.shape[1U], K = a3.shape[2U], N = b3.shape[2U];
  if ((b3.shape[0U] != batch) || (out3.shape[0U] != batch)) throw std::invalid_argument("batched matmul batch mismatch");
  if (b3.shape[1U] != K) throw std::invalid_argument("batched matmul inner dim mismatch");
  if ((out3.shape[1U] != M) || (out3.shape[2U] != N)) throw std::invalid_argument("MatMul output dimensions don't match inputs!");
  const Dev da = dev_of(a3, "matmul_batched"), db = dev_of(b3, "matmul_batched"), dc = dev_out(out3, "matmul_batched");
  const weedcu_mat am = mat_of(a3, 0U, a3.stride[0U]), bm = mat_of(b3, 0U, b3.stride[0U]), cm = mat_of(out3, 0U, out3.stride[0U]);
  throw_on_error(weedcu_matmul_real(da.ptr, &am, db.ptr, &bm, dc.ptr, &cm, M, K, N, batch, 0,
                                    backend_config().matmul_precision, dc.stream),
                 "matmul_batched");
}

// Token stride / output row stride. The reference reads stride[0] (src/ops/embedding.cpp:66-75),
// which is 0 when the leading extent is 1 (e.g. indices [1, T]) and then gathers token 0 only;
// for dense tensors the flat token index has stride 1, which is what is meant.
static tcapint flat_stride(const BaseTensor &t) {
  if (backend_config().ref_index_quirks) return t.stride[0U]; // defect D6 reproduced: [1, T] indices read token 0 for every slot
  tcapint expect = 1U;
  bool dense = true;
  for (size_t i = 0U; i < t.shape.size(); ++i) {
    if (t.shape[i] == 1U) continue;
    if (t.stride[i] != expect) dense = false;
    expect *= t.shape[i];
  }
  return dense ? 1U : t.stride[0U];
}
static tcapint row_stride_of_rows(const BaseTensor &t) { // all dims but the last form the token index
  if (backend_config().ref_index_quirks) return t.stride[0U]; // (and write every row to slot 0 when the leading extent is 1)
  BaseTensor lead;
  lead.shape.assign(t.shape.begin(), t.shape.end() - 1);
  lead.stride.assign(t.stride.begin(), t.stride.end() - 1);
  if (lead.shape.empty()) return t.stride[0U];
  return flat_stride(lead);
}
static const symint *sym_ptr(const SymbolTensor &s, const char *op) {
  if (s.storage->device != DeviceTag::GPU) throw std::domain_error(std::string(op) + ": indices must be GPU-resident");
  return s.device_ptr();
}
void embedding_gather(const SymbolTensor &indices, const Tensor &weight, Tensor &out) {
  validate_all_same_device({&indices, &weight, &out}, "embedding_gather");
  const Dev dw = dev_of(weight, "embedding_gather"), dout = dev_out(out, "embedding_gather");
  const tcapint D = weight.shape[1U], n = indices.get_broadcast_size();
  throw_on_error(weedcu_embedding_gather(sym_ptr(indices, "embedding_gather"), indices.offset, flat_stride(indices), n, dw.ptr,
                                         weight.offset, weight.stride[0U], weight.stride[1U], D, dout.ptr, out.offset,
                                         row_stride_of_rows(out), out.stride.back(), dout.stream),
                 "embedding_gather");
}
void embedding_scatter_add(Tensor &dW, const SymbolTensor &indices, const Tensor &dout) {
  validate_all_same_device({&dW, &indices, &dout}, "embedding_scatter_add");
  const Dev dw = dev_out(dW, "embedding_scatter_add"), dg = dev_of(dout, "embedding_scatter_add");
  const tcapint D = dW.shape[1U], n = indices.get_broadcast_size();
  throw_on_error(weedcu_embedding_scatter_add(dw.ptr, dW.offset, dW.stride[0U], dW.stride[1U], sym_ptr(indices, "embedding_scatter_add"),
                                              indices.offset, flat_stride(indices), n, D, dg.ptr, dout.offset, row_stride_of_rows(dout),
                                              dout.stride.back(), dw.stream),
                 "embedding_scatter_add");
}
SymbolTensorPtr argmax_last_token(const Tensor &logits) {
  if (logits.shape.size() != 3U) throw std::invalid_argument("argmax_last_token expects logits [B, T, V]");
  const tcapint B = logits.shape[0U], T = logits.shape[1U], V = logits.shape[2U];
  const Dev dl = dev_of(logits, "argmax_last_token");
  SymbolTensorPtr out = std::make_shared<SymbolTensor>(std::vector<tcapint>{B, 1U}, std::vector<tcapint>{1U, B}, false, DeviceTag::GPU,
                                                       logits.storage->get_device_id(), false);
  GpuIntStorage *os = dynamic_cast<GpuIntStorage *>(out->storage.get());
  if (!os) throw std::domain_error("argmax_last_token: output is not on the GPU");
  // a size-1 extent carries stride 0 (base_tensor.hpp:282-292): use the dense strides of [B, T, V]
  const tcapint sb = B > 1U ? logits.stride[0U] : 1U, st = T > 1U ? logits.stride[1U] : 0U, sv = V > 1U ? logits.stride[2U] : 0U;
  throw_on_error(weedcu_argmax_rows(dl.ptr, (uint64_t)logits.offset + (uint64_t)(T - 1U) * st, B, V, sb, sv, os->device_ptr_overwrite(), dl.stream),
                 "argmax_last_token");
  return out;
}
void triu_fill(Tensor &a, const complex &val, const tcapint diagonal) {
  if (a.shape.size() != 2U) throw std::invalid_argument("triu_fill requires a 2D tensor!");
  const Dev da = dev_out(a, "triu_fill");
  const weedcu_view av = a.view();
  throw_on_error(weedcu_triu_fill_real(da.ptr, &av, val.real(), diagonal, da.stream), "triu_fill");
}
} // namespace Weed
