// tensor.cpp — Tensor front-ends and reverse-mode autograd for the CUDA device.
//
// Same public behaviour as the reference's src/tensors/tensor.cpp (cited per function): each
// front-end allocates `out`, calls the Weed:: op, and records a Node whose closure accumulates into
// the parents' gradients; Tensor::backward walks the DFS topological order in reverse
// (tensor.cpp:371-401). Written fresh: one `Ctx` helper replaces the per-closure device/dtype
// juggling, because on this backend every tensor already lives on the GPU (cast() is a view copy).
//
// Fused device paths (backend_config().fused): Tensor::gelu is one kernel forward and one kernel
// backward instead of 9 + ~20 (tensor.cpp:841-851); the matmul node accumulates dA / dB inside the
// GEMM epilogue instead of tmp + add_in_place (tensor.cpp:1361-1400).
#include <unordered_map>
#include "weed_b200/ops.hpp"

#include <algorithm>
#include <cmath>
#include <unordered_set>

namespace Weed {
namespace {
thread_local bool g_skip_zero_fill = false;
struct SkipFillGuard {
  bool prev;
  SkipFillGuard() : prev(g_skip_zero_fill) { g_skip_zero_fill = true; }
  ~SkipFillGuard() { g_skip_zero_fill = prev; }
};
inline DeviceTag concrete(DeviceTag t) { return (t == DeviceTag::CPU) ? DeviceTag::CPU : DeviceTag::GPU; }
inline TensorPtr view_copy(const TensorPtr &t) { return std::make_shared<Tensor>(*t); }
std::vector<TensorPtr> grad_parents(const std::vector<TensorPtr> &parents) {
  std::vector<TensorPtr> out;
  for (const TensorPtr &p : parents)
    if (p->requires_grad) out.push_back(p);
  return out;
}
// Gradient of `t` viewed at t's full shape. In the fused configuration broadcast tensors keep their
// gradient at real extent (Tensor::make_gradient); ops that need it element-for-element with `t`
// materialise it here and settle_grad() sums it back, exactly the reference's own sequence.
TensorPtr full_grad(const TensorPtr &t) {
  TensorPtr g = view_copy(t->grad);
  if (g->get_broadcast_size() != t->get_broadcast_size()) {
    g->match_shape(t);
    g->materialize_broadcast();
  }
  return g;
}
void settle_grad(const TensorPtr &t, const TensorPtr &g) {
  t->grad = g;
  bool widened = false;
  for (size_t i = 0U; i < t->stride.size() && i < g->shape.size(); ++i)
    if (!t->stride[i] && g->shape.size() == t->shape.size() && g->shape[i] > 1U) widened = true;
  if (widened) t->reduce_grad_broadcast();
}
symint wrap_axis(symint axis, size_t rank) {
  while (axis < 0) axis += (symint)rank;
  return axis;
}
} // namespace

// ------------------------------------------------------------------------------ construction
Tensor::Tensor(const std::vector<tcapint> &shp, const std::vector<tcapint> &, const bool &rg, const bool &, const DType &dtype,
               const DeviceTag &dtag, const int64_t &did)
    : BaseTensor(shp, full_contiguous_stride(shp)), grad_node(nullptr), requires_grad(rg) {
  if (dtype == DType::INT) throw std::invalid_argument("Tensor cannot have DType::INT! (INT is only for SymbolTensor, not arithmetic Tensor.)");
  if (dtype == DType::COMPLEX) throw std::invalid_argument("Complex tensors are outside the CUDA backend's scope (SURVEY §8)");
  const tcapint size = get_size();
  if (concrete(dtag) == DeviceTag::GPU) {
    storage = std::make_shared<GpuRealStorage>(size, did);
    if (!g_skip_zero_fill) storage->FillZeros(); // reference zero-fills every dense allocation, tensor.cpp:204
  } else {
    storage = std::make_shared<CpuRealStorage>(size); // host staging, value-initialised
  }
}
Tensor::Tensor(const std::vector<real1> &val, const std::vector<tcapint> &shp, const bool &rg, const DeviceTag &dtag, const int64_t &did)
    : BaseTensor(shp, full_contiguous_stride(shp)), grad_node(nullptr), requires_grad(rg) {
  if (get_size() != val.size())
    throw std::invalid_argument("Tensor value initializer vector must have same size as implied by shape and stride!");
  if (concrete(dtag) == DeviceTag::GPU) storage = std::make_shared<GpuRealStorage>(val, did);
  else storage = std::make_shared<CpuRealStorage>(val);
}

real1 *Tensor::device_ptr() const {
  if (!storage || storage->device != DeviceTag::GPU) throw std::domain_error("Tensor is not GPU-resident");
  return static_cast<GpuRealStorage *>(storage.get())->device_ptr();
}
const real1 *Tensor::device_ptr_ro() const {
  if (!storage || storage->device != DeviceTag::GPU) throw std::domain_error("Tensor is not GPU-resident");
  return static_cast<GpuRealStorage *>(storage.get())->device_ptr_ro();
}
real1 *Tensor::device_ptr_accumulate(int &accumulate) const {
  if (!storage || storage->device != DeviceTag::GPU) throw std::domain_error("Tensor is not GPU-resident");
  GpuRealStorage *s = static_cast<GpuRealStorage *>(storage.get());
  if (s->zero_pending && covers_storage()) {
    accumulate = 0;
    return s->device_ptr_overwrite();
  }
  accumulate = 1;
  return s->device_ptr();
}
real1 *Tensor::device_ptr_accumulate_from(const real1 *&src, BufferPtr &keep) const {
  if (!storage || storage->device != DeviceTag::GPU) throw std::domain_error("Tensor is not GPU-resident");
  GpuRealStorage *s = static_cast<GpuRealStorage *>(storage.get());
  if (!s->zero_pending && !s->deferred_values && s->buffer_shared() && covers_storage()) {
    keep = s->buffer; // the other sharers' buffer: read from it, write the sums into a fresh private one
    src = reinterpret_cast<const real1 *>(keep->ptr);
    return s->device_ptr_overwrite();
  }
  int accumulate = 1;
  real1 *p = device_ptr_accumulate(accumulate);
  src = accumulate ? p : nullptr;
  return p;
}
bool BaseTensor::covers_storage() const {
  if (offset || !storage) return false;
  tcapint expect = 1U;
  for (size_t i = 0U; i < shape.size(); ++i) {
    if (shape[i] == 1U) continue;
    if (stride[i] != expect) return false;
    expect *= shape[i];
  }
  return expect == storage->size;
}
void *Tensor::stream() const {
  if (!storage || storage->device != DeviceTag::GPU) throw std::domain_error("Tensor is not GPU-resident");
  return static_cast<GpuRealStorage *>(storage.get())->dev->stream;
}

TensorPtr Tensor::clone(const TensorPtr &a) {
  TensorPtr z = zeros(a->shape, false, false, a->storage->dtype, a->storage->device, a->storage->get_device_id());
  return add(z, a);
}
void Tensor::upcast(const DType &dt) { storage = storage->Upcast(dt); }
TensorPtr Tensor::cast(const DeviceTag &dt) const {
  TensorPtr cp = std::make_shared<Tensor>(*this);
  if (dt == DeviceTag::CPU) cp->storage = cp->storage->cpu();
  else if (dt == DeviceTag::GPU) cp->storage = cp->storage->gpu();
  return cp;
}
void Tensor::cast_in_place(const DeviceTag &dt) {
  if (dt == DeviceTag::CPU) storage = storage->cpu();
  else if (dt == DeviceTag::GPU) storage = storage->gpu();
}
void Tensor::squeeze() { // tensor.hpp:153-164
  for (size_t i = 0U; i < shape.size(); ++i) {
    if (shape.size() == 1U) break;
    const size_t j = shape.size() - (i + 1U);
    if (shape[j] == 1U) {
      shape.erase(shape.begin() + (long)j);
      stride.erase(stride.begin() + (long)j);
    }
  }
}
void Tensor::squeeze(int64_t axis) {
  if (shape.size() == 1U) return;
  while (axis < 0) axis += (int64_t)shape.size();
  if (shape[(size_t)axis] != 1U) throw std::invalid_argument("Can only Tensor::squeeze() dimensions with size of 1!");
  shape.erase(shape.begin() + axis);
  stride.erase(stride.begin() + axis);
}
void Tensor::unsqueeze(int64_t axis) {
  while (axis < 0) axis += (int64_t)shape.size();
  shape.insert(shape.begin() + axis, 1U);
  stride.insert(stride.begin() + axis, 0U);
}

TensorPtr Tensor::zeros(const std::vector<tcapint> &shape, const bool &rg, const bool &s, const DType &dtype, const DeviceTag &dtag,
                        const int64_t &did) {
  TensorPtr z;
  {
    SkipFillGuard g;
    z = std::make_shared<Tensor>(shape, full_contiguous_stride(shape), rg, s, dtype, dtag, did);
  }
  z->storage->FillZeros();
  return z;
}
TensorPtr Tensor::ones_like(const std::vector<tcapint> &shape, const bool &rg, const bool &s, const DType &dtype, const DeviceTag &dtag,
                            const int64_t &did) {
  TensorPtr z;
  {
    SkipFillGuard g;
    z = std::make_shared<Tensor>(shape, full_contiguous_stride(shape), rg, s, dtype, dtag, did);
  }
  z->storage->FillOnes();
  return z;
}
// tensor.cpp:116-129 builds a sparse CPU one-hot; the device path keeps it dense and only uses it
// for the un-fused cross-entropy (small shapes). col-major index: t + tok*T.
TensorPtr Tensor::one_hot(const SymbolTensorPtr targets, const tcapint vocab_size) {
  const tcapint T = targets->get_broadcast_size();
  StoragePtr host = targets->storage->cpu();
  const std::vector<symint> &ids = static_cast<CpuIntStorage *>(host.get())->data;
  std::vector<real1> dense((size_t)T * vocab_size, ZERO_R1);
  for (tcapint t = 0U; t < T; ++t) {
    const tcapint tok = (tcapint)ids[targets->offset + t * targets->stride[0U]];
    dense[(size_t)t + (size_t)tok * T] = ONE_R1;
  }
  return std::make_shared<Tensor>(dense, std::vector<tcapint>{T, vocab_size}, false, targets->storage->device,
                                  targets->storage->get_device_id());
}
TensorPtr Tensor::make_gradient(const std::vector<tcapint> &shp, const bool &s, const DType &dtype, const DeviceTag &dtag, const int64_t did) {
  return zeros(shp, false, s, dtype, dtag, did);
}
void Tensor::make_gradient(const bool &) { // tensor.cpp:78-114 (device choice is not size-based here)
  if (!requires_grad) return;
  std::vector<tcapint> gshape = shape;
  if (backend_config().fused) {
    // A tensor broadcast by match_shape (a bias that became [B,T,F] with strides [0,0,1]) would get
    // a full [B,T,F] gradient in the reference, which every backward then fills, adds into and
    // sums back down (tensor.cpp:1117-1134): 5 passes over B*T*F per bias. The fused path keeps
    // the gradient at the tensor's real extent ([1,1,F]) and reduces contributions into it.
    for (size_t i = 0U; i < gshape.size(); ++i)
      if (!stride[i]) gshape[i] = 1U;
  }
  if (grad && grad->shape == gshape) return;
  grad = Tensor::make_gradient(gshape, false, storage->dtype, storage->device, storage->get_device_id());
}
TensorPtr Tensor::allocate_scalar_like(const Tensor &orig, const bool &rg) {
  return allocate_like(std::vector<tcapint>{1U}, std::vector<tcapint>{0U}, orig, orig.storage->dtype, rg, false);
}
TensorPtr Tensor::allocate_like(const Tensor &orig, const DType &dt, const bool &rg, const bool &s) {
  return allocate_like(orig.shape, full_contiguous_stride(orig.shape), orig, dt, rg, s);
}
TensorPtr Tensor::allocate_like(const std::vector<tcapint> &shp, const Tensor &orig, const DType &dt, const bool &rg, const bool &s) {
  return allocate_like(shp, full_contiguous_stride(shp), orig, dt, rg, s);
}
TensorPtr Tensor::allocate_like(const std::vector<tcapint> &shp, const std::vector<tcapint> &strd, const Tensor &orig, const DType &dt,
                                const bool &rg, const bool &s) {
  // "without Storage value initialization" (tensor.hpp:256-279): every caller overwrites the whole
  // buffer, so the device path skips the reference's redundant zero-fill pass (4 B/elem saved).
  SkipFillGuard g;
  return std::make_shared<Tensor>(shp, strd, rg, s, dt, orig.storage->device, orig.storage->get_device_id());
}
std::vector<TensorPtr> Tensor::chunk(TensorPtr a, const size_t &chunks, int64_t axis) {
  if (chunks == 0) throw std::invalid_argument("Tensor::chunk: chunks must be > 0");
  if (axis < 0) axis += (int64_t)a->shape.size();
  if (axis < 0 || axis >= (int64_t)a->shape.size()) throw std::invalid_argument("Tensor::chunk: axis out of range");
  const tcapint dim = a->shape[(size_t)axis];
  if (dim % chunks) throw std::invalid_argument("Tensor::chunk: dimension not divisible by chunks");
  const tcapint each = dim / (tcapint)chunks;
  std::vector<TensorPtr> out;
  for (size_t i = 0; i < chunks; ++i) {
    TensorPtr t = view_copy(a);
    t->shape[(size_t)axis] = each;
    t->offset += (tcapint)i * each * a->stride[(size_t)axis];
    out.push_back(t);
  }
  return out;
}
TensorPtr Tensor::contiguous(const TensorPtr a) { // tensor.hpp:319-331 (zeros + add); here: one strided copy
  if (is_contiguous(a->shape, a->stride)) return a;
  if (!backend_config().fused) {
    TensorPtr z = zeros(a->shape, false, false, a->storage->dtype, a->storage->device, a->storage->get_device_id());
    return add(z, a);
  }
  TensorPtr out = allocate_like(a->shape, *a, a->storage->dtype, a->requires_grad, false);
  // a broadcast dim (stride 0) keeps stride 0 in full_contiguous_stride only when extent is 1;
  // materialise everything else
  Weed::copy_broadcast(*out, *a);
  if (a->requires_grad) { // gradient flows straight through, like the reference's add node
    out->make_gradient();
    out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out)]() {
      TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
      if (!out) node_owner_lost();
      TensorPtr a_grad = view_copy(a->grad);
      TensorPtr out_grad = view_copy(out->grad);
      a_grad->match_shape(out_grad);
      a_grad->materialize_broadcast();
      Weed::add_in_place(*a_grad, *out_grad);
      a->grad = a_grad;
      a->reduce_grad_broadcast();
    });
  }
  return out;
}
TensorPtr Tensor::reshape(const TensorPtr a, const std::vector<symint> &s) {
  TensorPtr out = is_contiguous(a->shape, a->stride) ? view_copy(a) : contiguous(a);
  out->reshape(s);
  return out;
}
TensorPtr Tensor::transpose(const TensorPtr a) {
  TensorPtr out = view_copy(a);
  out->transpose();
  return out;
}
TensorPtr Tensor::transpose(const TensorPtr a, symint i, symint j) {
  TensorPtr out = view_copy(a);
  out->transpose(i, j);
  return out;
}
TensorPtr Tensor::flatten(const TensorPtr a, symint axis) {
  TensorPtr out = is_contiguous(a->shape, a->stride) ? view_copy(a) : contiguous(a);
  out->flatten(axis);
  return out;
}
TensorPtr Tensor::operator[](const tcapint &idx) const { // tensor.cpp:277-293
  if (idx >= shape.back()) throw std::invalid_argument("Tensor index out-of-range!");
  TensorPtr v = std::make_shared<Tensor>(*this);
  v->offset += idx * stride.back();
  v->shape.pop_back();
  v->stride.pop_back();
  if (v->shape.empty()) {
    v->shape = {1U};
    v->stride = {0U};
  }
  return v;
}

// ------------------------------------------------------------------------------ broadcasting
bool Tensor::match_shape(const TensorPtr a) { // tensor.cpp:306-332: align trailing dims, mutate in place
  if (shape.size() > a->shape.size()) return false;
  const size_t mine = shape.size(), theirs = a->shape.size();
  for (size_t i = 0U; i < mine; ++i) {
    const size_t m = mine - 1U - i, t = theirs - 1U - i;
    if ((shape[m] != a->shape[t]) && stride[m]) return false;
  }
  std::vector<tcapint> st(theirs, 0U);
  for (size_t i = 0U; i < mine; ++i) st[theirs - 1U - i] = stride[mine - 1U - i];
  shape = a->shape;
  stride = st;
  return true;
}
void Tensor::materialize_broadcast() { // tensor.cpp:334-353
  bool needs = false;
  for (size_t i = 0; i < shape.size(); ++i)
    if (shape[i] > 1U && stride[i] == 0) needs = true;
  if (!needs) return;
  TensorPtr tmp = Tensor::allocate_like(shape, *this, storage->dtype, requires_grad, false);
  Weed::copy_broadcast(*tmp, *this);
  *this = *tmp;
}
void Tensor::reduce_grad_broadcast() { // tensor.cpp:355-369
  if (!requires_grad || !grad)
    throw std::domain_error("Called Tensor::reduce_grad_broadcast() on a node instance without a gradient Tensor! (This should be "
                            "called only during autograd.)");
  for (symint i = (symint)stride.size() - 1; i >= 0; --i) {
    if (stride[(size_t)i]) continue;
    grad = sum(grad, i);
  }
}

// ------------------------------------------------------------------------------ backward
void Tensor::backward(TensorPtr loss) { // tensor.cpp:371-401
  if (!loss || !loss->requires_grad) return;
  loss->grad->storage->FillOnes();
  std::vector<NodePtr> topo;
  std::unordered_set<Node *> seen;
  // iterative post-order DFS (the reference recurses; deep transformer graphs would not fit the stack)
  struct Frame {
    NodePtr n;
    size_t next;
  };
  std::vector<Frame> stack;
  if (loss->grad_node) {
    seen.insert(loss->grad_node.get());
    stack.push_back({loss->grad_node, 0U});
  }
  while (!stack.empty()) {
    Frame &f = stack.back();
    if (f.next < f.n->parents.size()) {
      const TensorPtr &p = f.n->parents[f.next++];
      if (p && p->grad_node && !seen.count(p->grad_node.get())) {
        seen.insert(p->grad_node.get());
        stack.push_back({p->grad_node, 0U});
      }
    } else {
      topo.push_back(f.n);
      stack.pop_back();
    }
  }
  const std::function<void(Tensor *)> &on_final = backend_config().on_leaf_grad_final;
  if (!on_final) {
    for (auto it = topo.rbegin(); it != topo.rend(); ++it) (*it)->backward();
    return;
  }
  // execution index of the last node that touches each leaf (closures replace p->grad, so the
  // tensor — not its gradient buffer — is the key; SURVEY §7 hard part 8)
  std::unordered_map<Tensor *, size_t> last_use;
  size_t idx = 0;
  for (auto it = topo.rbegin(); it != topo.rend(); ++it, ++idx)
    for (const TensorPtr &p : (*it)->parents)
      if (p && !p->grad_node && p->requires_grad) last_use[p.get()] = idx;
  idx = 0;
  for (auto it = topo.rbegin(); it != topo.rend(); ++it, ++idx) {
    (*it)->backward();
    for (const TensorPtr &p : (*it)->parents)
      if (p && !p->grad_node && p->requires_grad) {
        auto lu = last_use.find(p.get());
        if (lu != last_use.end() && lu->second == idx) {
          last_use.erase(lu); // a node may list the same leaf twice
          on_final(p.get());
        }
      }
  }
}

// ------------------------------------------------------------------------------ softmax family
TensorPtr Tensor::softmax(const TensorPtr x, symint axis) {
  axis = wrap_axis(axis, x->shape.size());
  const bool rg = x->requires_grad;
  TensorPtr out = allocate_like(*x, x->storage->dtype, rg, false);
  Weed::softmax((tcapint)axis, *x, *out);
  if (rg) make_softmax_node(x, out, axis);
  return out;
}
void Tensor::make_softmax_node(TensorPtr x, TensorPtr out, symint axis) {
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{x}, [x, wout = std::weak_ptr<Tensor>(out), axis]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr x_grad = full_grad(x);
    Weed::softmax_grad((tcapint)axis, *x_grad, *out, *(out->grad));
    settle_grad(x, x_grad);
  });
}
TensorPtr Tensor::logsoftmax(const TensorPtr x, symint axis) {
  axis = wrap_axis(axis, x->shape.size());
  const bool rg = x->requires_grad;
  TensorPtr out = allocate_like(*x, x->storage->dtype, rg, false);
  Weed::logsoftmax((tcapint)axis, *x, *out);
  if (rg) make_logsoftmax_node(x, out, axis);
  return out;
}
void Tensor::make_logsoftmax_node(TensorPtr x, TensorPtr out, symint axis) {
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{x}, [x, wout = std::weak_ptr<Tensor>(out), axis]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr x_grad = full_grad(x);
    Weed::logsoftmax_grad((tcapint)axis, *x_grad, *out, *(out->grad));
    settle_grad(x, x_grad);
  });
}

// ------------------------------------------------------------------------------ slices
TensorPtr Tensor::slice(TensorPtr a, const int64_t &row) { // tensor.cpp:463-479
  const bool rg = a->requires_grad;
  TensorPtr out = view_copy(a);
  out->offset += (tcapint)row * a->stride[0U];
  out->shape.erase(out->shape.begin());
  out->stride.erase(out->stride.begin());
  if (rg) make_row_slice_node(a, out, (tcapint)row);
  return out;
}
void Tensor::make_row_slice_node(TensorPtr a, TensorPtr out, const tcapint &row) {
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out), row]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr a_grad = view_copy(a->grad);
    TensorPtr keep = a_grad->grad; // slicing must not register new nodes
    const bool rg = a_grad->requires_grad;
    a_grad->requires_grad = false;
    TensorPtr row_view = Tensor::slice(a_grad, row);
    a_grad->requires_grad = rg;
    (void)keep;
    Weed::add_in_place(*row_view, *(out->grad));
    a->grad = a_grad;
  });
}
TensorPtr Tensor::slice(TensorPtr a, int64_t axis, const tcapint &start, const tcapint &length) { // tensor.cpp:500-526
  while (axis < 0) axis += (int64_t)a->shape.size();
  if (axis >= (int64_t)a->shape.size()) throw std::invalid_argument("Tensor::slice: axis out of range");
  if (length <= 0 || start + length > a->shape[(size_t)axis]) throw std::invalid_argument("Tensor::slice: invalid range");
  const bool rg = a->requires_grad;
  TensorPtr out = view_copy(a);
  out->offset += start * a->stride[(size_t)axis];
  out->shape[(size_t)axis] = length;
  if (rg) make_slice_node(a, out, axis, start);
  return out;
}
void Tensor::make_slice_node(TensorPtr a, TensorPtr out, const int64_t &axis, const tcapint &start) {
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out), axis, start]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    // reference: zero tmp of a's shape, add dout into the window, add tmp into a_grad
    // (tensor.cpp:528-553). Equivalent and one pass: add dout into the window of a_grad directly.
    TensorPtr a_grad = view_copy(a->grad);
    TensorPtr out_grad = view_copy(out->grad);
    a_grad->match_shape(a);
    a_grad->materialize_broadcast();
    TensorPtr window = view_copy(a_grad);
    window->requires_grad = false;
    window->offset += start * window->stride[(size_t)axis];
    window->shape[(size_t)axis] = out_grad->shape[(size_t)axis];
    Weed::add_in_place(*window, *out_grad);
    a->grad = a_grad;
    a->reduce_grad_broadcast();
  });
}

// ------------------------------------------------------------------------------ reductions
TensorPtr Tensor::sum(TensorPtr a) {
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_scalar_like(*a, rg);
  Weed::sum(*a, *out);
  if (rg) make_sum_node(a, out);
  return out;
}
void Tensor::make_sum_node(TensorPtr a, TensorPtr out) { // tensor.cpp:568-581: da += dout (broadcast)
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr a_grad = view_copy(a->grad);
    TensorPtr out_grad = view_copy(out->grad);
    out_grad->match_shape(a_grad);
    Weed::add_in_place(*a_grad, *out_grad);
    a->grad = a_grad;
  });
}
TensorPtr Tensor::mean(TensorPtr a) {
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_scalar_like(*a, rg);
  Weed::mean(*a, *out);
  if (rg) make_mean_node(a, out);
  return out;
}
void Tensor::make_mean_node(TensorPtr a, TensorPtr out) { // tensor.cpp:596-612: da += dout / N
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr a_grad = view_copy(a->grad);
    TensorPtr out_grad = view_copy(out->grad);
    out_grad->match_shape(a_grad);
    TensorPtr s = SCALAR((real1)(ONE_R1 / (real1)a->get_broadcast_size()), out_grad);
    TensorPtr tmp = s * out_grad;
    Weed::add_in_place(*a_grad, *tmp);
    a->grad = a_grad;
  });
}
TensorPtr Tensor::sum(TensorPtr a, symint axis) { // tensor.cpp:614-653
  axis = wrap_axis(axis, a->shape.size());
  const size_t p_stride = a->stride[(size_t)axis];
  if (!p_stride || (a->shape[(size_t)axis] == 1U)) {
    a->shape[(size_t)axis] = 1U;
    return a;
  }
  a = contiguous(a);
  const bool rg = a->requires_grad;
  std::vector<tcapint> shp = a->shape, str = a->stride;
  shp[(size_t)axis] = 1U;
  str[(size_t)axis] = 0U;
  size_t j = (size_t)axis + 1;
  while ((j < str.size()) && !str[j]) ++j;
  if (j < str.size()) {
    const size_t o_stride = str[j] / p_stride;
    for (; j < str.size(); ++j) str[j] /= (tcapint)o_stride;
  }
  TensorPtr out = allocate_like(shp, str, *a, a->storage->dtype, rg, false);
  Weed::reduce((tcapint)axis, *a, *out);
  if (rg) make_sum_node(a, out, (tcapint)axis);
  return out;
}
void Tensor::make_sum_node(TensorPtr a, TensorPtr out, const tcapint &axis) { // tensor.cpp:655-680
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out), axis]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr dx = view_copy(a->grad);
    TensorPtr dy = view_copy(out->grad);
    if (dy->shape.size() < a->shape.size()) dy->unsqueeze(axis); // re-insert the reduced axis
    dx->match_shape(a);
    dx->materialize_broadcast();
    dy->match_shape(dx);
    Weed::reduce_grad(axis, *dx, *a, *dy);
    a->grad = dx;
    a->reduce_grad_broadcast();
  });
}
// max / min along an axis (reference tensor.cpp:705-785): the output tensor is laid out exactly like Tensor::sum(a, axis)'s
namespace {
TensorPtr axis_extremum(TensorPtr a, symint axis, bool is_min) {
  axis = wrap_axis(axis, a->shape.size());
  const size_t p_stride = a->stride[(size_t)axis];
  if (!p_stride || (a->shape[(size_t)axis] == 1U)) {
    a->shape[(size_t)axis] = 1U;
    return a;
  }
  a = Tensor::contiguous(a);
  const bool rg = a->requires_grad;
  std::vector<tcapint> shp = a->shape, str = a->stride;
  shp[(size_t)axis] = 1U;
  str[(size_t)axis] = 0U;
  size_t j = (size_t)axis + 1;
  while ((j < str.size()) && !str[j]) ++j;
  if (j < str.size()) {
    const size_t o_stride = str[j] / p_stride;
    for (; j < str.size(); ++j) str[j] /= (tcapint)o_stride;
  }
  TensorPtr out = Tensor::allocate_like(shp, str, *a, a->storage->dtype, rg, false);
  if (is_min) Weed::min((tcapint)axis, *a, *out);
  else Weed::max((tcapint)axis, *a, *out);
  if (rg) Tensor::make_match_node(a, out, (tcapint)axis);
  return out;
}
} // namespace
TensorPtr Tensor::max(TensorPtr a, symint axis) { return axis_extremum(a, axis, false); }
TensorPtr Tensor::min(TensorPtr a, symint axis) { return axis_extremum(a, axis, true); }
void Tensor::make_match_node(TensorPtr a, TensorPtr out, const tcapint &axis) { // tensor.cpp:787-814
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out), axis]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr dx = view_copy(a->grad);
    TensorPtr dy = view_copy(out->grad);
    if (dy->shape.size() < a->shape.size()) dy->unsqueeze(axis); // re-insert the reduced axis
    dx->match_shape(a);
    dx->materialize_broadcast();
    dy->match_shape(dx);
    Weed::match_grad(axis, *dx, *a, *dy, *out);
    a->grad = dx;
    a->reduce_grad_broadcast();
  });
}
TensorPtr Tensor::max(TensorPtr a) { // tensor.cpp:993-1022
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_scalar_like(*a, rg);
  Weed::max(*a, *out);
  if (rg) make_max_node(a, out);
  return out;
}
TensorPtr Tensor::min(TensorPtr a) { // tensor.cpp:1024-1053
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_scalar_like(*a, rg);
  Weed::min(*a, *out);
  if (rg) make_min_node(a, out);
  return out;
}
namespace {
void make_full_extremum_node(TensorPtr a, TensorPtr out, bool is_min) {
  out->make_gradient();
  a->make_gradient(true);
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out), is_min]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr a_grad = full_grad(a);
    TensorPtr out_grad = view_copy(out->grad);
    out_grad->match_shape(a_grad);
    if (is_min) Weed::min_grad(*a_grad, *a, *out_grad, *out);
    else Weed::max_grad(*a_grad, *a, *out_grad, *out);
    settle_grad(a, a_grad);
  });
}
} // namespace
void Tensor::make_max_node(TensorPtr a, TensorPtr out) { make_full_extremum_node(a, out, false); }
void Tensor::make_min_node(TensorPtr a, TensorPtr out) { make_full_extremum_node(a, out, true); }
TensorPtr Tensor::clamp(TensorPtr a, real1 lo, real1 hi) { // tensor.cpp:1055-1082
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_like(*a, a->storage->dtype, rg, false);
  Weed::clamp(*a, lo, hi, *out);
  if (rg) make_clamp_node(a, lo, hi, out);
  return out;
}
void Tensor::make_clamp_node(TensorPtr a, real1 lo, real1 hi, TensorPtr out) {
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, lo, hi, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr a_grad = full_grad(a);
    Weed::clamp_grad(*a_grad, *a, *(out->grad), lo, hi);
    settle_grad(a, a_grad);
  });
}
TensorPtr Tensor::mean(TensorPtr a, symint axis) { // tensor.cpp:682-693
  axis = wrap_axis(axis, a->shape.size());
  TensorPtr tmp = sum(a, axis);
  tmp->squeeze(axis);
  tmp = tmp / (real1)(a->shape[(size_t)axis]);
  tmp->unsqueeze(axis);
  return tmp;
}
TensorPtr Tensor::variance(TensorPtr a) { return ((a - mean(a)) ^ real1(2)) / (real1)(a->get_broadcast_size()); }
TensorPtr Tensor::variance(TensorPtr a, const tcapint &axis) {
  TensorPtr tmp = a - mean(a, (symint)axis);
  return mean(tmp * tmp, (symint)axis);
}

// ------------------------------------------------------------------------------ unary
namespace {
typedef void (*UnaryFwd)(const Tensor &, Tensor &);
typedef void (*UnaryBwd)(Tensor &, const Tensor &, const Tensor &);
// `uses_output`: the gradient kernel reads the forward OUTPUT (sigmoid, tanh) rather than the input
TensorPtr unary_op(TensorPtr a, UnaryFwd fwd, UnaryBwd bwd, bool uses_output) {
  const bool rg = a->requires_grad;
  TensorPtr out = Tensor::allocate_like(*a, a->storage->dtype, rg, false);
  fwd(*a, *out);
  if (rg) {
    out->make_gradient();
    out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out), bwd, uses_output]() {
      TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
      if (!out) node_owner_lost();
      TensorPtr a_grad = full_grad(a);
      bwd(*a_grad, uses_output ? *out : *a, *(out->grad));
      settle_grad(a, a_grad);
    });
  }
  return out;
}
} // namespace
TensorPtr Tensor::abs(TensorPtr a) { return unary_op(a, Weed::abs, Weed::abs_grad, false); }
TensorPtr Tensor::relu(TensorPtr a) { return unary_op(a, Weed::relu, Weed::relu_grad, false); }
TensorPtr Tensor::sigmoid(TensorPtr a) { return unary_op(a, Weed::sigmoid, Weed::sigmoid_grad, true); }
TensorPtr Tensor::tanh(TensorPtr a) { return unary_op(a, Weed::tanh, Weed::tanh_grad, true); }
TensorPtr Tensor::sin(TensorPtr a) { return unary_op(a, Weed::sin, Weed::sin_grad, false); }
TensorPtr Tensor::cos(TensorPtr a) { return unary_op(a, Weed::cos, Weed::cos_grad, false); }
#define WEED_NODE_ONLY(fn, bwd, src)                                                               \
  void Tensor::fn(TensorPtr a, TensorPtr out) {                                                    \
    out->make_gradient();                                                                          \
    out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out)]() { \
      TensorPtr out = wout.lock(); /* the node is owned by this tensor: a strong capture would be a cycle */ \
      if (!out) node_owner_lost(); \
      TensorPtr a_grad = full_grad(a);                                                             \
      bwd(*a_grad, *src, *(out->grad));                                                            \
      settle_grad(a, a_grad);                                                                      \
    });                                                                                            \
  }
WEED_NODE_ONLY(make_abs_node, Weed::abs_grad, a)
WEED_NODE_ONLY(make_relu_node, Weed::relu_grad, a)
WEED_NODE_ONLY(make_sigmoid_node, Weed::sigmoid_grad, out)
WEED_NODE_ONLY(make_tanh_node, Weed::tanh_grad, out)
WEED_NODE_ONLY(make_sin_node, Weed::sin_grad, a)
WEED_NODE_ONLY(make_cos_node, Weed::cos_grad, a)
#undef WEED_NODE_ONLY

void Tensor::make_gelu_node(TensorPtr a, TensorPtr out) { // the node unary_op gives the fused Tensor::gelu
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{a}, [a, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr a_grad = full_grad(a);
    Weed::gelu_grad(*a_grad, *a, *(out->grad));
    settle_grad(a, a_grad);
  });
}
TensorPtr Tensor::gelu(const TensorPtr x) { // tensor.cpp:841-851
  if (backend_config().fused) return unary_op(x, Weed::gelu, Weed::gelu_grad, false);
  const real1 k0 = real1(0.5), k1 = real1(0.044715), k2 = real1(0.7978845608028654);
  TensorPtr x3 = x * x * x;
  TensorPtr inner = k2 * (x + k1 * x3);
  TensorPtr t = Tensor::tanh(inner);
  return k0 * x * (Tensor::ones_like(x->shape, false, false, DType::REAL, x->storage->device, x->storage->get_device_id()) + t);
}

// ------------------------------------------------------------------------------ binary
namespace {
// Accumulate `contribution` into parent's gradient with the reference's broadcast handling:
// match_shape -> materialize_broadcast -> (+|-)= -> reduce_grad_broadcast (tensor.cpp:1117-1134).
bool dense_like(const Tensor &t, const std::vector<tcapint> &shape) {
  if (t.shape != shape) return false;
  tcapint expect = 1U;
  for (size_t i = 0U; i < shape.size(); ++i) {
    if (shape[i] == 1U) continue;
    if (t.stride[i] != expect) return false;
    expect *= shape[i];
  }
  return true;
}
// Fused accumulation into a broadcast parent (see Tensor::make_gradient): sum the contribution over
// the parent's broadcast dims (one reduction pass) and add the result into the real-extent gradient.
bool accumulate_reduced(const TensorPtr &parent, const TensorPtr &like, const Tensor &contribution, bool subtract) {
  TensorPtr pv = view_copy(parent);
  // the parent's own (match_shape-mutated) view says which dims are broadcast; `like` may have lost
  // or gained unit dims since (squeeze / unsqueeze mutate tensors in place)
  if (pv->shape.size() < like->shape.size() && !pv->match_shape(like)) return false;
  const size_t rank = pv->shape.size();
  bool any = false;
  size_t prefix = 0U;
  while (prefix < rank && (pv->shape[prefix] == 1U || !pv->stride[prefix])) ++prefix;
  bool prefix_only = true;
  std::vector<tcapint> rshape = pv->shape;
  for (size_t d = 0U; d < rank; ++d) {
    if (pv->shape[d] > 1U && !pv->stride[d]) {
      any = true;
      rshape[d] = 1U;
      if (d >= prefix) prefix_only = false;
    }
  }
  if (!any || !dense_like(contribution, contribution.shape) || contribution.get_broadcast_size() != pv->get_broadcast_size()) return false;
  tcapint rcount = 1U;
  for (tcapint s : rshape) rcount *= s;
  if (!parent->grad || parent->grad->get_broadcast_size() != rcount) return false;
  struct QuirkOff { // internal reductions always use the intended index order
    bool prev;
    QuirkOff() : prev(backend_config().ref_index_quirks) { backend_config().ref_index_quirks = false; }
    ~QuirkOff() { backend_config().ref_index_quirks = prev; }
  } guard;
  TensorPtr c = std::make_shared<Tensor>(contribution);
  c->requires_grad = false;
  c->grad = nullptr;
  c->grad_node = nullptr;
  if (prefix_only) { // leading dims are the broadcast ones (bias / gamma): one [P, rest] column sum
    tcapint P = 1U, rest = 1U;
    for (size_t d = 0U; d < rank; ++d) (d < prefix ? P : rest) *= pv->shape[d];
    c->BaseTensor::reshape({(symint)P, (symint)rest});
    c = Tensor::sum(c, 0);
  } else {
    c->BaseTensor::reshape(std::vector<symint>(pv->shape.begin(), pv->shape.end()));
    for (symint d = (symint)rank - 1; d >= 0; --d)
      if (pv->shape[(size_t)d] > 1U && !pv->stride[(size_t)d]) c = Tensor::sum(c, d);
  }
  TensorPtr g = view_copy(parent->grad);
  if (subtract) Weed::sub_in_place(*g, *c);
  else Weed::add_in_place(*g, *c);
  parent->grad = g;
  return true;
}
void accumulate(const TensorPtr &parent, const TensorPtr &like, const Tensor &contribution, bool subtract) {
  if (backend_config().fused && accumulate_reduced(parent, like, contribution, subtract)) return;
  TensorPtr g = view_copy(parent->grad);
  g->match_shape(like);
  if (g->get_broadcast_size() != contribution.get_broadcast_size()) {
    // real-extent gradient of a broadcast parent whose rank no longer lines up with `like`
    TensorPtr pv = view_copy(parent);
    g->match_shape(pv);
  }
  g->materialize_broadcast();
  if (subtract) Weed::sub_in_place(*g, contribution);
  else Weed::add_in_place(*g, contribution);
  parent->grad = g;
  parent->reduce_grad_broadcast();
}
// match_shape mutates its operand in place (tensor.cpp:306-332): after `y + bias` a bias Parameter
// permanently reads [B, T, F] with strides [0, 0, 1]. In the reference a later call with another T
// (multi-token prefill followed by single-token decode) then broadcasts the ACTIVATION to the stale
// extent and MultiHeadAttention::forward throws "Tensor::reshape(): sizes do not match" [measured;
// DESIGN.md defect D10]. Here an expanded Parameter dim (stride 0, extent > 1) that disagrees with
// its partner's extent is collapsed back to 1 before matching — a no-op whenever the reference's
// own behaviour is well defined (same extents as before).
void rebase_expanded_parameter(TensorPtr &p, const TensorPtr &other) {
  if (!dynamic_cast<Parameter *>(p.get())) return;
  const size_t mine = p->shape.size(), theirs = other->shape.size();
  for (size_t i = 0U; i < mine; ++i) {
    const size_t m = mine - 1U - i;
    if (p->stride[m] || p->shape[m] <= 1U) continue;
    if (i >= theirs || other->shape[theirs - 1U - i] != p->shape[m]) p->shape[m] = 1U;
  }
}
void prepare_binary(TensorPtr &a, TensorPtr &b, const char *what) {
  rebase_expanded_parameter(a, b);
  rebase_expanded_parameter(b, a);
  if (!a->match_shape(b) && !b->match_shape(a)) throw std::invalid_argument(std::string("Tensor shape mismatch in ") + what + "!");
}
} // namespace

TensorPtr Tensor::add(TensorPtr a, TensorPtr b) { // tensor.cpp:1084-1103
  const bool rg = a->requires_grad || b->requires_grad;
  prepare_binary(a, b, "add");
  TensorPtr out = Tensor::allocate_like(a->shape, *a, DType::REAL, rg, false);
  Weed::add(*a, *b, *out);
  if (rg) make_add_node(a, b, out);
  return out;
}
// d(a + b)/da = 1: the parent's gradient contribution IS out's gradient. When the parent is a non-leaf
// that only this node feeds (consumers == 1: nothing else will ever accumulate into its gradient) and
// its own gradient buffer is still an untouched lazy zero fill of the same dense layout, the parent
// simply adopts out's gradient buffer, read-only, instead of a 8 B/elem copy — the residual adds of
// a transformer block hand [B,T,d] gradients to the W_o / ff2 outputs this way. out's gradient is
// complete when its node runs and nothing writes that buffer afterwards (a parent with a second
// consumer, which WILL be accumulated into, still gets its own copy), so intermediate gradients
// stay readable. Leaves never adopt: their gradients persist across steps (zero_grad, optimiser).
static bool adopt_incoming_gradient(const TensorPtr &parent, const TensorPtr &out_grad) {
  if (!backend_config().fused || !backend_config().lazy_zero) return false;
  if (!parent->grad_node || parent->consumers != 1U || !parent->grad || parent.get() == out_grad.get()) return false;
  if (parent->view_owner) return false; // a view shares its owner's gradient tensor: replacing the view's pointer alone would orphan it
  const TensorPtr &g = parent->grad;
  if (g->storage->device != DeviceTag::GPU || out_grad->storage->device != DeviceTag::GPU || g->storage->dtype != DType::REAL) return false;
  if (g->shape != out_grad->shape || g->stride != out_grad->stride || g->shape != parent->shape) return false;
  if (g->offset || out_grad->offset || g->storage->size != out_grad->storage->size) return false;
  if (!dense_like(*g, g->shape) || g->get_broadcast_size() != g->storage->size) return false;
  GpuRealStorage *gs = static_cast<GpuRealStorage *>(g->storage.get());
  if (!gs->zero_pending) return false; // something already accumulated here
  TensorPtr adopted = view_copy(out_grad);
  adopted->requires_grad = g->requires_grad;
  parent->grad = adopted;
  return true;
}
void Tensor::make_add_node(TensorPtr a, TensorPtr b, TensorPtr out) {
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(grad_parents({a, b}), [a, b, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr out_grad = view_copy(out->grad);
    if (a->requires_grad && !adopt_incoming_gradient(a, out_grad)) accumulate(a, out_grad, *out_grad, false);
    if (b->requires_grad && !adopt_incoming_gradient(b, out_grad)) accumulate(b, out_grad, *out_grad, false);
  });
}
TensorPtr Tensor::sub(TensorPtr a, TensorPtr b) { // tensor.cpp:1404-1423
  const bool rg = a->requires_grad || b->requires_grad;
  prepare_binary(a, b, "sub");
  TensorPtr out = Tensor::allocate_like(a->shape, *a, DType::REAL, rg, false);
  Weed::sub(*a, *b, *out);
  if (rg) make_sub_node(a, b, out);
  return out;
}
void Tensor::make_sub_node(TensorPtr a, TensorPtr b, TensorPtr out) {
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(grad_parents({a, b}), [a, b, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr out_grad = view_copy(out->grad);
    if (a->requires_grad) accumulate(a, out_grad, *out_grad, false);
    if (b->requires_grad) accumulate(b, out_grad, *out_grad, true);
  });
}
TensorPtr Tensor::mul(TensorPtr a, TensorPtr b) { // tensor.cpp:1138-1157
  const bool rg = a->requires_grad || b->requires_grad;
  prepare_binary(a, b, "mul");
  TensorPtr out = Tensor::allocate_like(a->shape, *a, DType::REAL, rg, false);
  Weed::mul(*a, *b, *out);
  if (rg) make_mul_node(a, b, out);
  return out;
}
void Tensor::make_mul_node(TensorPtr a, TensorPtr b, TensorPtr out) { // tensor.cpp:1159-1202
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(grad_parents({a, b}), [a, b, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr out_grad = view_copy(out->grad);
    auto side = [&](const TensorPtr &p, const TensorPtr &other) {
      TensorPtr tmp = Tensor::allocate_like(out_grad->shape, *out_grad, DType::REAL, false, false);
      Weed::mul(*out_grad, *other, *tmp); // d(p*o)/dp = o
      accumulate(p, out_grad, *tmp, false);
    };
    if (a->requires_grad) side(a, b);
    if (b->requires_grad) side(b, a);
  });
}
TensorPtr Tensor::div(TensorPtr a, TensorPtr b) { // tensor.cpp:1458-1477
  const bool rg = a->requires_grad || b->requires_grad;
  prepare_binary(a, b, "div");
  TensorPtr out = Tensor::allocate_like(a->shape, *a, DType::REAL, rg, false);
  Weed::div(*a, *b, *out);
  if (rg) make_div_node(a, b, out);
  return out;
}
void Tensor::make_div_node(TensorPtr a, TensorPtr b, TensorPtr out) { // tensor.cpp:1479-1524
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(grad_parents({a, b}), [a, b, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    TensorPtr out_grad = view_copy(out->grad);
    if (a->requires_grad) { // da += dout / b
      TensorPtr tmp = Tensor::allocate_like(out_grad->shape, *out_grad, DType::REAL, false, false);
      Weed::div(*out_grad, *b, *tmp);
      accumulate(a, out_grad, *tmp, false);
    }
    if (b->requires_grad) { // db -= a / b^2   (as written in the reference: dout is not applied here)
      TensorPtr b_sqr = Tensor::allocate_like(b->shape, *b, DType::REAL, false, false);
      Weed::mul(*b, *b, *b_sqr);
      TensorPtr tmp = Tensor::allocate_like(a->shape, *a, DType::REAL, false, false);
      Weed::div(*a, *b_sqr, *tmp);
      accumulate(b, a, *tmp, true);
    }
  });
}

TensorPtr Tensor::pow(TensorPtr a, real1 p) {
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_like(*a, a->storage->dtype, rg, false);
  Weed::pow(*a, p, *out);
  if (rg) make_pow_node(a, p, out);
  return out;
}
void Tensor::make_pow_node(TensorPtr x, real1 p, TensorPtr y) { // tensor.cpp:1540-1573: dx += p * dy * y / x
  y->make_gradient();
  y->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{x}, [x, p, wy = std::weak_ptr<Tensor>(y)]() {
    TensorPtr y = wy.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!y) node_owner_lost();
    TensorPtr dy = view_copy(y->grad);
    TensorPtr _x = view_copy(x), _y = view_copy(y);
    _y->match_shape(_x);
    _x->match_shape(_y);
    dy->match_shape(_y);
    TensorPtr dy_y = Tensor::allocate_like(dy->shape, *dy, DType::REAL, false, false);
    Weed::mul(*dy, *_y, *dy_y);
    TensorPtr dy_y_p = SCALAR(p, dy_y) * dy_y;
    TensorPtr r = Tensor::allocate_like(dy_y_p->shape, *dy_y_p, DType::REAL, false, false);
    Weed::div(*dy_y_p, *_x, *r);
    accumulate(x, _y, *r, false);
  });
}
TensorPtr Tensor::exp(TensorPtr a, real1 b) {
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_like(*a, a->storage->dtype, rg, false);
  Weed::exp(*a, b, *out);
  if (rg) make_exp_node(a, (real1)std::log((real1_s)b), out);
  return out;
}
void Tensor::make_exp_node(TensorPtr x, real1 log_b, TensorPtr y) { // tensor.cpp:1589-1616: dx += log_b * dy * y
  y->make_gradient();
  y->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{x}, [x, log_b, wy = std::weak_ptr<Tensor>(y)]() {
    TensorPtr y = wy.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!y) node_owner_lost();
    TensorPtr dy = view_copy(y->grad);
    dy->match_shape(y);
    TensorPtr dy_v = SCALAR(log_b, dy) * dy;
    TensorPtr r = Tensor::allocate_like(dy_v->shape, *dy_v, DType::REAL, false, false);
    Weed::mul(*dy_v, *y, *r);
    accumulate(x, y, *r, false);
  });
}
TensorPtr Tensor::log(TensorPtr a, real1 b) {
  const bool rg = a->requires_grad;
  TensorPtr out = allocate_like(*a, a->storage->dtype, rg, false);
  Weed::log(*a, b, *out);
  if (rg) make_log_node(a, (real1)(ONE_R1 / std::log((real1_s)b)), out);
  return out;
}
void Tensor::make_log_node(TensorPtr x, real1 inv_log_b, TensorPtr y) { // tensor.cpp:1632-1659: dx += inv_log_b * dy / x
  y->make_gradient();
  y->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{x}, [x, inv_log_b, wy = std::weak_ptr<Tensor>(y)]() {
    TensorPtr y = wy.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!y) node_owner_lost();
    TensorPtr dy = view_copy(y->grad);
    dy->match_shape(x);
    TensorPtr dy_v = SCALAR(inv_log_b, dy) * dy;
    TensorPtr r = Tensor::allocate_like(dy_v->shape, *dy_v, DType::REAL, false, false);
    Weed::div(*dy_v, *x, *r);
    accumulate(x, x, *r, false);
  });
}

// ------------------------------------------------------------------------------ matmul
TensorPtr Tensor::matmul(TensorPtr a, TensorPtr b) { // tensor.cpp:1204-1326
  if (a->shape.size() < 2U) throw std::invalid_argument("Tensor::matmul requires a to have rank >= 2");
  const bool rg = a->requires_grad || b->requires_grad;

  if (a->shape.size() > 2U && b->shape.size() > 2U) {
    // N-D x N-D: `batch` independent products. The reference copies both operands contiguous and
    // loops on the host (tensor.cpp:1242-1269); its returned tensor carries NO grad_node, so no
    // gradient flows through a batched product (SURVEY §7 hard part 5(i)) — reproduced here.
    if (a->shape.size() != b->shape.size()) throw std::invalid_argument("batched matmul rank mismatch");
    const size_t rank = a->shape.size();
    for (size_t i = 0; i < rank - 2; ++i)
      if (a->shape[i] != b->shape[i]) throw std::invalid_argument("batched matmul batch mismatch");
    const symint M = (symint)a->shape[rank - 2], K = (symint)a->shape[rank - 1], K2 = (symint)b->shape[rank - 2],
                 N = (symint)b->shape[rank - 1];
    if (K != K2) throw std::invalid_argument("batched matmul inner dim mismatch");
    symint batch = 1;
    for (size_t i = 0; i < rank - 2; ++i) batch *= (symint)a->shape[i];
    TensorPtr a3 = reshape(a, {batch, M, K});
    TensorPtr b3 = reshape(b, {batch, K, N});
    std::vector<tcapint> out_shape(a->shape.begin(), a->shape.end() - 2);
    out_shape.push_back((tcapint)M);
    out_shape.push_back((tcapint)N);
    TensorPtr out = allocate_like(out_shape, full_contiguous_stride(out_shape), *a3, DType::REAL, rg, false);
    TensorPtr out3 = view_copy(out);
    out3->BaseTensor::reshape({batch, M, N});
    Weed::matmul_batched(*a3, *b3, *out3);
    return out;
  }

  const bool needs_flatten = (a->shape.size() > 2U);
  const symint K = (symint)a->shape.back();
  const symint M = (symint)a->shape[a->shape.size() - 2];
  const symint N = (symint)b->shape[1U];
  if ((symint)(b->shape[0U]) != K) throw std::invalid_argument("matmul dimension mismatch");
  if (b->shape.size() < 2U) b->unsqueeze(1U);
  symint batch = 1;
  for (size_t i = 0; i < a->shape.size() - 2; ++i) batch *= (symint)a->shape[i];
  TensorPtr a2 = a;
  if (needs_flatten) a2 = reshape(a, {batch * M, K});
  const tcapint as0 = a2->shape[0U], bs1 = b->shape[1U];
  TensorPtr out = allocate_like(std::vector<tcapint>{as0, bs1}, std::vector<tcapint>{1U, as0}, *a2, DType::REAL, rg, false);
  Weed::matmul(*a2, *b, *out);
  if (needs_flatten) {
    std::vector<symint> final_shape;
    for (size_t i = 0; i < a->shape.size() - 2; ++i) final_shape.push_back((symint)a->shape[i]);
    final_shape.push_back(M);
    final_shape.push_back(N);
    out = reshape(out, final_shape);
  }
  if (rg) make_matmul_node(a, b, out);
  return out;
}

// Fused Linear::forward (src/modules/linear.cpp): x W + bias as ONE tensor-core GEMM with the bias
// added in the epilogue, one autograd node instead of matmul + add. Returns nullptr when the fused
// kernel does not apply (the caller then composes the two ops like the reference).
TensorPtr Tensor::linear(TensorPtr a, TensorPtr w, TensorPtr bias, TensorPtr residual) {
  const BackendConfig &cfg = backend_config();
  if (!cfg.fused) return nullptr;
  if (a->shape.size() < 2U || w->shape.size() != 2U || (symint)w->shape[0U] != (symint)a->shape.back()) return nullptr;
  // <= 16 rows (decode steps, tiny batches) take the skinny kernel in either precision mode; larger
  // products need the bf16 tensor-core path with cached operands for the bias epilogue
  const bool skinny = a->get_broadcast_size() / a->shape.back() <= 16U;
  if (!skinny && (cfg.matmul_precision != WEEDCU_GEMM_BF16 || !cfg.operand_cache)) return nullptr;
  if (a->storage->device != DeviceTag::GPU || w->storage->device != DeviceTag::GPU) return nullptr;
  if (bias->storage->size != w->shape[1U] || bias->get_size() != w->shape[1U] || bias->storage->device != DeviceTag::GPU) return nullptr;
  if (residual) {
    // the residual must be the dense column-major tensor the result will be: same leading extents as `a`, last extent N
    if (residual->storage->device != DeviceTag::GPU || residual->shape.size() != a->shape.size() || !is_contiguous(residual->shape, residual->stride))
      return nullptr;
    for (size_t i = 0; i + 1U < a->shape.size(); ++i)
      if (residual->shape[i] != a->shape[i]) return nullptr;
    if (residual->shape.back() != w->shape[1U]) return nullptr;
  }
  const bool rg = a->requires_grad || w->requires_grad || bias->requires_grad || (residual && residual->requires_grad);
  const bool needs_flatten = (a->shape.size() > 2U);
  const symint K = (symint)a->shape.back(), M = (symint)a->shape[a->shape.size() - 2], N = (symint)w->shape[1U];
  symint batch = 1;
  for (size_t i = 0; i < a->shape.size() - 2; ++i) batch *= (symint)a->shape[i];
  TensorPtr a2 = a;
  if (needs_flatten) a2 = reshape(a, {batch * M, K});
  const tcapint as0 = a2->shape[0U];
  TensorPtr out = allocate_like(std::vector<tcapint>{as0, (tcapint)N}, std::vector<tcapint>{1U, as0}, *a2, DType::REAL, rg, false);
  // a vocabulary-sized output that takes part in autograd is the LM head in front of cross_entropy_loss: its epilogue leaves
  // the bf16 copy the backward reads and the log-sum-exp partials the loss reads; the 4 B/elem fp32 logits are not written
  const bool lm_head = !residual && !skinny && rg && cfg.epilogue_stats && (tcapint)N >= cfg.lm_head_min_cols;
  if (!(lm_head && Weed::matmul_bias_lse(*a2, *w, *bias, *out)) && !Weed::matmul_bias(*a2, *w, *bias, *out, residual.get())) return nullptr;
  return finish_linear(a, w, bias, out, rg, residual);
}

// gelu(x W + bias) with the activation in the GEMM epilogue: the pre-activation h keeps the Linear node, y = gelu(h) the
// GELU node, exactly as Linear::forward followed by Tensor::gelu builds them; nullptr when the fused kernel does not apply
TensorPtr Tensor::linear_gelu(TensorPtr a, TensorPtr w, TensorPtr bias) {
  const BackendConfig &cfg = backend_config();
  if (!cfg.fused || !cfg.epilogue_stats || cfg.matmul_precision != WEEDCU_GEMM_BF16) return nullptr;
  if (a->shape.size() < 2U || w->shape.size() != 2U || (symint)w->shape[0U] != (symint)a->shape.back()) return nullptr;
  if (a->storage->device != DeviceTag::GPU || w->storage->device != DeviceTag::GPU || bias->storage->device != DeviceTag::GPU) return nullptr;
  if (bias->storage->size != w->shape[1U] || bias->get_size() != w->shape[1U]) return nullptr;
  const bool rg = a->requires_grad || w->requires_grad || bias->requires_grad;
  const symint K = (symint)a->shape.back(), M = (symint)a->shape[a->shape.size() - 2], N = (symint)w->shape[1U];
  symint batch = 1;
  for (size_t i = 0; i < a->shape.size() - 2; ++i) batch *= (symint)a->shape[i];
  TensorPtr a2 = a;
  if (a->shape.size() > 2U) a2 = reshape(a, {batch * M, K});
  const tcapint as0 = a2->shape[0U];
  TensorPtr h = allocate_like(std::vector<tcapint>{as0, (tcapint)N}, std::vector<tcapint>{1U, as0}, *a2, DType::REAL, rg, false);
  TensorPtr y = allocate_like(std::vector<tcapint>{as0, (tcapint)N}, std::vector<tcapint>{1U, as0}, *a2, DType::REAL, rg, false);
  if (!Weed::matmul_bias_gelu(*a2, *w, *bias, *h, *y)) return nullptr;
  h = finish_linear(a, w, bias, h, rg);
  if (h->shape.size() != 2U) y->BaseTensor::reshape(std::vector<symint>(h->shape.begin(), h->shape.end()));
  if (rg) {
    make_gelu_node(h, y);
    // y's gradient is read by gelu_grad alone, which takes a bf16 copy: the ff2 Linear's dA product may write just that
    if (backend_config().bf16_act_grad && y->grad && y->grad->storage->device == DeviceTag::GPU) static_cast<GpuRealStorage *>(y->grad->storage.get())->accept_bf16_values = true;
  }
  return y;
}

// everything Tensor::linear does after the product: final shape, the Parameter mutation of `y + bias`, the node
TensorPtr Tensor::finish_linear(TensorPtr a, TensorPtr w, TensorPtr bias, TensorPtr out, bool rg, TensorPtr residual) {
  const bool needs_flatten = (a->shape.size() > 2U);
  const symint M = (symint)a->shape[a->shape.size() - 2], N = (symint)w->shape[1U];
  if (needs_flatten) {
    std::vector<symint> final_shape;
    for (size_t i = 0; i < a->shape.size() - 2; ++i) final_shape.push_back((symint)a->shape[i]);
    final_shape.push_back(M);
    final_shape.push_back(N);
    out = reshape(out, final_shape);
  }
  prepare_binary(out, bias, "add"); // the same match_shape mutation of the Parameter that `y + bias` performs
  if (rg) {
    out->make_gradient();
    std::vector<TensorPtr> parents{a, w, bias};
    if (residual) parents.push_back(residual);
    out->grad_node = std::make_shared<Node>(grad_parents(parents), [a, w, bias, residual, wout = std::weak_ptr<Tensor>(out)]() {
      TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
      if (!out) node_owner_lost();
      if (residual && residual->requires_grad) { // d(residual + y)/d residual = 1: what the add node of `x + Linear(...)` does
        TensorPtr out_grad = view_copy(out->grad);
        if (!adopt_incoming_gradient(residual, out_grad)) accumulate(residual, out_grad, *out_grad, false);
      }
      if (bias->requires_grad) {
        // the bias gradient (column sums of dY) rides on the pass that packs dY for the two GEMMs below
        TensorPtr out_grad = view_copy(out->grad);
        bool done = false;
        if (bias->grad && is_contiguous(out_grad->shape, out_grad->stride)) {
          TensorPtr dy2 = view_copy(out_grad);
          dy2->requires_grad = false;
          const tcapint N = out_grad->shape.back();
          dy2->BaseTensor::reshape({(symint)(out_grad->get_broadcast_size() / N), (symint)N});
          TensorPtr bg = view_copy(bias->grad);
          done = Weed::pack_with_column_sums(*dy2, *bg);
          if (done) bias->grad = bg;
        }
        if (!done) accumulate(bias, out_grad, *out_grad, false);
      }
      matmul_backward(a, w, out);
    });
  }
  return out;
}

std::vector<TensorPtr> Tensor::linear_grouped(TensorPtr a, const std::vector<TensorPtr> &ws, const std::vector<TensorPtr> &biases, bool bf16_only) {
  const BackendConfig &cfg = backend_config();
  const size_t G = ws.size();
  if (!cfg.fused || G < 2U || G > 3U || biases.size() != G) return {};
  if (a->shape.size() < 2U || a->storage->device != DeviceTag::GPU) return {};
  // a handful of rows (a decode step): one grouped skinny launch at fp32 in either precision mode; otherwise the grouped
  // tensor-core launch, which needs the bf16 operand shadows
  const bool skinny = a->get_broadcast_size() / a->shape.back() <= 16U;
  if (!skinny && (cfg.matmul_precision != WEEDCU_GEMM_BF16 || !cfg.operand_cache)) return {};
  for (size_t g = 0U; g < G; ++g) {
    const TensorPtr &w = ws[g], &bias = biases[g];
    if (!w || !bias || w->shape.size() != 2U || (symint)w->shape[0U] != (symint)a->shape.back() || w->shape != ws[0]->shape) return {};
    if (w->storage->device != DeviceTag::GPU || bias->storage->size != w->shape[1U] || bias->get_size() != w->shape[1U] ||
        bias->storage->device != DeviceTag::GPU)
      return {};
  }
  const bool needs_flatten = (a->shape.size() > 2U);
  const symint K = (symint)a->shape.back(), M = (symint)a->shape[a->shape.size() - 2], N = (symint)ws[0]->shape[1U];
  symint batch = 1;
  for (size_t i = 0; i < a->shape.size() - 2; ++i) batch *= (symint)a->shape[i];
  TensorPtr a2 = a;
  if (needs_flatten) a2 = reshape(a, {batch * M, K});
  const tcapint as0 = a2->shape[0U];
  std::vector<TensorPtr> outs(G);
  std::vector<bool> rgs(G);
  std::vector<const Tensor *> wp(G), bp(G);
  std::vector<Tensor *> op(G);
  for (size_t g = 0U; g < G; ++g) {
    rgs[g] = a->requires_grad || ws[g]->requires_grad || biases[g]->requires_grad;
    outs[g] = allocate_like(std::vector<tcapint>{as0, (tcapint)N}, std::vector<tcapint>{1U, as0}, *a2, DType::REAL, rgs[g], false);
    wp[g] = ws[g].get();
    bp[g] = biases[g].get();
    op[g] = outs[g].get();
  }
  if (!(skinny ? Weed::matmul_skinny_grouped(*a2, wp, bp, op) : Weed::matmul_bias_grouped(*a2, wp, bp, op, bf16_only))) return {};
  for (size_t g = 0U; g < G; ++g) outs[g] = finish_linear(a, ws[g], biases[g], outs[g], rgs[g]);
  return outs;
}

void Tensor::make_matmul_node(TensorPtr a, TensorPtr b, TensorPtr out) { // tensor.cpp:1328-1402
  out->make_gradient();
  out->grad_node = std::make_shared<Node>(grad_parents({a, b}), [a, b, wout = std::weak_ptr<Tensor>(out)]() {
    TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
    if (!out) node_owner_lost();
    matmul_backward(a, b, out);
  });
}
void Tensor::matmul_backward(TensorPtr a, TensorPtr b, TensorPtr out) {
  {
    TensorPtr out_grad = view_copy(out->grad);
    const bool needs_flatten = (a->shape.size() > 2U);
    const symint K = (symint)a->shape.back();
    const symint M = (symint)a->shape[a->shape.size() - 2];
    const symint N = (symint)b->shape[1U];
    symint batch = 1;
    for (size_t i = 0; i < a->shape.size() - 2; ++i) batch *= (symint)a->shape[i];
    TensorPtr a2 = a, out_grad2 = out_grad;
    if (needs_flatten) {
      a2 = reshape(a, {batch * M, K});
      out_grad2 = reshape(out_grad, {batch * M, N});
    }
    a2 = view_copy(a2);
    a2->requires_grad = false;
    out_grad2 = view_copy(out_grad2);
    out_grad2->requires_grad = false;
    const bool fuse = backend_config().fused;

    if (a->requires_grad) { // dA += dC * B^T
      TensorPtr a_grad = view_copy(a->grad);
      TensorPtr bt = transpose(b);
      bool done = false;
      if (fuse && is_contiguous(a_grad->shape, a_grad->stride) && a_grad->shape == a->shape) {
        bool dense = true;
        for (size_t i = 0; i < a_grad->shape.size(); ++i)
          if (a_grad->shape[i] > 1U && !a_grad->stride[i]) dense = false;
        if (dense) {
          TensorPtr g2 = view_copy(a_grad);
          g2->requires_grad = false;
          g2->BaseTensor::reshape({batch * M, K});
          Weed::matmul_accumulate(*out_grad2, *bt, *g2);
          done = true;
        }
      }
      if (!done) {
        TensorPtr tmp = Tensor::allocate_like(std::vector<tcapint>{(tcapint)(batch * M), (tcapint)K},
                                              std::vector<tcapint>{1U, (tcapint)(batch * M)}, *a2, DType::REAL, false, false);
        Weed::matmul(*out_grad2, *bt, *tmp);
        if (needs_flatten) {
          std::vector<symint> a_shape(a->shape.begin(), a->shape.end());
          tmp = reshape(tmp, a_shape);
        }
        Weed::add_in_place(*a_grad, *tmp);
      }
      a->grad = a_grad;
    }
    if (b->requires_grad) { // dB += A^T * dC
      TensorPtr b_grad = view_copy(b->grad);
      TensorPtr at = transpose(a2);
      if (fuse && b_grad->shape.size() == 2U && b_grad->stride[0U] && b_grad->stride[1U]) {
        b_grad->requires_grad = false;
        Weed::matmul_accumulate(*at, *out_grad2, *b_grad);
        b_grad->requires_grad = b->grad->requires_grad;
      } else {
        TensorPtr tmp = Tensor::allocate_like(b_grad->shape, *b_grad, DType::REAL, false, false);
        Weed::matmul(*at, *out_grad2, *tmp);
        Weed::add_in_place(*b_grad, *tmp);
      }
      b->grad = b_grad;
    }
  }
}
} // namespace Weed
