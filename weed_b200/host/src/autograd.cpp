// autograd.cpp — optimisers and losses on the CUDA device.
// Reference: include/autograd/adam.hpp:23-106, sgd.hpp:23-37, zero_grad.hpp:21-25, mse_loss.hpp,
// bci_with_logits_loss.hpp:20-23, cross_entropy_loss.hpp:21-34.
#include "weed_b200/autograd.hpp"

#include <cmath>

namespace Weed {
namespace {
constexpr tcapint kZeroInAdamMax = 4U << 20; // elements; see adam_collect
// A tensor whose view is exactly its whole storage in storage order (so a flat kernel may walk it)
bool covers_storage(const Tensor &t) {
  if (t.offset) return false;
  tcapint expect = 1U;
  for (size_t i = 0U; i < t.shape.size(); ++i) {
    if (t.shape[i] == 1U) continue;
    if (t.stride[i] != expect) return false;
    expect *= t.shape[i];
  }
  return expect == t.storage->size;
}
// Parameters are mutated by match_shape (a bias becomes [B,T,F] with strides [0,0,1], SURVEY §7
// hard part 5(ii)), so flat kernels go by storage: n = storage->size, gradient reduced to the same
// element count in the same order.
bool flat_pair(const Tensor &p, const Tensor &g) {
  return g.storage->size == p.storage->size && covers_storage(g) && p.storage->device == DeviceTag::GPU &&
         g.storage->device == DeviceTag::GPU;
}
} // namespace

void Adam::register_parameter(ParameterPtr p) {
  AdamState s;
  s.m = Tensor::zeros(p->shape, false, false, DType::REAL, p->storage->device, p->storage->get_device_id());
  s.v = Tensor::zeros(p->shape, false, false, DType::REAL, p->storage->device, p->storage->get_device_id());
  state[p] = s;
}

void adam_begin_step(Adam &opt, real1 &bias_correction1, real1 &bias_correction2) {
  opt.t += 1;
  bias_correction1 = (real1)(ONE_R1 - std::pow((real1_s)opt.beta1, (real1_s)opt.t));
  bias_correction2 = (real1)(ONE_R1 - std::pow((real1_s)opt.beta2, (real1_s)opt.t));
}

// fused path: every eligible parameter joins ONE multi-tensor launch (28 B/param; adam.hpp:84-104 issues ~15 ops and
// ~12 temporaries per parameter). Collecting takes the device pointers — which may issue pending lazy zero fills on
// the compute stream — so it is separate from the launch: the data-parallel path collects before it hands a bucket
// to the communication stream and launches there after the all-reduce.
void adam_collect(Adam &opt, const std::vector<ParameterPtr> &params, AdamBatch &b, std::vector<ParameterPtr> &slow) {
  const BackendConfig &cfg = backend_config();
  for (auto &p : params) {
    const auto it = opt.state.find(p);
    if (it == opt.state.end()) throw std::invalid_argument("Parameter passed to adam_step that was not registered with optimizer!");
    AdamState &s = it->second;
    TensorPtr g = p->grad;
    if (!g) throw std::invalid_argument("adam_step: parameter has no gradient");
    if (!(cfg.fused && flat_pair(*p, *g) && s.m->storage->size == p->storage->size && s.v->storage->size == p->storage->size)) {
      slow.push_back(p);
      continue;
    }
    {
      // never touched by any backward so far: the gradient and both moments are still pending lazy zero fills, and with
      // g = m = v = 0 the update (adam.hpp:84-104) leaves m, v and the parameter exactly as they are — nothing to launch
      // (the Q/K/V projections and the LayerNorm in front of them, which the reference's attention sends no gradient to)
      const GpuRealStorage *g0 = static_cast<const GpuRealStorage *>(g->storage.get());
      const GpuRealStorage *m0 = static_cast<const GpuRealStorage *>(s.m->storage.get()), *v0 = static_cast<const GpuRealStorage *>(s.v->storage.get());
      if (cfg.lazy_zero && g0->zero_pending && m0->zero_pending && v0->zero_pending) continue;
    }
    if (b.stream && b.stream != p->stream()) throw std::domain_error("adam_step: parameters live on different devices");
    b.stream = p->stream();
    b.p.push_back(p->device_ptr());
    // bf16 GEMM operand shadow of a weight that is a plain linear copy of the parameter (dense, leading
    // dimension a multiple of 8): the update kernel refreshes it in the same pass
    uint16_t *sh = nullptr;
    if (cfg.operand_cache && cfg.matmul_precision == WEEDCU_GEMM_BF16) {
      GpuRealStorage *ps = static_cast<GpuRealStorage *>(p->storage.get());
      for (size_t k = 0U; k < ps->shadows.size() && !sh; ++k) {
        const GpuRealStorage::Bf16Shadow &c = ps->shadows[k];
        if (c.offset == 0U && c.s_fast == 1U && c.s_slow == c.n_fast && (c.n_fast % 8U) == 0U && (uint64_t)c.n_fast * c.n_slow == ps->size) {
          sh = (uint16_t *)c.buf->ptr;
          b.refreshed.push_back({p->storage, c.buf});
        }
      }
    }
    b.shadow.push_back(sh);
    // a gradient still waiting for its lazy zero-fill was never touched by backward: pass "zeros"
    GpuRealStorage *gs = static_cast<GpuRealStorage *>(g->storage.get());
    b.g.push_back(gs->zero_pending ? nullptr : g->device_ptr_ro());
    // small gradients (everything but the embedding / LM-head matrices: the ones the next backward accumulates into with
    // split-K reduce-adds, column sums or LayerNorm partials) are zeroed by this kernel right after it has read them;
    // the big ones stay lazily zeroed, their first product of the next step overwrites them
    const bool zero_here = cfg.lazy_zero && !gs->zero_pending && !gs->buffer_shared() && p->storage->size <= kZeroInAdamMax;
    b.zero_grad.push_back(zero_here ? 1U : 0U);
    if (zero_here) b.zeroed.push_back(g->storage);
    b.m.push_back(s.m->device_ptr());
    b.v.push_back(s.v->device_ptr());
    b.n.push_back(p->storage->size);
  }
}
void adam_launch(Adam &opt, AdamBatch &b, real1 bias_correction1, real1 bias_correction2, void *stream) {
  if (b.p.empty()) return;
  throw_on_error(weedcu_adam_step_multi_zero((uint32_t)b.p.size(), b.p.data(), b.g.data(), b.m.data(), b.v.data(), b.n.data(), b.shadow.data(),
                                             b.zero_grad.data(), opt.lr, opt.beta1, opt.beta2, opt.eps, bias_correction1, bias_correction2,
                                             backend_config().grad_scale, stream ? stream : b.stream),
                 "adam_step");
  for (const StoragePtr &z : b.zeroed) { // contents changed (bf16 copies / column sums of the old values are stale) and are zero
    GpuRealStorage *gs = static_cast<GpuRealStorage *>(z.get());
    ++gs->version;
    gs->zero_version = gs->version;
  }
  // the shadows written by the kernel describe the parameter as it is now (device_ptr() in adam_collect moved the version)
  for (const auto &r : b.refreshed) {
    GpuRealStorage *ps = static_cast<GpuRealStorage *>(r.first.get());
    for (GpuRealStorage::Bf16Shadow &c : ps->shadows)
      if (c.buf == r.second) c.version = ps->version;
  }
}
// the reference's composition (adam.hpp:84-104), one parameter
void adam_slow(Adam &opt, const ParameterPtr &p, real1 bias_correction1, real1 bias_correction2) {
  const BackendConfig &cfg = backend_config();
  AdamState &s = opt.state.at(p);
  TensorPtr g = p->grad;
  if (cfg.grad_scale != ONE_R1) g = cfg.grad_scale * g;
  s.m = opt.beta1 * s.m + (ONE_R1 - opt.beta1) * g;
  s.v = opt.beta2 * s.v + (ONE_R1 - opt.beta2) * g * g;
  TensorPtr tmp = opt.lr * s.m / (bias_correction1 * (((s.v / bias_correction2) ^ ((real1)0.5)) + opt.eps));
  p->match_shape(tmp);
  tmp->match_shape(p);
  Weed::sub_in_place(*p, *tmp);
}

void adam_step(Adam &opt, const std::vector<ParameterPtr> &params) {
  real1 bc1, bc2;
  adam_begin_step(opt, bc1, bc2);
  AdamBatch batch;
  std::vector<ParameterPtr> slow;
  adam_collect(opt, params, batch, slow);
  for (const ParameterPtr &p : slow) adam_slow(opt, p, bc1, bc2);
  adam_launch(opt, batch, bc1, bc2, nullptr);
}

void sgd_step(const std::vector<ParameterPtr> &params, real1 lr) {
  const BackendConfig &cfg = backend_config();
  for (auto &p : params) {
    TensorPtr pg = p->grad;
    if (!pg) throw std::invalid_argument("sgd_step: parameter has no gradient");
    if (cfg.fused && flat_pair(*p, *pg)) {
      throw_on_error(weedcu_sgd_step(p->device_ptr(), pg->device_ptr_ro(), p->storage->size, lr, cfg.grad_scale, p->stream()), "sgd_step");
      continue;
    }
    TensorPtr tmp = (lr * cfg.grad_scale) * pg;
    tmp->match_shape(p);
    Weed::sub_in_place(*p, *tmp);
  }
}

void zero_grad(const std::vector<ParameterPtr> &params) {
  for (auto p : params)
    if (p->grad) p->grad->storage->FillZeros();
}

TensorPtr mse_loss(TensorPtr y_pred, TensorPtr y_true) { return Tensor::mean((y_pred - y_true) * (y_pred - y_true)); }

TensorPtr bci_with_logits_loss(TensorPtr logits, TensorPtr y_true) {
  return Tensor::relu(logits) - logits * y_true + Tensor::log(ONE_R1 + Tensor::exp(-ONE_R1 * Tensor::abs(logits)));
}

TensorPtr cross_entropy_loss(TensorPtr logits, SymbolTensorPtr targets) {
  const size_t rank = logits->shape.size();
  const tcapint V = logits->shape[rank - 1U];
  const tcapint rows = logits->get_broadcast_size() / V;
  if (backend_config().fused && logits->storage->device == DeviceTag::GPU && Tensor::is_contiguous(logits->shape, logits->stride) &&
      targets->get_broadcast_size() == rows && targets->stride[0U] == 1U) {
    // fused: one read of the logits forward (online log-sum-exp + exact int gather of the target
    // column), one elementwise kernel backward; never materialises log-softmax or a one-hot
    const bool rg = logits->requires_grad;
    TensorPtr loss = Tensor::allocate_scalar_like(*logits, rg);
    TensorPtr lse = Tensor::allocate_like(std::vector<tcapint>{rows}, *logits, DType::REAL, false, false);
    SymbolTensorPtr tg = targets->storage->device == DeviceTag::GPU ? targets : targets->cast(DeviceTag::GPU);
    const tcapint vs = logits->stride[rank - 1U];
    // logits straight from an LM head whose epilogue left the log-sum-exp partials: no pass over the logits at all
    if (!(vs == rows && Weed::cross_entropy_fwd_from_stats(*logits, *tg, *lse, *loss, rows, V)))
      throw_on_error(weedcu_cross_entropy_fwd(logits->device_ptr_ro(), logits->offset, rows, V, 1U, vs, tg->device_ptr() + tg->offset,
                                              lse->device_ptr(), loss->device_ptr(), logits->stream()),
                     "cross_entropy_loss");
    if (rg) {
      loss->make_gradient();
      loss->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{logits}, [logits, tg, lse, wloss = std::weak_ptr<Tensor>(loss), rows, V, vs]() {
        TensorPtr loss = wloss.lock(); // the node is owned by this tensor: a strong capture would be a cycle
        if (!loss) node_owner_lost();
        TensorPtr dl = std::make_shared<Tensor>(*(logits->grad));
        if (vs == rows && Tensor::is_contiguous(dl->shape, dl->stride) && dl->storage->device == DeviceTag::GPU &&
            Weed::cross_entropy_bwd_pack(*logits, *tg, *lse, *(loss->grad), *dl, rows, V)) {
          logits->grad = dl;
          return;
        }
        int accumulate = 1;
        real1 *dl_ptr = dl->device_ptr_accumulate(accumulate);
        throw_on_error(weedcu_cross_entropy_bwd(logits->device_ptr_ro(), logits->offset, rows, V, 1U, vs, tg->device_ptr() + tg->offset,
                                                lse->device_ptr_ro(), loss->grad->device_ptr_ro() + loss->grad->offset, dl_ptr, dl->offset,
                                                accumulate, logits->stream()),
                       "cross_entropy_loss backward");
        logits->grad = dl;
      });
    }
    return loss;
  }
  // reference composition (cross_entropy_loss.hpp:21-34): logits [1, T, V]
  const symint T = (symint)rows, Vs = (symint)V;
  TensorPtr lsm = Tensor::logsoftmax(logits, -1);
  lsm = Tensor::reshape(lsm, {T, Vs});
  TensorPtr oh = Tensor::one_hot(targets, V);
  TensorPtr selected = lsm * oh;
  TensorPtr gathered = Tensor::sum(selected, 1);
  return Tensor::mean(gathered) * real1(-1.0f);
}

void allreduce_gradients(const std::vector<ParameterPtr> &params, void *comm) {
  bool grouped = false; // one NCCL group: the per-parameter reductions are fused into few kernels
  for (const ParameterPtr &p : params) {
    if (!p->grad) continue;
    Tensor &g = *(p->grad);
    // untouched on this rank means untouched on every rank (same graph): zero everywhere, skip
    if (g.storage->device == DeviceTag::GPU && static_cast<GpuRealStorage *>(g.storage.get())->zero_pending) continue;
    if (!grouped) {
      throw_on_error(weedcu_nccl_group_start(), "allreduce_gradients");
      grouped = true;
    }
    throw_on_error(weedcu_nccl_allreduce_sum(comm, g.device_ptr(), g.storage->size, g.stream()), "allreduce_gradients");
  }
  if (grouped) throw_on_error(weedcu_nccl_group_end(), "allreduce_gradients");
}
GradientBuckets::GradientBuckets(void *c, size_t bytes) : comm(c), bucket_bytes(bytes) {
  // highest priority: the all-reduce kernels' blocks are scheduled ahead of the backward kernels' as SMs free up, so the
  // collective of a bucket runs WHILE the rest of backward does instead of queueing behind it
  throw_on_error(weedcu_stream_create_priority(&comm_stream, 1), "GradientBuckets");
  throw_on_error(weedcu_event_create(&ev_ready), "GradientBuckets");
  throw_on_error(weedcu_event_create(&ev_done), "GradientBuckets");
  throw_on_error(weedcu_event_create(&ev_head), "GradientBuckets");
}
GradientBuckets::~GradientBuckets() {
  backend_config().on_leaf_grad_final = nullptr;
  if (ev_ready) weedcu_event_destroy(ev_ready);
  if (ev_done) weedcu_event_destroy(ev_done);
  if (ev_head) weedcu_event_destroy(ev_head);
  if (comm_stream) weedcu_stream_destroy(comm_stream);
}
void GradientBuckets::begin() {
  pending.clear();
  pending_bytes = 0U;
  reduced.clear();
  tail.clear();
  backend_config().on_leaf_grad_final = [this](Tensor *leaf) { add(leaf); };
}
void GradientBuckets::begin(Adam &opt, const std::vector<ParameterPtr> &params) {
  begin();
  chained = &opt;
  owners.clear();
  slow.clear();
  for (const ParameterPtr &p : params) owners[p.get()] = p;
  adam_begin_step(opt, bc1, bc2);
}
void GradientBuckets::add(Tensor *leaf) {
  if (!leaf->grad || reduced.count(leaf)) return;
  if (chained && !owners.count(leaf)) return; // not one of the optimiser's parameters: nothing to exchange for it
  reduced.insert(leaf);
  pending.push_back(leaf);
  pending_bytes += (size_t)leaf->grad->storage->size * sizeof(real1);
  if (pending_bytes >= bucket_bytes) flush();
}
void GradientBuckets::flush() {
  if (pending.empty()) return;
  // device_ptr() materialises a pending lazy zero fill on the compute stream, so take the pointers
  // (the gradients' and, when the optimiser is chained, everything its update touches) before the
  // event that hands the bucket to the communication stream
  std::vector<std::pair<real1 *, size_t>> bufs;
  void *compute = nullptr;
  for (Tensor *leaf : pending) {
    Tensor &g = *(leaf->grad);
    bufs.push_back({g.device_ptr(), (size_t)g.storage->size});
    compute = g.stream();
  }
  AdamBatch batch;
  if (chained) {
    std::vector<ParameterPtr> ps;
    for (Tensor *leaf : pending) ps.push_back(owners.at(leaf));
    adam_collect(*chained, ps, batch, slow);
  }
  // (everything enqueued on the communication stream so far is complete at ev_head; this bucket becomes the tail of a split
  // update unless another one follows)
  tail.assign(pending.begin(), pending.end());
  throw_on_error(weedcu_event_record(ev_head, comm_stream), "GradientBuckets::flush");
  throw_on_error(weedcu_event_record(ev_ready, compute), "GradientBuckets::flush");
  throw_on_error(weedcu_stream_wait_event(comm_stream, ev_ready), "GradientBuckets::flush");
  // small gradients ride in one message
  std::vector<const real1 *> s_src;
  std::vector<real1 *> s_dst;
  std::vector<uint64_t> s_n;
  size_t s_total = 0U;
  if (small_elems)
    for (const auto &b : bufs)
      if (b.second < small_elems) s_total += (b.second + 3U) & ~(size_t)3U;
  if (s_total && s_total > staging_elems) {
    staging = static_cast<GpuRealStorage *>(pending.front()->grad->storage.get())->dev->MakeBuffer(sizeof(real1) * s_total);
    staging_elems = s_total;
  }
  const bool coalesce = s_total != 0U;
  if (coalesce) {
    size_t at = 0U;
    for (const auto &b : bufs)
      if (b.second < small_elems) {
        s_src.push_back(b.first);
        s_dst.push_back((real1 *)staging->ptr + at);
        s_n.push_back(b.second);
        at += (b.second + 3U) & ~(size_t)3U;
      }
    throw_on_error(weedcu_multi_copy((uint32_t)s_n.size(), s_src.data(), s_dst.data(), s_n.data(), comm_stream), "GradientBuckets::flush (gather)");
  }
  throw_on_error(weedcu_nccl_group_start(), "GradientBuckets::flush");
  for (const auto &b : bufs)
    if (!coalesce || b.second >= small_elems) throw_on_error(weedcu_nccl_allreduce_sum(comm, b.first, b.second, comm_stream), "GradientBuckets::flush");
  if (coalesce) throw_on_error(weedcu_nccl_allreduce_sum(comm, (real1 *)staging->ptr, s_total, comm_stream), "GradientBuckets::flush");
  throw_on_error(weedcu_nccl_group_end(), "GradientBuckets::flush");
  if (coalesce) // back to where the optimiser reads them (the padding between runs is summed too and ignored)
    throw_on_error(weedcu_multi_copy((uint32_t)s_n.size(), (const real1 *const *)s_dst.data(), const_cast<real1 *const *>((real1 **)s_src.data()), s_n.data(), comm_stream),
                   "GradientBuckets::flush (scatter)");
  // bucket reduced -> its parameters are updated right behind it on the communication stream, while the compute
  // stream carries on with the rest of backward (SURVEY 8e "fusion opportunity"): a parameter is only read by the
  // nodes that list it as a parent, and all of those were issued before ev_ready
  if (chained) adam_launch(*chained, batch, bc1, bc2, comm_stream);
  ++buckets_launched;
  pending.clear();
  pending_bytes = 0U;
}
void GradientBuckets::finish(const std::vector<ParameterPtr> &params) {
  finish_async(params);
  wait_all();
}
void GradientBuckets::wait_head() {
  if (tail_compute) throw_on_error(weedcu_stream_wait_event(tail_compute, ev_head), "GradientBuckets::wait_head");
}
void GradientBuckets::wait_all() {
  if (tail_compute) throw_on_error(weedcu_stream_wait_event(tail_compute, ev_done), "GradientBuckets::wait_all");
  if (chained_done.size()) {
    for (const ParameterPtr &p : chained_done) adam_slow(*chained_opt, p, bc1, bc2); // parameters the fused kernel cannot take: reference composition
    chained_done.clear();
  }
}
void GradientBuckets::finish_async(const std::vector<ParameterPtr> &params) {
  backend_config().on_leaf_grad_final = nullptr;
  void *compute = nullptr;
  std::vector<ParameterPtr> untouched;
  for (const ParameterPtr &p : params) {
    if (!p->grad) continue;
    compute = p->grad->stream();
    if (reduced.count(p.get())) continue;
    // not reached by this rank's backward walk: still reduce it unless it is untouched everywhere
    // (same graph on every rank: a pending zero fill here means a pending zero fill there)
    Tensor &g = *(p->grad);
    if (g.storage->device == DeviceTag::GPU && static_cast<GpuRealStorage *>(g.storage.get())->zero_pending) {
      untouched.push_back(p);
      continue;
    }
    reduced.insert(p.get());
    pending.push_back(p.get());
    pending_bytes += (size_t)g.storage->size * sizeof(real1);
  }
  flush(); // (the last bucket flushed — here or, when nothing is pending, during backward — is the tail: see flush())
  if (chained && !untouched.empty()) {
    // zero gradients still decay the moments (adam.hpp:84-104): one more launch behind the last bucket
    AdamBatch batch;
    adam_collect(*chained, untouched, batch, slow);
    if (compute) {
      throw_on_error(weedcu_event_record(ev_ready, compute), "GradientBuckets::finish");
      throw_on_error(weedcu_stream_wait_event(comm_stream, ev_ready), "GradientBuckets::finish");
    }
    adam_launch(*chained, batch, bc1, bc2, comm_stream);
  }
  tail_compute = compute;
  if (compute) throw_on_error(weedcu_event_record(ev_done, comm_stream), "GradientBuckets::finish");
  if (chained) {
    chained_opt = chained;
    chained_done = slow; // updated by wait_all(), once the compute stream has been told to wait for the exchange
    chained = nullptr;
    owners.clear();
    slow.clear();
  }
}
void broadcast_parameters(const std::vector<ParameterPtr> &params, void *comm, int root) {
  for (const ParameterPtr &p : params)
    throw_on_error(weedcu_nccl_broadcast(comm, p->device_ptr(), p->storage->size, root, p->stream()), "broadcast_parameters");
}
} // namespace Weed
