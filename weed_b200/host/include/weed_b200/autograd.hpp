// autograd.hpp — optimisers and losses (reference include/autograd/*.hpp), same signatures.
// With backend_config().fused (default) adam_step / sgd_step / cross_entropy_loss issue one fused
// kernel per parameter / per loss; with it off they compose the same tensor ops the reference does.
#pragma once
#include <unordered_set>
#include "weed_b200/ops.hpp"

namespace Weed {
struct AdamState {
  TensorPtr m; // first moment
  TensorPtr v; // second moment
};
struct Adam {
  real1 lr, beta1, beta2, eps;
  uint64_t t;
  std::unordered_map<ParameterPtr, AdamState> state;
  Adam(real1 l, real1 b1 = ADAM_BETA1_DEFAULT, real1 b2 = ADAM_BETA2_DEFAULT, real1 e = ADAM_EPSILON_DEFAULT)
      : lr(l), beta1(b1), beta2(b2), eps(e), t(0U) {}
  void register_parameter(ParameterPtr p);
  void register_parameters(const std::vector<ParameterPtr> &pv) {
    for (const ParameterPtr &p : pv) register_parameter(p);
  }
};
void adam_step(Adam &opt, const std::vector<ParameterPtr> &params);
// adam_step in pieces, for callers that update parameters group by group within one step (GradientBuckets):
//   adam_begin_step (t += 1, this step's bias corrections) ; per group: adam_collect -> adam_launch ; adam_slow for the
//   parameters adam_collect hands back because the fused multi-tensor kernel cannot take them
struct AdamBatch {
  std::vector<real1 *> p, m, v;
  std::vector<const real1 *> g;
  std::vector<uint64_t> n;
  std::vector<uint16_t *> shadow;
  std::vector<uint8_t> zero_grad;    // 1: the kernel zeroes this gradient after reading it (zero_grad fused into the update)
  std::vector<StoragePtr> zeroed;    // the gradient storages that will hold zeros once the launch has run
  std::vector<std::pair<StoragePtr, BufferPtr>> refreshed; // (parameter storage, shadow buffer) rewritten by the launch
  void *stream = nullptr;
};
void adam_begin_step(Adam &opt, real1 &bias_correction1, real1 &bias_correction2);
void adam_collect(Adam &opt, const std::vector<ParameterPtr> &params, AdamBatch &batch, std::vector<ParameterPtr> &slow);
void adam_launch(Adam &opt, AdamBatch &batch, real1 bias_correction1, real1 bias_correction2, void *stream);
void adam_slow(Adam &opt, const ParameterPtr &p, real1 bias_correction1, real1 bias_correction2);
void sgd_step(const std::vector<ParameterPtr> &params, real1 lr);
void zero_grad(const std::vector<ParameterPtr> &params);

TensorPtr mse_loss(TensorPtr y_pred, TensorPtr y_true);
TensorPtr bci_with_logits_loss(TensorPtr logits, TensorPtr y_true);
// logits [B, T, V] (the reference assumes B == 1, cross_entropy_loss.hpp:22-27; any B works here:
// rows = B*T), targets B*T token ids
TensorPtr cross_entropy_loss(TensorPtr logits, SymbolTensorPtr targets);

// Data-parallel training (no reference counterpart: Weed has no gradient exchange).
// One process per GPU; `comm` is an NCCL communicator from weedcu_nccl_init. Gradients are
// sum-all-reduced in place on the compute stream; the 1/world average is folded into the fused
// optimiser kernels through backend_config().grad_scale.
void allreduce_gradients(const std::vector<ParameterPtr> &params, void *comm);
// The same reduction overlapped with the backward walk: gradients are collected into buckets of
// ~bucket_bytes in the order they become final (Tensor::backward's on_leaf_grad_final hook: LM head
// first, embeddings last) and every full bucket is all-reduced as one NCCL group on a dedicated
// communication stream while the compute stream carries on with the remaining backward kernels.
//   GradientBuckets gb(comm);  gb.begin();  Tensor::backward(loss);  gb.finish(params);  adam_step(opt, params);
// or, with the optimiser chained onto the buckets:
//   gb.begin(opt, params);  Tensor::backward(loss);  gb.finish(params);
// finish() reduces what is left (including gradients no node touched on this rank but may have on
// another), and makes the compute stream wait for the communication stream.
struct GradientBuckets {
  void *comm;
  size_t bucket_bytes;
  void *comm_stream = nullptr, *ev_ready = nullptr, *ev_done = nullptr;
  std::vector<Tensor *> pending;
  size_t pending_bytes = 0U;
  std::unordered_set<Tensor *> reduced;
  uint64_t buckets_launched = 0U;
  // chained optimiser (begin(opt, params)): each bucket's parameters are updated by the fused Adam kernel on the
  // communication stream right behind the bucket's all-reduce, so only the last bucket's update is left on the
  // critical path; finish() then replaces adam_step for this step
  Adam *chained = nullptr;
  real1 bc1 = ONE_R1, bc2 = ONE_R1;
  std::unordered_map<Tensor *, ParameterPtr> owners;
  std::vector<ParameterPtr> slow;
  // gradients below small_elems travel together: gathered into `staging`, all-reduced as ONE message, scattered back (a
  // GPT-2-small step has ~60 bias / LayerNorm gradients of 768 .. 3072 floats: as separate collectives their per-message
  // latency, not their bytes, was most of the exchange at 8 GPUs)
  size_t small_elems = 65536U;
  BufferPtr staging;
  size_t staging_elems = 0U;
  // split update (finish_async): `tail` = the parameters of the last bucket that was flushed; everything before it is complete at ev_head
  void *ev_head = nullptr;
  std::vector<Tensor *> tail;
  void *tail_compute = nullptr;
  Adam *chained_opt = nullptr;
  std::vector<ParameterPtr> chained_done;
  explicit GradientBuckets(void *c, size_t bytes = 32U << 20);
  ~GradientBuckets();
  void begin();
  void begin(Adam &opt, const std::vector<ParameterPtr> &params);
  void add(Tensor *leaf);
  void flush();
  void finish(const std::vector<ParameterPtr> &params);
  // finish() without the final wait: the last bucket's all-reduce is in flight when this returns. wait_head() makes the
  // compute stream wait for every bucket before the last one, wait_all() for the last one too — a caller updates the
  // parameters outside `tail` between the two, behind which the last exchange hides.
  void finish_async(const std::vector<ParameterPtr> &params);
  void wait_head();
  void wait_all();
};
void broadcast_parameters(const std::vector<ParameterPtr> &params, void *comm, int root);
} // namespace Weed
