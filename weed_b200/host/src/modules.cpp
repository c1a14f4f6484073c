// modules.cpp — Linear / LayerNorm / Embedding / LearnedPositionalEncoding / MultiHeadAttention /
// TransformerEncoderLayer / Sequential for the CUDA device.
// Reference: src/modules/linear.cpp:21-100, layernorm.cpp:29-42, embedding.cpp:28-65,
// learned_positional_encoding.cpp:21-61, multihead_attention.cpp:145-356,
// transformer_encoder_layer.cpp:22-125, include/modules/sequential.hpp:23-84.
// forward() keeps the reference's composition when backend_config().fused is off; with it on,
// LayerNorm and the attention core are single fused device paths with the same results.
#include "weed_b200/modules.hpp"

#include <cmath>
#include <random>

namespace Weed {
namespace {
inline TensorPtr view_copy(const TensorPtr &t) { return std::make_shared<Tensor>(*t); }
bool dense_contiguous(const Tensor &t) {
  tcapint expect = 1U;
  for (size_t i = 0U; i < t.shape.size(); ++i) {
    if (t.shape[i] == 1U) continue;
    if (t.stride[i] != expect) return false;
    expect *= t.shape[i];
  }
  return true;
}
TensorPtr strided_view(const TensorPtr &base, const std::vector<tcapint> &shape, const std::vector<tcapint> &stride, tcapint offset) {
  TensorPtr v = view_copy(base);
  v->requires_grad = false;
  v->grad = nullptr;
  v->grad_node = nullptr;
  v->shape = shape;
  v->stride = stride;
  v->offset = offset;
  return v;
}
} // namespace

ParameterPtr MigrateGpu::pforward(const ParameterPtr p) {
  if (p->storage->is_gpu()) return p; // already resident: nothing to move on this backend
  ParameterPtr out = std::make_shared<Parameter>(*p);
  out->storage = out->storage->gpu();
  return out;
}
ParameterPtr MigrateCpu::pforward(const ParameterPtr p) {
  if (!p->storage->is_gpu()) return p;
  ParameterPtr out = std::make_shared<Parameter>(*p);
  out->storage = out->storage->cpu();
  return out;
}

// ------------------------------------------------------------------------------------- Linear
Linear::Linear(tcapint in_f, tcapint out_f, bool use_bias, bool init_rand, DType dtype, DeviceTag device, int64_t device_id)
    : Module(LINEAR_T), in_features(in_f), out_features(out_f) {
  if (dtype != DType::REAL) throw std::invalid_argument("Linear: only DType::REAL is supported on the CUDA device");
  const std::vector<tcapint> shape{in_f, out_f};
  if (init_rand) { // Xavier-uniform, linear.cpp:29-45
    std::random_device rd;
    std::mt19937 gen(rd());
    const real1_s lim = (real1_s)(std::sqrt(6.0 / (in_f + out_f)));
    std::uniform_real_distribution<real1_s> dis(-lim, lim);
    std::vector<real1> init((size_t)in_f * out_f);
    for (auto &v : init) v = (real1)dis(gen);
    weight = std::make_shared<Parameter>(init, shape, device, device_id);
  } else {
    weight = std::make_shared<Parameter>(shape, std::vector<tcapint>{1U, in_f}, true, dtype, device, device_id);
    weight->storage->FillZeros();
  }
  if (use_bias) {
    bias = std::make_shared<Parameter>(std::vector<tcapint>{out_f}, std::vector<tcapint>{1U}, true, dtype, device, device_id);
    bias->storage->FillZeros();
  }
}
void Linear::migrate_cpu() {
  MigrateCpu mc;
  weight = mc.pforward(weight);
  if (bias) bias = mc.pforward(bias);
}
void Linear::migrate_gpu() {
  MigrateGpu mg;
  weight = mg.pforward(weight);
  if (bias) bias = mg.pforward(bias);
}
TensorPtr Linear::forward(const TensorPtr x) {
  if (bias) {
    TensorPtr fused = Tensor::linear(x, weight, bias);
    if (fused) return fused;
  }
  TensorPtr y = x >> weight;
  if (bias) y = y + bias;
  return y;
}
TensorPtr Linear::forward_add(const TensorPtr x, const TensorPtr residual) {
  if (bias && backend_config().fused) {
    TensorPtr fused = Tensor::linear(x, weight, bias, residual);
    if (fused) return fused;
  }
  return residual + forward(x);
}
std::vector<ParameterPtr> Linear::parameters() {
  if (bias) return {weight, bias};
  return {weight};
}

// ------------------------------------------------------------------------------------- LayerNorm
LayerNorm::LayerNorm(const tcapint &f, const DeviceTag &dtag, const real1 &e, const int64_t &did) : Module(LAYERNORM_T), features(f), eps(e) {
  gamma = std::make_shared<Parameter>(std::vector<real1>(f, ONE_R1), std::vector<tcapint>{1U, 1U, f}, dtag, did);
  beta = std::make_shared<Parameter>(std::vector<real1>(f, ZERO_R1), std::vector<tcapint>{1U, 1U, f}, dtag, did);
}
void LayerNorm::migrate_cpu() {
  MigrateCpu mc;
  gamma = mc.pforward(gamma);
  beta = mc.pforward(beta);
}
void LayerNorm::migrate_gpu() {
  MigrateGpu mg;
  gamma = mg.pforward(gamma);
  beta = mg.pforward(beta);
}
TensorPtr LayerNorm::forward(const TensorPtr x) {
  const BackendConfig &cfg = backend_config();
  const size_t rank = x->shape.size();
  // the reference's axis reduce permutes its outputs when two non-axis extents exceed 1 (defect D1); with at most one
  // (B == 1, or rank 2) its indexing is the intended one and the fused kernels reproduce it in the faithful mode too
  size_t wide = 0U;
  for (size_t i = 0U; i + 1U < rank; ++i)
    if (x->shape[i] > 1U) ++wide;
  const bool fusable = cfg.fused && (!cfg.ref_index_quirks || wide <= 1U) && x->storage->device == DeviceTag::GPU && rank >= 2U &&
                       x->shape[rank - 1U] == features && dense_contiguous(*x) && gamma->storage->size == features &&
                       beta->storage->size == features;
  if (!fusable) { // the reference's composition, layernorm.cpp:29-42
    TensorPtr xc = x - Tensor::mean(x, -1);
    TensorPtr y = xc / ((Tensor::mean(xc * xc, -1) + eps) ^ real1(0.5f));
    return y * gamma + beta;
  }
  const tcapint rows = x->get_broadcast_size() / features;
  const bool rg = x->requires_grad || gamma->requires_grad || beta->requires_grad;
  TensorPtr y = Tensor::allocate_like(x->shape, *x, DType::REAL, rg, false);
  TensorPtr mean = Tensor::allocate_like(std::vector<tcapint>{rows}, *x, DType::REAL, false, false);
  TensorPtr rstd = Tensor::allocate_like(std::vector<tcapint>{rows}, *x, DType::REAL, false, false);
  Weed::layernorm_forward(*x, rows, features, *gamma, *beta, eps, *y, *mean, *rstd);
  if (rg) {
    ParameterPtr g = gamma, b = beta;
    std::vector<TensorPtr> parents;
    for (const TensorPtr &p : std::vector<TensorPtr>{x, g, b})
      if (p->requires_grad) parents.push_back(p);
    const tcapint F = features;
    y->make_gradient();
    y->grad_node = std::make_shared<Node>(parents, [x, g, b, wy = std::weak_ptr<Tensor>(y), mean, rstd, rows, F]() {
      TensorPtr y = wy.lock(); // the node is owned by this tensor: a strong capture would be a cycle
      if (!y) node_owner_lost();
      // one kernel: dx += ..., dgamma += sum_rows dy*xhat, dbeta += sum_rows dy (16 B/elem)
      TensorPtr dx, dg, db;
      if (x->requires_grad) {
        dx = view_copy(x->grad);
        dx->match_shape(x);
        dx->materialize_broadcast();
      } else {
        dx = Tensor::zeros(x->shape, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
      }
      auto param_grad = [&](const ParameterPtr &p) -> TensorPtr {
        if (!p->requires_grad) return nullptr;
        TensorPtr pg = view_copy(p->grad);
        if (pg->storage->size != F) { // gradient still carries broadcast dims: reduce it first
          p->grad = pg;
          p->reduce_grad_broadcast();
          pg = view_copy(p->grad);
        }
        return pg;
      };
      dg = param_grad(g);
      db = param_grad(b);
      // dx = (what dx already holds) + this contribution; a dx that shares the residual gradient's buffer copy-on-write
      // (the add node in front of this LayerNorm) is read from there and written to a private buffer: no copy
      const real1 *dx_src = nullptr;
      BufferPtr keep;
      real1 *dx_ptr = dx->device_ptr_accumulate_from(dx_src, keep);
      throw_on_error(weedcu_layernorm_bwd_from(x->device_ptr_ro() + x->offset, y->grad->device_ptr_ro() + y->grad->offset, rows, F,
                                               g->device_ptr_ro() + g->offset, mean->device_ptr_ro(), rstd->device_ptr_ro(),
                                               dx_src ? dx_src + dx->offset : nullptr, dx_ptr + dx->offset,
                                               dg ? dg->device_ptr() + dg->offset : nullptr, db ? db->device_ptr() + db->offset : nullptr,
                                               backend_config().layernorm_exact_grad ? 1 : 0 /* reference chain */, x->stream()),
                     "LayerNorm backward");
      if (x->requires_grad) x->grad = dx;
      if (dg) g->grad = dg;
      if (db) b->grad = db;
    });
  }
  return y;
}

// ------------------------------------------------------------------------------------- Embedding
Embedding::Embedding(const tcapint &vocab, const tcapint &dim, const DType &dtype, const DeviceTag &dtag, int64_t did)
    : Module(EMBEDDING_T), num_embeddings(vocab), embedding_dim(dim),
      weight(std::make_shared<Parameter>(std::vector<tcapint>{vocab, dim}, std::vector<tcapint>{1, vocab}, true, dtype, dtag, did)) {}
void Embedding::migrate_cpu() {
  MigrateCpu mc;
  weight = mc.pforward(weight);
}
void Embedding::migrate_gpu() {
  MigrateGpu mg;
  weight = mg.pforward(weight);
}
TensorPtr Embedding::forward(const SymbolTensorPtr indices_) { // embedding.cpp:28-65
  SymbolTensorPtr indices = indices_->storage->device == weight->storage->device ? indices_ : indices_->cast(weight->storage->device);
  std::vector<tcapint> out_shape = indices->shape;
  out_shape.push_back(embedding_dim);
  TensorPtr out = Tensor::allocate_like(out_shape, Tensor::full_contiguous_stride(out_shape), *weight, DType::REAL, weight->requires_grad, false);
  // faithful mode: with [1, T] indices the reference writes slot 0 only (D6) and the rest keeps the zero fill of its allocation
  if (backend_config().ref_index_quirks) out->storage->FillZeros();
  Weed::embedding_gather(*indices, *weight, *out);
  if (weight->requires_grad) {
    ParameterPtr w = weight;
    out->make_gradient();
    out->grad_node = std::make_shared<Node>(std::vector<TensorPtr>{w}, [indices, w, wout = std::weak_ptr<Tensor>(out)]() {
      TensorPtr out = wout.lock(); // the node is owned by this tensor: a strong capture would be a cycle
      if (!out) node_owner_lost();
      TensorPtr dW = view_copy(w->grad);
      TensorPtr dout = view_copy(out->grad);
      dW->match_shape(w);
      dW->materialize_broadcast();
      Weed::embedding_scatter_add(*dW, *indices, *dout);
      w->grad = dW;
      w->reduce_grad_broadcast();
    });
  }
  return out;
}

// ------------------------------------------------------------------------------------- positional
LearnedPositionalEncoding::LearnedPositionalEncoding(const tcapint &max_len_, const tcapint &d_model_, const DeviceTag &dtag)
    : Module(LEARNED_POSITIONAL_ENCODING_T), max_len(max_len_), d_model(d_model_) {
  std::random_device rd;
  std::mt19937 gen(rd());
  std::uniform_real_distribution<real1_s> dis(real1_s(0), real1_s(0.01));
  std::vector<real1> init((size_t)max_len * d_model);
  for (auto &v : init) v = (real1)dis(gen);
  pos_encoding = std::make_shared<Parameter>(init, std::vector<tcapint>{1U, max_len, d_model}, dtag);
}
void LearnedPositionalEncoding::migrate_cpu() {
  MigrateCpu mc;
  pos_encoding = mc.pforward(pos_encoding);
}
void LearnedPositionalEncoding::migrate_gpu() {
  MigrateGpu mg;
  pos_encoding = mg.pforward(pos_encoding);
}
TensorPtr LearnedPositionalEncoding::forward(const TensorPtr x) {
  const tcapint T = x->shape[1];
  if (T > max_len) throw std::invalid_argument("Input sequence length exceeds maximum positional encoding length!");
  return x + Tensor::slice(pos_encoding, 1, 0, T);
}

// ------------------------------------------------------------------------------------- attention
MultiHeadAttention::MultiHeadAttention(tcapint d_model_, tcapint num_heads_, tcapint num_kv_heads_, tcapint head_dim_, DeviceTag dtag,
                                       RoPEPtr r, real1_f mask_val_, const int64_t did, const bool _use_kv_cache, int kv_quant_bits_)
    : Module(MULTIHEAD_ATTENTION_T), d_model((symint)d_model_), num_heads((symint)num_heads_),
      num_kv_heads((symint)(num_kv_heads_ ? num_kv_heads_ : num_heads_)), head_dim((symint)(!head_dim_ ? d_model_ / num_heads_ : head_dim_)),
      mask_val(mask_val_), W_q(std::make_shared<Linear>(d_model_, d_model_, true, true, DType::REAL, dtag, did)),
      W_k(std::make_shared<Linear>(d_model_, d_model_, true, true, DType::REAL, dtag, did)),
      W_v(std::make_shared<Linear>(d_model_, d_model_, true, true, DType::REAL, dtag, did)),
      W_o(std::make_shared<Linear>(d_model_, d_model_, true, true, DType::REAL, dtag, did)), rope(r), use_kv_cache(_use_kv_cache),
      kv_quant_bits(kv_quant_bits_) {
  if (d_model % num_heads) throw std::invalid_argument("d_model must be divisible by num_heads");
  _register_params();
  if (mask_val == ZERO_R1) mask_val = -1.701411835e38f; // -2^127, multihead_attention.hpp:106-114
}
void MultiHeadAttention::_register_params() {
  param_vector = W_q->parameters();
  auto add = [&](const std::vector<ParameterPtr> &q) { param_vector.insert(param_vector.end(), q.begin(), q.end()); };
  add(W_k->parameters());
  add(W_v->parameters());
  add(W_o->parameters());
}
void MultiHeadAttention::train() {
  for (auto &l : {W_q, W_k, W_v, W_o}) l->train();
}
void MultiHeadAttention::eval() {
  for (auto &l : {W_q, W_k, W_v, W_o}) l->eval();
}
void MultiHeadAttention::reset_cache() {
  k_cache = nullptr;
  v_cache = nullptr;
  cache_len = 0U;
  max_seq_len = 0U;
}
void MultiHeadAttention::migrate_cpu() {
  for (auto &l : {W_q, W_k, W_v, W_o}) l->migrate_cpu();
}
void MultiHeadAttention::migrate_gpu() {
  for (auto &l : {W_q, W_k, W_v, W_o}) l->migrate_gpu();
}

TensorPtr MultiHeadAttention::forward(const TensorPtr x) { // multihead_attention.cpp:145-356
  const symint B = (symint)x->shape[0], T = (symint)x->shape[1];
  // the three projections read the same x: one grouped tensor-core launch when the bf16 path applies
  // (each output keeps the node Linear::forward would give it), otherwise three Linear::forward calls
  TensorPtr Q, K, V;
  if (W_q->bias && W_k->bias && W_v->bias) {
    // when the fused tensor-core attention core follows, it is the projections' only reader and takes their bf16 copies
    const BackendConfig &c0 = backend_config();
    const bool core_reads_bf16 = c0.fused && c0.matmul_precision == WEEDCU_GEMM_BF16 && !use_kv_cache && !rope && num_kv_heads == num_heads && head_dim == 64 &&
                                 (x->shape[0] % 8U) == 0U && x->shape.size() == 3U && (x->shape[1] % 8U) == 0U && x->shape[1] >= 64U;
    const std::vector<TensorPtr> qkv =
        Tensor::linear_grouped(x, {W_q->weight, W_k->weight, W_v->weight}, {W_q->bias, W_k->bias, W_v->bias}, core_reads_bf16);
    if (qkv.size() == 3U) {
      Q = qkv[0];
      K = qkv[1];
      V = qkv[2];
    }
  }
  if (!Q) {
    Q = W_q->forward(x);
    K = W_k->forward(x);
    V = W_v->forward(x);
  }
  if (use_kv_cache && kv_quant_bits > 0)
    throw std::domain_error("4-bit TurboQuant KV cache (multihead_attention.cpp:205-277) is host-loop code outside this backend's "
                            "scope; construct with kv_quant_bits = 0 or set use_kv_cache = false");
  const BackendConfig &cfg = backend_config();
  // (RoPE rotates Q and K between the projections and the scores: those layers take the reference's composition below)
  const bool fuse = cfg.fused && !use_kv_cache && !rope && (num_kv_heads == num_heads) && dense_contiguous(*Q) && dense_contiguous(*K) &&
                    dense_contiguous(*V);
  TensorPtr out;
  if (fuse) {
    // Attention core with no autograd edges, exactly like the reference (its batched matmul returns
    // a tensor without grad_node, tensor.cpp:1253-1271) but laid out for the GPU: each (b,h) pair
    // becomes a contiguous [T, hd] matrix, scores/probabilities are [T, T] per pair, the scale +
    // causal mask + softmax chain (multihead_attention.cpp:319-334) is one fused kernel.
    const tcapint Bu = (tcapint)B, Tu = (tcapint)T, H = (tcapint)num_heads, hd = (tcapint)head_dim, BH = Bu * H;
    if (cfg.matmul_precision == WEEDCU_GEMM_BF16) {
      // one entry point: head relayout fused with the bf16 operand conversion, probabilities
      // emitted as bf16 straight into the P V product (include/weedcu.h: weedcu_attention_fwd)
      out = Tensor::allocate_like(std::vector<tcapint>{Bu, Tu, H * hd}, *x, DType::REAL, false, false);
      // the relayout back to [B, T, C] also leaves the bf16 operand of the W_o product (no pack pass)
      const OutputShadow os = (Bu % 4U) == 0U ? begin_output_shadow(*out, H * hd) : OutputShadow();
      int rc = Weed::attention_forward_bf16(*Q, *K, *V, *out, os.ptr, Bu, Tu, H, hd, std::sqrt((real1)head_dim), mask_val, (T > 1) ? 1 : 0);
      if (rc == 0) {
        end_output_shadow(os);
        return fuse_residual ? W_o->forward_add(out, fuse_residual) : W_o->forward(out);
      }
      if (rc != WEEDCU_ENOSUP) throw_on_error(rc, "attention");
    }
    auto to_heads = [&](const TensorPtr &lin) { // [B,T,(h,j)] -> [T, hd, B, H] contiguous
      TensorPtr dst = Tensor::allocate_like(std::vector<tcapint>{Tu, hd, Bu, H}, *lin, DType::REAL, false, false);
      TensorPtr src = strided_view(lin, {Tu, hd, Bu, H}, {Bu, Bu * Tu * H, 1U, Bu * Tu}, lin->offset);
      Weed::copy_broadcast(*dst, *src);
      return dst;
    };
    TensorPtr Qc = to_heads(Q), Kc = to_heads(K), Vc = to_heads(V);
    Q = K = V = nullptr;
    TensorPtr q3 = strided_view(Qc, {BH, Tu, hd}, {Tu * hd, 1U, Tu}, 0U);
    TensorPtr kt3 = strided_view(Kc, {BH, hd, Tu}, {Tu * hd, Tu, 1U}, 0U); // K^T view
    TensorPtr v3 = strided_view(Vc, {BH, Tu, hd}, {Tu * hd, 1U, Tu}, 0U);
    TensorPtr scores = Tensor::allocate_like(std::vector<tcapint>{Tu, Tu, BH}, *Qc, DType::REAL, false, false);
    TensorPtr s3 = strided_view(scores, {BH, Tu, Tu}, {Tu * Tu, 1U, Tu}, 0U);
    Weed::matmul_batched(*q3, *kt3, *s3);
    throw_on_error(weedcu_attn_softmax_real(scores->device_ptr(), scores->device_ptr(), BH, Tu, Tu, std::sqrt((real1)head_dim), mask_val,
                                            (T > 1) ? 1 : 0, /*batch_fastest=*/0, scores->stream()),
                   "attention softmax");
    TensorPtr oc = Tensor::allocate_like(std::vector<tcapint>{Tu, hd, Bu, H}, *Qc, DType::REAL, false, false);
    TensorPtr o3 = strided_view(oc, {BH, Tu, hd}, {Tu * hd, 1U, Tu}, 0U);
    Weed::matmul_batched(*s3, *v3, *o3);
    // back to (B, T, num_heads*head_dim): out[b,t,h,j] = oc[t,j,b,h]
    out = Tensor::allocate_like(std::vector<tcapint>{Bu, Tu, H * hd}, *x, DType::REAL, false, false);
    TensorPtr dst = strided_view(out, {Bu, Tu, H, hd}, {1U, Bu, Bu * Tu, Bu * Tu * H}, 0U);
    TensorPtr src = strided_view(oc, {Bu, Tu, H, hd}, {Tu * hd, 1U, Tu * hd * Bu, Tu}, 0U);
    Weed::copy_broadcast(*dst, *src);
    return fuse_residual ? W_o->forward_add(out, fuse_residual) : W_o->forward(out);
  }

  if (cfg.fused && use_kv_cache && !rope && kv_quant_bits == 0 && num_kv_heads == num_heads && head_dim <= 64 && dense_contiguous(*Q) &&
      dense_contiguous(*K) && dense_contiguous(*V) && x->storage->device == DeviceTag::GPU) {
    // Float KV cache (multihead_attention.cpp:169-199, 278-287) with the attention over it as ONE
    // entry: append K, V to their slots, scores against the cache in place, softmax, P V, output
    // already in [B, T, H*hd] (include/weedcu.h: weedcu_attention_decode). The reference's ~15 ops
    // copy the whole cache contiguous twice per call.
    const tcapint Bu = (tcapint)B, Tu = (tcapint)T, H = (tcapint)num_heads, hd = (tcapint)head_dim;
    if (!k_cache) {
      if (!max_seq_len) max_seq_len = 2048U;
      cache_len = 0U;
      const std::vector<tcapint> cs{Bu, H, max_seq_len, hd};
      k_cache = Tensor::zeros(cs, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
      v_cache = Tensor::zeros(cs, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
    }
    if (k_cache->shape[0U] != Bu) throw std::invalid_argument("KV cache was allocated for another batch size; call reset_cache()");
    if (cache_len + Tu > max_seq_len) throw std::invalid_argument("KV cache is full (slice out of range)");
    out = Tensor::allocate_like(std::vector<tcapint>{Bu, Tu, H * hd}, *x, DType::REAL, false, false);
    const int rc = weedcu_attention_decode(Q->device_ptr_ro() + Q->offset, K->device_ptr_ro() + K->offset, V->device_ptr_ro() + V->offset,
                                           k_cache->device_ptr(), v_cache->device_ptr(), out->device_ptr(), Bu, Tu, H, hd, max_seq_len, cache_len,
                                           std::sqrt((real1)head_dim), mask_val, (T > 1) ? 1 : 0, x->stream());
    if (rc == 0) {
      cache_len += Tu;
      return fuse_residual ? W_o->forward_add(out, fuse_residual) : W_o->forward(out);
    }
    if (rc != WEEDCU_ENOSUP) throw_on_error(rc, "attention decode");
  }

  Q = Tensor::reshape(Q, std::vector<symint>{B, T, num_heads, head_dim});
  K = Tensor::reshape(K, std::vector<symint>{B, T, num_kv_heads, head_dim});
  V = Tensor::reshape(V, std::vector<symint>{B, T, num_kv_heads, head_dim});
  Q = Tensor::transpose(Q, 1, 2);
  K = Tensor::transpose(K, 1, 2);
  V = Tensor::transpose(V, 1, 2);
  if (rope) { // optional rotary embedding (Qwen), multihead_attention.cpp:163-167
    Q = rope->forward(Q);
    K = rope->forward(K);
  }

  if (use_kv_cache) { // float cache, multihead_attention.cpp:169-199,278-287
    const tcapint T_new = (tcapint)T;
    if (!k_cache) {
      if (!max_seq_len) max_seq_len = rope ? rope->max_seq_len : 2048U;
      cache_len = 0U;
      const std::vector<tcapint> cs{(tcapint)B, (tcapint)num_kv_heads, max_seq_len, (tcapint)head_dim};
      k_cache = Tensor::zeros(cs, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
      v_cache = Tensor::zeros(cs, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
    }
    TensorPtr k_slot = Tensor::slice(k_cache, 2, cache_len, T_new);
    TensorPtr v_slot = Tensor::slice(v_cache, 2, cache_len, T_new);
    Weed::add_in_place(*k_slot, *K);
    Weed::add_in_place(*v_slot, *V);
    cache_len += T_new;
    K = Tensor::slice(k_cache, 2, 0, cache_len);
    V = Tensor::slice(v_cache, 2, 0, cache_len);
  }
  if (num_kv_heads < num_heads) { // GQA broadcast, multihead_attention.cpp:290-311
    const symint groups = num_heads / num_kv_heads;
    const tcapint T_k = (tcapint)K->shape[2];
    const std::vector<tcapint> rs{(tcapint)B, (tcapint)num_heads, T_k, (tcapint)head_dim};
    TensorPtr K_rep = Tensor::zeros(rs, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
    TensorPtr V_rep = Tensor::zeros(rs, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
    for (symint g = 0; g < groups; ++g) {
      TensorPtr K_slice = Tensor::slice(K_rep, 1, (tcapint)(g * num_kv_heads), (tcapint)num_kv_heads);
      TensorPtr V_slice = Tensor::slice(V_rep, 1, (tcapint)(g * num_kv_heads), (tcapint)num_kv_heads);
      Weed::add_in_place(*K_slice, *K);
      Weed::add_in_place(*V_slice, *V);
    }
    K = K_rep;
    V = V_rep;
  }
  TensorPtr Kt = Tensor::transpose(K, -2, -1);
  TensorPtr scores = Q >> Kt;
  scores = scores / real1(std::sqrt((real1)head_dim));
  if (T > 1) {
    const tcapint T_q = (tcapint)T, T_k = use_kv_cache ? (tcapint)K->shape[2] : T_q;
    TensorPtr mask = Tensor::zeros({T_q, T_k}, false, false, DType::REAL, x->storage->device, x->storage->get_device_id());
    Weed::triu_fill(*mask, mask_val);
    scores = scores + mask;
  }
  TensorPtr weights = Tensor::softmax(scores, -1);
  out = weights >> V;
  out = Tensor::transpose(out, 1, 2);
  const symint attn_dim = (symint)num_heads * (symint)head_dim;
  out = Tensor::reshape(out, {B, T, attn_dim});
  return fuse_residual ? W_o->forward_add(out, fuse_residual) : W_o->forward(out);
}

// ------------------------------------------------------------------------------------- encoder layer
TransformerEncoderLayer::TransformerEncoderLayer(const tcapint &d_model_, const tcapint &num_heads_, const tcapint &d_ff_,
                                                 const DeviceTag &dtag, const ActivationFunctionType &afn, const int64_t &did)
    : Module(TRANSFORMER_ENCODER_LAYER_T), d_model(d_model_), d_ff(d_ff_), num_heads(num_heads_),
      self_attn(std::make_shared<MultiHeadAttention>(d_model_, num_heads_, num_heads_, 0U, dtag, nullptr, ZERO_R1, did)),
      ff1(std::make_shared<Linear>(d_model_, d_ff_, true, true, DType::REAL, dtag, did)),
      ff2(std::make_shared<Linear>(d_ff_, d_model_, true, true, DType::REAL, dtag, did)),
      norm1(std::make_shared<LayerNorm>(d_model_, dtag, FP_NORM_EPSILON, did)),
      norm2(std::make_shared<LayerNorm>(d_model_, dtag, FP_NORM_EPSILON, did)) {
  switch (afn) {
  case SIGMOID_FN: activation = std::make_shared<Sigmoid>(); break;
  case TANH_FN: activation = std::make_shared<Tanh>(); break;
  case RELU_FN: activation = std::make_shared<ReLU>(); break;
  case SWIGLU_FN: activation = std::make_shared<SwiGLU>(); break; // (as the reference: a default-constructed SwiGLU without projections)
  case GELU_FN:
  default: activation = std::make_shared<GeLU>();
  }
  _register_params();
}
void TransformerEncoderLayer::_register_params() {
  param_vector = self_attn->parameters();
  auto add = [&](const std::vector<ParameterPtr> &q) { param_vector.insert(param_vector.end(), q.begin(), q.end()); };
  add(ff1->parameters());
  add(ff2->parameters());
  add(norm1->parameters());
  add(norm2->parameters());
}
void TransformerEncoderLayer::train() {
  self_attn->train(); ff1->train(); ff2->train(); norm1->train(); norm2->train(); activation->train();
}
void TransformerEncoderLayer::eval() {
  self_attn->eval(); ff1->eval(); ff2->eval(); norm1->eval(); norm2->eval(); activation->eval();
}
void TransformerEncoderLayer::migrate_cpu() {
  self_attn->migrate_cpu(); ff1->migrate_cpu(); ff2->migrate_cpu(); norm1->migrate_cpu(); norm2->migrate_cpu();
}
void TransformerEncoderLayer::migrate_gpu() {
  self_attn->migrate_gpu(); ff1->migrate_gpu(); ff2->migrate_gpu(); norm1->migrate_gpu(); norm2->migrate_gpu();
}
// Pre-norm block, transformer_encoder_layer.cpp:63-125. The reference's per-sublayer
// migrate_gpu()/migrate_cpu() "telescoping" is offload for small VRAM; with 180 GB of HBM3e the
// parameters simply stay resident.
TensorPtr MultiHeadAttention::forward_add(const TensorPtr x, const TensorPtr residual) {
  struct Guard {
    TensorPtr &slot;
    ~Guard() { slot = nullptr; }
  } guard{fuse_residual};
  fuse_residual = residual;
  return forward(x);
}
TensorPtr TransformerEncoderLayer::forward(const TensorPtr x_) {
  TensorPtr x = x_->storage->device == DeviceTag::GPU ? x_ : x_->cast(DeviceTag::GPU);
  TensorPtr x1 = norm1->forward(x);
  // x + self_attn(...) and x1 + ff2(...): the two residual adds (transformer_encoder_layer.cpp:63-125) ride in the
  // epilogue of the W_o / ff2 product when the fused path applies; forward_add composes Linear + add otherwise
  x1 = self_attn->forward_add(x1, x);
  TensorPtr ff = norm2->forward(x1);
  TensorPtr act;
  // ff1 followed by GELU: the activation rides in the epilogue of ff1's product (Tensor::linear_gelu)
  if (ff1->bias && dynamic_cast<GeLU *>(activation.get())) act = Tensor::linear_gelu(ff, ff1->weight, ff1->bias);
  if (!act) act = activation->forward(ff1->forward(ff));
  return ff2->forward_add(act, x1);
}

// ------------------------------------------------------------------------------------- Sequential
Sequential::Sequential(const std::vector<ModulePtr> &l) : Module(SEQUENTIAL_T), layers(l) {
  for (const ModulePtr &m : layers) {
    const std::vector<ParameterPtr> p = m->parameters();
    param_vector.insert(param_vector.end(), p.begin(), p.end());
  }
}
TensorPtr Sequential::forward(const TensorPtr x) {
  TensorPtr tmp = x;
  for (const ModulePtr &m : layers) tmp = m->forward(tmp);
  return tmp;
}
TensorPtr Sequential::forward(const SymbolTensorPtr x) {
  if (layers.empty()) return std::make_shared<Tensor>();
  TensorPtr tmp = layers[0]->forward(x);
  for (size_t i = 1U; i < layers.size(); ++i) tmp = layers[i]->forward(tmp);
  return tmp;
}
} // namespace Weed
