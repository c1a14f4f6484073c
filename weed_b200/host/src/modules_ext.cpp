// modules_ext.cpp — the remaining real-dtype module families of the reference on the CUDA device:
// Dropout, PositionalEncoding, RoPE, RMSNorm, SwiGLU, QwenDecoderLayer, GRU, LSTM. Each forward() is the reference's
// own composition of Tensor:: front-ends (cited per function), so autograd and the device kernels underneath are the ones
// the hot-path modules use; nothing here adds a kernel.
#include "weed_b200/modules.hpp"

#include <cmath>
#include <random>

namespace Weed {
namespace {
const DeviceTag kDev = DeviceTag::GPU; // DEFAULT_DEVICE on this backend
} // namespace

// ------------------------------------------------------------------------------------- Dropout
Dropout::Dropout(real1 prob) : Module(DROPOUT_T), p(prob), training(true), mask(nullptr) {
  if ((p < ZERO_R1) || (p >= ONE_R1))
    throw std::invalid_argument("Dropout probability must be at least 0.0 and cannot be greater than or equal to 1.0!");
}
TensorPtr Dropout::forward(const TensorPtr x) { // src/modules/dropout.cpp:17-43
  if (!training || p == ZERO_R1) return x;
  std::uniform_real_distribution<real1_s> dis((real1_s)ZERO_R1, (real1_s)ONE_R1);
  std::random_device rd;
  std::mt19937 gen(rd());
  // mask[i] = 1 with probability 1 - p (the reference builds it sparse on the host; here dense, uploaded once)
  const tcapint sz = x->get_broadcast_size();
  std::vector<real1> m(sz, ZERO_R1);
  for (tcapint n = 0; n < sz; ++n)
    if (dis(gen) > p) m[n] = ONE_R1;
  mask = std::make_shared<Tensor>(m, x->shape, false, x->storage->device, x->storage->get_device_id());
  return (x * mask) / real1(ONE_R1 - p);
}

// ------------------------------------------------------------------------------------- PositionalEncoding
PositionalEncoding::PositionalEncoding(tcapint max_seq_len_, tcapint d_model_, real1_f pos_val_, DeviceTag device)
    : Module(POSITIONAL_ENCODING_T), max_seq_len(max_seq_len_), d_model(d_model_), pos_val(pos_val_) {
  // src/modules/positional_encoding.cpp:27-52, index arithmetic included: the table is filled at pos * d_model + i and
  // then read as a column-major [max_seq_len, d_model] tensor
  const tcapint max_i = d_model >> 1U;
  std::vector<real1> values((size_t)max_seq_len * d_model);
  for (tcapint pos = 0; pos < max_seq_len; ++pos) {
    for (tcapint i = 0; i < max_i; ++i) {
      const tcapint idx = pos * d_model + (i << 1U);
      const real1 coeff = (real1)(1.0 / std::pow(pos_val, ((real1)(i << 1U)) / d_model));
      values[idx] = std::cos(coeff * pos);
      values[idx + 1U] = std::sin(coeff * pos);
    }
    if (d_model & 1U) {
      const tcapint idx = pos * d_model + (max_i << 1U);
      const real1 div = (real1)std::pow(pos_val, (2.0 * max_i) / d_model);
      values[idx] = std::cos(pos / div);
    }
  }
  pe = std::make_shared<Parameter>(values, std::vector<tcapint>{max_seq_len, d_model}, device);
  pe->eval(); // never requires_grad
}
void PositionalEncoding::migrate_cpu() {
  MigrateCpu mc;
  pe = mc.pforward(pe);
}
void PositionalEncoding::migrate_gpu() {
  MigrateGpu mg;
  pe = mg.pforward(pe);
}
TensorPtr PositionalEncoding::forward(const TensorPtr x) { // :53-60
  const tcapint T = x->shape[1U];
  return x + Tensor::slice(pe, 0, 0, T);
}

// ------------------------------------------------------------------------------------- RoPE
RoPE::RoPE(const tcapint &head_dim_, const tcapint &max_seq_len_, const real1_f &base_)
    : Module(ROPE_T), head_dim(head_dim_), max_seq_len(max_seq_len_), base(base_) {
  _build_tables();
}
void RoPE::_build_tables() { // src/modules/rope.cpp:19-50: theta_i = base^(-2i / head_dim), both elements of a pair share cos / sin
  const tcapint half = head_dim >> 1U;
  std::vector<real1> cos_data((size_t)max_seq_len * head_dim), sin_data((size_t)max_seq_len * head_dim);
  for (tcapint pos = 0U; pos < max_seq_len; ++pos) {
    for (tcapint i = 0U; i < half; ++i) {
      const real1_f theta = (real1_f)std::pow(base, -2.0f * (real1_f)i / (real1_f)head_dim);
      const real1_f angle = (real1_f)pos * theta;
      const real1 c = (real1)std::cos(angle), s = (real1)std::sin(angle);
      cos_data[pos + (size_t)(2U * i) * max_seq_len] = c; // column-major [max_seq_len, head_dim]
      cos_data[pos + (size_t)(2U * i + 1U) * max_seq_len] = c;
      sin_data[pos + (size_t)(2U * i) * max_seq_len] = s;
      sin_data[pos + (size_t)(2U * i + 1U) * max_seq_len] = s;
    }
  }
  cos_table = std::make_shared<Tensor>(cos_data, std::vector<tcapint>{max_seq_len, head_dim}, false, kDev);
  sin_table = std::make_shared<Tensor>(sin_data, std::vector<tcapint>{max_seq_len, head_dim}, false, kDev);
}
TensorPtr RoPE::_rotate_half(const TensorPtr x) { // :52-79: out = [-x[..., half:], x[..., :half]] on [B, H, T, head_dim]
  const tcapint B = x->shape[0U], H = x->shape[1U], T = x->shape[2U], D = head_dim, half = head_dim >> 1U;
  TensorPtr out = Tensor::zeros({B, H, T, D}, x->requires_grad, false, x->storage->dtype, x->storage->device, x->storage->get_device_id());
  TensorPtr x0 = Tensor::slice(x, 3, 0U, half), x1 = Tensor::slice(x, 3, half, half);
  TensorPtr out0 = Tensor::slice(out, 3, 0U, half), out1 = Tensor::slice(out, 3, half, half);
  Weed::add_in_place(*out0, *(x1 * real1(-1.0f)));
  Weed::add_in_place(*out1, *x0);
  return out;
}
TensorPtr RoPE::forward(const TensorPtr x) { // :81-99: x * cos + rotate_half(x) * sin, positions 0 .. T-1 of this call
  const symint T = (symint)x->shape[2U];
  TensorPtr c = Tensor::slice(cos_table, 0, 0U, (tcapint)T), s = Tensor::slice(sin_table, 0, 0U, (tcapint)T);
  TensorPtr cos_b = Tensor::reshape(c, {1, 1, T, (symint)head_dim}), sin_b = Tensor::reshape(s, {1, 1, T, (symint)head_dim});
  return x * cos_b + _rotate_half(x) * sin_b;
}

// ------------------------------------------------------------------------------------- RMSNorm
RMSNorm::RMSNorm(const tcapint &hidden_size_, const symint &axis_) : Module(RMS_NORM_T), axis(axis_), hidden_size(hidden_size_) {
  weight = std::make_shared<Parameter>(std::vector<tcapint>{hidden_size}, std::vector<tcapint>{1U}, false);
  weight->storage->FillOnes();
}
TensorPtr RMSNorm::forward(const TensorPtr x) { // include/modules/rms_norm.hpp:37-41
  return (x / ((Tensor::mean(x * x, axis) + real1(FP_NORM_EPSILON)) ^ real1(0.5f))) * weight;
}

// ------------------------------------------------------------------------------------- SwiGLU
SwiGLU::SwiGLU(const tcapint &hidden_size_, const tcapint &intermediate_size_)
    : Module(SWIGLU_T), hidden_size(hidden_size_), intermediate_size(intermediate_size_) {
  gate_proj = std::make_shared<Linear>(hidden_size, intermediate_size, false);
  up_proj = std::make_shared<Linear>(hidden_size, intermediate_size, false);
  down_proj = std::make_shared<Linear>(intermediate_size, hidden_size, false);
}
void SwiGLU::_register_params() {
  param_vector.clear();
  for (const LinearPtr &l : {gate_proj, up_proj, down_proj}) {
    const std::vector<ParameterPtr> q = l->parameters();
    param_vector.insert(param_vector.end(), q.begin(), q.end());
  }
}
void SwiGLU::train() {
  for (const LinearPtr &l : {gate_proj, up_proj, down_proj}) l->train();
}
void SwiGLU::eval() {
  for (const LinearPtr &l : {gate_proj, up_proj, down_proj}) l->eval();
}
void SwiGLU::migrate_cpu() {
  for (const LinearPtr &l : {gate_proj, up_proj, down_proj}) l->migrate_cpu();
}
void SwiGLU::migrate_gpu() {
  for (const LinearPtr &l : {gate_proj, up_proj, down_proj}) l->migrate_gpu();
}
TensorPtr SwiGLU::forward(const TensorPtr x) { // include/modules/swiglu.hpp:78-84: SiLU(gate) * up, then down
  TensorPtr gate = gate_proj->forward(x), up = up_proj->forward(x);
  TensorPtr activated = gate * Tensor::sigmoid(gate) * up;
  return down_proj->forward(activated);
}

// ------------------------------------------------------------------------------------- QwenDecoderLayer
QwenDecoderLayer::QwenDecoderLayer(const tcapint &d_model_, const tcapint &num_heads_, const tcapint &num_kv_heads_, const tcapint &d_ff_,
                                   const tcapint &max_seq_len, const real1_f &rope_base, const real1_f &, const int64_t &did)
    : Module(QWEN_DECODER_LAYER_T), d_model(d_model_), num_heads(num_heads_), num_kv_heads(num_kv_heads_) {
  const tcapint head_dim = d_model_ / num_heads_;
  RoPEPtr rope = std::make_shared<RoPE>(head_dim, max_seq_len, rope_base);
  self_attn = std::make_shared<MultiHeadAttention>(d_model_, num_heads_, num_kv_heads_, head_dim, DEFAULT_DEVICE, rope, ZERO_R1, did);
  mlp = std::make_shared<SwiGLU>(d_model_, d_ff_);
  input_layernorm = std::make_shared<RMSNorm>(d_model_, -1);
  post_attention_layernorm = std::make_shared<RMSNorm>(d_model_, -1);
}
void QwenDecoderLayer::_register_params() {
  param_vector.clear();
  auto add = [&](const std::vector<ParameterPtr> &q) { param_vector.insert(param_vector.end(), q.begin(), q.end()); };
  add(self_attn->parameters());
  add(mlp->parameters());
  add(input_layernorm->parameters());
  add(post_attention_layernorm->parameters());
}
void QwenDecoderLayer::train() {
  self_attn->train();
  mlp->train();
  input_layernorm->train();
  post_attention_layernorm->train();
  for (auto &p : param_vector) p->train();
}
void QwenDecoderLayer::eval() {
  self_attn->eval();
  mlp->eval();
  input_layernorm->eval();
  post_attention_layernorm->eval();
  for (auto &p : param_vector) p->eval();
}
void QwenDecoderLayer::migrate_cpu() {
  self_attn->migrate_cpu();
  mlp->migrate_cpu();
  input_layernorm->migrate_cpu();
  post_attention_layernorm->migrate_cpu();
}
void QwenDecoderLayer::migrate_gpu() {
  self_attn->migrate_gpu();
  mlp->migrate_gpu();
  input_layernorm->migrate_gpu();
  post_attention_layernorm->migrate_gpu();
}
TensorPtr QwenDecoderLayer::forward(const TensorPtr x) { // include/modules/qwen_decoder_layer.hpp:104-116
  TensorPtr residual = x;
  TensorPtr h = self_attn->forward(input_layernorm->forward(x));
  h = h + residual;
  residual = h;
  h = mlp->forward(post_attention_layernorm->forward(h));
  return h + residual;
}

// ------------------------------------------------------------------------------------- GRU / LSTM
namespace {
// the recurrent state starts as [H] and is broadcast to [B, H] on first use (gru.cpp:18-22, lstm.cpp:18-27)
void expand_state(TensorPtr &state, const TensorPtr &x) {
  if (state->shape.size() != 1U) return;
  state->shape.insert(state->shape.begin(), x->shape[0U]);
  state->stride.insert(state->stride.begin(), 0U);
  state->materialize_broadcast();
}
} // namespace
GRU::GRU(tcapint in, tcapint hid, DeviceTag dtag)
    : Module(GRU_T), input_dim(in), hidden_dim(hid), W_x(std::make_shared<Linear>(in, 3 * hid, true, true, DType::REAL, dtag)),
      W_h(std::make_shared<Linear>(hid, 3 * hid, true, true, DType::REAL, dtag)), state(Tensor::zeros({hidden_dim})) {}
std::vector<ParameterPtr> GRU::parameters() {
  std::vector<ParameterPtr> px = W_x->parameters();
  const std::vector<ParameterPtr> ph = W_h->parameters();
  px.insert(px.end(), ph.begin(), ph.end());
  return px;
}
void GRU::train() {
  W_x->train();
  W_h->train();
}
void GRU::eval() {
  W_x->eval();
  W_h->eval();
}
void GRU::migrate_cpu() {
  W_x->migrate_cpu();
  W_h->migrate_cpu();
}
void GRU::migrate_gpu() {
  W_x->migrate_gpu();
  W_h->migrate_gpu();
}
TensorPtr GRU::forward(const TensorPtr x) { // src/modules/gru.cpp:17-43
  expand_state(state, x);
  TensorPtr z = W_x->forward(x) + W_h->forward(state);
  const std::vector<TensorPtr> zc = Tensor::chunk(z, 3, -1);
  TensorPtr z_t = Tensor::sigmoid(zc[0U]), r_t = Tensor::sigmoid(zc[1U]);
  TensorPtr h_tilde = Tensor::tanh(zc[2U] + W_h->forward(r_t * state)); // (as written: W_h is 3H wide, the sum broadcasts — reference behaviour)
  return (Tensor::ones_like(z_t->shape) - z_t) * state + z_t * h_tilde;
}
LSTM::LSTM(tcapint in, tcapint hid, DeviceTag dtag, const int64_t &did)
    : Module(LSTM_T), input_dim(in), hidden_dim(hid), W_x(std::make_shared<Linear>(in, 4 * hid, true, true, DType::REAL, dtag, did)),
      W_h(std::make_shared<Linear>(hid, 4 * hid, true, true, DType::REAL, dtag, did)),
      state{Tensor::zeros(std::vector<tcapint>{hidden_dim}), Tensor::zeros(std::vector<tcapint>{hidden_dim})} {}
std::vector<ParameterPtr> LSTM::parameters() {
  std::vector<ParameterPtr> px = W_x->parameters();
  const std::vector<ParameterPtr> ph = W_h->parameters();
  px.insert(px.end(), ph.begin(), ph.end());
  return px;
}
void LSTM::train() {
  W_x->train();
  W_h->train();
}
void LSTM::eval() {
  W_x->eval();
  W_h->eval();
}
void LSTM::migrate_cpu() {
  W_x->migrate_cpu();
  W_h->migrate_cpu();
}
void LSTM::migrate_gpu() {
  W_x->migrate_gpu();
  W_h->migrate_gpu();
}
TensorPtr LSTM::forward(const TensorPtr x) { // src/modules/lstm.cpp:17-55
  expand_state(state.h, x);
  expand_state(state.c, x);
  TensorPtr z = W_x->forward(x) + W_h->forward(state.h);
  const std::vector<TensorPtr> zc = Tensor::chunk(z, 4, -1);
  TensorPtr f = Tensor::sigmoid(zc[0U]), i = Tensor::sigmoid(zc[1U]), g = Tensor::tanh(zc[2U]), o = Tensor::sigmoid(zc[3U]);
  TensorPtr c = f * state.c + i * g;
  TensorPtr h = o * Tensor::tanh(c);
  state.h = h;
  state.c = c;
  return h;
}
} // namespace Weed
