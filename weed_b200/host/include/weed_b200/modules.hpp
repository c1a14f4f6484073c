// modules.hpp — the transformer / MLP module set on the hot path (reference include/modules/*.hpp),
// same class names, public fields, constructor argument order and forward() semantics.
#pragma once
#include "weed_b200/autograd.hpp"

namespace Weed {
enum ModuleType {
  NONE_MODULE_TYPE = 0, SEQUENTIAL_T = 1, LINEAR_T = 2, RELU_T = 3, SIGMOID_T = 4, TANH_T = 5, DROPOUT_T = 6,
  LAYERNORM_T = 7, EMBEDDING_T = 8, GRU_T = 9, LSTM_T = 10, MIGRATE_CPU_T = 11, MIGRATE_GPU_T = 12, SOFTMAX_T = 13,
  LOGSOFTMAX_T = 14, QRACK_NEURON_T = 15, QRACK_NEURON_LAYER_T = 16, MULTIHEAD_ATTENTION_T = 17,
  TRANSFORMER_ENCODER_LAYER_T = 18, GELU_T = 19, MEAN_T = 20, MIN_T = 21, MAX_T = 22, RESHAPE_T = 23, VARIANCE_T = 24,
  STDDEV_T = 25, POSITIONAL_ENCODING_T = 26, MEAN_CENTER_T = 27, FLATTEN_T = 28, LEARNED_POSITIONAL_ENCODING_T = 29,
  RMS_NORM_T = 30, ROPE_T = 31, SWIGLU_T = 32, QWEN_DECODER_LAYER_T = 33
};

struct Module;
typedef std::shared_ptr<Module> ModulePtr;
struct Module {
  ModuleType mtype;
  Module(ModuleType t) : mtype(t) {}
  virtual ~Module() {}
  virtual TensorPtr forward(const TensorPtr) = 0;
  virtual TensorPtr forward(const SymbolTensorPtr) {
    throw std::domain_error("Embedding::forward(x) takes a Tensor, not a SymbolTensor!");
  }
  virtual std::vector<ParameterPtr> parameters() { return std::vector<ParameterPtr>(); }
  virtual void train() {
    for (const auto &p : parameters()) p->train();
  }
  virtual void eval() {
    for (const auto &p : parameters()) p->eval();
  }
  virtual void migrate_cpu() {}
  virtual void migrate_gpu() {}
  virtual void set_max_kv_seq_len(tcapint) {}
  virtual void reset_cache() {}
  // checkpoint format of the reference (src/modules/module.cpp:53-375): module type tag, then the module's own fields
  virtual void save(std::ostream &) const;
  static ModulePtr load(std::istream &);
  static void write_module_type(std::ostream &out, const ModuleType &x);
  static void read_module_type(std::istream &in, ModuleType &x);
};

#define WEED_UNARY_MODULE(Name, TypeTag, expr)                                                     \
  struct Name : public Module {                                                                    \
    Name() : Module(TypeTag) {}                                                                    \
    TensorPtr forward(const TensorPtr x) override { return expr; }                                 \
  };                                                                                               \
  typedef std::shared_ptr<Name> Name##Ptr;
WEED_UNARY_MODULE(ReLU, RELU_T, Tensor::relu(x))
WEED_UNARY_MODULE(Sigmoid, SIGMOID_T, Tensor::sigmoid(x))
WEED_UNARY_MODULE(Tanh, TANH_T, Tensor::tanh(x))
WEED_UNARY_MODULE(GeLU, GELU_T, Tensor::gelu(x))
#undef WEED_UNARY_MODULE

// modules whose only state is an axis (reference include/modules/{softmax,logsoftmax,mean,max,min,variance,stddev,
// mean_center,flatten}.hpp): forward is one Tensor:: front-end, save() writes the type tag and the axis
#define WEED_AXIS_MODULE(Name, TypeTag, default_axis, expr)                                        \
  struct Name : public Module {                                                                    \
    symint axis;                                                                                   \
    Name(const symint &a = default_axis) : Module(TypeTag), axis(a) {}                             \
    TensorPtr forward(const TensorPtr x) override { return expr; }                                 \
    void save(std::ostream &os) const override;                                                    \
  };                                                                                               \
  typedef std::shared_ptr<Name> Name##Ptr;
WEED_AXIS_MODULE(Softmax, SOFTMAX_T, -1, Tensor::softmax(x, axis))
WEED_AXIS_MODULE(LogSoftmax, LOGSOFTMAX_T, -1, Tensor::logsoftmax(x, axis))
WEED_AXIS_MODULE(Mean, MEAN_T, 0, Tensor::mean(x, axis))
WEED_AXIS_MODULE(Max, MAX_T, -1, Tensor::max(x, axis))
WEED_AXIS_MODULE(Min, MIN_T, -1, Tensor::min(x, axis))
WEED_AXIS_MODULE(Variance, VARIANCE_T, 0, Tensor::variance(x, (tcapint)axis))
WEED_AXIS_MODULE(Stddev, STDDEV_T, 0, Tensor::stddev(x, (tcapint)axis))
WEED_AXIS_MODULE(MeanCenter, MEAN_CENTER_T, 0, x - Tensor::mean(x, axis))
WEED_AXIS_MODULE(Flatten, FLATTEN_T, -1, Tensor::flatten(x, axis))
#undef WEED_AXIS_MODULE

struct Reshape : public Module { // reference include/modules/reshape.hpp
  std::vector<symint> shape;
  Reshape(const std::vector<symint> &s) : Module(RESHAPE_T), shape(s) {}
  TensorPtr forward(const TensorPtr x) override { return Tensor::reshape(x, shape); }
  void save(std::ostream &os) const override;
};
typedef std::shared_ptr<Reshape> ReshapePtr;

// reference include/modules/dropout.hpp, src/modules/dropout.cpp: Bernoulli(1 - p) mask from std::random_device, y = x * mask / (1 - p)
struct Dropout : public Module {
  real1 p;
  bool training;
  TensorPtr mask;
  Dropout() : Module(DROPOUT_T), p(ZERO_R1), training(true) {}
  Dropout(real1 prob);
  void train() override {
    Module::train();
    training = true;
  }
  void eval() override {
    Module::eval();
    training = false;
  }
  TensorPtr forward(const TensorPtr x) override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<Dropout> DropoutPtr;

// parameter migration helpers (reference include/modules/migrate_gpu.hpp / migrate_cpu.hpp)
struct MigrateGpu : public Module {
  MigrateGpu() : Module(MIGRATE_GPU_T) {}
  TensorPtr forward(const TensorPtr x) override { return x->cast(DeviceTag::GPU); }
  ParameterPtr pforward(const ParameterPtr p);
};
struct MigrateCpu : public Module {
  MigrateCpu() : Module(MIGRATE_CPU_T) {}
  TensorPtr forward(const TensorPtr x) override { return x->cast(DeviceTag::CPU); }
  ParameterPtr pforward(const ParameterPtr p);
};

typedef std::shared_ptr<MigrateGpu> MigrateGpuPtr;
typedef std::shared_ptr<MigrateCpu> MigrateCpuPtr;

struct Linear : public Module {
  tcapint in_features, out_features;
  ParameterPtr weight; // (in_features, out_features), column-major
  ParameterPtr bias;   // (out_features) or null
  Linear() : Module(LINEAR_T) {}
  Linear(tcapint in_f, tcapint out_f, bool use_bias = true, bool init_rand = true, DType dtype = DType::REAL,
         DeviceTag device = DeviceTag::DEFAULT_DEVICE, int64_t device_id = -1);
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr x) override;
  // residual + forward(x): one kernel when the fused tensor-core path applies (the residual rides in the GEMM
  // epilogue), otherwise exactly Tensor::add(residual, forward(x)) as TransformerEncoderLayer::forward composes it
  TensorPtr forward_add(const TensorPtr x, const TensorPtr residual);
  std::vector<ParameterPtr> parameters() override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<Linear> LinearPtr;

struct LayerNorm : Module {
  tcapint features;
  real1 eps;
  ParameterPtr gamma, beta; // shape [1, 1, features]
  LayerNorm() : Module(LAYERNORM_T) {}
  LayerNorm(const tcapint &f, const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const real1 &e = FP_NORM_EPSILON,
            const int64_t &did = -1);
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr x) override;
  std::vector<ParameterPtr> parameters() override { return {gamma, beta}; }
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<LayerNorm> LayerNormPtr;

struct Embedding : public Module {
  tcapint num_embeddings, embedding_dim;
  ParameterPtr weight; // [vocab, dim] strides (1, vocab)
  Embedding() : Module(EMBEDDING_T) {}
  Embedding(const tcapint &vocab, const tcapint &dim, const DType &dtype = DType::REAL,
            const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, int64_t did = -1);
  TensorPtr forward(const TensorPtr) override { throw std::domain_error("Embedding::forward(x) takes a SymbolTensor, not a Tensor!"); }
  TensorPtr forward(const SymbolTensorPtr t) override;
  std::vector<ParameterPtr> parameters() override { return {weight}; }
  void migrate_cpu() override;
  void migrate_gpu() override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<Embedding> EmbeddingPtr;

struct LearnedPositionalEncoding : public Module {
  tcapint max_len, d_model;
  ParameterPtr pos_encoding; // (1, max_len, d_model)
  LearnedPositionalEncoding() : Module(LEARNED_POSITIONAL_ENCODING_T) {}
  LearnedPositionalEncoding(const tcapint &max_len_, const tcapint &d_model_, const DeviceTag &dtag = DEFAULT_DEVICE);
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr x) override;
  std::vector<ParameterPtr> parameters() override { return {pos_encoding}; }
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<LearnedPositionalEncoding> LearnedPositionalEncodingPtr;

// reference include/modules/positional_encoding.hpp, src/modules/positional_encoding.cpp: fixed sinusoidal table
struct PositionalEncoding : public Module {
  tcapint max_seq_len, d_model;
  real1_f pos_val;
  ParameterPtr pe; // [max_seq_len, d_model], never requires_grad
  PositionalEncoding(tcapint max_seq_len, tcapint d_model, real1_f pos_val_ = 8192.0, DeviceTag device = DEFAULT_DEVICE);
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr x) override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<PositionalEncoding> PositionalEncodingPtr;

// reference include/modules/rope.hpp, src/modules/rope.cpp: rotary position embedding on [B, H, T, head_dim]
struct RoPE : public Module {
  tcapint head_dim, max_seq_len;
  real1_f base;
  TensorPtr cos_table, sin_table; // [max_seq_len, head_dim]
  RoPE() : Module(ROPE_T), head_dim(0U), max_seq_len(0U), base(10000.0f) {}
  RoPE(const tcapint &head_dim_, const tcapint &max_seq_len_ = 2048U, const real1_f &base_ = 10000.0f);
  void _build_tables();
  TensorPtr _rotate_half(const TensorPtr x);
  TensorPtr forward(const TensorPtr x) override;
  void save(std::ostream &os) const override;
};
typedef std::shared_ptr<RoPE> RoPEPtr;

// reference include/modules/rms_norm.hpp:25-49: x / (mean(x*x, axis) + eps)^0.5 * weight
struct RMSNorm : public Module {
  symint axis;
  tcapint hidden_size;
  ParameterPtr weight;
  RMSNorm() : Module(RMS_NORM_T), axis(-1), hidden_size(0U) {}
  RMSNorm(const tcapint &hidden_size_, const symint &axis_ = -1);
  std::vector<ParameterPtr> parameters() override { return {weight}; }
  TensorPtr forward(const TensorPtr x) override;
  void save(std::ostream &os) const override;
};
typedef std::shared_ptr<RMSNorm> RMSNormPtr;

// reference include/modules/swiglu.hpp:20-90: down(silu(gate(x)) * up(x)), three bias-free Linear layers
struct SwiGLU : public Module {
  tcapint hidden_size, intermediate_size;
  LinearPtr gate_proj, up_proj, down_proj;
  std::vector<ParameterPtr> param_vector;
  SwiGLU() : Module(SWIGLU_T), hidden_size(0U), intermediate_size(0U) {}
  SwiGLU(const tcapint &hidden_size_, const tcapint &intermediate_size_);
  void _register_params();
  std::vector<ParameterPtr> parameters() override { return param_vector; }
  void train() override;
  void eval() override;
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr x) override;
  void save(std::ostream &os) const override;
};
typedef std::shared_ptr<SwiGLU> SwiGLUPtr;

// reference include/modules/gru.hpp, src/modules/gru.cpp:17-43
struct GRU : public Module {
  tcapint input_dim, hidden_dim;
  LinearPtr W_x, W_h; // x -> 3H, h -> 3H
  TensorPtr state;
  GRU() : Module(GRU_T), input_dim(0U), hidden_dim(0U) {}
  GRU(tcapint in, tcapint hid, DeviceTag dtag = DeviceTag::DEFAULT_DEVICE);
  std::vector<ParameterPtr> parameters() override;
  void train() override;
  void eval() override;
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr) override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<GRU> GRUPtr;

// reference include/modules/lstm.hpp, src/modules/lstm.cpp:17-55
struct LSTMState {
  TensorPtr h, c;
};
struct LSTM : public Module {
  tcapint input_dim, hidden_dim;
  LinearPtr W_x, W_h; // input -> 4H, hidden -> 4H
  LSTMState state;
  LSTM() : Module(LSTM_T), input_dim(0U), hidden_dim(0U) {}
  LSTM(tcapint in, tcapint hid, DeviceTag dtag = DEFAULT_DEVICE, const int64_t &did = -1);
  std::vector<ParameterPtr> parameters() override;
  void train() override;
  void eval() override;
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr) override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<LSTM> LSTMPtr;

struct MultiHeadAttention : public Module {
  symint d_model, num_heads, num_kv_heads, head_dim;
  real1_f mask_val;
  LinearPtr W_q, W_k, W_v, W_o;
  RoPEPtr rope;
  bool use_kv_cache;
  TensorPtr k_cache, v_cache; // float KV cache (kv_quant_bits == 0), [B, H_kv, max_seq_len, hd]
  tcapint cache_len = 0U;
  tcapint max_seq_len = 0U;
  int kv_quant_bits = 0;
  std::vector<ParameterPtr> param_vector;

  MultiHeadAttention()
      : Module(MULTIHEAD_ATTENTION_T), d_model(0), num_heads(0), num_kv_heads(0), head_dim(0), mask_val(ZERO_R1),
        use_kv_cache(false) {}
  MultiHeadAttention(tcapint d_model_, tcapint num_heads_, tcapint num_kv_heads_ = 0, tcapint head_dim_ = 0U,
                     DeviceTag dtag = DEFAULT_DEVICE, RoPEPtr r = nullptr, real1_f mask_val_ = ZERO_R1, const int64_t did = -1,
                     const bool _use_kv_cache = true, int kv_quant_bits_ = 4);
  std::vector<ParameterPtr> parameters() override { return param_vector; }
  void train() override;
  void eval() override;
  void set_max_kv_seq_len(tcapint m) override { max_seq_len = m; }
  void reset_cache() override;
  void migrate_cpu() override;
  void migrate_gpu() override;
  TensorPtr forward(const TensorPtr x) override;
  // residual + forward(x) with the add folded into the output projection (Linear::forward_add)
  TensorPtr forward_add(const TensorPtr x, const TensorPtr residual);
  TensorPtr fuse_residual; // set for the duration of forward_add: the W_o projection adds it
  void save(std::ostream &) const override;
  void _register_params(); // (the reference's load leaves param_vector empty; here a loaded module can resume training)
};
typedef std::shared_ptr<MultiHeadAttention> MultiHeadAttentionPtr;

struct TransformerEncoderLayer : public Module {
  tcapint d_model, d_ff, num_heads;
  MultiHeadAttentionPtr self_attn;
  LinearPtr ff1, ff2;
  LayerNormPtr norm1, norm2;
  ModulePtr activation;
  std::vector<ParameterPtr> param_vector;
  TransformerEncoderLayer() : Module(TRANSFORMER_ENCODER_LAYER_T) {}
  TransformerEncoderLayer(const tcapint &d_model_, const tcapint &num_heads_, const tcapint &d_ff_,
                          const DeviceTag &dtag = DEFAULT_DEVICE, const ActivationFunctionType &afn = GELU_FN,
                          const int64_t &did = -1);
  std::vector<ParameterPtr> parameters() override { return param_vector; }
  void train() override;
  void eval() override;
  void migrate_cpu() override;
  void migrate_gpu() override;
  void set_max_kv_seq_len(tcapint m) override { self_attn->set_max_kv_seq_len(m); }
  void reset_cache() override { self_attn->reset_cache(); }
  TensorPtr forward(const TensorPtr x) override;
  void save(std::ostream &) const override;
  void _register_params();
};
typedef std::shared_ptr<TransformerEncoderLayer> TransformerEncoderLayerPtr;

// reference include/modules/qwen_decoder_layer.hpp:24-123: pre-norm (RMSNorm) attention with RoPE + SwiGLU MLP, residuals
struct QwenDecoderLayer : public Module {
  tcapint d_model, num_heads, num_kv_heads;
  MultiHeadAttentionPtr self_attn;
  SwiGLUPtr mlp;
  RMSNormPtr input_layernorm, post_attention_layernorm;
  std::vector<ParameterPtr> param_vector;
  QwenDecoderLayer() : Module(QWEN_DECODER_LAYER_T), d_model(0U), num_heads(0U), num_kv_heads(0U) {}
  QwenDecoderLayer(const tcapint &d_model_, const tcapint &num_heads_, const tcapint &num_kv_heads_, const tcapint &d_ff_,
                   const tcapint &max_seq_len = 2048U, const real1_f &rope_base = 10000.0f, const real1_f &eps = 1e-6f, const int64_t &did = -1);
  void _register_params();
  void train() override;
  void eval() override;
  void migrate_cpu() override;
  void migrate_gpu() override;
  void set_max_kv_seq_len(tcapint m) override { self_attn->set_max_kv_seq_len(m); }
  void reset_cache() override { self_attn->reset_cache(); }
  std::vector<ParameterPtr> parameters() override { return param_vector; }
  TensorPtr forward(const TensorPtr x) override;
  void save(std::ostream &os) const override;
};
typedef std::shared_ptr<QwenDecoderLayer> QwenDecoderLayerPtr;

struct Sequential : public Module {
  std::vector<ModulePtr> layers;
  std::vector<ParameterPtr> param_vector;
  Sequential(const std::vector<ModulePtr> &l);
  void train() override { for (const ModulePtr &m : layers) m->train(); }
  void eval() override { for (const ModulePtr &m : layers) m->eval(); }
  void migrate_cpu() override { for (const ModulePtr &m : layers) m->migrate_cpu(); }
  void migrate_gpu() override { for (const ModulePtr &m : layers) m->migrate_gpu(); }
  void set_max_kv_seq_len(tcapint m) override { for (const ModulePtr &md : layers) md->set_max_kv_seq_len(m); }
  void reset_cache() override { for (const ModulePtr &m : layers) m->reset_cache(); }
  TensorPtr forward(const TensorPtr x) override;
  TensorPtr forward(const SymbolTensorPtr x) override;
  std::vector<ParameterPtr> parameters() override { return param_vector; }
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<Sequential> SequentialPtr;
} // namespace Weed
