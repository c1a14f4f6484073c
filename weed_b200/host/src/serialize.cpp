// serialize.cpp — Weed's checkpoint format on the CUDA device.
//
// Byte layout as the reference writes it (include/common/serializer.hpp:25-99: raw little-endian fields, no framing):
//   Storage    src/storage/storage.cpp:25-119        StorageType (4) | device id (8) | element count (4) | elements
//   Parameter  src/tensors/parameter.cpp:16-56       device id (4) | offset (4) | rank (4) | rank x (shape, stride) | Storage
//   Module     src/modules/module.cpp:53-375 + each module's save(): ModuleType (4) | the module's own fields, children in order
// A device storage is read back to the host for writing (src/storage/gpu_real_storage.cpp:35-50). With
// BackendConfig::save_portable (default) it is tagged REAL_CPU_DENSE / device id -1, i.e. exactly what the reference's CPU
// build writes for the same weights; loading accepts the CPU and the GPU tags and always places the data on the device,
// because this backend has no CPU compute path.
#include "weed_b200/modules.hpp"

#include <istream>
#include <ostream>

namespace Weed {
namespace {
template <typename T> void put(std::ostream &out, const T &x) { out.write(reinterpret_cast<const char *>(&x), sizeof(T)); }
template <typename T> void get(std::istream &in, T &x) {
  in.read(reinterpret_cast<char *>(&x), sizeof(T));
  if (!in) throw std::domain_error("Unexpected end of stream while loading a Weed checkpoint!");
}
} // namespace

void Serializer::write_bool(std::ostream &out, const bool &x) { put(out, x); }
void Serializer::read_bool(std::istream &in, bool &x) { get(in, x); }
void Serializer::write_tcapint(std::ostream &out, const tcapint &x) { put(out, x); }
void Serializer::read_tcapint(std::istream &in, tcapint &x) { get(in, x); }
void Serializer::write_symint(std::ostream &out, const symint &x) { put(out, x); }
void Serializer::read_symint(std::istream &in, symint &x) { get(in, x); }
void Serializer::write_int64(std::ostream &out, const symint &x) { put(out, (int64_t)x); }
void Serializer::read_int64(std::istream &in, symint &x) {
  int64_t wide;
  get(in, wide);
  x = (symint)(wide & 0xffffffffLL); // the high word of a reference-written file is not part of the value
}
void Serializer::write_size_t(std::ostream &out, const size_t &x) { put(out, x); }
void Serializer::read_size_t(std::istream &in, size_t &x) { get(in, x); }
void Serializer::write_real(std::ostream &out, const real1 &x) { put(out, x); }
void Serializer::read_real(std::istream &in, real1 &x) { get(in, x); }
void Serializer::write_real1_f(std::ostream &out, const real1_f &x) { put(out, x); }
void Serializer::read_real1_f(std::istream &in, real1_f &x) { get(in, x); }

// ------------------------------------------------------------------------------------- Storage
void Storage::write_storage_type(std::ostream &out, const StorageType &x) { put(out, x); }
void Storage::read_storage_type(std::istream &in, StorageType &x) { get(in, x); }
void Storage::save(std::ostream &os) const { // header only; the typed storages append their elements
  write_storage_type(os, stype);
  Serializer::write_int64(os, (symint)get_device_id());
  Serializer::write_tcapint(os, size);
}
namespace {
template <typename T> void write_elements(std::ostream &os, const std::vector<T> &v) {
  os.write(reinterpret_cast<const char *>(v.data()), (std::streamsize)(sizeof(T) * v.size()));
}
void write_header(std::ostream &os, StorageType stype, int64_t did, tcapint size) {
  Storage::write_storage_type(os, stype);
  Serializer::write_int64(os, (symint)did);
  Serializer::write_tcapint(os, size);
}
} // namespace
void CpuRealStorage::save(std::ostream &os) const {
  Storage::save(os);
  write_elements(os, data);
}
void CpuIntStorage::save(std::ostream &os) const {
  Storage::save(os);
  write_elements(os, data);
}
void GpuRealStorage::save(std::ostream &os) const {
  const bool portable = backend_config().save_portable;
  write_header(os, portable ? REAL_CPU_DENSE : REAL_GPU_DENSE, portable ? -1 : get_device_id(), size);
  StoragePtr host = const_cast<GpuRealStorage *>(this)->cpu(); // one blocking read-back (materialises deferred / lazily zeroed values)
  write_elements(os, static_cast<CpuRealStorage *>(host.get())->data);
}
void GpuIntStorage::save(std::ostream &os) const {
  const bool portable = backend_config().save_portable;
  write_header(os, portable ? INT_CPU_DENSE : INT_GPU_DENSE, portable ? -1 : get_device_id(), size);
  StoragePtr host = const_cast<GpuIntStorage *>(this)->cpu();
  write_elements(os, static_cast<CpuIntStorage *>(host.get())->data);
}
StoragePtr Storage::load(std::istream &is) {
  StorageType stype;
  read_storage_type(is, stype);
  symint did;
  Serializer::read_int64(is, did);
  tcapint size;
  Serializer::read_tcapint(is, size);
  switch (stype) {
  case StorageType::REAL_CPU_DENSE:
  case StorageType::REAL_GPU_DENSE: {
    std::vector<real1> v(size);
    is.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(sizeof(real1) * (size_t)size));
    if (!is) throw std::domain_error("Unexpected end of stream in Storage::load!");
    return std::make_shared<GpuRealStorage>(v, stype == StorageType::REAL_GPU_DENSE ? (int64_t)did : (int64_t)-1);
  }
  case StorageType::INT_CPU_DENSE:
  case StorageType::INT_GPU_DENSE: {
    std::vector<symint> v(size);
    is.read(reinterpret_cast<char *>(v.data()), (std::streamsize)(sizeof(symint) * (size_t)size));
    if (!is) throw std::domain_error("Unexpected end of stream in Storage::load!");
    return std::make_shared<GpuIntStorage>(v, stype == StorageType::INT_GPU_DENSE ? (int64_t)did : (int64_t)-1);
  }
  case StorageType::COMPLEX_CPU_DENSE:
  case StorageType::COMPLEX_GPU_DENSE:
  case StorageType::REAL_CPU_SPARSE:
  case StorageType::COMPLEX_CPU_SPARSE:
    throw std::domain_error("Storage::load: complex and sparse storages are outside the CUDA backend's scope (SURVEY §8)");
  case StorageType::NONE_STORAGE_TYPE:
  default:
    throw std::domain_error("Can't recognize StorageType in Storage::load!");
  }
}

// ------------------------------------------------------------------------------------- Parameter
void Parameter::save(std::ostream &out) { // src/tensors/parameter.cpp:16-30 (un-broadcasts the match_shape-mutated view first)
  Serializer::write_symint(out, backend_config().save_portable ? (symint)-1 : (symint)storage->get_device_id());
  for (size_t i = 0U; i < stride.size(); ++i)
    if (!stride[i]) shape[i] = 1U;
  Serializer::write_tcapint(out, offset);
  Serializer::write_tcapint(out, (tcapint)shape.size());
  for (size_t i = 0U; i < shape.size(); ++i) {
    Serializer::write_tcapint(out, shape[i]);
    Serializer::write_tcapint(out, stride[i]);
  }
  storage->save(out);
}
ParameterPtr Parameter::load(std::istream &in) { // :31-56 (the stored offset is read and, as in the reference, not applied)
  symint did;
  Serializer::read_symint(in, did);
  tcapint offset;
  Serializer::read_tcapint(in, offset);
  tcapint sz;
  Serializer::read_tcapint(in, sz);
  if (sz > WEEDCU_MAX_RANK) throw std::domain_error("Parameter::load: rank exceeds the device layer's maximum");
  std::vector<tcapint> shape(sz), stride(sz);
  for (size_t i = 0U; i < shape.size(); ++i) {
    Serializer::read_tcapint(in, shape[i]);
    Serializer::read_tcapint(in, stride[i]);
  }
  StoragePtr storage = Storage::load(in);
  ParameterPtr p = std::make_shared<Parameter>(std::vector<real1>{ZERO_R1}, std::vector<tcapint>{1U}, storage->device, storage->get_device_id());
  p->shape = shape;
  p->stride = stride;
  p->storage = storage;
  if (p->get_size() > storage->size) throw std::domain_error("Parameter::load: the view does not fit its storage");
  return p;
}

// ------------------------------------------------------------------------------------- Module::save
void Module::write_module_type(std::ostream &out, const ModuleType &x) { put(out, x); }
void Module::read_module_type(std::istream &in, ModuleType &x) { get(in, x); }
void Module::save(std::ostream &os) const { write_module_type(os, mtype); }

#define WEED_AXIS_SAVE(Name)                                                                       \
  void Name::save(std::ostream &os) const {                                                        \
    Module::save(os);                                                                              \
    Serializer::write_symint(os, axis);                                                            \
  }
WEED_AXIS_SAVE(Softmax)
WEED_AXIS_SAVE(LogSoftmax)
WEED_AXIS_SAVE(Mean)
WEED_AXIS_SAVE(Max)
WEED_AXIS_SAVE(Min)
WEED_AXIS_SAVE(Variance)
WEED_AXIS_SAVE(Stddev)
WEED_AXIS_SAVE(MeanCenter)
WEED_AXIS_SAVE(Flatten)
#undef WEED_AXIS_SAVE

void Reshape::save(std::ostream &os) const { // include/modules/reshape.hpp
  Module::save(os);
  Serializer::write_tcapint(os, (tcapint)shape.size());
  for (size_t i = 0U; i < shape.size(); ++i) Serializer::write_symint(os, shape[i]);
}
void Dropout::save(std::ostream &os) const { // src/modules/dropout.cpp
  Module::save(os);
  Serializer::write_real(os, p);
  Serializer::write_bool(os, training);
}
void Sequential::save(std::ostream &os) const { // src/modules/sequential.cpp:16-22
  Module::save(os);
  Serializer::write_tcapint(os, (tcapint)(layers.size()));
  for (size_t i = 0U; i < layers.size(); ++i) layers[i]->save(os);
}
void Linear::save(std::ostream &os) const { // src/modules/linear.cpp:109-118
  Module::save(os);
  Serializer::write_tcapint(os, in_features);
  Serializer::write_tcapint(os, out_features);
  weight->save(os);
  Serializer::write_bool(os, !!bias);
  if (bias) bias->save(os);
}
void LayerNorm::save(std::ostream &os) const { // src/modules/layernorm.cpp:44-50
  Module::save(os);
  Serializer::write_tcapint(os, features);
  Serializer::write_real(os, eps);
  gamma->save(os);
  beta->save(os);
}
void Embedding::save(std::ostream &os) const { // src/modules/embedding.cpp:66-71
  Module::save(os);
  Serializer::write_tcapint(os, num_embeddings);
  Serializer::write_tcapint(os, embedding_dim);
  weight->save(os);
}
void LearnedPositionalEncoding::save(std::ostream &os) const { // src/modules/learned_positional_encoding.cpp:63-68
  Module::save(os);
  Serializer::write_tcapint(os, max_len);
  Serializer::write_tcapint(os, d_model);
  pos_encoding->save(os);
}
void PositionalEncoding::save(std::ostream &os) const { // src/modules/positional_encoding.cpp (the table is rebuilt on load)
  Module::save(os);
  Serializer::write_tcapint(os, max_seq_len);
  Serializer::write_tcapint(os, d_model);
  Serializer::write_real1_f(os, pos_val);
}
void RoPE::save(std::ostream &os) const { // src/modules/rope.cpp (cos / sin tables are rebuilt on load)
  Module::save(os);
  Serializer::write_tcapint(os, head_dim);
  Serializer::write_tcapint(os, max_seq_len);
  Serializer::write_real1_f(os, base);
}
void RMSNorm::save(std::ostream &os) const { // include/modules/rms_norm.hpp:42-47
  Module::save(os);
  Serializer::write_symint(os, axis);
  Serializer::write_tcapint(os, hidden_size);
  weight->save(os);
}
void SwiGLU::save(std::ostream &os) const { // src/modules/swiglu.cpp
  Module::save(os);
  Serializer::write_tcapint(os, hidden_size);
  Serializer::write_tcapint(os, intermediate_size);
  gate_proj->save(os);
  up_proj->save(os);
  down_proj->save(os);
}
void GRU::save(std::ostream &os) const { // src/modules/gru.cpp:44-50
  Module::save(os);
  Serializer::write_tcapint(os, input_dim);
  Serializer::write_tcapint(os, hidden_dim);
  W_x->save(os);
  W_h->save(os);
}
void LSTM::save(std::ostream &os) const { // src/modules/lstm.cpp:56-62
  Module::save(os);
  Serializer::write_tcapint(os, input_dim);
  Serializer::write_tcapint(os, hidden_dim);
  W_x->save(os);
  W_h->save(os);
}
void MultiHeadAttention::save(std::ostream &os) const { // src/modules/multihead_attention.cpp:358-375
  Module::save(os);
  Serializer::write_real1_f(os, mask_val);
  Serializer::write_symint(os, d_model);
  Serializer::write_symint(os, num_heads);
  Serializer::write_symint(os, num_kv_heads);
  Serializer::write_symint(os, head_dim);
  Serializer::write_bool(os, use_kv_cache);
  Serializer::write_symint(os, (symint)kv_quant_bits);
  W_q->save(os);
  W_k->save(os);
  W_v->save(os);
  W_o->save(os);
  Serializer::write_bool(os, (bool)rope);
  if (rope) rope->save(os);
}
void TransformerEncoderLayer::save(std::ostream &os) const { // src/modules/transformer_encoder_layer.cpp:126-137
  Module::save(os);
  Serializer::write_tcapint(os, d_model);
  Serializer::write_tcapint(os, d_ff);
  Serializer::write_tcapint(os, num_heads);
  self_attn->save(os);
  ff1->save(os);
  ff2->save(os);
  norm1->save(os);
  norm2->save(os);
  activation->save(os);
}
void QwenDecoderLayer::save(std::ostream &os) const { // src/modules/qwen_decoder_layer.cpp
  Module::save(os);
  Serializer::write_tcapint(os, d_model);
  Serializer::write_tcapint(os, num_heads);
  Serializer::write_tcapint(os, num_kv_heads);
  self_attn->save(os);
  mlp->save(os);
  input_layernorm->save(os);
  post_attention_layernorm->save(os);
}

// ------------------------------------------------------------------------------------- Module::load
namespace {
template <typename T> std::shared_ptr<T> load_as(std::istream &is, const char *what) {
  std::shared_ptr<T> m = std::dynamic_pointer_cast<T>(Module::load(is));
  if (!m) throw std::domain_error(std::string("Module::load: expected a ") + what + " sub-module");
  return m;
}
symint read_axis(std::istream &is) {
  symint axis;
  Serializer::read_symint(is, axis);
  return axis;
}
} // namespace

ModulePtr Module::load(std::istream &is) { // src/modules/module.cpp:58-375, one case per module type
  ModuleType mtype;
  read_module_type(is, mtype);
  switch (mtype) {
  case SEQUENTIAL_T: {
    tcapint sz;
    Serializer::read_tcapint(is, sz);
    std::vector<ModulePtr> mv;
    mv.reserve(sz);
    for (tcapint i = 0U; i < sz; ++i) mv.push_back(load(is));
    return std::make_shared<Sequential>(mv);
  }
  case LINEAR_T: {
    LinearPtr l = std::make_shared<Linear>();
    Serializer::read_tcapint(is, l->in_features);
    Serializer::read_tcapint(is, l->out_features);
    l->weight = Parameter::load(is);
    bool is_bias;
    Serializer::read_bool(is, is_bias);
    if (is_bias) l->bias = Parameter::load(is);
    return l;
  }
  case GELU_T: return std::make_shared<GeLU>();
  case RELU_T: return std::make_shared<ReLU>();
  case SIGMOID_T: return std::make_shared<Sigmoid>();
  case TANH_T: return std::make_shared<Tanh>();
  case SWIGLU_T: {
    SwiGLUPtr s = std::make_shared<SwiGLU>();
    Serializer::read_tcapint(is, s->hidden_size);
    Serializer::read_tcapint(is, s->intermediate_size);
    s->gate_proj = load_as<Linear>(is, "Linear");
    s->up_proj = load_as<Linear>(is, "Linear");
    s->down_proj = load_as<Linear>(is, "Linear");
    s->_register_params();
    return s;
  }
  case DROPOUT_T: {
    DropoutPtr d = std::make_shared<Dropout>();
    Serializer::read_real(is, d->p);
    Serializer::read_bool(is, d->training);
    return d;
  }
  case EMBEDDING_T: {
    EmbeddingPtr e = std::make_shared<Embedding>();
    Serializer::read_tcapint(is, e->num_embeddings);
    Serializer::read_tcapint(is, e->embedding_dim);
    e->weight = Parameter::load(is);
    return e;
  }
  case LAYERNORM_T: {
    LayerNormPtr l = std::make_shared<LayerNorm>();
    Serializer::read_tcapint(is, l->features);
    Serializer::read_real(is, l->eps);
    l->gamma = Parameter::load(is);
    l->beta = Parameter::load(is);
    return l;
  }
  case GRU_T: {
    GRUPtr g = std::make_shared<GRU>();
    Serializer::read_tcapint(is, g->input_dim);
    Serializer::read_tcapint(is, g->hidden_dim);
    g->W_x = load_as<Linear>(is, "Linear");
    g->W_h = load_as<Linear>(is, "Linear");
    g->state = Tensor::zeros({g->hidden_dim});
    return g;
  }
  case LSTM_T: {
    LSTMPtr l = std::make_shared<LSTM>();
    Serializer::read_tcapint(is, l->input_dim);
    Serializer::read_tcapint(is, l->hidden_dim);
    l->W_x = load_as<Linear>(is, "Linear");
    l->W_h = load_as<Linear>(is, "Linear");
    l->state = LSTMState{Tensor::zeros(std::vector<tcapint>{l->hidden_dim}), Tensor::zeros(std::vector<tcapint>{l->hidden_dim})};
    return l;
  }
  case MIGRATE_CPU_T: return std::make_shared<MigrateCpu>();
  case MIGRATE_GPU_T: return std::make_shared<MigrateGpu>();
  case MEAN_CENTER_T: return std::make_shared<MeanCenter>(read_axis(is));
  case SOFTMAX_T: return std::make_shared<Softmax>(read_axis(is));
  case LOGSOFTMAX_T: return std::make_shared<LogSoftmax>(read_axis(is));
  case FLATTEN_T: return std::make_shared<Flatten>(read_axis(is));
  case MEAN_T: return std::make_shared<Mean>(read_axis(is));
  case MAX_T: return std::make_shared<Max>(read_axis(is));
  case MIN_T: return std::make_shared<Min>(read_axis(is));
  case VARIANCE_T: return std::make_shared<Variance>(read_axis(is));
  case STDDEV_T: return std::make_shared<Stddev>(read_axis(is));
  case RMS_NORM_T: {
    RMSNormPtr r = std::make_shared<RMSNorm>();
    Serializer::read_symint(is, r->axis);
    Serializer::read_tcapint(is, r->hidden_size);
    r->weight = Parameter::load(is);
    return r;
  }
  case RESHAPE_T: {
    tcapint sz;
    Serializer::read_tcapint(is, sz);
    std::vector<symint> shape(sz);
    for (tcapint i = 0U; i < sz; ++i) Serializer::read_symint(is, shape[i]);
    return std::make_shared<Reshape>(shape);
  }
  case ROPE_T: {
    RoPEPtr r = std::make_shared<RoPE>();
    Serializer::read_tcapint(is, r->head_dim);
    Serializer::read_tcapint(is, r->max_seq_len);
    Serializer::read_real1_f(is, r->base);
    r->_build_tables();
    return r;
  }
  case MULTIHEAD_ATTENTION_T: {
    MultiHeadAttentionPtr m = std::make_shared<MultiHeadAttention>();
    Serializer::read_real1_f(is, m->mask_val);
    Serializer::read_symint(is, m->d_model);
    Serializer::read_symint(is, m->num_heads);
    Serializer::read_symint(is, m->num_kv_heads);
    Serializer::read_symint(is, m->head_dim);
    Serializer::read_bool(is, m->use_kv_cache);
    symint kv_quant_bits_tmp = 0;
    Serializer::read_symint(is, kv_quant_bits_tmp);
    m->kv_quant_bits = (int)kv_quant_bits_tmp;
    m->W_q = load_as<Linear>(is, "Linear");
    m->W_k = load_as<Linear>(is, "Linear");
    m->W_v = load_as<Linear>(is, "Linear");
    m->W_o = load_as<Linear>(is, "Linear");
    bool has_rope;
    Serializer::read_bool(is, has_rope);
    if (has_rope) m->rope = load_as<RoPE>(is, "RoPE");
    m->_register_params();
    return m;
  }
  case TRANSFORMER_ENCODER_LAYER_T: {
    TransformerEncoderLayerPtr t = std::make_shared<TransformerEncoderLayer>();
    Serializer::read_tcapint(is, t->d_model);
    Serializer::read_tcapint(is, t->d_ff);
    Serializer::read_tcapint(is, t->num_heads);
    t->self_attn = load_as<MultiHeadAttention>(is, "MultiHeadAttention");
    t->ff1 = load_as<Linear>(is, "Linear");
    t->ff2 = load_as<Linear>(is, "Linear");
    t->norm1 = load_as<LayerNorm>(is, "LayerNorm");
    t->norm2 = load_as<LayerNorm>(is, "LayerNorm");
    t->activation = load(is);
    t->_register_params();
    return t;
  }
  case POSITIONAL_ENCODING_T: {
    tcapint max_seq_len, d_model;
    Serializer::read_tcapint(is, max_seq_len);
    Serializer::read_tcapint(is, d_model);
    real1_f pos_val;
    Serializer::read_real1_f(is, pos_val);
    return std::make_shared<PositionalEncoding>(max_seq_len, d_model, pos_val);
  }
  case LEARNED_POSITIONAL_ENCODING_T: {
    LearnedPositionalEncodingPtr l = std::make_shared<LearnedPositionalEncoding>();
    Serializer::read_tcapint(is, l->max_len);
    Serializer::read_tcapint(is, l->d_model);
    l->pos_encoding = Parameter::load(is);
    return l;
  }
  case QWEN_DECODER_LAYER_T: {
    QwenDecoderLayerPtr q = std::make_shared<QwenDecoderLayer>();
    Serializer::read_tcapint(is, q->d_model);
    Serializer::read_tcapint(is, q->num_heads);
    Serializer::read_tcapint(is, q->num_kv_heads);
    q->self_attn = load_as<MultiHeadAttention>(is, "MultiHeadAttention");
    q->mlp = load_as<SwiGLU>(is, "SwiGLU");
    q->input_layernorm = load_as<RMSNorm>(is, "RMSNorm");
    q->post_attention_layernorm = load_as<RMSNorm>(is, "RMSNorm");
    q->_register_params();
    return q;
  }
  case QRACK_NEURON_T:
  case QRACK_NEURON_LAYER_T:
    throw std::domain_error("Module::load: Qrack modules are outside the CUDA backend's scope (SURVEY §8)");
  case NONE_MODULE_TYPE:
  default:
    throw std::domain_error("Can't recognize ModuleType " + std::to_string(mtype) + " in Module::load!");
  }
}
} // namespace Weed
