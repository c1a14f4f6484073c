// weed_harness.cpp — ONE client source, compiled twice:
//   (1) against the UNMODIFIED reference (vm6502q/weed CPU build, oracle/_ref/libweed_ref.a)
//       -> oracle/_ref/libweed_ref_harness.so          (test infrastructure: the live oracle)
//   (2) against this repo's host library (weed_b200/host, -DWEED_B200)
//       -> weed_b200/libweed_b200_harness.so           (the product, driven by tests and bench.py)
// It only uses Weed's public API (tensors/tensor.hpp, modules/*.hpp, autograd/*.hpp) — that the
// same file builds against both trees is the source-level proof of the drop-in claim — and exposes
// a small handle-based C-ABI in the style of the reference's own src/shared_api.cpp:79-449
// (integer handles, int error codes, last-error string).
#include "common/serializer.hpp" // (modules/max.hpp and friends use Serializer without including it)
#include "autograd/adam.hpp"
#include "autograd/bci_with_logits_loss.hpp"
#include "autograd/cross_entropy_loss.hpp"
#include "autograd/mse_loss.hpp"
#include "autograd/sgd.hpp"
#include "autograd/zero_grad.hpp"
#include "modules/embedding.hpp"
#include "modules/gelu.hpp"
#include "modules/layernorm.hpp"
#include "modules/learned_positional_encoding.hpp"
#include "modules/linear.hpp"
#include "modules/multihead_attention.hpp"
#include "modules/relu.hpp"
#include "modules/sequential.hpp"
#include "modules/sigmoid.hpp"
#include "modules/tanh.hpp"
#include "modules/dropout.hpp"
#include "modules/gru.hpp"
#include "modules/lstm.hpp"
#include "modules/max.hpp"
#include "modules/mean.hpp"
#include "modules/min.hpp"
#include "modules/positional_encoding.hpp"
#include "modules/qwen_decoder_layer.hpp"
#include "modules/rms_norm.hpp"
#include "modules/rope.hpp"
#include "modules/softmax.hpp"
#include "modules/swiglu.hpp"
#include "modules/transformer_encoder_layer.hpp"
#include "tensors/real_tensor.hpp"
#include "tensors/symbol_tensor.hpp"

#include <chrono>
#include <fstream>
#include <cstring>
#include <iterator>
#include <map>
#include <string>

using namespace Weed;

namespace {
DeviceTag g_dtag = DeviceTag::CPU;
std::string g_error;
int64_t g_next = 1;
std::map<int64_t, TensorPtr> g_tensors;
std::map<int64_t, SymbolTensorPtr> g_symbols;
std::map<int64_t, ModulePtr> g_modules;
std::map<int64_t, std::shared_ptr<Adam>> g_adams;

int64_t put(const TensorPtr &t) {
  g_tensors[g_next] = t;
  return g_next++;
}
TensorPtr T(int64_t h) {
  auto it = g_tensors.find(h);
  if (it == g_tensors.end()) throw std::invalid_argument("bad tensor handle");
  return it->second;
}
ModulePtr M(int64_t h) {
  auto it = g_modules.find(h);
  if (it == g_modules.end()) throw std::invalid_argument("bad module handle");
  return it->second;
}
std::vector<tcapint> shape_vec(int rank, const uint32_t *shape) { return std::vector<tcapint>(shape, shape + rank); }

#define WH_TRY(...)                                                                                \
  try {                                                                                            \
    __VA_ARGS__;                                                                                   \
  } catch (const std::exception &e) {                                                              \
    g_error = e.what();                                                                            \
    return -1;                                                                                     \
  }
} // namespace

extern "C" {

const char *wh_last_error() { return g_error.c_str(); }
const char *wh_backend() {
#ifdef WEED_B200
  return "weed_b200";
#else
  return "reference";
#endif
}

// device_tag: 2 = CPU, 3 = GPU (Weed::DeviceTag values)
int wh_init(int device_tag) {
  WH_TRY({
    g_dtag = (device_tag == 3) ? DeviceTag::GPU : DeviceTag::CPU;
#ifdef WEED_B200
    if (g_dtag == DeviceTag::GPU) WEED_GPU_SINGLETON.InitOCL();
#endif
    return 0;
  })
}
int wh_reset() {
  g_tensors.clear();
  g_symbols.clear();
  g_modules.clear();
  g_adams.clear();
  return 0;
}
// handle bookkeeping for callers that build several models in one process (bench.py's parity legs): everything created
// since `mark` is dropped
int64_t wh_mark() { return g_next; }
int wh_release_since(int64_t mark) {
  auto drop = [mark](auto &m) {
    for (auto it = m.begin(); it != m.end();) it = (it->first >= mark) ? m.erase(it) : std::next(it);
  };
  drop(g_tensors);
  drop(g_symbols);
  drop(g_adams);
  drop(g_modules);
  return 0;
}
int wh_free(int64_t h) {
  g_tensors.erase(h);
  g_symbols.erase(h);
  return 0;
}
// backend switches of this repo's host library; accepted and ignored by the reference build
int wh_config(const char *name, double value) {
#ifdef WEED_B200
  WH_TRY({
    const std::string n(name);
    BackendConfig &c = backend_config();
    if (n == "fused") c.fused = value != 0;
    else if (n == "ref_index_quirks") c.ref_index_quirks = value != 0;
    else if (n == "matmul_precision") c.matmul_precision = (int)value;
    else if (n == "grad_scale") c.grad_scale = (real1)value;
    else if (n == "layernorm_exact_grad") c.layernorm_exact_grad = value != 0;
    else if (n == "bf16_act_grad") c.bf16_act_grad = value != 0;
    else if (n == "operand_cache") c.operand_cache = value != 0;
    else if (n == "lazy_zero") c.lazy_zero = value != 0;
    else if (n == "defer_grads") c.defer_grads = value != 0;
    else if (n == "cow_grads") c.cow_grads = value != 0;
    else if (n == "pdl") weedcu_set_pdl(value != 0 ? 1 : 0);
    else if (n == "epilogue_stats") c.epilogue_stats = value != 0;
    else if (n == "lm_head_min_cols") c.lm_head_min_cols = (tcapint)value;
    else throw std::invalid_argument("unknown config key");
    return 0;
  })
#else
  (void)name;
  (void)value;
  return 0;
#endif
}
int wh_sync() {
#ifdef WEED_B200
  WH_TRY({
    if (g_dtag == DeviceTag::GPU) WEED_GPU_SINGLETON.GetWeedDevice(-1)->clFinish();
    return 0;
  })
#else
  return 0;
#endif
}
// the cudaStream_t all device work of this library is issued on (0 for the reference build)
void *wh_stream() {
#ifdef WEED_B200
  try {
    return WEED_GPU_SINGLETON.GetWeedDevice(-1)->stream;
  } catch (...) {
    return nullptr;
  }
#else
  return nullptr;
#endif
}

// ------------------------------------------------------------------------------- tensors
int64_t wh_tensor(const float *data, uint32_t n, int rank, const uint32_t *shape, int requires_grad) {
  WH_TRY({
    std::vector<real1> v(data, data + n);
    return put(std::make_shared<Tensor>(v, shape_vec(rank, shape), requires_grad != 0, g_dtag));
  })
}
int64_t wh_scalar(float value, int requires_grad) {
  WH_TRY({ return put(std::make_shared<Tensor>((real1)value, requires_grad != 0, g_dtag)); })
}
int64_t wh_symbol(const int32_t *data, uint32_t n, int rank, const uint32_t *shape) {
  WH_TRY({
    std::vector<symint> v(data, data + n);
    g_symbols[g_next] = std::make_shared<SymbolTensor>(v, shape_vec(rank, shape), false, g_dtag);
    return g_next++;
  })
}
// Overwrite an existing SymbolTensor from a host buffer (bench.py passes pinned memory: this is
// the per-step host->device input copy of the end-to-end measurement). Asynchronous on the device
// stream for the CUDA build; a plain copy for the reference build.
int wh_symbol_upload(int64_t h, const int32_t *src, uint32_t n) {
  WH_TRY({
    SymbolTensorPtr s = g_symbols.at(h);
    if (s->storage->size != n) throw std::invalid_argument("wh_symbol_upload: element count mismatch");
#ifdef WEED_B200
    if (s->storage->device == DeviceTag::GPU) {
      GpuIntStorage *gs = static_cast<GpuIntStorage *>(s->storage.get());
      throw_on_error(weedcu_memcpy_h2d(gs->device_ptr(), src, sizeof(int32_t) * (size_t)n, gs->dev->stream), "wh_symbol_upload");
      return 0;
    }
#endif
    IntStorage *is = static_cast<IntStorage *>(s->storage.get());
    for (uint32_t i = 0; i < n; ++i) is->write(i, src[i]);
    return 0;
  })
}
// an explicit (offset, shape, stride) view on the storage of an existing tensor
int64_t wh_view(int64_t h, uint32_t offset, int rank, const uint32_t *shape, const uint32_t *stride) {
  WH_TRY({
    TensorPtr v = std::make_shared<Tensor>(*T(h));
    v->offset = offset;
    v->shape = shape_vec(rank, shape);
    v->stride = shape_vec(rank, stride);
    v->grad = nullptr;
    v->grad_node = nullptr;
    return put(v);
  })
}
int wh_info(int64_t h, int *rank, uint32_t *shape, uint32_t *stride, uint32_t *offset, uint32_t *storage_size, int *requires_grad) {
  WH_TRY({
    TensorPtr t = T(h);
    *rank = (int)t->shape.size();
    for (size_t i = 0; i < t->shape.size() && i < 8; ++i) {
      shape[i] = t->shape[i];
      stride[i] = t->stride[i];
    }
    *offset = t->offset;
    *storage_size = t->storage->size;
    *requires_grad = t->requires_grad ? 1 : 0;
    return 0;
  })
}
// logical values: element i of the flat column-major index space, through the view
int wh_read(int64_t h, float *out, uint32_t capacity, uint32_t *count) {
  WH_TRY({
    TensorPtr t = T(h);
    const tcapint n = t->get_broadcast_size();
    *count = n;
    if (n > capacity) throw std::invalid_argument("wh_read: buffer too small");
#ifdef WEED_B200
    const std::vector<real1> v = to_host_logical(*t);
    std::memcpy(out, v.data(), sizeof(float) * n);
#else
    TensorPtr c = t->cast(DeviceTag::CPU);
    RealTensor rt(*c);
    for (tcapint i = 0; i < n; ++i) out[i] = rt[i];
#endif
    return 0;
  })
}
// raw storage contents in storage order
int wh_read_storage(int64_t h, float *out, uint32_t capacity, uint32_t *count) {
  WH_TRY({
    TensorPtr t = T(h);
    StoragePtr s = t->storage->cpu();
    const tcapint n = s->size;
    *count = n;
    if (n > capacity) throw std::invalid_argument("wh_read_storage: buffer too small");
    RealStorage *rs = static_cast<RealStorage *>(s.get());
    for (tcapint i = 0; i < n; ++i) out[i] = (*rs)[i];
    return 0;
  })
}
int64_t wh_grad(int64_t h) {
  WH_TRY({
    TensorPtr t = T(h);
    if (!t->grad) return (int64_t)0;
    return put(t->grad);
  })
}
int wh_backward(int64_t h) {
  WH_TRY({
    Tensor::backward(T(h));
    return 0;
  })
}

// Tensor:: front-ends by name. ins: tensor handles; f: float params; i: int params.
int64_t wh_op(const char *name, const int64_t *ins, int n_in, const float *f, int n_f, const int32_t *iv, int n_i) {
  WH_TRY({
    const std::string op(name);
    auto in = [&](int k) {
      if (k >= n_in) throw std::invalid_argument("wh_op: missing tensor argument");
      return T(ins[k]);
    };
    auto fp = [&](int k) {
      if (k >= n_f) throw std::invalid_argument("wh_op: missing float argument");
      return (real1)f[k];
    };
    auto ip = [&](int k) {
      if (k >= n_i) throw std::invalid_argument("wh_op: missing int argument");
      return (symint)iv[k];
    };
    TensorPtr r;
    if (op == "add") r = Tensor::add(in(0), in(1));
    else if (op == "sub") r = Tensor::sub(in(0), in(1));
    else if (op == "mul") r = Tensor::mul(in(0), in(1));
    else if (op == "div") r = Tensor::div(in(0), in(1));
    else if (op == "matmul") r = Tensor::matmul(in(0), in(1));
    else if (op == "add_scalar") r = in(0) + fp(0);
    else if (op == "mul_scalar") r = in(0) * fp(0);
    else if (op == "div_scalar") r = in(0) / fp(0);
    else if (op == "rsub_scalar") r = fp(0) - in(0);
    else if (op == "relu") r = Tensor::relu(in(0));
    else if (op == "sigmoid") r = Tensor::sigmoid(in(0));
    else if (op == "tanh") r = Tensor::tanh(in(0));
    else if (op == "gelu") r = Tensor::gelu(in(0));
    else if (op == "abs") r = Tensor::abs(in(0));
    else if (op == "pow") r = Tensor::pow(in(0), fp(0));
    else if (op == "exp") r = (n_f > 0) ? Tensor::exp(in(0), fp(0)) : Tensor::exp(in(0));
    else if (op == "log") r = (n_f > 0) ? Tensor::log(in(0), fp(0)) : Tensor::log(in(0));
    else if (op == "sum") r = Tensor::sum(in(0));
    else if (op == "mean") r = Tensor::mean(in(0));
    else if (op == "sum_axis") r = Tensor::sum(in(0), ip(0));
    else if (op == "mean_axis") r = Tensor::mean(in(0), ip(0));
    else if (op == "max") r = Tensor::max(in(0));
    else if (op == "min") r = Tensor::min(in(0));
    else if (op == "max_axis") r = Tensor::max(in(0), ip(0));
    else if (op == "min_axis") r = Tensor::min(in(0), ip(0));
    else if (op == "clamp") r = Tensor::clamp(in(0), fp(0), fp(1));
    else if (op == "sin") r = Tensor::sin(in(0));
    else if (op == "cos") r = Tensor::cos(in(0));
    else if (op == "softmax") r = Tensor::softmax(in(0), ip(0));
    else if (op == "logsoftmax") r = Tensor::logsoftmax(in(0), ip(0));
    else if (op == "transpose") r = (n_i >= 2) ? Tensor::transpose(in(0), ip(0), ip(1)) : Tensor::transpose(in(0));
    else if (op == "contiguous") r = Tensor::contiguous(in(0));
    else if (op == "slice") r = Tensor::slice(in(0), (int64_t)ip(0), (tcapint)ip(1), (tcapint)ip(2));
    else if (op == "row_slice") r = Tensor::slice(in(0), (int64_t)ip(0));
    else if (op == "reshape") {
      std::vector<symint> s(iv, iv + n_i);
      r = Tensor::reshape(in(0), s);
    } else if (op == "mse_loss") r = mse_loss(in(0), in(1));
    else if (op == "bci_with_logits_loss") r = bci_with_logits_loss(in(0), in(1));
    else throw std::invalid_argument("wh_op: unknown op '" + op + "'");
    return put(r);
  })
}
int64_t wh_cross_entropy(int64_t logits, int64_t targets) {
  WH_TRY({
    auto it = g_symbols.find(targets);
    if (it == g_symbols.end()) throw std::invalid_argument("bad symbol handle");
    return put(cross_entropy_loss(T(logits), it->second));
  })
}

// ------------------------------------------------------------------------------- modules
int64_t wh_module(const char *kind, const int64_t *args, int n) {
  WH_TRY({
    const std::string k(kind);
    auto a = [&](int i) {
      if (i >= n) throw std::invalid_argument("wh_module: missing argument");
      return args[i];
    };
    ModulePtr m;
    if (k == "linear") m = std::make_shared<Linear>((tcapint)a(0), (tcapint)a(1), a(2) != 0, true, DType::REAL, g_dtag);
    else if (k == "layernorm") m = std::make_shared<LayerNorm>((tcapint)a(0), g_dtag);
    else if (k == "embedding") m = std::make_shared<Embedding>((tcapint)a(0), (tcapint)a(1), DType::REAL, g_dtag);
    else if (k == "posenc") m = std::make_shared<LearnedPositionalEncoding>((tcapint)a(0), (tcapint)a(1), g_dtag);
    else if (k == "mha") {
      // use_kv_cache = false, kv_quant_bits = 0: the deterministic configuration (SURVEY §7 hard part 6)
      m = std::make_shared<MultiHeadAttention>((tcapint)a(0), (tcapint)a(1), 0U, 0U, g_dtag, nullptr, ZERO_R1, -1, false, 0);
    } else if (k == "encoder") {
      TransformerEncoderLayerPtr e = std::make_shared<TransformerEncoderLayer>((tcapint)a(0), (tcapint)a(1), (tcapint)a(2), g_dtag);
      e->self_attn->use_kv_cache = false; // public fields; constructor defaults are kv-cache + 4-bit quant
      e->self_attn->kv_quant_bits = 0;
      m = e;
    } else if (k == "rmsnorm") m = std::make_shared<RMSNorm>((tcapint)a(0));
    else if (k == "swiglu") {
      SwiGLUPtr sw = std::make_shared<SwiGLU>((tcapint)a(0), (tcapint)a(1));
      sw->_register_params(); // the constructor leaves param_vector empty (include/modules/swiglu.hpp:35-44)
      m = sw;
    } else if (k == "rope") m = std::make_shared<RoPE>((tcapint)a(0), (tcapint)a(1));
    else if (k == "qwen") {
      QwenDecoderLayerPtr q = std::make_shared<QwenDecoderLayer>((tcapint)a(0), (tcapint)a(1), (tcapint)a(2), (tcapint)a(3), (tcapint)a(4));
      q->self_attn->use_kv_cache = false; // constructor defaults: kv cache + 4-bit quantisation (host code with random_device)
      q->self_attn->kv_quant_bits = 0;
      q->mlp->_register_params();
      q->_register_params();
      m = q;
    } else if (k == "gru") m = std::make_shared<GRU>((tcapint)a(0), (tcapint)a(1), g_dtag);
    else if (k == "lstm") m = std::make_shared<LSTM>((tcapint)a(0), (tcapint)a(1), g_dtag);
    else if (k == "posenc_fixed") m = std::make_shared<PositionalEncoding>((tcapint)a(0), (tcapint)a(1), 8192.0, g_dtag);
    else if (k == "dropout") m = std::make_shared<Dropout>((real1)a(0) / (real1)1000);
    else if (k == "softmax") m = std::make_shared<Softmax>((symint)a(0));
    else if (k == "mean") m = std::make_shared<Mean>((symint)a(0));
    else if (k == "max") m = std::make_shared<Max>((symint)a(0));
    else if (k == "min") m = std::make_shared<Min>((symint)a(0));
    else if (k == "gelu") m = std::make_shared<GeLU>();
    else if (k == "relu") m = std::make_shared<ReLU>();
    else if (k == "tanh") m = std::make_shared<Tanh>();
    else if (k == "sigmoid") m = std::make_shared<Sigmoid>();
    else if (k == "sequential") {
      std::vector<ModulePtr> layers;
      for (int i = 0; i < n; ++i) layers.push_back(M(args[i]));
      m = std::make_shared<Sequential>(layers);
    } else throw std::invalid_argument("wh_module: unknown kind '" + k + "'");
    g_modules[g_next] = m;
    return g_next++;
  })
}
// checkpoint round trip through Module::save / Module::load (src/modules/module.cpp:53-375)
int wh_module_save(int64_t mh, const char *path) {
  WH_TRY({
    std::ofstream o(path, std::ios::binary);
    if (!o) throw std::invalid_argument("wh_module_save: cannot open the file");
    M(mh)->save(o);
    o.close();
    return 0;
  })
}
int64_t wh_module_load(const char *path) {
  WH_TRY({
    std::ifstream i(path, std::ios::binary);
    if (!i) throw std::invalid_argument("wh_module_load: cannot open the file");
    ModulePtr m = Module::load(i);
    i.close();
    g_modules[g_next] = m;
    return g_next++;
  })
}
int wh_module_set(int64_t mh, const char *field, int64_t value) {
  WH_TRY({
    const std::string f(field);
    ModulePtr m = M(mh);
    MultiHeadAttention *att = dynamic_cast<MultiHeadAttention *>(m.get());
    if (TransformerEncoderLayer *e = dynamic_cast<TransformerEncoderLayer *>(m.get())) att = e->self_attn.get();
    if (f == "train") value ? m->train() : m->eval();
    else if (f == "reset_cache") m->reset_cache();
    else if (f == "max_kv_seq_len") m->set_max_kv_seq_len((tcapint)value);
    else if (att && f == "use_kv_cache") att->use_kv_cache = value != 0;
    else if (att && f == "kv_quant_bits") att->kv_quant_bits = (int)value;
    else throw std::invalid_argument("wh_module_set: unknown field");
    return 0;
  })
}
int wh_param_count(int64_t mh) {
  WH_TRY({ return (int)M(mh)->parameters().size(); })
}
// storage element count of parameter i (what wh_param_set expects)
int64_t wh_param_size(int64_t mh, int i) {
  WH_TRY({ return (int64_t)M(mh)->parameters().at((size_t)i)->storage->size; })
}
int64_t wh_param(int64_t mh, int i) {
  WH_TRY({ return put(M(mh)->parameters().at((size_t)i)); })
}
// inject weights: replace the parameter's storage contents (storage order), keeping its view
int wh_param_set(int64_t mh, int i, const float *data, uint32_t n) {
  WH_TRY({
    ParameterPtr p = M(mh)->parameters().at((size_t)i);
    if (p->storage->size != n) throw std::invalid_argument("wh_param_set: element count mismatch");
    std::vector<real1> v(data, data + n);
    Tensor tmp(v, std::vector<tcapint>{n}, false, g_dtag);
    p->storage = tmp.storage;
    return 0;
  })
}
int64_t wh_forward(int64_t mh, int64_t x) {
  WH_TRY({ return put(M(mh)->forward(T(x))); })
}
int64_t wh_forward_symbol(int64_t mh, int64_t s) {
  WH_TRY({
    auto it = g_symbols.find(s);
    if (it == g_symbols.end()) throw std::invalid_argument("bad symbol handle");
    return put(M(mh)->forward(it->second));
  })
}
// Greedy decode step: arg-max token of the last position of logits [B, T, V] as a symbol handle [B, 1].
// This repo's host library computes it on the device (Weed::argmax_last_token); the reference has no
// arg-max, so its build reads the logits back and scans them on the host (what a Python client does).
int64_t wh_argmax_last(int64_t logits) {
  WH_TRY({
    TensorPtr lg = T(logits);
#ifdef WEED_B200
    SymbolTensorPtr s = argmax_last_token(*lg);
#else
    const tcapint B = lg->shape[0U], Tn = lg->shape[1U], V = lg->shape[2U];
    TensorPtr c = lg->cast(DeviceTag::CPU);
    RealTensor rt(*c);
    std::vector<symint> best(B, 0);
    for (tcapint b = 0U; b < B; ++b) {
      real1 bv = rt[b + B * (Tn - 1U)];
      for (tcapint v = 1U; v < V; ++v) {
        const real1 x = rt[b + B * ((Tn - 1U) + Tn * v)]; // flat column-major index of (b, T-1, v)
        if (x > bv) {
          bv = x;
          best[b] = (symint)v;
        }
      }
    }
    SymbolTensorPtr s = std::make_shared<SymbolTensor>(best, std::vector<tcapint>{B, 1U}, false, g_dtag);
#endif
    g_symbols[g_next] = s;
    return g_next++;
  })
}
int wh_read_symbol(int64_t h, int32_t *out, uint32_t n) {
  WH_TRY({
    auto it = g_symbols.find(h);
    if (it == g_symbols.end()) throw std::invalid_argument("bad symbol handle");
    SymbolTensorPtr c = it->second->cast(DeviceTag::CPU);
    if (c->get_broadcast_size() < n) throw std::invalid_argument("wh_read_symbol: symbol is smaller than n");
    for (uint32_t i = 0; i < n; ++i) out[i] = (int32_t)(*static_cast<IntStorage *>(c->storage.get()))[c->get_storage_index(i)];
    return 0;
  })
}
// mutating helpers used by the example workloads (binary_addition_transformer.cpp:128-131)
int wh_squeeze(int64_t h, int axis) {
  WH_TRY({
    T(h)->squeeze(axis);
    return 0;
  })
}

// ------------------------------------------------------------------------------- optimisers
int64_t wh_adam(float lr, float b1, float b2, float eps, int64_t module) {
  WH_TRY({
    std::shared_ptr<Adam> o = std::make_shared<Adam>((real1)lr, (real1)b1, (real1)b2, (real1)eps);
    o->register_parameters(M(module)->parameters());
    g_adams[g_next] = o;
    return g_next++;
  })
}
// first (which = 0) or second (which = 1) moment of parameter i of `module` in optimiser `opt`, as a tensor handle
int64_t wh_adam_moment(int64_t opt, int64_t module, int i, int which) {
  WH_TRY({
    ParameterPtr p = M(module)->parameters().at((size_t)i);
    const AdamState &st = g_adams.at(opt)->state.at(p);
    return put(which ? st.v : st.m);
  })
}
int wh_adam_step(int64_t opt, int64_t module) {
  WH_TRY({
    adam_step(*g_adams.at(opt), M(module)->parameters());
    return 0;
  })
}
int wh_sgd_step(int64_t module, float lr) {
  WH_TRY({
    sgd_step(M(module)->parameters(), (real1)lr);
    return 0;
  })
}
int wh_zero_grad(int64_t module) {
  WH_TRY({
    zero_grad(M(module)->parameters());
    return 0;
  })
}

// ------------------------------------------------------------------------------- data parallel
// (this repo's host library only; the reference has no gradient exchange)
#ifdef WEED_B200
namespace {
void *g_comm = nullptr;
int g_world = 1;
bool g_dp_active = true;
std::unique_ptr<GradientBuckets> g_buckets;
} // namespace
#endif
int wh_dp_load(const char *libnccl_path) {
#ifdef WEED_B200
  return weedcu_nccl_load(libnccl_path);
#else
  (void)libnccl_path;
  return -1;
#endif
}
int wh_dp_unique_id(void *id128) {
#ifdef WEED_B200
  return weedcu_nccl_unique_id(id128);
#else
  (void)id128;
  return -1;
#endif
}
int wh_dp_init(const void *id128, int rank, int world) {
#ifdef WEED_B200
  WH_TRY({
    throw_on_error(weedcu_nccl_init(id128, rank, world, &g_comm), "wh_dp_init");
    g_world = world;
    backend_config().grad_scale = ONE_R1 / (real1)world;
    return 0;
  })
#else
  (void)id128;
  (void)rank;
  (void)world;
  return -1;
#endif
}
// switch the gradient exchange off / on again for this process (bench.py --check-dp: rank 0 repeats the run alone on
// the global batch); the 1/world gradient scale follows
int wh_dp_set_active(int active) {
#ifdef WEED_B200
  g_dp_active = active != 0;
  backend_config().grad_scale = (g_dp_active && g_comm && g_world > 1) ? ONE_R1 / (real1)g_world : ONE_R1;
  return 0;
#else
  (void)active;
  return -1;
#endif
}
int wh_dp_broadcast_params(int64_t model) {
#ifdef WEED_B200
  WH_TRY({
    if (g_comm) broadcast_parameters(M(model)->parameters(), g_comm, 0);
    return 0;
  })
#else
  (void)model;
  return 0;
#endif
}

// One full training step on token input: forward, cross-entropy, backward, (gradient all-reduce,)
// Adam, zero_grad — the loop body of the reference's train_step (src/shared_api.cpp:356-424) with
// Adam instead of SGD. Returns the loss tensor handle (read it with wh_read; reading syncs).
int64_t wh_train_step_tokens(int64_t model, int64_t opt, int64_t tokens, int64_t targets) {
  WH_TRY({
    // WH_TIMING=1 prints the host time of each phase (diagnostic for launch-bound steps)
    static const bool timing = getenv("WH_TIMING") != nullptr;
    auto now = [] { return std::chrono::steady_clock::now(); };
    auto us = [](std::chrono::steady_clock::time_point a, std::chrono::steady_clock::time_point b) {
      return std::chrono::duration<double, std::micro>(b - a).count();
    };
    const auto t0 = now();
    ModulePtr m = M(model);
#ifdef WEED_B200
    // WH_DP_EVENTS=1: device timeline of the step's phases on the compute stream (diagnostic; printed two steps late so that
    // reading the events never blocks the step being issued)
    static const bool ev_on = getenv("WH_DP_EVENTS") && atoi(getenv("WH_DP_EVENTS")) != 0;
    static void *evs[3][5] = {{nullptr}};
    static uint64_t ev_step = 0;
    void *cstream = nullptr;
    if (ev_on) {
      cstream = m->parameters().front()->stream();
      if (!evs[0][0])
        for (auto &row : evs)
          for (void *&e : row) throw_on_error(weedcu_event_create(&e), "events");
      if (ev_step >= 2) {
        void **old = evs[(ev_step - 2) % 3];
        float a = 0, b = 0, c = 0, d = 0;
        weedcu_event_sync(old[4]);
        weedcu_event_elapsed_ms(old[0], old[1], &a);
        weedcu_event_elapsed_ms(old[1], old[2], &b);
        weedcu_event_elapsed_ms(old[2], old[3], &c);
        weedcu_event_elapsed_ms(old[3], old[4], &d);
        static const int ev_rank = getenv("RANK") ? atoi(getenv("RANK")) : 0;
        if (ev_rank == 0) fprintf(stderr, "[wh events] forward+loss %.3f ms, backward %.3f, exchange tail %.3f, adam+zero %.3f\n", a, b, c, d);
      }
      throw_on_error(weedcu_event_record(evs[ev_step % 3][0], cstream), "events");
    }
#endif
    TensorPtr logits = m->forward(g_symbols.at(tokens));
    const auto t1 = now();
    TensorPtr loss = cross_entropy_loss(logits, g_symbols.at(targets));
    const auto t2 = now();
    const std::vector<ParameterPtr> params = m->parameters();
#ifdef WEED_B200
    if (ev_on) throw_on_error(weedcu_event_record(evs[ev_step % 3][1], cstream), "events");
#endif
    bool chained = false;
#ifdef WEED_B200
    // data parallel: bucketed gradient all-reduce on a communication stream, overlapped with the
    // rest of the backward walk (WH_DP_OVERLAP=0: one grouped all-reduce after backward)
    static const bool overlap = !(getenv("WH_DP_OVERLAP") && atoi(getenv("WH_DP_OVERLAP")) == 0);
    const bool dp = g_comm && g_world > 1 && g_dp_active;
    if (dp && overlap) {
      if (!g_buckets) {
        const char *bb = getenv("WH_DP_BUCKET_BYTES");
        g_buckets.reset(bb ? new GradientBuckets(g_comm, (size_t)atoll(bb)) : new GradientBuckets(g_comm));
      }
      // WH_DP_CHAIN_ADAM=1: the fused Adam update of each bucket's parameters runs on the communication stream right behind
      // the bucket's all-reduce (GradientBuckets::begin(opt, params)). Off by default: measured on 2 x B200 the chained step
      // takes 10.52 ms against 10.38 ms — the step is throughput-bound, so an update that overlaps backward only moves its
      // HBM traffic into backward, and the ~20 extra launches delay the next bucket's all-reduce
      static const bool chain = getenv("WH_DP_CHAIN_ADAM") && atoi(getenv("WH_DP_CHAIN_ADAM")) != 0;
      chained = chain;
      if (chain) g_buckets->begin(*g_adams.at(opt), params);
      else g_buckets->begin();
    }
#endif
    Tensor::backward(loss);
    const auto t3 = now();
#ifdef WEED_B200
    if (ev_on) throw_on_error(weedcu_event_record(evs[ev_step % 3][2], cstream), "events");
    // The parameters outside the last flushed bucket (the embedding's gradient, 154 MB at the GPT-2 shape) are updated while
    // that bucket is still being exchanged: 10.60 -> 10.38 ms/step on 8 x B200, 9.86 -> 9.77 on 2 (WH_DP_SPLIT_ADAM=0: one
    // update after the whole exchange)
    static const bool split = !(getenv("WH_DP_SPLIT_ADAM") && atoi(getenv("WH_DP_SPLIT_ADAM")) == 0);
    bool updated = false;
    if (dp && overlap && !chained && split) {
      g_buckets->finish_async(params);
      std::unordered_set<Tensor *> in_tail(g_buckets->tail.begin(), g_buckets->tail.end());
      std::vector<ParameterPtr> head, tail;
      for (const ParameterPtr &p : params) (in_tail.count(p.get()) ? tail : head).push_back(p);
      Adam &o = *g_adams.at(opt);
      real1 bc1, bc2;
      adam_begin_step(o, bc1, bc2);
      std::vector<ParameterPtr> slow;
      AdamBatch hb, tb;
      adam_collect(o, head, hb, slow);
      adam_collect(o, tail, tb, slow);
      g_buckets->wait_head();
      static void *sev[4] = {nullptr, nullptr, nullptr, nullptr};
      static uint64_t scount = 0;
      if (ev_on) {
        if (!sev[0])
          for (void *&e : sev) throw_on_error(weedcu_event_create(&e), "events");
        else if (scount % 4 == 3) { // previous step's: head update, and the last bucket's exchange as the communication stream saw it
          float a = 0, b = 0, c = 0;
          weedcu_event_sync(sev[2]);
          weedcu_event_elapsed_ms(sev[0], sev[1], &a);
          weedcu_event_elapsed_ms(g_buckets->ev_head, g_buckets->ev_done, &b);
          weedcu_event_elapsed_ms(sev[0], sev[2], &c);
          fprintf(stderr, "[wh split] adam(head) %.3f ms, last bucket on the comm stream %.3f ms, adam(head) start -> adam(tail) end %.3f ms\n", a, b, c);
        }
        ++scount;
        throw_on_error(weedcu_event_record(sev[0], cstream), "events");
      }
      adam_launch(o, hb, bc1, bc2, nullptr);
      if (ev_on) throw_on_error(weedcu_event_record(sev[1], cstream), "events");
      g_buckets->wait_all();
      adam_launch(o, tb, bc1, bc2, nullptr);
      if (ev_on) throw_on_error(weedcu_event_record(sev[2], cstream), "events");
      for (const ParameterPtr &p : slow) adam_slow(o, p, bc1, bc2);
      updated = true;
    } else if (dp && overlap) g_buckets->finish(params);
    else if (dp) allreduce_gradients(params, g_comm);
#else
    const bool updated = false;
#endif
    const auto t4 = now();
#ifdef WEED_B200
    if (ev_on) throw_on_error(weedcu_event_record(evs[ev_step % 3][3], cstream), "events");
#endif
    if (!chained && !updated) adam_step(*g_adams.at(opt), params);
    const auto t5 = now();
    zero_grad(params);
    m->reset_cache();
#ifdef WEED_B200
    if (ev_on) {
      throw_on_error(weedcu_event_record(evs[ev_step % 3][4], cstream), "events");
      ++ev_step;
    }
#endif
    const auto t6 = now();
    if (timing)
      fprintf(stderr, "[wh] forward %.0f us, loss %.0f, backward %.0f, allreduce %.0f, adam %.0f, zero_grad %.0f\n", us(t0, t1), us(t1, t2),
              us(t2, t3), us(t3, t4), us(t4, t5), us(t5, t6));
    return put(loss);
  })
}

double wh_wall_seconds() {
  return std::chrono::duration<double>(std::chrono::steady_clock::now().time_since_epoch()).count();
}

} // extern "C"
