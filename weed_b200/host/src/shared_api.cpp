// shared_api.cpp — the reference's C API (src/shared_api.cpp:79-449) over this backend's host library: module ids index a
// table of (module, last result, error latch); one mutex per module plus a table mutex, as the reference serialises it
// (:21-45). Every entry catches C++ exceptions, prints what() like the reference and latches an error code for
// get_error(): 1 = the module call failed, 2 = invalid argument / unknown id.
#include "shared_api.hpp"

#include "weed_b200/modules.hpp"

#include <fstream>
#include <iostream>
#include <mutex>

using namespace Weed;

namespace {
struct ModuleResult {
  std::mutex mtx;
  ModulePtr m;
  TensorPtr t;
  int error;
  explicit ModuleResult(ModulePtr a) : m(a), t(nullptr), error(0) {}
};
std::mutex table_mutex;
int meta_error = 0;
std::vector<std::unique_ptr<ModuleResult>> table;

// the entry of `mid`, or null (and meta_error = 2) when the id is not live; `lock` then holds the module's mutex
ModuleResult *acquire(uintw mid, std::unique_lock<std::mutex> &lock) {
  std::lock_guard<std::mutex> meta(table_mutex);
  if ((mid >= table.size()) || !table[mid]) {
    std::cout << "Invalid argument: module ID not found!" << std::endl;
    meta_error = 2;
    return nullptr;
  }
  lock = std::unique_lock<std::mutex>(table[mid]->mtx);
  return table[mid].get();
}
// column-major dense extents of a C-API input: stride[i] = prod(shape[0..i)) (src/shared_api.cpp:173-190)
tcapint dense_extent(uintw n, const uintw *shape, std::vector<tcapint> &sh) {
  sh.resize(n);
  tcapint stride = 1U, max_index = 0U;
  for (size_t i = 0U; i < n; ++i) {
    sh[i] = (tcapint)shape[i];
    max_index += (sh[i] - 1U) * stride;
    stride *= sh[i];
  }
  return n ? max_index + 1U : 0U;
}
const TensorPtr result_of(ModuleResult *r) {
  if (!r->t) {
    std::cout << "Invalid argument: module result tensor not found!" << std::endl;
    std::lock_guard<std::mutex> meta(table_mutex);
    meta_error = 2;
  }
  return r->t;
}
} // namespace

extern "C" {
int get_error(const uintw mid) {
  std::lock_guard<std::mutex> meta(table_mutex);
  if (meta_error) {
    meta_error = 0;
    return 2;
  }
  if ((mid >= table.size()) || !table[mid]) {
    std::cout << "Invalid argument: module ID not found!" << std::endl;
    return 2;
  }
  const int e = table[mid]->error;
  table[mid]->error = 0;
  return e;
}

uintw load_module(const char *f) {
  std::lock_guard<std::mutex> meta(table_mutex);
  ModulePtr m;
  try {
    std::ifstream i(f, std::ios::binary);
    if (!i) throw std::invalid_argument(std::string("load_module: cannot open ") + f);
    m = Module::load(i);
    i.close();
    m->eval();
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    meta_error = 1;
    return 0U;
  }
  uintw id = 0U;
  while ((id < table.size()) && table[id]) ++id;
  if (id == table.size()) table.push_back(std::unique_ptr<ModuleResult>(new ModuleResult(m)));
  else table[id] = std::unique_ptr<ModuleResult>(new ModuleResult(m));
  return id;
}

void save_module(uintw mid, const char *f) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  try {
    std::ofstream o(f, std::ios::binary);
    r->m->train();
    r->m->save(o);
    o.close();
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    std::lock_guard<std::mutex> meta(table_mutex);
    meta_error = 1;
  }
}

void free_module(uintw mid) {
  {
    std::unique_lock<std::mutex> lock;
    if (!acquire(mid, lock)) return;
  } // (the module's mutex is released before its entry is destroyed)
  std::lock_guard<std::mutex> meta(table_mutex);
  if ((mid < table.size()) && table[mid]) table[mid] = nullptr;
}

void forward(uintw mid, uintw dtype, uintw n, uintw *shape, double *d) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  TensorPtr x;
  try {
    if (dtype != 1U) throw std::invalid_argument("forward: only real (dtype 1) inputs exist on the CUDA backend");
    std::vector<tcapint> sh;
    const tcapint count = dense_extent(n, shape, sh);
    std::vector<real1> v(count);
    for (size_t i = 0U; i < count; ++i) v[i] = (real1)d[i];
    x = std::make_shared<Tensor>(v, sh);
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    std::lock_guard<std::mutex> meta(table_mutex);
    meta_error = 2;
    return;
  }
  try {
    r->t = Tensor::contiguous(r->m->forward(x));
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    r->error = 1;
  }
}

void forward_int(uintw mid, uintw, uintw n, uintw *shape, intw *d) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  SymbolTensorPtr x;
  try {
    std::vector<tcapint> sh;
    const tcapint count = dense_extent(n, shape, sh);
    std::vector<symint> v(count);
    for (size_t i = 0U; i < count; ++i) v[i] = (symint)d[i];
    x = std::make_shared<SymbolTensor>(v, sh);
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    std::lock_guard<std::mutex> meta(table_mutex);
    meta_error = 2;
    return;
  }
  try {
    r->t = Tensor::contiguous(r->m->forward(x));
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    r->error = 1;
  }
}

uintw get_result_index_count(uintw mid) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return 0U;
  const TensorPtr t = result_of(r);
  return t ? (uintw)t->shape.size() : 0U;
}
void get_result_dims(uintw mid, uintw *shape, uintw *stride) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  const TensorPtr t = result_of(r);
  if (!t) return;
  for (size_t i = 0U; i < t->shape.size(); ++i) {
    shape[i] = t->shape[i];
    stride[i] = t->stride[i];
  }
}
uintw get_result_size(uintw mid) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return 0U;
  const TensorPtr t = result_of(r);
  return t ? (uintw)t->storage->size : 0U;
}
uintw get_result_offset(uintw mid) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return 0U;
  const TensorPtr t = result_of(r);
  return t ? (uintw)t->offset : 0U;
}
uintw get_result_type(uintw mid) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return 0U;
  const TensorPtr t = result_of(r);
  return t ? (uintw)t->storage->dtype : 0U;
}
void get_result(uintw mid, double *d) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  const TensorPtr t = result_of(r);
  if (!t) return;
  const std::vector<real1> host = to_host(*t); // one blocking read-back of the whole storage
  for (size_t i = 0U; i < host.size(); ++i) d[i] = (double)host[i];
}

// one SGD step on token input (src/shared_api.cpp:356-424): train(), forward, cross-entropy over target_ids, backward,
// sgd_step, eval()
void train_step(uintw mid, uintw n, uintw *shape, intw *input_ids, uintw n_target, intw *target_ids, double learning_rate) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  try {
    r->m->train();
    std::vector<tcapint> sh;
    const tcapint count = dense_extent(n, shape, sh);
    std::vector<symint> v(count);
    for (size_t i = 0U; i < count; ++i) v[i] = (symint)input_ids[i];
    SymbolTensorPtr x = std::make_shared<SymbolTensor>(v, sh);
    std::vector<symint> tgt(n_target);
    for (tcapint i = 0U; i < n_target; ++i) tgt[i] = (symint)target_ids[i];
    SymbolTensorPtr targets = std::make_shared<SymbolTensor>(tgt, std::vector<tcapint>{(tcapint)n_target});
    TensorPtr logits = r->m->forward(x);
    TensorPtr loss = cross_entropy_loss(logits, targets);
    Tensor::backward(loss);
    sgd_step(r->m->parameters(), real1(learning_rate));
    r->m->eval();
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    r->error = 1;
  }
}

void reset_kv_cache(uintw mid) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  r->error = 0;
  try {
    r->m->reset_cache();
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    r->error = 1;
  }
}
void set_max_kv_seq_len(uintw mid, uintw m) {
  std::unique_lock<std::mutex> lock;
  ModuleResult *r = acquire(mid, lock);
  if (!r) return;
  r->error = 0;
  try {
    r->m->set_max_kv_seq_len((tcapint)m);
  } catch (const std::exception &ex) {
    std::cout << ex.what() << std::endl;
    r->error = 1;
  }
}
} // extern "C"
