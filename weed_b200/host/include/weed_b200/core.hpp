// core.hpp — host side of the B200 backend: Weed's type system, Storage hierarchy, GpuDevice /
// CUDAEngine and the BaseTensor / Tensor / Parameter / SymbolTensor view types, re-written from
// scratch with the reference's public names and signatures so client code written against
// vm6502q/weed compiles unchanged (tools/harness/weed_harness.cpp is built against BOTH).
//
// What differs from the reference (by design):
//   * ENABLE_GPU comes from WEED_ENABLE_CUDA; the engine singleton is CUDAEngine and a "buffer" is
//     an owning device pointer from the stream-ordered pool instead of a cl::Buffer
//     (reference include/storage/gpu_storage.hpp:25-102, include/devices/gpu_device.hpp:30-287).
//   * Placement is never size-based: DEFAULT_DEVICE means GPU and get_dtag_by_presidence() returns
//     GPU (the reference bounces tensors below GSTRIDE to the CPU, src/tensors/base_tensor.cpp:56-75;
//     SURVEY §7 hard part 4). CPU storages exist only as host staging for upload / read-back.
//   * Real dtype only on the device path (complex/sparse are outside SURVEY §8).
#pragma once

#include "weedcu.h"

#include <complex>
#include <cstddef>
#include <cstdint>
#include <cstring>
#include <functional>
#include <iosfwd>
#include <limits>
#include <memory>
#include <mutex>
#include <stdexcept>
#include <string>
#include <unordered_map>
#include <vector>

#define WEED_ENABLE_CUDA 1
#define ENABLE_GPU 1
#define WEED_FPPOW 5
#define WEED_TCAPPOW 5

#define tcapint uint32_t
#define symint int32_t
#define tlenint uint32_t

namespace Weed {
typedef float real1;
typedef float real1_f;
typedef float real1_s;
typedef std::complex<real1> complex;

#define ZERO_R1 0.0f
#define ONE_R1 1.0f
#define HALF_R1 0.5f
#define ZERO_R1_F 0.0f
#define ONE_R1_F 1.0f
const complex ONE_CMPLX = complex(ONE_R1, ZERO_R1); // reference include/common/weed_types.hpp:206-208
const complex ZERO_CMPLX = complex(ZERO_R1, ZERO_R1);
const complex I_CMPLX = complex(ZERO_R1, ONE_R1);
constexpr real1 PI_R1 = (real1)3.14159265358979323846;
constexpr real1 E_R1 = (real1)2.71828182845904523536;
constexpr real1 ADAM_BETA1_DEFAULT = (real1)0.9;
constexpr real1 ADAM_BETA2_DEFAULT = (real1)0.999;
constexpr real1 ADAM_EPSILON_DEFAULT = (real1)1e-8;
// reference include/common/weed_types.hpp:213-214
constexpr real1 FP_NORM_EPSILON = (real1)(std::numeric_limits<real1>::epsilon() / 4);

// reference include/enums/*.hpp — values are part of the serialised format, kept identical
enum DeviceTag { NONE_DEVICE = 0, DEFAULT_DEVICE = 1, CPU = 2, GPU = 3 };
enum DType { NONE_DTYPE = 0, REAL = 1, COMPLEX = 2, INT = 3, DEFAULT_DTYPE = REAL };
enum StorageType {
  NONE_STORAGE_TYPE = 0, REAL_CPU_DENSE = 1, REAL_GPU_DENSE = 2, COMPLEX_CPU_DENSE = 3,
  COMPLEX_GPU_DENSE = 4, INT_CPU_DENSE = 5, INT_GPU_DENSE = 6, REAL_CPU_SPARSE = 7,
  COMPLEX_CPU_SPARSE = 8
};
enum ActivationFunctionType { NONE_FN = 0, SIGMOID_FN = 1, TANH_FN = 2, RELU_FN = 3, GELU_FN = 4, SWIGLU_FN = 5 };

struct bad_alloc : public std::bad_alloc {
  std::string m;
  bad_alloc(const std::string &message) : m(message) {}
  const char *what() const noexcept override { return m.c_str(); }
};

// ---------------------------------------------------------------------------------------------
// Backend switches (process-global, like the reference's pfControl / engine singletons).
struct Tensor;
struct BackendConfig {
  // Fused device kernels behind Tensor::gelu, LayerNorm::forward, adam_step, sgd_step,
  // cross_entropy_loss, the attention core and matmul-backward accumulation. Off = every op is
  // issued exactly as the reference composes it (one kernel per Weed:: op).
  bool fused = true;
  // Reproduce the reference CPU loops' index decomposition in reduce / reduce_grad
  // (src/ops/reduce.cpp:17-31,84-99) bit-for-bit, including its permutation of rank>=3 outputs.
  bool ref_index_quirks = false;
  // WEEDCU_GEMM_FP32 (FpMath parity path) or WEEDCU_GEMM_BF16 (tcgen05 tensor cores)
  int matmul_precision = WEEDCU_GEMM_FP32;
  // data-parallel: gradients are averaged over this many ranks inside the optimiser kernels
  real1 grad_scale = ONE_R1;
  // keep bf16 operand shadows on their storages (pack once per write, not once per GEMM)
  bool operand_cache = true;
  // FillZeros() on device buffers is deferred until something reads the buffer (fused mode only)
  bool lazy_zero = true;
  // gradients whose only readers are bf16 GEMM operands + column sums (dlogits of the fused cross-entropy,
  // the GELU input gradient) are not written in fp32 until something reads them (GpuStorage::deferred_values)
  bool defer_grads = true;
  // a gradient whose first contribution is a plain copy of another complete gradient shares its buffer
  // copy-on-write instead (GpuStorage::cow)
  bool cow_grads = true;
  // the tensor-core GEMM's extended epilogue (weedcu_gemm_bf16_ex) leaves what the consumer of a Linear output reads:
  // LayerNorm row partials behind the residual products, GELU + the bf16 operand copy behind ff1, log-sum-exp partials and
  // bf16-only logits behind the LM head, bf16-only Q / K / V
  bool epilogue_stats = true;
  // a Linear with at least this many output features whose output takes part in autograd is treated as an LM head: bf16-only
  // logits + log-sum-exp partials from the GEMM epilogue, fp32 logits only if something reads them
  tcapint lm_head_min_cols = 4096U;
  // Module::save writes device storages as REAL_CPU_DENSE / device id -1 — byte-compatible with what the reference's CPU
  // build writes for the same weights, and loadable by it. Off: REAL_GPU_DENSE + the device id, as the reference's GPU
  // build does (src/storage/gpu_real_storage.cpp:35-50). Loading accepts both and always places the data on the device.
  bool save_portable = true;
  // LayerNorm backward: false = what the reference's autograd chain computes (its div node drops dout on the denominator
  // branch, tensor.cpp:1506-1521, leaving a term that does not scale with the upstream gradient); true = the analytic
  // gradient. bench.py --check-dp uses `true` to test the gradient exchange: with the reference chain a sharded batch and
  // a whole batch differ by that unscaled term, whatever exchanges the gradients.
  bool layernorm_exact_grad = false;
  // the ff2 Linear's dA is written as its bf16 copy alone when GELU backward is its only reader (tensor.cpp: linear_gelu)
  bool bf16_act_grad = true;
  // Tensor::backward calls this for every leaf tensor (no grad_node, requires_grad: the Parameters)
  // right after the LAST node that lists it as a parent has run, i.e. when its gradient is final;
  // data-parallel training hangs the bucketed all-reduce on it (autograd.hpp: GradientBuckets)
  std::function<void(Tensor *)> on_leaf_grad_final;
};
BackendConfig &backend_config();

void throw_on_error(int rc, const char *what);
// an autograd closure found its output tensor destroyed (only possible for a stack copy of a graph tensor): never silent
[[noreturn]] void node_owner_lost();

// ---------------------------------------------------------------------------------------------
// Device layer (reference include/devices/gpu_device.hpp, include/common/oclengine.hpp)
struct DeviceBuffer {
  void *ptr;
  size_t bytes;
  void *stream;
  DeviceBuffer(void *p, size_t b, void *s) : ptr(p), bytes(b), stream(s) {}
  ~DeviceBuffer(); // stream-ordered free: safe while kernels are still in flight
  DeviceBuffer(const DeviceBuffer &) = delete;
  DeviceBuffer &operator=(const DeviceBuffer &) = delete;
};
typedef std::shared_ptr<DeviceBuffer> BufferPtr;

struct GpuDevice {
  int64_t deviceID;
  void *stream; // in-order compute stream == the reference's FIFO of QueueItems
  size_t totalAlloc = 0;
  size_t maxAlloc = (size_t)-1;
  std::mutex allocMutex;

  GpuDevice(int64_t did);
  void Bind() const; // cudaSetDevice

  BufferPtr MakeBuffer(size_t bytes, const void *host_ptr = nullptr);
  // blocking device->host read (reference LockSync, gpu_device.cpp:388-420)
  bool LockSync(BufferPtr buffer, size_t bytes, void *dst, bool allow_lock = false);
  void CopyBuffer(const BufferPtr &dst, const BufferPtr &src, size_t bytes); // stream-ordered device-to-device copy
  void UnlockSync(BufferPtr, void *) {}
  void ClearRealBuffer(BufferPtr buffer, size_t n);
  void FillOnesReal(BufferPtr buffer, size_t n);
  void FillValueReal(BufferPtr buffer, size_t n, real1 v);
  void ClearIntBuffer(BufferPtr buffer, size_t n);
  void FillOnesInt(BufferPtr buffer, size_t n);
  void FillValueInt(BufferPtr buffer, size_t n, symint v);
  real1 GetReal(BufferPtr buffer, tcapint idx);
  void SetReal(real1 v, BufferPtr buffer, tcapint idx);
  symint GetInt(BufferPtr buffer, tcapint idx);
  void SetInt(symint v, BufferPtr buffer, tcapint idx);
  void clFinish(bool hard = false);
  void AddAlloc(size_t sz);
  void SubtractAlloc(size_t sz);
};
typedef std::shared_ptr<GpuDevice> GpuDevicePtr;

struct CUDAEngine {
  static CUDAEngine &Instance();
  static void InitOCL() { Instance(); } // name kept: reference test_main.cpp:91-93 calls it
  int GetDeviceCount();
  int64_t GetDefaultDeviceID() { return default_device; }
  void SetDefaultDeviceID(int64_t did);
  GpuDevicePtr GetWeedDevice(int64_t did = -1); // -1 = default; ids wrap modulo device count
  size_t GetActiveAllocSize(int64_t did);

private:
  CUDAEngine();
  std::vector<GpuDevicePtr> devices;
  int64_t default_device = 0;
  std::mutex mtx;
};
#define WEED_GPU_SINGLETON (CUDAEngine::Instance())

// ---------------------------------------------------------------------------------------------
// Storage (reference include/storage/*.hpp)
struct Storage;
typedef std::shared_ptr<Storage> StoragePtr;

struct Storage : public std::enable_shared_from_this<Storage> {
  StorageType stype;
  DeviceTag device;
  DType dtype;
  tcapint size;
  Storage(StorageType st, DeviceTag dt, DType ty, tcapint n) : stype(st), device(dt), dtype(ty), size(n) {
    if (!size) throw std::invalid_argument("Storage must have size of at least 1!");
  }
  virtual ~Storage() {}
  virtual tcapint get_sparse_size() const { return size; }
  virtual bool is_sparse() const { return false; }
  virtual StoragePtr get_ptr() { return shared_from_this(); }
  virtual int64_t get_device_id() const { return -1; }
  virtual void FillZeros() = 0;
  virtual void FillOnes() = 0;
  virtual StoragePtr Upcast(const DType &dt);
  virtual bool is_gpu() = 0;
  virtual StoragePtr cpu() = 0;
  virtual StoragePtr gpu(const int64_t &did = -1) = 0;
  // Checkpoint format of the reference (src/storage/storage.cpp:25-119): storage type, device id (8 bytes), element count,
  // then the elements. GPU storages are read back to the host first (src/storage/gpu_real_storage.cpp:35-50).
  virtual void save(std::ostream &) const;
  static StoragePtr load(std::istream &);
  static void write_storage_type(std::ostream &out, const StorageType &x);
  static void read_storage_type(std::istream &in, StorageType &x);
};

// reference include/common/serializer.hpp:25-99: raw little-endian fields, no framing
struct Serializer {
  static void write_bool(std::ostream &out, const bool &x);
  static void read_bool(std::istream &in, bool &x);
  static void write_tcapint(std::ostream &out, const tcapint &x);
  static void read_tcapint(std::istream &in, tcapint &x);
  static void write_symint(std::ostream &out, const symint &x);
  static void read_symint(std::istream &in, symint &x);
  // the reference writes sizeof(int64_t) bytes starting at a 4-byte symint (serializer.hpp:51-56): the low word is the
  // value, the high word whatever followed it on the stack. Here: the sign-extended value; readers use the low word.
  static void write_int64(std::ostream &out, const symint &x);
  static void read_int64(std::istream &in, symint &x);
  static void write_size_t(std::ostream &out, const size_t &x);
  static void read_size_t(std::istream &in, size_t &x);
  static void write_real(std::ostream &out, const real1 &x);
  static void read_real(std::istream &in, real1 &x);
  static void write_real1_f(std::ostream &out, const real1_f &x);
  static void read_real1_f(std::istream &in, real1_f &x);
};

template <typename T> struct TypedStorage : Storage {
  TypedStorage(StorageType st, DeviceTag dt, tcapint n)
      : Storage(st, dt, std::is_same<T, real1>::value ? DType::REAL : DType::INT, n) {}
  virtual T operator[](const tcapint &idx) const = 0;
  virtual void write(const tcapint &idx, const T &val) = 0;
  virtual void add(const tcapint &idx, const T &val) = 0;
  virtual void FillValue(const T &v) = 0;
  void FillZeros() override { FillValue(T(0)); }
  void FillOnes() override { FillValue(T(1)); }
};
typedef TypedStorage<real1> RealStorage;
typedef TypedStorage<symint> IntStorage;
typedef std::shared_ptr<RealStorage> RealStoragePtr;
typedef std::shared_ptr<IntStorage> IntStoragePtr;

template <typename T> struct CpuStorage : TypedStorage<T> {
  std::vector<T> data; // host staging only; no compute runs on it in this backend
  CpuStorage(StorageType st, tcapint n) : TypedStorage<T>(st, DeviceTag::CPU, n), data(n) {}
  CpuStorage(StorageType st, const std::vector<T> &v) : TypedStorage<T>(st, DeviceTag::CPU, (tcapint)v.size()), data(v) {}
  T operator[](const tcapint &idx) const override { return data.at(idx); }
  void write(const tcapint &idx, const T &val) override { data.at(idx) = val; }
  void add(const tcapint &idx, const T &val) override { data.at(idx) += val; }
  void FillValue(const T &v) override { std::fill(data.begin(), data.end(), v); }
  bool is_gpu() override { return false; }
  StoragePtr cpu() override { return Storage::get_ptr(); }
};
struct CpuRealStorage : CpuStorage<real1> {
  CpuRealStorage(tcapint n) : CpuStorage<real1>(REAL_CPU_DENSE, n) {}
  CpuRealStorage(const std::vector<real1> &v) : CpuStorage<real1>(REAL_CPU_DENSE, v) {}
  StoragePtr gpu(const int64_t &did = -1) override;
  void save(std::ostream &) const override;
};
struct CpuIntStorage : CpuStorage<symint> {
  CpuIntStorage(tcapint n) : CpuStorage<symint>(INT_CPU_DENSE, n) {}
  CpuIntStorage(const std::vector<symint> &v) : CpuStorage<symint>(INT_CPU_DENSE, v) {}
  StoragePtr gpu(const int64_t &did = -1) override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<CpuRealStorage> CpuRealStoragePtr;
typedef std::shared_ptr<CpuIntStorage> CpuIntStoragePtr;

template <typename T> struct GpuStorage : TypedStorage<T> {
  GpuDevicePtr dev;
  mutable BufferPtr buffer;
  // Copy-on-write sharing of the device buffer between storages (BackendConfig::cow_grads): the first
  // contribution to a lazily zeroed gradient that is a plain copy of another complete gradient (d(a + b)/da = 1
  // for a parent with several consumers) shares that gradient's buffer instead of copying 8 B/elem. Every
  // access goes through device_ptr*(): reads see the shared buffer, the first write gives the writer a
  // private buffer (copied, or fresh when the write overwrites everything), so neither storage can observe
  // the other's later writes. Storages that share hold the same token; alone again == not shared.
  mutable std::shared_ptr<char> cow;
  bool buffer_shared() const { return cow && cow.use_count() > 1; }
  void unshare(bool keep_contents) const {
    if (!buffer_shared()) {
      cow.reset();
      return;
    }
    BufferPtr fresh = dev->MakeBuffer(sizeof(T) * (size_t)TypedStorage<T>::size);
    if (keep_contents) dev->CopyBuffer(fresh, buffer, sizeof(T) * (size_t)TypedStorage<T>::size);
    buffer = fresh;
    cow.reset();
  }
  // this storage := src's current contents, without a copy. Both must be the same size on the same device.
  void share_buffer_from(const GpuStorage<T> &src) {
    src.materialize();
    if (!src.cow) src.cow = std::make_shared<char>(0);
    cow = src.cow;
    buffer = src.buffer;
    deferred_values = nullptr;
    zero_pending = false;
    ++version;
  }
  GpuStorage(StorageType st, tcapint n, int64_t did, bool alloc = true) : TypedStorage<T>(st, DeviceTag::GPU, n) {
    dev = CUDAEngine::Instance().GetWeedDevice(did);
    dev->AddAlloc(sizeof(T) * (size_t)n);
    if (alloc) buffer = dev->MakeBuffer(sizeof(T) * (size_t)n);
  }
  GpuStorage(StorageType st, const std::vector<T> &val, int64_t did)
      : TypedStorage<T>(st, DeviceTag::GPU, (tcapint)val.size()) {
    dev = CUDAEngine::Instance().GetWeedDevice(did);
    dev->AddAlloc(sizeof(T) * val.size());
    buffer = dev->MakeBuffer(sizeof(T) * val.size(), val.data());
  }
  virtual ~GpuStorage() { dev->SubtractAlloc(sizeof(T) * (size_t)TypedStorage<T>::size); }
  int64_t get_device_id() const override { return dev->deviceID; }
  // Lazy zero-fill (fused mode): FillZeros() on a gradient only marks the buffer; the fill is issued
  // by the first access that does not overwrite the whole buffer, and skipped entirely when the
  // first consumer overwrites it (matmul / copy instead of accumulate) or nothing ever touches it.
  // `version` counts potential writes; the bf16 operand shadows of GpuRealStorage key on it.
  // the only reader of these values takes their bf16 copy (the gradient of a fused Linear + GELU output: gelu_grad): a
  // tensor-core product that overwrites the whole storage may then write that copy alone and defer the fp32 values
  bool accept_bf16_values = false;
  mutable bool zero_pending = false;
  mutable uint64_t version = 0;
  // the buffer is KNOWN to hold zeros while zero_version == version (the fused Adam kernel zeroes small gradients right
  // after reading them): FillZeros() is then a no-op, and accumulating kernels add into real zeros — no fill launch
  uint64_t zero_version = ~(uint64_t)0;
  // Deferred values (fused mode, BackendConfig::defer_grads): a producer that already left everything its
  // consumers read — the bf16 GEMM operand copy and the column sums — may skip writing the fp32 values and
  // register how to compute them instead. The first access that needs the fp32 buffer runs it (a full
  // overwrite that does not count as a write: the shadows stay valid); a full overwrite or fill drops it.
  mutable std::function<void()> deferred_values;
  void materialize() const {
    if (deferred_values) {
      std::function<void()> f;
      f.swap(deferred_values);
      zero_pending = false;
      unshare(false);
      f();
      return;
    }
    if (!zero_pending) return;
    zero_pending = false;
    unshare(false);
    const_cast<GpuStorage<T> *>(this)->fill_now(T(0));
  }
  virtual void fill_now(const T &v) = 0;
  T *device_ptr() const { // read/write access
    materialize();
    unshare(true);
    ++version;
    return reinterpret_cast<T *>(buffer->ptr);
  }
  const T *device_ptr_ro() const { // read-only access: does not invalidate shadows
    materialize();
    return reinterpret_cast<const T *>(buffer->ptr);
  }
  T *device_ptr_overwrite() const { // the caller overwrites every element
    deferred_values = nullptr;
    zero_pending = false;
    unshare(false);
    ++version;
    return reinterpret_cast<T *>(buffer->ptr);
  }
  void write(const tcapint &, const T &) override { throw std::domain_error("Don't use GPU-based Storage::write()!"); }
  void add(const tcapint &, const T &) override { throw std::domain_error("Don't use GPU-based Storage::add()!"); }
  bool is_gpu() override { return true; }
  StoragePtr gpu(const int64_t & = -1) override { return Storage::get_ptr(); }
};
struct GpuRealStorage : GpuStorage<real1> {
  GpuRealStorage(const tcapint &n, int64_t did, const bool &alloc = true) : GpuStorage<real1>(REAL_GPU_DENSE, n, did, alloc) {}
  GpuRealStorage(const std::vector<real1> &val, const int64_t &did = -1) : GpuStorage<real1>(REAL_GPU_DENSE, val, did) {}
  void FillValue(const real1 &v) override;
  void fill_now(const real1 &v) override { dev->FillValueReal(buffer, size, v); }
  real1 operator[](const tcapint &idx) const override {
    if (idx >= size) throw std::invalid_argument("GpuStorage::operator[] argument out-of-bounds!");
    materialize();
    return dev->GetReal(buffer, idx);
  }
  StoragePtr cpu() override;
  void save(std::ostream &) const override;
  // bf16 copies of matrix views of this storage, packed for the tensor-core GEMM (ops.cpp)
  struct Bf16Shadow {
    BufferPtr buf;
    uint64_t version;
    tcapint offset, n_fast, n_slow, s_fast, s_slow;
  };
  std::vector<Bf16Shadow> shadows;
  // column sums of the [n_fast, n_slow] matrix this storage holds, left by a producer that had the
  // values in registers anyway (the fused cross-entropy backward); valid while colsum_version == version
  BufferPtr colsum;
  uint64_t colsum_version = 0U;
  tcapint colsum_n = 0U;
  // per-row statistics of the [rows, cols] matrix this storage holds, left as per-column-tile partials by the epilogue of
  // the GEMM that produced it (weedcu_gemm_bf16_ex): kind 1 = (mean, M2) for LayerNorm::forward, kind 2 = (max, sum exp)
  // for cross_entropy_loss; valid while row_stats_version == version
  // the bf16 operands of the product that wrote this storage's bf16 copy without writing the fp32 values (the LM head's
  // logits): what the deferred fp32 values and the cross-entropy's target logit are recomputed from
  struct GemmSource {
    BufferPtr a, b;
    int a_major = 0, b_major = 0;
    uint64_t lda = 0U, ldb = 0U;
    uint32_t M = 0U, N = 0U, K = 0U;
    StoragePtr w_storage, bias_storage; // the weight whose shadow `b` is (rewritten in place by the optimiser) and the bias
    uint64_t w_version = 0U, bias_version = 0U;
    tcapint bias_offset = 0U;
  };
  std::shared_ptr<GemmSource> gemm_source;
  BufferPtr row_stats;
  int row_stats_kind = 0;
  uint32_t row_stats_tiles = 0U, row_stats_tile_cols = 0U;
  tcapint row_stats_rows = 0U, row_stats_cols = 0U;
  uint64_t row_stats_version = 0U;
};
struct GpuIntStorage : GpuStorage<symint> {
  GpuIntStorage(const tcapint &n, int64_t did, const bool &alloc = true) : GpuStorage<symint>(INT_GPU_DENSE, n, did, alloc) {}
  GpuIntStorage(const std::vector<symint> &val, const int64_t &did = -1) : GpuStorage<symint>(INT_GPU_DENSE, val, did) {}
  void FillValue(const symint &v) override {
    zero_pending = false;
    ++version;
    dev->FillValueInt(buffer, size, v);
  }
  void fill_now(const symint &v) override { dev->FillValueInt(buffer, size, v); }
  symint operator[](const tcapint &idx) const override {
    if (idx >= size) throw std::invalid_argument("GpuStorage::operator[] argument out-of-bounds!");
    return dev->GetInt(buffer, idx);
  }
  StoragePtr cpu() override;
  void save(std::ostream &) const override;
};
typedef std::shared_ptr<GpuRealStorage> GpuRealStoragePtr;
typedef std::shared_ptr<GpuIntStorage> GpuIntStoragePtr;

// ---------------------------------------------------------------------------------------------
// Tensors (reference include/tensors/base_tensor.hpp, tensor.hpp, parameter.hpp, symbol_tensor.hpp)
struct Node;
typedef std::shared_ptr<Node> NodePtr;
struct BaseTensor;
typedef std::shared_ptr<BaseTensor> BaseTensorPtr;

struct BaseTensor {
  StoragePtr storage;
  tcapint offset;
  std::vector<tcapint> shape;
  std::vector<tcapint> stride;

  BaseTensor() : storage(nullptr), offset(0U) {}
  BaseTensor(const std::vector<tcapint> &shp, const std::vector<tcapint> &strd) : storage(nullptr), offset(0U), shape(shp), stride(strd) {
    validate_constructor();
  }
  virtual ~BaseTensor() {}

  void copy(const BaseTensor &cp) {
    storage = cp.storage;
    offset = cp.offset;
    shape = cp.shape;
    stride = cp.stride;
  }
  void validate_constructor();
  tcapint get_size() const;            // 1 + sum (shape-1)*stride : span in storage
  tcapint get_broadcast_size() const;  // prod shape
  bool is_contiguous() const { return !offset && is_contiguous(shape, stride); }
  bool is_scalar() const;
  bool covers_storage() const; // the view addresses every storage element exactly once
  tcapint get_storage_index(const tcapint &idx) const;
  void reshape(const std::vector<symint> &s);
  void transpose();
  void transpose(symint i, symint j);
  void flatten(symint axis);
  static bool is_contiguous(const std::vector<tcapint> &shp, const std::vector<tcapint> &s);
  static std::vector<tcapint> full_contiguous_stride(const std::vector<tcapint> &shp);
  static DType get_dtype_by_presidence(const std::vector<BaseTensorPtr> &v);
  static DeviceTag get_dtag_by_presidence(const std::vector<BaseTensorPtr> &v);
  // weedcu view of this tensor (offset, shape, stride), rank <= 8
  weedcu_view view() const;
};

struct SymbolTensor;
typedef std::shared_ptr<SymbolTensor> SymbolTensorPtr;
struct SymbolTensor : BaseTensor {
  SymbolTensor(const std::vector<tcapint> &shp, const std::vector<tcapint> &strd, const bool &rg = false,
               const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1, const bool &s = true);
  SymbolTensor(const std::vector<symint> &val, const std::vector<tcapint> &shp, const bool &rg = false,
               const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1);
  SymbolTensor(const SymbolTensor &orig) { copy(orig); }
  void copy(const SymbolTensor &cp) { BaseTensor::copy(cp); }
  SymbolTensorPtr cast(const DeviceTag &dt) const;
  using BaseTensor::reshape;
  static SymbolTensorPtr reshape(const SymbolTensorPtr a, const std::vector<symint> &s);
  using BaseTensor::transpose;
  static SymbolTensorPtr transpose(const SymbolTensorPtr a);
  static SymbolTensorPtr transpose(const SymbolTensorPtr a, symint i, symint j);
  using BaseTensor::flatten;
  static SymbolTensorPtr flatten(const SymbolTensorPtr a, const symint &axis);
  const symint *device_ptr() const;
};

struct Tensor;
typedef std::shared_ptr<Tensor> TensorPtr;

#define SCALAR(v, o) std::make_shared<Weed::Tensor>(v, false, o->storage->device, o->storage->get_device_id())

struct Tensor : public BaseTensor, public std::enable_shared_from_this<Tensor> {
  NodePtr grad_node;
  TensorPtr grad;
  bool requires_grad = false;
  // A copy of a tensor that has a grad_node (reshape / transpose / flatten / operator[] / chunk views) shares that node, and
  // the node's closure reaches its output through a weak pointer (a strong one would be a cycle, DESIGN.md defect D7). The
  // copy therefore keeps the node's owner alive: if only the view survives, backward still finds the output tensor.
  TensorPtr view_owner;
  // how many autograd Nodes list this tensor as a parent (= how many gradient contributions it will
  // receive); a non-leaf with exactly one lets Tensor::add's backward hand it the incoming gradient
  // buffer instead of a copy (tensor.cpp: adopt_incoming_gradient)
  uint32_t consumers = 0U;

  Tensor() {}
  Tensor(const std::vector<tcapint> &shp, const std::vector<tcapint> &str, const bool &rg = false, const bool &s = true,
         const DType &dtype = DType::REAL, const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1);
  Tensor(const std::vector<real1> &val, const std::vector<tcapint> &shp, const bool &rg = false,
         const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1);
  Tensor(const real1 &val, const bool &rg = false, const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1)
      : Tensor(std::vector<real1>{val}, std::vector<tcapint>{1U}, rg, dtag, did) {}
  Tensor(const Tensor &orig) : BaseTensor() { copy(orig); }
  Tensor &operator=(const Tensor &orig) { copy(orig); return *this; }

  void copy(const Tensor &cp) {
    BaseTensor::copy(cp);
    grad_node = cp.grad_node;
    grad = cp.grad;
    requires_grad = cp.requires_grad;
    view_owner = nullptr;
    if (cp.grad_node && &cp != this) view_owner = cp.view_owner ? cp.view_owner : cp.owner_ptr();
  }
  // shared_ptr to this tensor when it is owned by one (null for a stack object)
  TensorPtr owner_ptr() const {
    try {
      return std::const_pointer_cast<Tensor>(shared_from_this());
    } catch (const std::bad_weak_ptr &) {
      return nullptr;
    }
  }
  static TensorPtr clone(const TensorPtr &a);
  void make_gradient(const bool &force_sparse = false);
  bool match_shape(const TensorPtr a);
  void materialize_broadcast();
  void reduce_grad_broadcast();
  TensorPtr operator[](const tcapint &idx) const;
  void upcast(const DType &dt);
  TensorPtr cast(const DeviceTag &dt) const;
  void cast_in_place(const DeviceTag &dt);
  void squeeze();
  void squeeze(int64_t axis);
  void unsqueeze(int64_t axis);

  static TensorPtr zeros(const std::vector<tcapint> &shape, const bool &rg = false, const bool &s = true,
                         const DType &dtype = DType::REAL, const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1);
  static TensorPtr ones_like(const std::vector<tcapint> &shape, const bool &rg = false, const bool &s = true,
                             const DType &dtype = DType::REAL, const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1);
  static TensorPtr one_hot(const SymbolTensorPtr targets, const tcapint vocab_size);
  static TensorPtr make_gradient(const std::vector<tcapint> &shp, const bool &s, const DType &dtype, const DeviceTag &dtag, const int64_t did);
  static TensorPtr allocate_scalar_like(const Tensor &orig, const bool &rg);
  static TensorPtr allocate_like(const Tensor &orig, const DType &dt, const bool &rg, const bool &s);
  static TensorPtr allocate_like(const std::vector<tcapint> &shape, const Tensor &orig, const DType &dt, const bool &rg, const bool &s);
  static TensorPtr allocate_like(const std::vector<tcapint> &shape, const std::vector<tcapint> &stride, const Tensor &orig,
                                 const DType &dt, const bool &rg, const bool &s);
  static std::vector<TensorPtr> chunk(TensorPtr a, const size_t &chunks, int64_t axis = -1);
  static TensorPtr contiguous(const TensorPtr a);
  using BaseTensor::reshape;
  static TensorPtr reshape(const TensorPtr a, const std::vector<symint> &s);
  using BaseTensor::transpose;
  static TensorPtr transpose(const TensorPtr a);
  static TensorPtr transpose(const TensorPtr a, symint i, symint j);
  using BaseTensor::flatten;
  static TensorPtr flatten(const TensorPtr a, symint axis);

  static void backward(const TensorPtr loss);

  static TensorPtr softmax(const TensorPtr x, symint axis);
  static void make_softmax_node(TensorPtr x, TensorPtr out, symint axis);
  static TensorPtr logsoftmax(const TensorPtr x, symint axis);
  static void make_logsoftmax_node(TensorPtr x, TensorPtr out, symint axis);
  static TensorPtr slice(TensorPtr a, const int64_t &row);
  static void make_row_slice_node(TensorPtr a, TensorPtr out, const tcapint &row);
  static TensorPtr slice(TensorPtr a, int64_t axis, const tcapint &start, const tcapint &length);
  static void make_slice_node(TensorPtr a, TensorPtr out, const int64_t &axis, const tcapint &start);
  static TensorPtr sum(TensorPtr a);
  static void make_sum_node(TensorPtr a, TensorPtr out);
  static TensorPtr mean(TensorPtr a);
  static void make_mean_node(TensorPtr a, TensorPtr out);
  static TensorPtr mean(TensorPtr a, symint axis);
  static TensorPtr variance(TensorPtr a);
  static TensorPtr variance(TensorPtr a, const tcapint &axis);
  static TensorPtr stddev(TensorPtr a) { return pow(variance(a), real1(0.5)); }
  static TensorPtr stddev(TensorPtr a, const tcapint &axis) { return pow(variance(a, axis), real1(0.5)); }
  static TensorPtr sum(TensorPtr a, symint axis);
  static void make_sum_node(TensorPtr a, TensorPtr out, const tcapint &axis);
  static TensorPtr max(TensorPtr a, symint axis);
  static TensorPtr min(TensorPtr a, symint axis);
  static void make_match_node(TensorPtr a, TensorPtr out, const tcapint &axis);
  static TensorPtr max(TensorPtr a);
  static void make_max_node(TensorPtr a, TensorPtr out);
  static TensorPtr min(TensorPtr a);
  static void make_min_node(TensorPtr a, TensorPtr out);
  static TensorPtr clamp(TensorPtr a, real1 lo, real1 hi);
  static void make_clamp_node(TensorPtr a, real1 lo, real1 hi, TensorPtr out);
  static TensorPtr abs(TensorPtr a);
  static void make_abs_node(TensorPtr a, TensorPtr out);
  static TensorPtr sigmoid(TensorPtr a);
  static void make_sigmoid_node(TensorPtr a, TensorPtr out);
  static TensorPtr tanh(TensorPtr a);
  static void make_tanh_node(TensorPtr a, TensorPtr out);
  static TensorPtr gelu(const TensorPtr x);
  static void make_gelu_node(TensorPtr a, TensorPtr out);
  // gelu(a w + bias) with the activation in the tensor-core GEMM's epilogue (same two autograd nodes as Linear::forward +
  // Tensor::gelu), or nullptr when that kernel does not apply
  static TensorPtr linear_gelu(TensorPtr a, TensorPtr w, TensorPtr bias);
  static TensorPtr relu(TensorPtr a);
  static void make_relu_node(TensorPtr a, TensorPtr out);
  static TensorPtr sin(TensorPtr a);
  static void make_sin_node(TensorPtr a, TensorPtr out);
  static TensorPtr cos(TensorPtr a);
  static void make_cos_node(TensorPtr a, TensorPtr out);
  static TensorPtr add(TensorPtr a, TensorPtr b);
  static void make_add_node(TensorPtr a, TensorPtr b, TensorPtr out);
  static TensorPtr mul(TensorPtr a, TensorPtr b);
  static void make_mul_node(TensorPtr a, TensorPtr b, TensorPtr out);
  static TensorPtr matmul(TensorPtr a, TensorPtr b);
  static void make_matmul_node(TensorPtr a, TensorPtr b, TensorPtr out);
  static void matmul_backward(TensorPtr a, TensorPtr b, TensorPtr out);
  // fused x W + bias (+ residual: the `x + Linear(...)` of a transformer block in the same kernel), or nullptr
  static TensorPtr linear(TensorPtr a, TensorPtr w, TensorPtr bias, TensorPtr residual = nullptr);
  // the same for two or three Linear layers reading one input (W_q / W_k / W_v): one grouped launch,
  // each output with exactly the autograd node Tensor::linear would give it; empty when not eligible
  static TensorPtr finish_linear(TensorPtr a, TensorPtr w, TensorPtr bias, TensorPtr out, bool rg, TensorPtr residual = nullptr);
  static std::vector<TensorPtr> linear_grouped(TensorPtr a, const std::vector<TensorPtr> &ws, const std::vector<TensorPtr> &biases, bool bf16_only = false);
  static TensorPtr sub(TensorPtr a, TensorPtr b);
  static void make_sub_node(TensorPtr a, TensorPtr b, TensorPtr out);
  static TensorPtr div(TensorPtr a, TensorPtr b);
  static void make_div_node(TensorPtr a, TensorPtr b, TensorPtr out);
  static TensorPtr pow(TensorPtr a, real1 p);
  static void make_pow_node(TensorPtr a, real1 p, TensorPtr out);
  static TensorPtr exp(TensorPtr a, real1 b = E_R1);
  static void make_exp_node(TensorPtr a, real1 log_b, TensorPtr out);
  static TensorPtr log(TensorPtr a, real1 b = E_R1);
  static void make_log_node(TensorPtr a, real1 inv_log_b, TensorPtr out);

  // device pointer of element 0 of the underlying storage (GPU tensors only)
  real1 *device_ptr() const;          // read/write access (bumps the storage version)
  const real1 *device_ptr_ro() const; // read-only access
  // gradient destination: accumulate = 0 (and no zero-fill is issued) when the storage is lazily
  // zeroed and this view covers all of it, so the kernel may store instead of add
  real1 *device_ptr_accumulate(int &accumulate) const;
  // as device_ptr_accumulate for kernels that can read the old values from one buffer and write the sums to another:
  // a copy-on-write shared gradient is then accumulated into out of place (src = the shared buffer, kept alive by `keep`)
  // instead of being copied first. src == nullptr: nothing to add to (plain store).
  real1 *device_ptr_accumulate_from(const real1 *&src, BufferPtr &keep) const;
  void *stream() const;
};

inline TensorPtr operator+(TensorPtr l, TensorPtr r) { return Tensor::add(l, r); }
inline TensorPtr operator+(real1 l, TensorPtr r) { return Tensor::add(SCALAR(l, r), r); }
inline TensorPtr operator+(TensorPtr l, real1 r) { return r + l; }
inline TensorPtr operator-(TensorPtr l, TensorPtr r) { return Tensor::sub(l, r); }
inline TensorPtr operator-(real1 l, TensorPtr r) { return Tensor::sub(SCALAR(l, r), r); }
inline TensorPtr operator-(TensorPtr l, real1 r) { return Tensor::sub(l, SCALAR(r, l)); }
inline TensorPtr operator*(TensorPtr l, TensorPtr r) { return Tensor::mul(l, r); }
inline TensorPtr operator*(real1 l, TensorPtr r) { return Tensor::mul(SCALAR(l, r), r); }
inline TensorPtr operator*(TensorPtr l, real1 r) { return r * l; }
inline TensorPtr operator/(TensorPtr l, TensorPtr r) { return Tensor::div(l, r); }
inline TensorPtr operator/(real1 l, TensorPtr r) { return Tensor::div(SCALAR(l, r), r); }
inline TensorPtr operator/(TensorPtr l, real1 r) { return Tensor::div(l, SCALAR(r, l)); }
inline TensorPtr operator>>(TensorPtr l, TensorPtr r) { return Tensor::matmul(l, r); }
inline TensorPtr operator<<(TensorPtr r, TensorPtr l) { return Tensor::matmul(l, r); }
inline TensorPtr operator^(TensorPtr base, real1 power) { return Tensor::pow(base, power); }
inline TensorPtr operator^(real1 base, TensorPtr power) { return Tensor::exp(power, base); }

// reference include/autograd/node.hpp:22-41
struct Node {
  std::vector<TensorPtr> parents;
  std::function<void()> backward;
  Node(const std::vector<TensorPtr> &p, const std::function<void()> &b) : parents(p), backward(b) {
    for (auto &t : parents) {
      t->make_gradient();
      ++t->consumers;
      // a view of a graph tensor shares that tensor's gradient: a contribution through the view is one for the owner too
      if (t->view_owner) ++t->view_owner->consumers;
    }
  }
};

struct Parameter;
typedef std::shared_ptr<Parameter> ParameterPtr;
struct Parameter : Tensor {
  Parameter(const std::vector<tcapint> &shp, const std::vector<tcapint> &str, const bool &s = true, const DType &dtype = DType::REAL,
            const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE, const int64_t &did = -1)
      : Tensor(shp, str, true, s, DType::REAL, dtag, did) {}
  Parameter(const std::vector<real1> &val, const std::vector<tcapint> &shp, const DeviceTag &dtag = DeviceTag::DEFAULT_DEVICE,
            const int64_t &did = -1)
      : Tensor(val, shp, true, dtag, did) {}
  void train() { requires_grad = true; }
  void eval() {
    requires_grad = false;
    grad = nullptr;
  }
  void save(std::ostream &out);
  static ParameterPtr load(std::istream &in);
};

// flat accessors (reference include/tensors/real_tensor.hpp, real_scalar.hpp). On a GPU tensor each
// element access is a blocking 4-byte read, exactly like the reference's GpuRealStorage::operator[].
struct RealTensor : public Tensor {
  RealTensor(const Tensor &orig) : Tensor(orig) {
    if (storage->dtype != DType::REAL) throw std::domain_error("RealTensor constructor must copy from a real-valued generic Tensor!");
  }
  real1 operator[](const tcapint &idx) const { return (*static_cast<RealStorage *>(storage.get()))[get_storage_index(idx)]; }
};
struct Scalar : public Tensor {
  Scalar(const real1 &v, const bool &rg = false, DeviceTag dtag = DeviceTag::DEFAULT_DEVICE, int64_t did = -1) : Tensor(v, rg, dtag, did) {}
};
struct RealScalar : public Scalar {
  RealScalar(const real1 &v, const bool &rg = false, DeviceTag dtag = DeviceTag::DEFAULT_DEVICE, int64_t did = -1) : Scalar(v, rg, dtag, did) {}
  real1 get_item() const { return (*static_cast<RealStorage *>(storage.get()))[offset]; }
};
typedef std::shared_ptr<RealScalar> RealScalarPtr;

// Whole-tensor read-back helpers (one blocking copy instead of per-element reads)
std::vector<real1> to_host(const Tensor &t);             // storage contents, storage order
std::vector<real1> to_host_logical(const Tensor &t);     // flat logical (column-major) order through the view
} // namespace Weed
