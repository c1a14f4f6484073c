// device_storage.cpp — CUDAEngine / GpuDevice / Storage / BaseTensor / SymbolTensor.
// Replaces reference src/common/oclengine.cpp, src/devices/gpu_device.cpp and the GPU half of
// src/storage/*.cpp: a device is a CUDA stream + the stream-ordered pool, a buffer is an owning
// device pointer. No per-launch host wait (the reference blocks on writeArgsEvent for every kernel,
// gpu_device.cpp:296-305); the only host syncs are explicit reads (operator[], cpu(), save()).
#include "weed_b200/core.hpp"

#include <algorithm>
#include <cstdlib>
#include <iostream>

namespace Weed {

BackendConfig &backend_config() {
  static BackendConfig cfg = [] {
    BackendConfig c;
    if (const char *e = getenv("WEED_B200_FUSED")) c.fused = atoi(e) != 0;
    if (const char *e = getenv("WEED_REF_INDEX_QUIRKS")) c.ref_index_quirks = atoi(e) != 0;
    if (const char *e = getenv("WEED_B200_MATMUL")) {
      const std::string s(e);
      c.matmul_precision = (s == "bf16") ? WEEDCU_GEMM_BF16 : WEEDCU_GEMM_FP32;
    }
    if (const char *e = getenv("WEED_B200_OPERAND_CACHE")) c.operand_cache = atoi(e) != 0;
    if (const char *e = getenv("WEED_B200_LAZY_ZERO")) c.lazy_zero = atoi(e) != 0;
    if (const char *e = getenv("WEED_B200_DEFER_GRADS")) c.defer_grads = atoi(e) != 0;
    if (const char *e = getenv("WEED_B200_COW_GRADS")) c.cow_grads = atoi(e) != 0;
    if (const char *e = getenv("WEED_B200_EPILOGUE")) c.epilogue_stats = atoi(e) != 0;
    return c;
  }();
  return cfg;
}

// Non-zero weedcu_* return -> the exception types the reference throws at the same points
// (bad_alloc: include/devices/pool_item.hpp:32-40; runtime_error: gpu_device.cpp:137-178).
void node_owner_lost() {
  throw std::logic_error("autograd: the output tensor of a graph node was destroyed before backward reached it (a view of it "
                         "outlived its owner without a shared_ptr; DESIGN.md defect D7)");
}
void throw_on_error(int rc, const char *what) {
  if (rc == 0) return;
  const std::string msg = std::string(what) + ": " + weedcu_error_string(rc);
  if (rc == WEEDCU_EINVAL) throw std::invalid_argument(msg);
  if (rc == 2 /* cudaErrorMemoryAllocation */) throw bad_alloc(msg);
  throw std::runtime_error(msg);
}

DeviceBuffer::~DeviceBuffer() {
  if (ptr) weedcu_free(ptr, stream);
}

// --------------------------------------------------------------------------------- GpuDevice
GpuDevice::GpuDevice(int64_t did) : deviceID(did), stream(nullptr) {
  Bind();
  throw_on_error(weedcu_stream_create(&stream), "GpuDevice stream");
  if (const char *e = getenv("WEED_MAX_ALLOC_MB")) maxAlloc = ((size_t)atoll(e)) << 20U; // oclengine.cpp:521-559
}
void GpuDevice::Bind() const {
  int cur = -1;
  weedcu_get_device(&cur);
  if (cur != (int)deviceID) throw_on_error(weedcu_set_device((int)deviceID), "cudaSetDevice");
}
BufferPtr GpuDevice::MakeBuffer(size_t bytes, const void *host_ptr) {
  Bind();
  void *p = nullptr;
  throw_on_error(weedcu_malloc(&p, bytes, stream), "GpuDevice::MakeBuffer");
  BufferPtr b = std::make_shared<DeviceBuffer>(p, bytes, stream);
  if (host_ptr) throw_on_error(weedcu_memcpy_h2d(p, host_ptr, bytes, stream), "GpuDevice::MakeBuffer upload");
  return b;
}
bool GpuDevice::LockSync(BufferPtr buffer, size_t bytes, void *dst, bool) {
  Bind();
  throw_on_error(weedcu_memcpy_d2h(dst, buffer->ptr, bytes, stream), "GpuDevice::LockSync");
  throw_on_error(weedcu_stream_sync(stream), "GpuDevice::LockSync");
  return false;
}
void GpuDevice::ClearRealBuffer(BufferPtr b, size_t n) { FillValueReal(b, n, ZERO_R1); }
void GpuDevice::FillOnesReal(BufferPtr b, size_t n) { FillValueReal(b, n, ONE_R1); }
void GpuDevice::FillValueReal(BufferPtr b, size_t n, real1 v) {
  Bind();
  throw_on_error(weedcu_fill_real((float *)b->ptr, n, v, stream), "FillValueReal");
}
void GpuDevice::ClearIntBuffer(BufferPtr b, size_t n) { FillValueInt(b, n, 0); }
void GpuDevice::FillOnesInt(BufferPtr b, size_t n) { FillValueInt(b, n, 1); }
void GpuDevice::FillValueInt(BufferPtr b, size_t n, symint v) {
  Bind();
  throw_on_error(weedcu_fill_int((int32_t *)b->ptr, n, v, stream), "FillValueInt");
}
real1 GpuDevice::GetReal(BufferPtr b, tcapint idx) {
  real1 v;
  Bind();
  throw_on_error(weedcu_memcpy_d2h(&v, (const float *)b->ptr + idx, sizeof(real1), stream), "GetReal");
  throw_on_error(weedcu_stream_sync(stream), "GetReal");
  return v;
}
void GpuDevice::SetReal(real1 v, BufferPtr b, tcapint idx) {
  Bind();
  throw_on_error(weedcu_memcpy_h2d((float *)b->ptr + idx, &v, sizeof(real1), stream), "SetReal");
}
symint GpuDevice::GetInt(BufferPtr b, tcapint idx) {
  symint v;
  Bind();
  throw_on_error(weedcu_memcpy_d2h(&v, (const int32_t *)b->ptr + idx, sizeof(symint), stream), "GetInt");
  throw_on_error(weedcu_stream_sync(stream), "GetInt");
  return v;
}
void GpuDevice::SetInt(symint v, BufferPtr b, tcapint idx) {
  Bind();
  throw_on_error(weedcu_memcpy_h2d((int32_t *)b->ptr + idx, &v, sizeof(symint), stream), "SetInt");
}
void GpuDevice::clFinish(bool) { throw_on_error(weedcu_stream_sync(stream), "clFinish"); }
void GpuDevice::AddAlloc(size_t sz) { // reference gpu_device.hpp:116-125
  std::lock_guard<std::mutex> lock(allocMutex);
  totalAlloc += sz;
  if (totalAlloc > maxAlloc) {
    totalAlloc -= sz;
    throw bad_alloc("VRAM limits exceeded in GpuDevice::AddAlloc()");
  }
}
void GpuDevice::SubtractAlloc(size_t sz) {
  std::lock_guard<std::mutex> lock(allocMutex);
  totalAlloc = (sz > totalAlloc) ? 0 : totalAlloc - sz;
}

// --------------------------------------------------------------------------------- CUDAEngine
CUDAEngine::CUDAEngine() {
  int n = 0;
  throw_on_error(weedcu_device_count(&n), "CUDAEngine: device discovery");
  if (n <= 0) throw std::runtime_error("CUDAEngine: no CUDA device (this backend has no CPU fallback)");
  devices.resize((size_t)n);
  int cur = 0;
  weedcu_get_device(&cur); // one process per GPU: the launcher picks the device with cudaSetDevice
  default_device = cur;
  if (const char *e = getenv("WEED_CUDA_DEFAULT_DEVICE")) default_device = atoll(e) % n; // cf. WEED_OCL_DEFAULT_DEVICE
}
CUDAEngine &CUDAEngine::Instance() {
  static CUDAEngine inst;
  return inst;
}
int CUDAEngine::GetDeviceCount() { return (int)devices.size(); }
void CUDAEngine::SetDefaultDeviceID(int64_t did) { default_device = did % (int64_t)devices.size(); }
GpuDevicePtr CUDAEngine::GetWeedDevice(int64_t did) {
  const int64_t n = (int64_t)devices.size();
  if (did < 0) did = default_device;
  did %= n; // ids wrap, reference oclengine.cpp:47-61
  std::lock_guard<std::mutex> lock(mtx);
  if (!devices[(size_t)did]) devices[(size_t)did] = std::make_shared<GpuDevice>(did);
  return devices[(size_t)did];
}
size_t CUDAEngine::GetActiveAllocSize(int64_t did) { return GetWeedDevice(did)->totalAlloc; }

// --------------------------------------------------------------------------------- Storage
StoragePtr Storage::Upcast(const DType &dt) {
  if (dt == DType::COMPLEX) throw std::invalid_argument("Complex storage is outside the CUDA backend's scope (SURVEY §8)");
  return get_ptr();
}

StoragePtr CpuRealStorage::gpu(const int64_t &did) { return std::make_shared<GpuRealStorage>(data, did); }
StoragePtr CpuIntStorage::gpu(const int64_t &did) { return std::make_shared<GpuIntStorage>(data, did); }
void GpuDevice::CopyBuffer(const BufferPtr &dst, const BufferPtr &src, size_t bytes) {
  Bind();
  throw_on_error(weedcu_memcpy_d2d(dst->ptr, src->ptr, bytes, stream), "GpuDevice::CopyBuffer");
}
void GpuRealStorage::FillValue(const real1 &v) {
  if (v == ZERO_R1 && zero_version == version && !zero_pending && !deferred_values) return; // already zero, nothing wrote since
  ++version;
  deferred_values = nullptr;
  if (v == ZERO_R1 && backend_config().fused && backend_config().lazy_zero) { // lazy: see GpuStorage::zero_pending
    zero_pending = true;
    return;
  }
  zero_pending = false;
  unshare(false);
  dev->FillValueReal(buffer, size, v);
}
StoragePtr GpuRealStorage::cpu() {
  materialize();
  CpuRealStoragePtr cp = std::make_shared<CpuRealStorage>(size);
  dev->LockSync(buffer, sizeof(real1) * (size_t)size, cp->data.data(), false);
  return cp;
}
StoragePtr GpuIntStorage::cpu() {
  CpuIntStoragePtr cp = std::make_shared<CpuIntStorage>(size);
  dev->LockSync(buffer, sizeof(symint) * (size_t)size, cp->data.data(), false);
  return cp;
}

// --------------------------------------------------------------------------------- BaseTensor
void BaseTensor::validate_constructor() {
  if (shape.size() != stride.size()) throw std::invalid_argument("Tensor shape vector must have same length as stride vector!");
  if ((shape.size() == 1U) && (shape[0U] == 1U)) stride[0U] = 0U;
}
tcapint BaseTensor::get_size() const {
  if (shape.empty()) return 0U;
  tcapint last = 0U;
  for (size_t i = 0U; i < shape.size(); ++i) last += (shape[i] - 1U) * stride[i];
  return last + 1U;
}
tcapint BaseTensor::get_broadcast_size() const {
  if (shape.empty()) return 0U;
  tcapint n = 1U;
  for (tcapint s : shape) n *= s;
  return n;
}
bool BaseTensor::is_scalar() const {
  if (shape.empty()) return false;
  for (size_t i = 0U; i < shape.size(); ++i)
    if (shape[i] != 1U && stride[i] != 0U) return false;
  return true;
}
tcapint BaseTensor::get_storage_index(const tcapint &idx) const {
  if (is_scalar()) return offset;
  tcapint rem = idx, at = offset;
  for (size_t i = 0U; (i < shape.size()) && rem; ++i) {
    at += (rem % shape[i]) * stride[i];
    rem /= shape[i];
  }
  if (rem) throw std::invalid_argument("Tensor index out-of-range!");
  return at;
}
void BaseTensor::reshape(const std::vector<symint> &s) {
  if (!is_contiguous(shape, stride)) throw std::domain_error("Can't reshape BaseTensor that isn't contiguous!");
  const tcapint total = get_size();
  std::vector<tcapint> dims(s.size());
  int infer = -1;
  tcapint known = 1U;
  for (size_t i = 0U; i < s.size(); ++i) {
    if (s[i] < 0) {
      if (infer >= 0) throw std::invalid_argument("Tensor::reshape(): only one -1 dimension allowed");
      infer = (int)i;
    } else {
      dims[i] = (tcapint)s[i];
      known *= (tcapint)s[i];
    }
  }
  if (infer >= 0) {
    if (!known || (total % known)) throw std::invalid_argument("Tensor::reshape(): cannot infer dimension size");
    dims[(size_t)infer] = total / known;
  }
  tcapint n = 1U;
  for (tcapint d : dims) n *= d;
  if (n != total) throw std::invalid_argument("Tensor::reshape(): sizes do not match");
  shape = dims;
  stride = full_contiguous_stride(dims);
}
void BaseTensor::transpose() {
  if (shape.size() > 2U) throw std::invalid_argument("Tensor::transpose is only for 2D tensors (and vectors and covectors)!");
  if (shape.size() == 1U) { // column vector -> row vector
    shape = {1U, shape[0U]};
    stride = {0U, stride[0U]};
  } else {
    std::swap(shape[0U], shape[1U]);
    std::swap(stride[0U], stride[1U]);
  }
}
void BaseTensor::transpose(symint i, symint j) {
  while (i < 0) i += (symint)shape.size();
  while (j < 0) j += (symint)shape.size();
  if (i != j) {
    std::swap(shape[(size_t)i], shape[(size_t)j]);
    std::swap(stride[(size_t)i], stride[(size_t)j]);
  }
}
void BaseTensor::flatten(symint axis) {
  while (axis < 0) axis += (symint)shape.size();
  if (axis < 1) throw std::invalid_argument("Can't flatten axis 0!");
  if ((tcapint)axis >= shape.size()) throw std::invalid_argument("Flatten axis is greater than highest index!");
  std::vector<symint> shp(shape.begin(), shape.end());
  shp[(size_t)axis - 1U] *= shp[(size_t)axis];
  shp.erase(shp.begin() + axis);
  reshape(shp);
}
bool BaseTensor::is_contiguous(const std::vector<tcapint> &shp, const std::vector<tcapint> &s) {
  tcapint expect = 1U;
  for (size_t i = 0U; i < s.size(); ++i) {
    if (!s[i]) continue;
    if (s[i] != expect) return false;
    expect *= shp[i];
  }
  return true;
}
std::vector<tcapint> BaseTensor::full_contiguous_stride(const std::vector<tcapint> &shp) {
  std::vector<tcapint> st(shp.size());
  tcapint acc = 1U;
  for (size_t i = 0U; i < shp.size(); ++i) {
    st[i] = (shp[i] == 1U) ? 0U : acc;
    acc *= shp[i];
  }
  return st;
}
DType BaseTensor::get_dtype_by_presidence(const std::vector<BaseTensorPtr> &) { return DType::REAL; }
// Reference: size-based (src/tensors/base_tensor.cpp:56-75). Here the device path never bounces
// to the host: everything runs where the GPU backend lives.
DeviceTag BaseTensor::get_dtag_by_presidence(const std::vector<BaseTensorPtr> &) { return DeviceTag::GPU; }

weedcu_view BaseTensor::view() const {
  if (shape.size() > WEEDCU_MAX_RANK) throw std::invalid_argument("Tensor rank exceeds WEEDCU_MAX_RANK");
  weedcu_view v;
  v.offset = offset;
  v.rank = (int32_t)shape.size();
  for (int d = 0; d < WEEDCU_MAX_RANK; ++d) {
    v.shape[d] = d < v.rank ? shape[(size_t)d] : 1U;
    v.stride[d] = d < v.rank ? stride[(size_t)d] : 0U;
  }
  return v;
}

// --------------------------------------------------------------------------------- SymbolTensor
static DeviceTag resolve_tag(DeviceTag t) { return (t == DeviceTag::CPU) ? DeviceTag::CPU : DeviceTag::GPU; }

SymbolTensor::SymbolTensor(const std::vector<tcapint> &shp, const std::vector<tcapint> &, const bool &, const DeviceTag &dtag,
                           const int64_t &did, const bool &)
    : BaseTensor(shp, full_contiguous_stride(shp)) {
  const tcapint n = get_size();
  if (resolve_tag(dtag) == DeviceTag::GPU) storage = std::make_shared<GpuIntStorage>(n, did);
  else storage = std::make_shared<CpuIntStorage>(n);
  storage->FillZeros();
}
SymbolTensor::SymbolTensor(const std::vector<symint> &val, const std::vector<tcapint> &shp, const bool &, const DeviceTag &dtag,
                           const int64_t &did)
    : BaseTensor(shp, full_contiguous_stride(shp)) {
  if (get_size() != val.size())
    throw std::invalid_argument("Tensor value initializer vector must have same size as implied by shape and stride!");
  if (resolve_tag(dtag) == DeviceTag::GPU) storage = std::make_shared<GpuIntStorage>(val, did);
  else storage = std::make_shared<CpuIntStorage>(val);
}
SymbolTensorPtr SymbolTensor::cast(const DeviceTag &dt) const {
  SymbolTensorPtr cp = std::make_shared<SymbolTensor>(*this);
  if (dt == DeviceTag::CPU) cp->storage = cp->storage->cpu();
  else if (dt == DeviceTag::GPU) cp->storage = cp->storage->gpu();
  return cp;
}
SymbolTensorPtr SymbolTensor::reshape(const SymbolTensorPtr a, const std::vector<symint> &s) {
  SymbolTensorPtr out = std::make_shared<SymbolTensor>(*a);
  out->reshape(s);
  return out;
}
SymbolTensorPtr SymbolTensor::transpose(const SymbolTensorPtr a) {
  SymbolTensorPtr out = std::make_shared<SymbolTensor>(*a);
  out->transpose();
  return out;
}
SymbolTensorPtr SymbolTensor::transpose(const SymbolTensorPtr a, symint i, symint j) {
  SymbolTensorPtr out = std::make_shared<SymbolTensor>(*a);
  out->transpose(i, j);
  return out;
}
SymbolTensorPtr SymbolTensor::flatten(const SymbolTensorPtr a, const symint &axis) {
  SymbolTensorPtr out = std::make_shared<SymbolTensor>(*a);
  out->flatten(axis);
  return out;
}
const symint *SymbolTensor::device_ptr() const {
  if (!storage || storage->device != DeviceTag::GPU) throw std::domain_error("SymbolTensor is not GPU-resident");
  return static_cast<GpuIntStorage *>(storage.get())->device_ptr();
}

// --------------------------------------------------------------------------------- read-back
std::vector<real1> to_host(const Tensor &t) {
  StoragePtr s = t.storage->cpu();
  return static_cast<CpuRealStorage *>(s.get())->data;
}
std::vector<real1> to_host_logical(const Tensor &t) {
  const std::vector<real1> raw = to_host(t);
  const tcapint n = t.get_broadcast_size();
  std::vector<real1> out(n);
  for (tcapint i = 0U; i < n; ++i) out[i] = raw[t.get_storage_index(i)];
  return out;
}
} // namespace Weed
