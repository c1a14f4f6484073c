// runtime.cu — device, stream, event, memory-pool and copy entry points of weedcu.h.
// Replaces OCLEngine device discovery (reference include/common/oclengine.hpp:249-395) and the
// buffer / queue half of GpuDevice (reference src/devices/gpu_device.cpp:34-76,250-312,388-447).
// One in-order compute stream per device gives the same ordering as the reference's FIFO of
// QueueItems (gpu_device.cpp:180-248) without its per-launch host wait (gpu_device.cpp:296-305).
#include "common.cuh"
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <mutex>
#include <unordered_map>
#include <vector>

namespace weedcu {

static std::atomic<uint64_t> g_launches{0};
static std::mutex g_mutex;
static cudaStream_t g_default_stream[64] = {nullptr};
static bool g_default_owned[64] = {false};
static bool g_pool_configured[64] = {false};

void count_launch(int n) { g_launches.fetch_add((uint64_t)n, std::memory_order_relaxed); }

int after_launch() {
  count_launch(1);
  return (int)cudaGetLastError();
}

struct ProfRec {
  int cls;
  cudaEvent_t e0, e1;
  double work;
};
static bool g_prof_on = false;
static std::vector<ProfRec> g_prof;
bool prof_on() { return g_prof_on; }
int prof_begin(int cls, cudaStream_t st, double work) {
  ProfRec r;
  r.cls = cls;
  r.work = work;
  if (cudaEventCreate(&r.e0) != cudaSuccess || cudaEventCreate(&r.e1) != cudaSuccess) return -1;
  note_stream_op();
  cudaEventRecord(r.e0, st);
  g_prof.push_back(r);
  return (int)g_prof.size() - 1;
}
void prof_end(int idx, cudaStream_t st) {
  note_stream_op();
  if (idx >= 0 && idx < (int)g_prof.size()) cudaEventRecord(g_prof[(size_t)idx].e1, st);
}

static int current_device() {
  int d = 0;
  cudaGetDevice(&d);
  return (d >= 0 && d < 64) ? d : 0;
}

static void configure_pool(int dev) {
  if (g_pool_configured[dev]) return;
  cudaMemPool_t pool;
  if (cudaDeviceGetDefaultMemPool(&pool, dev) == cudaSuccess) {
    // Keep freed blocks cached in the pool: allocation in the autograd hot loop must not
    // touch the driver (180 GB of HBM3e; the host tracks its own budget).
    uint64_t threshold = UINT64_MAX;
    cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &threshold);
  }
  g_pool_configured[dev] = true;
}

cudaStream_t resolve_stream(void *s) {
  if (s) return (cudaStream_t)s;
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_mutex);
  if (!g_default_stream[dev]) {
    cudaStream_t st;
    if (cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking) != cudaSuccess) return nullptr;
    g_default_stream[dev] = st;
    g_default_owned[dev] = true;
    configure_pool(dev);
  }
  return g_default_stream[dev];
}

} // namespace weedcu

using namespace weedcu;

extern "C" {

int weedcu_device_count(int *count) {
  if (!count) return WEEDCU_EINVAL;
  WCU_CHECK(cudaGetDeviceCount(count));
  return 0;
}
int weedcu_set_device(int device) {
  WCU_CHECK(cudaSetDevice(device));
  return 0;
}
int weedcu_get_device(int *device) {
  if (!device) return WEEDCU_EINVAL;
  WCU_CHECK(cudaGetDevice(device));
  return 0;
}
int weedcu_device_info(int device, char *name, int name_len, uint64_t *total_mem, int *sm_count,
                       int *cc_major, int *cc_minor) {
  cudaDeviceProp p;
  WCU_CHECK(cudaGetDeviceProperties(&p, device));
  if (name && name_len > 0) {
    strncpy(name, p.name, (size_t)name_len - 1);
    name[name_len - 1] = 0;
  }
  if (total_mem) *total_mem = (uint64_t)p.totalGlobalMem;
  if (sm_count) *sm_count = p.multiProcessorCount;
  if (cc_major) *cc_major = p.major;
  if (cc_minor) *cc_minor = p.minor;
  return 0;
}
const char *weedcu_error_string(int code) {
  if (code == 0) return "ok";
  if (code == WEEDCU_EINVAL) return "weedcu: invalid argument";
  if (code == WEEDCU_ENOSUP) return "weedcu: unsupported configuration";
  if (code == WEEDCU_ENCCL) return "weedcu: NCCL not loaded";
  if (code >= 1000) return "weedcu: NCCL error";
  return cudaGetErrorString((cudaError_t)code);
}
void *weedcu_default_stream(void) { return (void *)resolve_stream(nullptr); }
int weedcu_set_default_stream(void *stream) {
  const int dev = current_device();
  std::lock_guard<std::mutex> lock(g_mutex);
  if (g_default_stream[dev] && g_default_owned[dev]) cudaStreamDestroy(g_default_stream[dev]);
  g_default_stream[dev] = (cudaStream_t)stream;
  g_default_owned[dev] = false;
  configure_pool(dev);
  return 0;
}
int weedcu_stream_create(void **stream) {
  if (!stream) return WEEDCU_EINVAL;
  cudaStream_t st;
  WCU_CHECK(cudaStreamCreateWithFlags(&st, cudaStreamNonBlocking));
  configure_pool(current_device());
  *stream = (void *)st;
  return 0;
}
int weedcu_stream_create_priority(void **stream, int high) {
  if (!stream) return WEEDCU_EINVAL;
  int least = 0, greatest = 0;
  WCU_CHECK(cudaDeviceGetStreamPriorityRange(&least, &greatest)); // numerically lower = higher priority
  cudaStream_t st;
  WCU_CHECK(cudaStreamCreateWithPriority(&st, cudaStreamNonBlocking, high ? greatest : least));
  configure_pool(current_device());
  *stream = (void *)st;
  return 0;
}
int weedcu_stream_destroy(void *stream) {
  WCU_CHECK(cudaStreamDestroy((cudaStream_t)stream));
  return 0;
}
int weedcu_stream_sync(void *stream) {
  WCU_CHECK(cudaStreamSynchronize(resolve_stream(stream)));
  return 0;
}
int weedcu_stream_wait_event(void *stream, void *event) {
  note_stream_op();
  WCU_CHECK(cudaStreamWaitEvent(resolve_stream(stream), (cudaEvent_t)event, 0));
  return 0;
}
int weedcu_event_create(void **event) {
  if (!event) return WEEDCU_EINVAL;
  cudaEvent_t e;
  WCU_CHECK(cudaEventCreate(&e));
  *event = (void *)e;
  return 0;
}
int weedcu_event_destroy(void *event) {
  WCU_CHECK(cudaEventDestroy((cudaEvent_t)event));
  return 0;
}
int weedcu_event_record(void *event, void *stream) {
  note_stream_op();
  WCU_CHECK(cudaEventRecord((cudaEvent_t)event, resolve_stream(stream)));
  return 0;
}
int weedcu_event_sync(void *event) {
  WCU_CHECK(cudaEventSynchronize((cudaEvent_t)event));
  return 0;
}
int weedcu_event_elapsed_ms(void *start, void *stop, float *ms) {
  if (!ms) return WEEDCU_EINVAL;
  WCU_CHECK(cudaEventElapsedTime(ms, (cudaEvent_t)start, (cudaEvent_t)stop));
  return 0;
}
static double g_malloc_ms = 0.0, g_free_ms = 0.0;
static uint64_t g_mallocs = 0, g_frees = 0;
static inline double now_ms() {
  return std::chrono::duration<double, std::milli>(std::chrono::steady_clock::now().time_since_epoch()).count();
}
} // extern "C"

// Stream-ordered caching allocator in front of cudaMallocAsync. A training step allocates the same
// few hundred sizes every iteration; cudaMallocAsync costs ~10 us of host time per call on this
// path, a free-list hit costs ~0.1 us. A block freed on stream S is handed out again only for
// stream S, so reuse is ordered after every kernel that was enqueued before the free (the same
// guarantee cudaFreeAsync gives). WEEDCU_POOL=0 bypasses the cache.
namespace weedcu {
namespace {
struct PoolKey {
  cudaStream_t stream;
  size_t bytes;
  bool operator==(const PoolKey &o) const { return stream == o.stream && bytes == o.bytes; }
};
struct PoolKeyHash {
  size_t operator()(const PoolKey &k) const { return std::hash<size_t>()(k.bytes) ^ (std::hash<void *>()((void *)k.stream) << 1); }
};
std::mutex g_pool_mutex;
std::unordered_map<PoolKey, std::vector<void *>, PoolKeyHash> g_pool_free;
std::unordered_map<void *, size_t> g_pool_live; // ptr -> rounded size
const bool g_pool_on = [] {
  const char *e = getenv("WEEDCU_POOL");
  return !(e && atoi(e) == 0);
}();
inline size_t pool_round(size_t bytes) {
  if (bytes < 512) return 512;
  if (bytes <= (1u << 20)) return (bytes + 511) & ~(size_t)511;
  return (bytes + ((1u << 16) - 1)) & ~(size_t)((1u << 16) - 1);
}
void pool_release_cached_locked() {
  for (auto &kv : g_pool_free)
    for (void *p : kv.second) cudaFreeAsync(p, kv.first.stream);
  g_pool_free.clear();
}
} // namespace

static int g_pdl_on = -1;
bool pdl_enabled() {
  int &on = g_pdl_on;
  if (on < 0) {
    // on by default (WEEDCU_PDL=0 launches plainly). History: with launch_dependents issued BEFORE griddepcontrol.wait,
    // not-yet-started kernels piled up behind one another (a chain fill -> cross-entropy backward read a stale value,
    // one run did not finish); waiting first fixed both
    const char *e = getenv("WEEDCU_PDL");
    on = (e && e[0] == '0') ? 0 : 1;
  }
  return on != 0;
}

// One issuing thread per process (the reference's model): plain globals. The edge is only taken when the
// previous launch of this library went to the same stream, so a second stream (the communication stream, another
// device's compute stream) never inherits a chain it is not part of.
static bool g_prev_is_kernel = false;
static cudaStream_t g_prev_stream = nullptr;
static int g_kernel_class = 0;
void note_stream_op() { g_prev_is_kernel = false; }
void note_kernel_class(int cls) { g_kernel_class = cls; }
bool pdl_take_edge(cudaStream_t st) {
  // WEEDCU_PDL_CLASSES: bit c set = launches of profiling class c (weedcu.h WEEDCU_PROF_*) may take the edge
  static long mask = -2;
  if (mask == -2) {
    const char *e = getenv("WEEDCU_PDL_CLASSES");
    mask = e ? strtol(e, nullptr, 0) : -1;
  }
  // WEEDCU_PDL_EDGE="p,c": only launches of class c that directly follow a launch of class p take the edge (bisecting)
  static int edge_p = -2, edge_c = -2;
  static int prev_class = 0;
  if (edge_p == -2) {
    const char *e = getenv("WEEDCU_PDL_EDGE");
    edge_p = edge_c = -1;
    if (e) sscanf(e, "%d,%d", &edge_p, &edge_c);
  }
  bool take = g_prev_is_kernel && g_prev_stream == st && pdl_enabled() && ((mask >> (g_kernel_class & 31)) & 1L);
  g_prev_stream = st;
  if (edge_p >= 0 && !(prev_class == edge_p && g_kernel_class == edge_c)) take = false;
  prev_class = g_kernel_class;
  g_prev_is_kernel = true;
  return take;
}

void ensure_dynamic_smem(const void *kernel, int bytes) {
  static std::mutex m;
  static std::unordered_map<const void *, int> done;
  std::lock_guard<std::mutex> lock(m);
  auto it = done.find(kernel);
  if (it != done.end() && it->second >= bytes) return;
  cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes);
  done[kernel] = bytes;
}

cudaError_t pool_alloc(void **ptr, size_t bytes, cudaStream_t st) {
  if (!g_pool_on) { note_stream_op(); return cudaMallocAsync(ptr, bytes ? bytes : 16, st); }
  const size_t sz = pool_round(bytes);
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  auto it = g_pool_free.find(PoolKey{st, sz});
  if (it != g_pool_free.end() && !it->second.empty()) {
    *ptr = it->second.back();
    it->second.pop_back();
    g_pool_live[*ptr] = sz;
    return cudaSuccess;
  }
  note_stream_op();
  cudaError_t e = cudaMallocAsync(ptr, sz, st);
  if (e == cudaErrorMemoryAllocation) { // give the cached blocks back to the driver and retry once
    (void)cudaGetLastError();
    pool_release_cached_locked();
    cudaStreamSynchronize(st);
    e = cudaMallocAsync(ptr, sz, st);
  }
  if (e == cudaSuccess) g_pool_live[*ptr] = sz;
  return e;
}
cudaError_t pool_free(void *ptr, cudaStream_t st) {
  if (!ptr) return cudaSuccess;
  if (!g_pool_on) { note_stream_op(); return cudaFreeAsync(ptr, st); }
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  auto it = g_pool_live.find(ptr);
  if (it == g_pool_live.end()) { note_stream_op(); return cudaFreeAsync(ptr, st); } // not ours
  g_pool_free[PoolKey{st, it->second}].push_back(ptr);
  g_pool_live.erase(it);
  return cudaSuccess;
}
} // namespace weedcu

extern "C" {
int weedcu_malloc(void **ptr, size_t bytes, void *stream) {
  if (!ptr) return WEEDCU_EINVAL;
  const double t0 = now_ms();
  const cudaError_t e = pool_alloc(ptr, bytes, resolve_stream(stream));
  g_malloc_ms += now_ms() - t0;
  ++g_mallocs;
  return (int)e;
}
int weedcu_free(void *ptr, void *stream) {
  if (!ptr) return 0;
  const double t0 = now_ms();
  const cudaError_t e = pool_free(ptr, resolve_stream(stream));
  g_free_ms += now_ms() - t0;
  ++g_frees;
  return (int)e;
}
/* hand every cached block back to the driver (between phases with different shapes, or tests) */
int weedcu_pool_trim(void) {
  std::lock_guard<std::mutex> lock(g_pool_mutex);
  pool_release_cached_locked();
  return 0;
}
// host-side time spent inside the pool allocator (diagnostics for launch-bound steps)
int weedcu_host_stats(double *malloc_ms, uint64_t *mallocs, double *free_ms, uint64_t *frees) {
  if (malloc_ms) *malloc_ms = g_malloc_ms;
  if (mallocs) *mallocs = g_mallocs;
  if (free_ms) *free_ms = g_free_ms;
  if (frees) *frees = g_frees;
  return 0;
}
int weedcu_mem_info(uint64_t *free_bytes, uint64_t *total_bytes) {
  size_t f = 0, t = 0;
  WCU_CHECK(cudaMemGetInfo(&f, &t));
  if (free_bytes) *free_bytes = f;
  if (total_bytes) *total_bytes = t;
  return 0;
}
int weedcu_host_alloc(void **ptr, size_t bytes) {
  if (!ptr) return WEEDCU_EINVAL;
  WCU_CHECK(cudaHostAlloc(ptr, bytes ? bytes : 16, cudaHostAllocDefault));
  return 0;
}
int weedcu_host_free(void *ptr) {
  if (!ptr) return 0;
  WCU_CHECK(cudaFreeHost(ptr));
  return 0;
}
int weedcu_memcpy_h2d(void *dst, const void *src, size_t bytes, void *stream) {
  if (!bytes) return 0;
  note_stream_op();
  WCU_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyHostToDevice, resolve_stream(stream)));
  return 0;
}
int weedcu_memcpy_d2h(void *dst, const void *src, size_t bytes, void *stream) {
  if (!bytes) return 0;
  note_stream_op();
  WCU_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToHost, resolve_stream(stream)));
  return 0;
}
int weedcu_memcpy_d2d(void *dst, const void *src, size_t bytes, void *stream) {
  if (!bytes) return 0;
  note_stream_op();
  WCU_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, resolve_stream(stream)));
  return 0;
}
int weedcu_prof_enable(int on) {
  if (on) {
    for (auto &p : g_prof) {
      cudaEventDestroy(p.e0);
      cudaEventDestroy(p.e1);
    }
    g_prof.clear();
  }
  g_prof_on = on != 0;
  return 0;
}
int weedcu_prof_read(int cls, double *total_ms, uint64_t *launches, double *work) {
  if (!total_ms || !launches || !work) return WEEDCU_EINVAL;
  *total_ms = 0.0;
  *launches = 0;
  *work = 0.0;
  for (auto &p : g_prof) {
    if (p.cls != cls) continue;
    WCU_CHECK(cudaEventSynchronize(p.e1));
    float ms = 0.0f;
    WCU_CHECK(cudaEventElapsedTime(&ms, p.e0, p.e1));
    *total_ms += ms;
    *launches += 1;
    *work += p.work;
  }
  return 0;
}
int weedcu_set_pdl(int on) {
  const int prev = pdl_enabled() ? 1 : 0;
  if (on >= 0) {
    g_pdl_on = on ? 1 : 0;
    note_stream_op(); // the next launch starts a fresh chain
  }
  return prev;
}
int weedcu_launch_count(uint64_t *count) {
  if (!count) return WEEDCU_EINVAL;
  *count = g_launches.load();
  return 0;
}

} // extern "C"
