// common.cuh — shared device/host helpers for the weedcu kernels (sm_100a only).
#pragma once
#include "weedcu.h"
#include <cuda_runtime.h>
#include <stdint.h>

#define WCU_CHECK(expr)                                                                            \
  do {                                                                                             \
    cudaError_t _e = (expr);                                                                       \
    if (_e != cudaSuccess) return (int)_e;                                                         \
  } while (0)

namespace weedcu {

constexpr int kMaxRank = WEEDCU_MAX_RANK;
constexpr int kNumSMs = 148; // B200: 2 dies x 74 SMs; grids are sized in multiples of this

cudaStream_t resolve_stream(void *s);
void count_launch(int n = 1);
int after_launch(); // cudaGetLastError -> return code, bumps the launch counter
// stream-ordered caching allocator (runtime.cu); kernels' workspaces come from it too
cudaError_t pool_alloc(void **ptr, size_t bytes, cudaStream_t st);
cudaError_t pool_free(void *ptr, cudaStream_t st);
// cudaFuncSetAttribute(MaxDynamicSharedMemorySize) once per kernel, not once per launch
void ensure_dynamic_smem(const void *kernel, int bytes);

// Programmatic dependent launch: every kernel of this library is launched with the programmatic-stream-
// serialization attribute and starts with pdl_grid_sync(): it waits until the PREVIOUS grid has completed
// and flushed (griddepcontrol.wait) before touching memory, and then lets the NEXT kernel of the stream be
// scheduled while this one runs (griddepcontrol.launch_dependents). Launch latency, block scheduling and per-kernel set-up (barrier init, TMEM allocation,
// tensor-map prefetch) thereby overlap the tail of the preceding kernel instead of following its
// completion; stream order of every memory access is unchanged. WEEDCU_PDL=0 launches plainly.
bool pdl_enabled();
// Only a kernel that directly follows another kernel OF THIS LIBRARY takes the attribute: after a
// memset / memcpy / event / allocation on the stream (note_stream_op) the next launch is a plain one,
// so its ordering against those operations is the ordinary stream order.
void note_stream_op();
bool pdl_take_edge(cudaStream_t st); // true when the previous operation of this library was one of our kernels on the SAME stream; marks "kernel" for the next
__device__ __forceinline__ void pdl_grid_sync() {
  // wait first, then release the dependents: a kernel's successor may be scheduled while it runs, but never while
  // it is itself still waiting (chains of not-yet-started kernels piling up behind one another)
  asm volatile("griddepcontrol.wait;" ::: "memory");
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
}
template <typename... KArgs, typename... Args>
static inline void launch_k(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args &&...args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute attr;
  attr.id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr.val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = &attr;
  cfg.numAttrs = pdl_take_edge(st) ? 1u : 0u;
  (void)cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...); // errors surface in after_launch()
}

// Optional per-class event timing (weedcu_prof_*). A scope brackets the launches issued while it
// is alive; when profiling is off it costs one predictable branch.
bool prof_on();
void note_kernel_class(int cls); // the class of the launches that follow (WEEDCU_PDL_CLASSES bisects by it)
int prof_begin(int cls, cudaStream_t st, double work);
void prof_end(int idx, cudaStream_t st);
struct ProfScope {
  int idx;
  cudaStream_t st;
  ProfScope(int cls, cudaStream_t s, double work) : idx(-1), st(s) {
    note_kernel_class(cls);
    if (prof_on()) idx = prof_begin(cls, s, work);
  }
  ~ProfScope() {
    if (idx >= 0) prof_end(idx, st);
  }
};

// ---------------------------------------------------------------------------------------------
// Collapsed multi-operand index space. All operands share `shape`; each has its own strides.
// Offsets are folded into the base pointers on the host.
template <int NOPS> struct IndexSpace {
  int rank;
  uint32_t n; // total elements (tcapint is uint32 in the reference build)
  uint32_t shape[kMaxRank];
  uint32_t stride[NOPS][kMaxRank];
};

// Merge adjacent dims that are jointly contiguous for every operand and drop extent-1 dims.
template <int NOPS>
static inline bool build_index_space(const weedcu_view *const *views, IndexSpace<NOPS> &sp) {
  const int rank = views[0]->rank;
  if (rank <= 0 || rank > kMaxRank) return false;
  for (int o = 1; o < NOPS; ++o) {
    if (views[o]->rank != rank) return false;
    for (int d = 0; d < rank; ++d)
      if (views[o]->shape[d] != views[0]->shape[d]) return false;
  }
  uint64_t n = 1;
  int r = 0;
  for (int d = 0; d < rank; ++d) {
    const uint32_t ext = views[0]->shape[d];
    if (ext == 0) return false;
    n *= ext;
    if (ext == 1) continue;
    bool merge = (r > 0);
    if (merge) {
      for (int o = 0; o < NOPS; ++o) {
        const uint32_t st = views[o]->stride[d];
        if ((uint64_t)sp.stride[o][r - 1] * sp.shape[r - 1] != st) { merge = false; break; }
      }
    }
    if (merge) {
      sp.shape[r - 1] *= ext;
    } else {
      sp.shape[r] = ext;
      for (int o = 0; o < NOPS; ++o) sp.stride[o][r] = views[o]->stride[d];
      ++r;
    }
  }
  if (n > 0xffffffffull) return false;
  if (r == 0) { // all extents 1
    sp.shape[0] = 1;
    for (int o = 0; o < NOPS; ++o) sp.stride[o][0] = 0;
    r = 1;
  }
  for (int d = r; d < kMaxRank; ++d) {
    sp.shape[d] = 1;
    for (int o = 0; o < NOPS; ++o) sp.stride[o][d] = 0;
  }
  sp.rank = r;
  sp.n = (uint32_t)n;
  return true;
}

// ---------------------------------------------------------------------------------------------
__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}
__device__ __forceinline__ float warp_max(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
  return v;
}
// block-wide sum; `red` is >= 32 floats of shared memory. Result valid in every thread.
__device__ __forceinline__ float block_sum(float v, float *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_sum(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : 0.0f;
  t = warp_sum(t);
  return t;
}
__device__ __forceinline__ float block_max(float v, float *red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  float t = (lane < nw) ? red[lane] : -INFINITY;
  t = warp_max(t);
  return t;
}

static inline unsigned grid_for(uint64_t work_items, unsigned block, unsigned max_waves = 16) {
  uint64_t g = (work_items + block - 1) / block;
  const uint64_t cap = (uint64_t)kNumSMs * max_waves;
  if (g > cap) g = cap;
  if (g == 0) g = 1;
  return (unsigned)g;
}

static inline bool aligned16(const void *p) { return (((uintptr_t)p) & 15u) == 0; }

} // namespace weedcu
