// nccl_dp.cpp — data-parallel collectives for gradient averaging (no reference counterpart: Weed has
// no gradient exchange). Thin C-ABI over NCCL, resolved at run time with dlopen so the
// library loads on boxes without NCCL and the NCCL version is the one torch ships (2.28.x).
#include "weedcu.h"
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <stdint.h>
#include <string.h>

namespace weedcu {
cudaStream_t resolve_stream(void *s);
void count_launch(int n);
void note_stream_op(); // the next kernel of this library launches plainly (no programmatic-dependent-launch edge)
}

namespace {
struct UniqueId { char internal[128]; };
typedef int (*GetUniqueIdFn)(UniqueId *);
typedef int (*CommInitRankFn)(void **, int, UniqueId, int);
typedef int (*CommDestroyFn)(void *);
typedef int (*AllReduceFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*BroadcastFn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*GroupFn)();
void *g_lib = nullptr;
GetUniqueIdFn p_get_id = nullptr;
CommInitRankFn p_init = nullptr;
CommDestroyFn p_destroy = nullptr;
AllReduceFn p_allreduce = nullptr;
BroadcastFn p_bcast = nullptr;
GroupFn p_group_start = nullptr, p_group_end = nullptr;
constexpr int kNcclFloat = 7, kNcclSum = 0;
inline int wrap(int r) { return r == 0 ? 0 : 1000 + r; }
} // namespace

extern "C" {

int weedcu_nccl_load(const char *path) {
  if (g_lib) return 0;
  const char *candidates[] = {path, "libnccl.so.2", "libnccl.so", nullptr};
  for (int i = 0; i < 3 && !g_lib; ++i)
    if (candidates[i] && candidates[i][0]) g_lib = dlopen(candidates[i], RTLD_NOW | RTLD_GLOBAL);
  if (!g_lib) return WEEDCU_ENCCL;
  p_get_id = (GetUniqueIdFn)dlsym(g_lib, "ncclGetUniqueId");
  p_init = (CommInitRankFn)dlsym(g_lib, "ncclCommInitRank");
  p_destroy = (CommDestroyFn)dlsym(g_lib, "ncclCommDestroy");
  p_allreduce = (AllReduceFn)dlsym(g_lib, "ncclAllReduce");
  p_bcast = (BroadcastFn)dlsym(g_lib, "ncclBroadcast");
  p_group_start = (GroupFn)dlsym(g_lib, "ncclGroupStart");
  p_group_end = (GroupFn)dlsym(g_lib, "ncclGroupEnd");
  if (!p_get_id || !p_init || !p_destroy || !p_allreduce || !p_bcast) {
    g_lib = nullptr;
    return WEEDCU_ENCCL;
  }
  return 0;
}
int weedcu_nccl_unique_id(void *id128) {
  if (!g_lib) return WEEDCU_ENCCL;
  if (!id128) return WEEDCU_EINVAL;
  UniqueId id;
  const int r = p_get_id(&id);
  memcpy(id128, &id, 128);
  return wrap(r);
}
int weedcu_nccl_init(const void *id128, int rank, int world, void **comm) {
  if (!g_lib) return WEEDCU_ENCCL;
  if (!id128 || !comm) return WEEDCU_EINVAL;
  UniqueId id;
  memcpy(&id, id128, 128);
  return wrap(p_init(comm, world, id, rank));
}
int weedcu_nccl_destroy(void *comm) {
  if (!g_lib) return WEEDCU_ENCCL;
  return wrap(p_destroy(comm));
}
int weedcu_nccl_group_start(void) {
  if (!g_lib || !p_group_start) return WEEDCU_ENCCL;
  return wrap(p_group_start());
}
int weedcu_nccl_group_end(void) {
  if (!g_lib || !p_group_end) return WEEDCU_ENCCL;
  weedcu::note_stream_op();
  return wrap(p_group_end());
}
int weedcu_nccl_allreduce_sum(void *comm, float *buf, uint64_t n, void *stream) {
  if (!g_lib) return WEEDCU_ENCCL;
  weedcu::count_launch(1);
  weedcu::note_stream_op();
  return wrap(p_allreduce(buf, buf, (size_t)n, kNcclFloat, kNcclSum, comm, weedcu::resolve_stream(stream)));
}
int weedcu_nccl_broadcast(void *comm, float *buf, uint64_t n, int root, void *stream) {
  if (!g_lib) return WEEDCU_ENCCL;
  weedcu::count_launch(1);
  weedcu::note_stream_op();
  return wrap(p_bcast(buf, buf, (size_t)n, kNcclFloat, root, comm, weedcu::resolve_stream(stream)));
}

} // extern "C"
