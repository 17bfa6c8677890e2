// Multi-GPU plumbing of the frame-sharded / CFG-split forward: a thin wrapper over NCCL, one communicator
// per exchange group.  The reference has no inference-time parallelism at all; its only collective use is accelerate's DDP
// in train.py.  NCCL is bound at run time (dlopen of libnccl.so.2 -- the instance PyTorch already loaded when the caller is
// a torch process), so the library itself has no link-time dependency on it and single-GPU users never touch it.
//
// Exchanges on the data path (all asynchronous on the caller's stream, CUDA-graph capturable):
//   * all_reduce_sum_f32: the (sum, sum of squares) of the 45 cross-frame GroupNorms of ResnetBlock3D / conv_norm_out
//     (resnet.py:166,177; unet.py:474) across the frame shards: 64 floats per batch element, latency-bound.
//   * all_to_all_rows: frames <-> pixels re-sharding around each motion module (motion_module.py:138-162): rank s holds
//     frames [s F/k, (s+1) F/k) x all pixels before, all F frames x pixels [s HW/k, (s+1) HW/k) after -- the temporal
//     attention couples frames per pixel, everything inside the module is per pixel.  Grouped ncclSend/ncclRecv.
//   * all_gather: the two 0.5 MB/k noise predictions of a CFG pair before the combine (pipeline_videoswap.py:578-580).
#include <dlfcn.h>
#include <nccl.h>     // types and enums only: every function is resolved with dlsym

#include "../../include/videoswap_b200.h"
#include "common.cuh"
#include "kernels.h"

namespace vs {
namespace {

struct NcclApi {
  void* handle = nullptr;
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  ncclResult_t (*AllReduce)(const void*, void*, size_t, ncclDataType_t, ncclRedOp_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Send)(const void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*Recv)(void*, size_t, ncclDataType_t, int, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*GroupStart)() = nullptr;
  ncclResult_t (*GroupEnd)() = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
};
NcclApi g_nccl;

int load_nccl() {
  if (g_nccl.handle) return 0;
  // RTLD_NOLOAD first: reuse the instance already mapped into the process (torch's), then a fresh load by soname
  void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_LOCAL);
  if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_LOCAL);
  VS_REQUIRE(h != nullptr, "multi-GPU mode needs NCCL: dlopen(libnccl.so.2) failed: %s", dlerror());
#define VS_SYM(field, name)                                                       \
  *reinterpret_cast<void**>(&g_nccl.field) = dlsym(h, name);                      \
  VS_REQUIRE(g_nccl.field != nullptr, "NCCL symbol %s not found", name)
  VS_SYM(GetUniqueId, "ncclGetUniqueId");
  VS_SYM(CommInitRank, "ncclCommInitRank");
  VS_SYM(CommDestroy, "ncclCommDestroy");
  VS_SYM(AllReduce, "ncclAllReduce");
  VS_SYM(AllGather, "ncclAllGather");
  VS_SYM(Send, "ncclSend");
  VS_SYM(Recv, "ncclRecv");
  VS_SYM(GroupStart, "ncclGroupStart");
  VS_SYM(GroupEnd, "ncclGroupEnd");
  VS_SYM(GetErrorString, "ncclGetErrorString");
#undef VS_SYM
  g_nccl.handle = h;
  return 0;
}

#define VS_CHECK_NCCL(expr)                                                                              \
  do {                                                                                                   \
    ncclResult_t _r = (expr);                                                                            \
    if (_r != ncclSuccess) {                                                                             \
      vs::set_error("%s:%d NCCL error %s: %s", __FILE__, __LINE__, #expr, g_nccl.GetErrorString(_r));   \
      return 1;                                                                                          \
    }                                                                                                    \
  } while (0)

}  // namespace
}  // namespace vs

struct vs_comm {
  ncclComm_t comm = nullptr;
  int rank = 0, nranks = 1;
};

namespace vs {

int comm_rank(const vs_comm* c) { return c ? c->rank : 0; }
int comm_size(const vs_comm* c) { return c ? c->nranks : 1; }

int comm_all_reduce_sum_f32(vs_comm* c, cudaStream_t st, float* buf, size_t n) {
  if (!c || c->nranks == 1) return 0;
  VS_CHECK_NCCL(g_nccl.AllReduce(buf, buf, n, ncclFloat32, ncclSum, c->comm, st));
  count_launch(1);
  return 0;
}

int comm_all_gather(vs_comm* c, cudaStream_t st, const void* send, void* recv, size_t bytes) {
  VS_REQUIRE(c != nullptr, "all_gather: null communicator");
  VS_CHECK_NCCL(g_nccl.AllGather(send, recv, bytes, ncclInt8, c->comm, st));
  count_launch(1);
  return 0;
}

// src: [n_outer][nranks][chunk] -> peer j receives src[:, j, :];  dst: [nranks][n_outer][chunk] with dst[j] = what peer j
// sent.  With (n_outer = local frames, chunk = HW/k * C) this is frames -> pixels; called with src/dst roles swapped
// ("gather" = 1: src is [nranks][n_outer][chunk], dst [n_outer][nranks][chunk]) it is the way back.
int comm_all_to_all_rows(vs_comm* c, cudaStream_t st, const __half* src, __half* dst, int n_outer, size_t chunk, int gather) {
  VS_REQUIRE(c != nullptr && c->nranks > 1, "all_to_all: needs a communicator with > 1 rank");
  const int k = c->nranks;
  VS_CHECK_NCCL(g_nccl.GroupStart());
  for (int j = 0; j < k; ++j) {
    for (int o = 0; o < n_outer; ++o) {
      const __half* s = gather ? src + ((size_t)j * n_outer + o) * chunk : src + ((size_t)o * k + j) * chunk;
      __half* d = gather ? dst + ((size_t)o * k + j) * chunk : dst + ((size_t)j * n_outer + o) * chunk;
      VS_CHECK_NCCL(g_nccl.Send(s, chunk, ncclFloat16, j, c->comm, st));
      VS_CHECK_NCCL(g_nccl.Recv(d, chunk, ncclFloat16, j, c->comm, st));
    }
  }
  VS_CHECK_NCCL(g_nccl.GroupEnd());
  count_launch(1);
  return 0;
}

}  // namespace vs

using namespace vs;

extern "C" int vs_comm_unique_id(void* out128) {
  VS_REQUIRE(out128 != nullptr, "vs_comm_unique_id: null output");
  if (int e = load_nccl()) return e;
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId is 128 bytes");
  VS_CHECK_NCCL(g_nccl.GetUniqueId(reinterpret_cast<ncclUniqueId*>(out128)));
  return 0;
}

extern "C" int vs_comm_create(const void* id128, int rank, int nranks, vs_comm** out) {
  VS_REQUIRE(id128 && out && nranks >= 1 && rank >= 0 && rank < nranks, "vs_comm_create: bad arguments");
  if (int e = load_nccl()) return e;
  ncclUniqueId id;
  memcpy(&id, id128, sizeof(id));
  vs_comm* c = new vs_comm();
  c->rank = rank; c->nranks = nranks;
  ncclResult_t r = g_nccl.CommInitRank(&c->comm, nranks, id, rank);
  if (r != ncclSuccess) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); delete c; return 1; }
  *out = c;
  return 0;
}

extern "C" void vs_comm_destroy(vs_comm* c) {
  if (!c) return;
  if (c->comm && g_nccl.CommDestroy) g_nccl.CommDestroy(c->comm);
  delete c;
}

extern "C" int vs_comm_all_gather(vs_comm* c, void* stream, const void* d_send, void* d_recv, size_t bytes_per_rank) {
  return comm_all_gather(c, (cudaStream_t)stream, d_send, d_recv, bytes_per_rank);
}

extern "C" int vs_comm_all_reduce_sum_f32(vs_comm* c, void* stream, float* d_buf, size_t n) {
  VS_REQUIRE(c != nullptr, "vs_comm_all_reduce_sum_f32: null communicator");
  return comm_all_reduce_sum_f32(c, (cudaStream_t)stream, d_buf, n);
}
