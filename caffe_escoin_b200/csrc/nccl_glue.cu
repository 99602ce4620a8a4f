// nccl_glue.cu -- the one exchange step of the path: gradient all-reduce + 1/N scale over the flat diff
// buffer (reference NCCL<Dtype>::on_gradients_ready, src/caffe/parallel.cpp:238-256).  NCCL is resolved
// with dlopen at first use so the library has no link-time NCCL dependency and shares whatever NCCL the
// host process (e.g. torch's bundled libnccl.so.2) already loaded.
#include <dlfcn.h>

#include <cstring>
#include <mutex>

#include "common.cuh"

namespace escort {

__global__ void scale_kernel(float *__restrict__ x, size_t n, float s) {
  const size_t n4 = n / 4;
  float4 *x4 = reinterpret_cast<float4 *>(x);
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
    float4 v = x4[i];
    v.x *= s; v.y *= s; v.z *= s; v.w *= s;
    x4[i] = v;
  }
  for (size_t i = n4 * 4 + (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x)
    x[i] *= s;
}

__global__ void scale_kernel_unaligned(float *__restrict__ x, size_t n, float s) {
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) x[i] *= s;
}

struct nccl_unique_id { char internal[128]; };  // ncclUniqueId (nccl.h: NCCL_UNIQUE_ID_BYTES = 128), passed by value
typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_broadcast_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef int (*nccl_get_unique_id_fn)(nccl_unique_id *);
typedef int (*nccl_comm_init_rank_fn)(void **, int, nccl_unique_id, int);
typedef int (*nccl_comm_destroy_fn)(void *);
typedef const char *(*nccl_errstr_fn)(int);

struct NcclApi {
  nccl_allreduce_fn allreduce = nullptr;
  nccl_broadcast_fn broadcast = nullptr;
  nccl_get_unique_id_fn get_unique_id = nullptr;
  nccl_comm_init_rank_fn comm_init_rank = nullptr;
  nccl_comm_destroy_fn comm_destroy = nullptr;
  nccl_errstr_fn errstr = nullptr;
};

// resolved once per process (std::call_once: Caffe runs one host thread per GPU)
static const NcclApi *nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void *h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);  // the copy the host process already loaded, if any (same soname)
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) return;
    api.allreduce = (nccl_allreduce_fn)dlsym(h, "ncclAllReduce");
    api.broadcast = (nccl_broadcast_fn)dlsym(h, "ncclBroadcast");
    api.get_unique_id = (nccl_get_unique_id_fn)dlsym(h, "ncclGetUniqueId");
    api.comm_init_rank = (nccl_comm_init_rank_fn)dlsym(h, "ncclCommInitRank");
    api.comm_destroy = (nccl_comm_destroy_fn)dlsym(h, "ncclCommDestroy");
    api.errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
  });
  return &api;
}

static int nccl_fail(const char *what, int nrc) {
  const NcclApi *a = nccl_api();
  set_last_error(std::string(what) + " failed: " + (a->errstr ? a->errstr(nrc) : "?"));
  return ESCORT_ENCCL;
}
#define ESCORT_NCCL_NEED(fn, what)                                          \
  do {                                                                      \
    if (!nccl_api()->fn) {                                                  \
      set_last_error(std::string(what) + ": libnccl.so.2 not found");      \
      return ESCORT_ENCCL;                                                  \
    }                                                                       \
  } while (0)

}  // namespace escort

using namespace escort;

extern "C" int escort_allreduce_grads(void *comm, float *flat, size_t count, float scale, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(flat || count == 0, "escort_allreduce_grads: null buffer");
  if (count == 0) return 0;
  if (comm) {
    ESCORT_NCCL_NEED(allreduce, "escort_allreduce_grads");
    // ncclFloat32 = 7, ncclSum = 0 (nccl.h, stable since NCCL 2.0)
    const int nrc = nccl_api()->allreduce(flat, flat, count, 7, 0, comm, stream);
    if (nrc != 0) return nccl_fail("ncclAllReduce", nrc);
  }
  if (scale != 1.0f) {
    const int blocks = (int)std::min<size_t>((count / 4 + 255) / 256 + 1, 148 * 8);
    if ((reinterpret_cast<uintptr_t>(flat) & 15) == 0) scale_kernel<<<blocks, 256, 0, stream>>>(flat, count, scale);
    else scale_kernel_unaligned<<<blocks, 256, 0, stream>>>(flat, count, scale);  // a layer's slice of the flat buffer
    ESCORT_LAUNCH_CHECK();
  }
  return 0;
}

extern "C" int escort_broadcast(void *comm, float *buf, size_t count, int root, escort_stream_t stream_) {
  ESCORT_REQUIRE(comm && (buf || count == 0), "escort_broadcast: bad arguments");
  if (count == 0) return 0;
  ESCORT_NCCL_NEED(broadcast, "escort_broadcast");
  const int nrc = nccl_api()->broadcast(buf, buf, count, 7, root, comm, (cudaStream_t)stream_);
  return nrc == 0 ? 0 : nccl_fail("ncclBroadcast", nrc);
}

extern "C" int escort_comm_unique_id(void *id128) {
  ESCORT_REQUIRE(id128, "escort_comm_unique_id: null buffer");
  ESCORT_NCCL_NEED(get_unique_id, "escort_comm_unique_id");
  const int nrc = nccl_api()->get_unique_id(reinterpret_cast<nccl_unique_id *>(id128));
  return nrc == 0 ? 0 : nccl_fail("ncclGetUniqueId", nrc);
}

extern "C" int escort_comm_init_rank(void **comm_out, int nranks, const void *id128, int rank) {
  ESCORT_REQUIRE(comm_out && id128 && nranks > 0 && rank >= 0 && rank < nranks, "escort_comm_init_rank: bad arguments");
  ESCORT_NCCL_NEED(comm_init_rank, "escort_comm_init_rank");
  nccl_unique_id id;
  memcpy(&id, id128, sizeof id);
  const int nrc = nccl_api()->comm_init_rank(comm_out, nranks, id, rank);
  return nrc == 0 ? 0 : nccl_fail("ncclCommInitRank", nrc);
}

extern "C" int escort_comm_destroy(void *comm) {
  if (!comm) return 0;
  ESCORT_NCCL_NEED(comm_destroy, "escort_comm_destroy");
  const int nrc = nccl_api()->comm_destroy(comm);
  return nrc == 0 ? 0 : nccl_fail("ncclCommDestroy", nrc);
}
