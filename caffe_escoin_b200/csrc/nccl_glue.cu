// nccl_glue.cu -- the one exchange step of the path: gradient all-reduce + 1/N scale over the flat diff
// buffer (reference NCCL<Dtype>::on_gradients_ready, src/caffe/parallel.cpp:238-256).  NCCL is resolved
// with dlopen at first use so the library has no link-time NCCL dependency and shares whatever NCCL the
// host process (e.g. torch's bundled libnccl.so.2) already loaded.
#include <dlfcn.h>

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

typedef int (*nccl_allreduce_fn)(const void *, void *, size_t, int, int, void *, cudaStream_t);
typedef const char *(*nccl_errstr_fn)(int);
static nccl_allreduce_fn g_allreduce = nullptr;
static nccl_errstr_fn g_errstr = nullptr;

static int resolve_nccl() {
  if (g_allreduce) return 0;
  void *h = dlopen(nullptr, RTLD_NOW);  // already loaded by the host process?
  void *sym = h ? dlsym(h, "ncclAllReduce") : nullptr;
  if (!sym) {
    h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    sym = h ? dlsym(h, "ncclAllReduce") : nullptr;
  }
  if (!sym) {
    set_last_error("escort_allreduce_grads: libnccl.so.2 not found");
    return ESCORT_ENCCL;
  }
  g_allreduce = (nccl_allreduce_fn)sym;
  g_errstr = (nccl_errstr_fn)dlsym(h, "ncclGetErrorString");
  return 0;
}

}  // namespace escort

using namespace escort;

extern "C" int escort_allreduce_grads(void *comm, float *flat, size_t count, float scale, escort_stream_t stream_) {
  cudaStream_t stream = (cudaStream_t)stream_;
  ESCORT_REQUIRE(flat || count == 0, "escort_allreduce_grads: null buffer");
  if (count == 0) return 0;
  if (comm) {
    int rc = resolve_nccl();
    if (rc) return rc;
    // ncclFloat32 = 7, ncclSum = 0 (nccl.h, stable since NCCL 2.0)
    const int nrc = g_allreduce(flat, flat, count, 7, 0, comm, stream);
    if (nrc != 0) {
      set_last_error(std::string("ncclAllReduce failed: ") + (g_errstr ? g_errstr(nrc) : "?"));
      return ESCORT_ENCCL;
    }
  }
  if (scale != 1.0f) {
    ESCORT_REQUIRE((reinterpret_cast<uintptr_t>(flat) & 15) == 0, "escort_allreduce_grads: buffer must be 16-byte aligned");
    const int blocks = (int)std::min<size_t>((count / 4 + 255) / 256 + 1, 148 * 8);
    scale_kernel<<<blocks, 256, 0, stream>>>(flat, count, scale);
    ESCORT_LAUNCH_CHECK();
  }
  return 0;
}
