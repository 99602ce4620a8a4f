// caffemodel.cu -- C ABI of the weight on-disk path: `.caffemodel` -> per-layer blobs -> (the
// caller's) WeightAlign, and a magnitude pruning step.  Host code only; the parser is host/escort_caffemodel.hpp.
// Replaces, for this path, Net::CopyTrainedLayersFrom(const string) (src/caffe/net.cpp:785-821, ReadNetParamsFromBinaryFileOrDie
// + Blob::FromProto) -- protobuf itself is not in this image.
#include <algorithm>
#include <cmath>
#include <cstring>
#include <vector>

#include "../host/escort_caffemodel.hpp"
#include "common.cuh"

struct escort_caffemodel {
  escort_host::CaffeModel m;
};

using namespace escort;

extern "C" ESCORT_API int escort_caffemodel_open(const char *path, escort_caffemodel **out) {
  ESCORT_REQUIRE(path && out, "escort_caffemodel_open: null argument");
  escort_caffemodel *h = new escort_caffemodel();
  if (!escort_host::ReadCaffeModel(path, h->m)) {
    set_last_error(("escort_caffemodel_open: " + h->m.error).c_str());
    delete h;
    *out = nullptr;
    return ESCORT_EINVAL;
  }
  *out = h;
  return 0;
}

extern "C" ESCORT_API int escort_caffemodel_close(escort_caffemodel *m) {
  delete m;
  return 0;
}

extern "C" ESCORT_API int escort_caffemodel_save(const escort_caffemodel *m, const char *path) {
  ESCORT_REQUIRE(m && path, "escort_caffemodel_save: null argument");
  if (!escort_host::WriteCaffeModel(path, m->m)) {
    set_last_error("escort_caffemodel_save: cannot write the file");
    return ESCORT_EINVAL;
  }
  return 0;
}

extern "C" ESCORT_API int escort_caffemodel_num_layers(const escort_caffemodel *m) { return m ? (int)m->m.layers.size() : ESCORT_EINVAL; }

extern "C" ESCORT_API int escort_caffemodel_find(const escort_caffemodel *m, const char *layer_name) {
  if (!m || !layer_name) return ESCORT_EINVAL;
  for (size_t i = 0; i < m->m.layers.size(); ++i)   // the name match of Net::CopyTrainedLayersFrom, net.cpp:790-796
    if (m->m.layers[i].name == layer_name) return (int)i;
  return ESCORT_EINVAL;
}

extern "C" ESCORT_API int escort_caffemodel_layer(const escort_caffemodel *m, int layer, escort_layer_info *info) {
  ESCORT_REQUIRE(m && info && layer >= 0 && layer < (int)m->m.layers.size(), "escort_caffemodel_layer: bad arguments");
  const escort_host::CmLayer &L = m->m.layers[layer];
  memset(info, 0, sizeof(*info));
  info->name = L.name.c_str();
  info->type = L.type.c_str();
  info->num_blobs = (int)L.blobs.size();
  info->is_conv = L.has_conv;
  info->is_inner_product = L.has_ip;
  info->num_output = L.num_output; info->bias_term = L.bias_term; info->group = L.group;
  info->kernel_h = L.kernel_h; info->kernel_w = L.kernel_w; info->stride_h = L.stride_h; info->stride_w = L.stride_w;
  info->pad_h = L.pad_h; info->pad_w = L.pad_w; info->dilation = L.dilation;
  return 0;
}

extern "C" ESCORT_API int escort_caffemodel_blob(escort_caffemodel *m, int layer, int blob, int *ndim, long *shape8, float **data_host,
                                                 long *count) {
  ESCORT_REQUIRE(m && layer >= 0 && layer < (int)m->m.layers.size(), "escort_caffemodel_blob: bad layer");
  escort_host::CmLayer &L = m->m.layers[layer];
  ESCORT_REQUIRE(blob >= 0 && blob < (int)L.blobs.size(), "escort_caffemodel_blob: bad blob index");
  escort_host::CmBlob &B = L.blobs[blob];
  ESCORT_REQUIRE(B.shape.size() <= 8, "escort_caffemodel_blob: more than 8 axes");
  if (ndim) *ndim = (int)B.shape.size();
  if (shape8) for (size_t i = 0; i < B.shape.size(); ++i) shape8[i] = B.shape[i];
  if (data_host) *data_host = B.data.data();   // owned by the model; writable (pruning), valid until close
  if (count) *count = (long)B.data.size();
  return 0;
}

// Magnitude pruning to a target sparsity: the `count * sparsity` smallest |w| become exactly 0 (ties at the threshold
// are kept), which is what the SkimCaffe checkpoints named in run.sh:13 look like to WeightAlign (it packs `!= 0`).
extern "C" ESCORT_API int escort_prune_magnitude(float *weights_host, long count, double sparsity, float *threshold_out, long *nnz_out) {
  ESCORT_REQUIRE(weights_host && count >= 0 && sparsity >= 0.0 && sparsity <= 1.0, "escort_prune_magnitude: bad arguments");
  const long k = (long)std::floor(sparsity * (double)count);
  float thr = 0.f;
  if (k > 0) {
    std::vector<float> mag((size_t)count);
    for (long i = 0; i < count; ++i) mag[i] = std::fabs(weights_host[i]);
    std::nth_element(mag.begin(), mag.begin() + (k - 1), mag.end());
    thr = mag[k - 1];   // the k-th smallest magnitude: everything strictly below it, and it, goes
    long removed = 0;
    for (long i = 0; i < count; ++i)
      if (std::fabs(weights_host[i]) < thr) { weights_host[i] = 0.f; ++removed; }
    for (long i = 0; i < count && removed < k; ++i)   // ties at the threshold, in index order, until k are gone
      if (weights_host[i] != 0.f && std::fabs(weights_host[i]) == thr) { weights_host[i] = 0.f; ++removed; }
  }
  long nnz = 0;
  for (long i = 0; i < count; ++i) nnz += weights_host[i] != 0.f;
  if (threshold_out) *threshold_out = thr;
  if (nnz_out) *nnz_out = nnz;
  return 0;
}
