// escort_caffemodel.hpp -- the weight on-disk path of SURVEY section 8(f4): a `.caffemodel` (binary NetParameter) read
// into per-layer blobs, the way Net::CopyTrainedLayersFrom hands them to WeightAlign
// (src/caffe/net.cpp:785-821 -> Blob::FromProto, src/caffe/blob.cpp:466-520), and written back (pruning tool).
//
// protobuf is not in this image, so this is a reader / writer of the protobuf WIRE FORMAT for exactly the fields of
// src/caffe/proto/caffe.proto that path touches (field numbers cited at each use):
//   NetParameter   { name = 1, layers = 2 (V1LayerParameter), layer = 100 (LayerParameter) }
//   LayerParameter { name = 1, type = 2, blobs = 7, convolution_param = 106, inner_product_param = 117 }
//   V1LayerParameter { name = 4, type = 5 (enum), blobs = 6, convolution_param = 10, inner_product_param = 17 }
//   BlobProto      { num = 1, channels = 2, height = 3, width = 4, data = 5, diff = 6, shape = 7, double_data = 8 }
//   BlobShape      { dim = 1 }
//   ConvolutionParameter { num_output = 1, bias_term = 2, pad = 3, kernel_size = 4, group = 5, stride = 6, pad_h = 9,
//                          pad_w = 10, kernel_h = 11, kernel_w = 12, stride_h = 13, stride_w = 14, dilation = 18 }
//   InnerProductParameter { num_output = 1, bias_term = 2 }
// Every other field is kept as raw bytes and written back unchanged, so a model survives read -> prune -> write.
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

namespace escort_host {

struct CmBlob {
  std::vector<long> shape;    // Blob::FromProto: legacy (num, channels, height, width) if any of them is present, else shape.dim
  std::vector<float> data;    // `data`, or `double_data` narrowed to float (blob.cpp:490-500)
  bool legacy_dims = false;   // how the shape was stored (kept for the writer)
};

struct CmLayer {
  std::string name, type;     // V1 layers: type is the enum value as text ("V1:4" = CONVOLUTION)
  bool v1 = false;
  std::vector<CmBlob> blobs;
  // ConvolutionParameter / InnerProductParameter (0 / defaults when absent)
  bool has_conv = false, has_ip = false;
  int num_output = 0, bias_term = 1, group = 1;
  int kernel_h = 0, kernel_w = 0, stride_h = 1, stride_w = 1, pad_h = 0, pad_w = 0, dilation = 1;
  std::vector<std::pair<uint32_t, std::string>> other;  // (tag, raw bytes incl. length prefix where applicable) of untouched fields
};

struct CaffeModel {
  std::string name;
  std::vector<CmLayer> layers;
  std::vector<std::pair<uint32_t, std::string>> other;
  std::string error;
};

namespace wire {
struct Reader {
  const uint8_t *p, *end;
  bool ok = true;
  Reader(const void *b, size_t n) : p((const uint8_t *)b), end((const uint8_t *)b + n) {}
  bool more() const { return ok && p < end; }
  uint64_t varint() {
    uint64_t v = 0;
    for (int shift = 0; shift < 70; shift += 7) {
      if (p >= end) { ok = false; return 0; }
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7f) << (shift < 64 ? shift : 63);
      if (!(b & 0x80)) return v;
    }
    ok = false;
    return 0;
  }
  // one field: tag -> (field number, wire type); the payload of a length-delimited field as a sub-range
  bool field(uint32_t &num, int &wt, uint64_t &scalar, const uint8_t *&sub, size_t &sublen, const uint8_t *&raw_begin) {
    raw_begin = p;
    const uint64_t tag = varint();
    if (!ok) return false;
    num = (uint32_t)(tag >> 3);
    wt = (int)(tag & 7);
    scalar = 0; sub = nullptr; sublen = 0;
    switch (wt) {
      case 0: scalar = varint(); break;
      case 1: if (end - p < 8) { ok = false; break; } memcpy(&scalar, p, 8); p += 8; break;
      case 5: if (end - p < 4) { ok = false; break; } { uint32_t v; memcpy(&v, p, 4); scalar = v; } p += 4; break;
      case 2: {
        const uint64_t n = varint();
        if (!ok || (uint64_t)(end - p) < n) { ok = false; break; }
        sub = p; sublen = (size_t)n; p += n;
        break;
      }
      default: ok = false;  // groups (3 / 4) do not occur in caffe.proto
    }
    return ok;
  }
};
inline void put_varint(std::string &o, uint64_t v) {
  while (v >= 0x80) { o.push_back((char)(v | 0x80)); v >>= 7; }
  o.push_back((char)v);
}
inline void put_tag(std::string &o, uint32_t num, int wt) { put_varint(o, ((uint64_t)num << 3) | (uint64_t)wt); }
inline void put_bytes(std::string &o, uint32_t num, const std::string &b) { put_tag(o, num, 2); put_varint(o, b.size()); o += b; }
inline void put_uint(std::string &o, uint32_t num, uint64_t v) { put_tag(o, num, 0); put_varint(o, v); }
}  // namespace wire

inline bool parse_blob(const uint8_t *b, size_t n, CmBlob &out) {
  wire::Reader r(b, n);
  long legacy[4] = {0, 0, 0, 0};
  bool has_legacy = false;
  std::vector<double> dd;
  while (r.more()) {
    uint32_t num; int wt; uint64_t s; const uint8_t *sub; size_t sl; const uint8_t *raw;
    if (!r.field(num, wt, s, sub, sl, raw)) return false;
    if (num >= 1 && num <= 4 && wt == 0) { legacy[num - 1] = (long)(int32_t)s; has_legacy = true; }   // num, channels, height, width
    else if (num == 5) {                                                                             // repeated float data [packed]
      if (wt == 2) { const size_t k = sl / 4; const size_t o = out.data.size(); out.data.resize(o + k); memcpy(out.data.data() + o, sub, k * 4); }
      else if (wt == 5) { float f; uint32_t v = (uint32_t)s; memcpy(&f, &v, 4); out.data.push_back(f); }
    } else if (num == 8) {                                                                           // repeated double double_data [packed]
      if (wt == 2) { const size_t k = sl / 8; const size_t o = dd.size(); dd.resize(o + k); memcpy(dd.data() + o, sub, k * 8); }
      else if (wt == 1) { double d; memcpy(&d, &s, 8); dd.push_back(d); }
    } else if (num == 7 && wt == 2) {                                                                // BlobShape shape { repeated int64 dim = 1 [packed] }
      wire::Reader rs(sub, sl);
      while (rs.more()) {
        uint32_t n2; int w2; uint64_t s2; const uint8_t *sub2; size_t sl2; const uint8_t *raw2;
        if (!rs.field(n2, w2, s2, sub2, sl2, raw2)) return false;
        if (n2 != 1) continue;
        if (w2 == 0) out.shape.push_back((long)s2);
        else if (w2 == 2) { wire::Reader rd(sub2, sl2); while (rd.more()) out.shape.push_back((long)rd.varint()); if (!rd.ok) return false; }
      }
    }  // diff (6), double_diff (9): not part of the weight path
  }
  if (!r.ok) return false;
  if (has_legacy) { out.shape.assign(legacy, legacy + 4); out.legacy_dims = true; }   // blob.cpp:469-478
  if (!dd.empty()) { out.data.resize(dd.size()); for (size_t i = 0; i < dd.size(); ++i) out.data[i] = (float)dd[i]; }  // :490-494
  if (out.shape.empty() && !out.data.empty()) out.shape.push_back((long)out.data.size());
  long count = 1;
  for (long d : out.shape) count *= d;
  return (size_t)count == out.data.size();  // CHECK_EQ(count_, proto.data_size()), blob.cpp:496
}

inline void parse_conv(const uint8_t *b, size_t n, CmLayer &L) {
  wire::Reader r(b, n);
  int kernel = 0, stride = 0, pad = -1;
  auto rep = [&](int wt, uint64_t s, const uint8_t *sub, size_t sl) -> int {  // first element of a repeated uint32 (2-D convs)
    if (wt == 0) return (int)s;
    wire::Reader rr(sub, sl);
    return rr.more() ? (int)rr.varint() : 0;
  };
  while (r.more()) {
    uint32_t num; int wt; uint64_t s; const uint8_t *sub; size_t sl; const uint8_t *raw;
    if (!r.field(num, wt, s, sub, sl, raw)) return;
    switch (num) {
      case 1: L.num_output = (int)s; break;
      case 2: L.bias_term = (int)s; break;
      case 3: if (pad < 0) pad = rep(wt, s, sub, sl); break;
      case 4: if (!kernel) kernel = rep(wt, s, sub, sl); break;
      case 5: L.group = (int)s; break;
      case 6: if (!stride) stride = rep(wt, s, sub, sl); break;
      case 9: L.pad_h = (int)s; break;
      case 10: L.pad_w = (int)s; break;
      case 11: L.kernel_h = (int)s; break;
      case 12: L.kernel_w = (int)s; break;
      case 13: L.stride_h = (int)s; break;
      case 14: L.stride_w = (int)s; break;
      case 18: L.dilation = rep(wt, s, sub, sl); break;
      default: break;
    }
  }
  if (kernel) L.kernel_h = L.kernel_w = kernel;     // base_conv_layer.cpp:46-95: kernel_size wins over kernel_h / kernel_w being absent
  if (stride) L.stride_h = L.stride_w = stride;
  if (pad >= 0) L.pad_h = L.pad_w = pad;
}

inline bool parse_layer(const uint8_t *b, size_t n, bool v1, CmLayer &L) {
  const uint32_t F_NAME = v1 ? 4 : 1, F_TYPE = v1 ? 5 : 2, F_BLOBS = v1 ? 6 : 7, F_CONV = v1 ? 10 : 106, F_IP = v1 ? 17 : 117;
  L.v1 = v1;
  wire::Reader r(b, n);
  while (r.more()) {
    uint32_t num; int wt; uint64_t s; const uint8_t *sub; size_t sl; const uint8_t *raw;
    if (!r.field(num, wt, s, sub, sl, raw)) return false;
    if (num == F_NAME && wt == 2) L.name.assign((const char *)sub, sl);
    else if (num == F_TYPE && wt == 2 && !v1) L.type.assign((const char *)sub, sl);
    else if (num == F_TYPE && wt == 0 && v1) L.type = "V1:" + std::to_string((int)s);
    else if (num == F_BLOBS && wt == 2) {
      CmBlob blob;
      if (!parse_blob(sub, sl, blob)) return false;
      L.blobs.push_back(std::move(blob));
    } else {
      if (num == F_CONV && wt == 2) { L.has_conv = true; parse_conv(sub, sl, L); }
      if (num == F_IP && wt == 2) {
        L.has_ip = true;
        wire::Reader ri(sub, sl);
        while (ri.more()) {
          uint32_t n2; int w2; uint64_t s2; const uint8_t *sub2; size_t sl2; const uint8_t *raw2;
          if (!ri.field(n2, w2, s2, sub2, sl2, raw2)) break;
          if (n2 == 1) L.num_output = (int)s2;
          if (n2 == 2) L.bias_term = (int)s2;
        }
      }
      L.other.emplace_back(num, std::string((const char *)raw, (const char *)r.p));  // raw bytes incl. tag: written back verbatim
    }
  }
  return r.ok;
}

inline bool ParseCaffeModel(const void *buf, size_t n, CaffeModel &m) {
  wire::Reader r(buf, n);
  while (r.more()) {
    uint32_t num; int wt; uint64_t s; const uint8_t *sub; size_t sl; const uint8_t *raw;
    if (!r.field(num, wt, s, sub, sl, raw)) break;
    if (num == 1 && wt == 2) m.name.assign((const char *)sub, sl);
    else if ((num == 100 || num == 2) && wt == 2) {   // LayerParameter layer = 100 / V1LayerParameter layers = 2
      CmLayer L;
      if (!parse_layer(sub, sl, num == 2, L)) { m.error = "malformed layer " + std::to_string(m.layers.size()); return false; }
      m.layers.push_back(std::move(L));
    } else m.other.emplace_back(num, std::string((const char *)raw, (const char *)r.p));
  }
  if (!r.ok) m.error = "malformed NetParameter (truncated or not a binary proto)";
  return r.ok;
}

inline bool ReadCaffeModel(const std::string &path, CaffeModel &m) {
  FILE *f = fopen(path.c_str(), "rb");
  if (!f) { m.error = "cannot open " + path; return false; }
  std::string buf;
  char tmp[1 << 16];
  size_t k;
  while ((k = fread(tmp, 1, sizeof(tmp), f)) > 0) buf.append(tmp, k);
  fclose(f);
  return ParseCaffeModel(buf.data(), buf.size(), m);
}

inline std::string SerializeCaffeModel(const CaffeModel &m) {
  std::string out;
  if (!m.name.empty()) wire::put_bytes(out, 1, m.name);
  for (const auto &kv : m.other) out += kv.second;
  for (const CmLayer &L : m.layers) {
    const uint32_t F_NAME = L.v1 ? 4 : 1, F_TYPE = L.v1 ? 5 : 2, F_BLOBS = L.v1 ? 6 : 7;
    std::string lb;
    wire::put_bytes(lb, F_NAME, L.name);
    if (L.v1) { if (L.type.rfind("V1:", 0) == 0) wire::put_uint(lb, F_TYPE, (uint64_t)atoi(L.type.c_str() + 3)); }
    else wire::put_bytes(lb, F_TYPE, L.type);
    for (const auto &kv : L.other) lb += kv.second;
    for (const CmBlob &B : L.blobs) {
      std::string bb;
      if (B.legacy_dims) {
        for (int i = 0; i < 4; ++i) wire::put_uint(bb, 1 + i, (uint64_t)(i < (int)B.shape.size() ? B.shape[i] : 1));
      } else {
        std::string dims, sh;
        for (long d : B.shape) wire::put_varint(dims, (uint64_t)d);
        wire::put_bytes(sh, 1, dims);
        wire::put_bytes(bb, 7, sh);
      }
      wire::put_bytes(bb, 5, std::string((const char *)B.data.data(), B.data.size() * 4));
      wire::put_bytes(lb, F_BLOBS, bb);
    }
    wire::put_bytes(out, L.v1 ? 2 : 100, lb);
  }
  return out;
}

inline bool WriteCaffeModel(const std::string &path, const CaffeModel &m) {
  const std::string out = SerializeCaffeModel(m);
  FILE *f = fopen(path.c_str(), "wb");
  if (!f) return false;
  const bool ok = fwrite(out.data(), 1, out.size(), f) == out.size();
  fclose(f);
  return ok;
}

}  // namespace escort_host
