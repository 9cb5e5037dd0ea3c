// caffemodel.cpp -- read / write .caffemodel files (binary protobuf NetParameter) without libprotobuf.
// Only the fields Net::CopyTrainedLayersFrom (caffe_3d/src/caffe/net.cpp:852-883) and Net::ToProto
// (:885-904) touch are interpreted:
//   NetParameter   { name = 1; layer = 100 (LayerParameter); layers = 2 (V1, rejected) }   caffe.proto:62-99
//   LayerParameter { name = 1; type = 2; blobs = 7 (BlobProto) }                           caffe.proto:282-301
//   BlobProto      { shape = 7 {dim = 1 packed int64}; data = 5 packed float; diff = 6;
//                    legacy num/channels/height/width = 1..4 }                             caffe.proto:5-20
// Matching is by layer name; blob counts and shapes must agree (kReshape = false, net.cpp:867-874),
// where legacy 4-D dims are compared after stripping leading 1s as Blob::ShapeEquals does (blob.cpp:411-437).
#include <cstdint>
#include <cstring>
#include <fstream>
#include <sstream>
#include <stdexcept>

#include "net.hpp"

namespace eco {
namespace {

struct Reader {
  const uint8_t* p;
  const uint8_t* end;
  bool ok() const { return p < end; }
  uint64_t varint() {
    uint64_t v = 0;
    int shift = 0;
    while (p < end) {
      const uint8_t b = *p++;
      v |= (uint64_t)(b & 0x7F) << shift;
      if (!(b & 0x80)) return v;
      shift += 7;
      if (shift > 63) break;
    }
    throw std::runtime_error("caffemodel: truncated varint");
  }
  Reader sub() {
    const uint64_t n = varint();
    if ((uint64_t)(end - p) < n) throw std::runtime_error("caffemodel: truncated length-delimited field");
    Reader r{p, p + n};
    p += n;
    return r;
  }
  void skip(int wire) {
    switch (wire) {
      case 0: varint(); break;
      case 1: if (end - p < 8) throw std::runtime_error("caffemodel: truncated"); p += 8; break;
      case 2: sub(); break;
      case 5: if (end - p < 4) throw std::runtime_error("caffemodel: truncated"); p += 4; break;
      default: throw std::runtime_error("caffemodel: unsupported wire type");
    }
  }
};

struct BlobData {
  std::vector<long long> shape;
  bool legacy = false;
  std::vector<float> data;
};

BlobData read_blob(Reader r) {
  BlobData b;
  long long legacy[4] = {0, 0, 0, 0};
  bool has_legacy = false, has_shape = false;
  while (r.ok()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 7 && wire == 2) {
      Reader s = r.sub();
      has_shape = true;
      while (s.ok()) {
        const uint64_t k2 = s.varint();
        if ((k2 >> 3) == 1 && (k2 & 7) == 2) {
          Reader d = s.sub();
          while (d.ok()) b.shape.push_back((long long)d.varint());
        } else if ((k2 >> 3) == 1 && (k2 & 7) == 0) {
          b.shape.push_back((long long)s.varint());
        } else {
          s.skip((int)(k2 & 7));
        }
      }
    } else if (field == 5 && wire == 2) {
      Reader d = r.sub();
      const size_t n = (size_t)(d.end - d.p) / 4;
      const size_t old = b.data.size();
      b.data.resize(old + n);
      std::memcpy(b.data.data() + old, d.p, n * 4);
    } else if (field == 5 && wire == 5) {
      if (r.end - r.p < 4) throw std::runtime_error("truncated caffemodel (float field)");
      float f;
      std::memcpy(&f, r.p, 4);
      r.p += 4;
      b.data.push_back(f);
    } else if (field >= 1 && field <= 4 && wire == 0) {
      legacy[field - 1] = (long long)r.varint();
      has_legacy = true;
    } else if (field == 8 && wire == 2) {  // double_data (packed)
      Reader d = r.sub();
      const size_t n = (size_t)(d.end - d.p) / 8;
      for (size_t i = 0; i < n; ++i) {
        double v;
        std::memcpy(&v, d.p + 8 * i, 8);
        b.data.push_back((float)v);
      }
    } else {
      r.skip(wire);
    }
  }
  if (!has_shape && has_legacy) {
    b.legacy = true;
    b.shape.assign(legacy, legacy + 4);
  }
  return b;
}

bool shape_equals(const BlobData& src, const std::vector<int>& dst) {
  if (!src.legacy) {
    if (src.shape.size() != dst.size()) return false;
    for (size_t i = 0; i < dst.size(); ++i)
      if (src.shape[i] != dst[i]) return false;
    return true;
  }
  // legacy 4-D: compare as [num, channels, height, width] with the target left-padded by 1s (blob.cpp:411-437)
  if (dst.size() > 4) return false;
  long long d4[4] = {1, 1, 1, 1};
  for (size_t i = 0; i < dst.size(); ++i) d4[4 - dst.size() + i] = dst[i];
  for (int i = 0; i < 4; ++i)
    if (src.shape[i] != d4[i]) return false;
  return true;
}

void put_varint(std::string& o, uint64_t v) {
  while (v >= 0x80) {
    o.push_back((char)((v & 0x7F) | 0x80));
    v >>= 7;
  }
  o.push_back((char)v);
}
void put_key(std::string& o, int field, int wire) { put_varint(o, ((uint64_t)field << 3) | (uint64_t)wire); }
void put_bytes(std::string& o, int field, const std::string& s) {
  put_key(o, field, 2);
  put_varint(o, s.size());
  o.append(s);
}

}  // namespace

void Net::copy_from(const std::string& path) {
  if (params_dev_newer_) sync_params_to_host();
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("Could not open " + path);
  std::string buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  Reader r{reinterpret_cast<const uint8_t*>(buf.data()), reinterpret_cast<const uint8_t*>(buf.data()) + buf.size()};
  while (r.ok()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 2 && wire == 2)
      throw std::runtime_error("caffemodel uses V1 'layers'; upgrade it with the reference's upgrade_net_proto_binary");
    if (field != 100 || wire != 2) {
      r.skip(wire);
      continue;
    }
    Reader lr = r.sub();
    std::string lname;
    std::vector<BlobData> blobs;
    while (lr.ok()) {
      const uint64_t k2 = lr.varint();
      const int f2 = (int)(k2 >> 3), w2 = (int)(k2 & 7);
      if (f2 == 1 && w2 == 2) {
        Reader s = lr.sub();
        lname.assign(reinterpret_cast<const char*>(s.p), (size_t)(s.end - s.p));
      } else if (f2 == 7 && w2 == 2) {
        blobs.push_back(read_blob(lr.sub()));
      } else {
        lr.skip(w2);
      }
    }
    // "Ignoring source layer" when the name is unknown (net.cpp:860-863)
    OrigLayer* target = nullptr;
    for (auto& L : layers_)
      if (L.name == lname) { target = &L; break; }  // first match, as the reference's linear search (net.cpp:856-859)
    if (!target) continue;
    if (target->params.size() != blobs.size()) {
      std::ostringstream o;
      o << "Incompatible number of blobs for layer " << lname << ": " << target->params.size() << " vs " << blobs.size();
      throw std::runtime_error(o.str());
    }
    for (size_t i = 0; i < blobs.size(); ++i) {
      ParamBlob& pb = target->params[i];
      if (!shape_equals(blobs[i], pb.shape) || blobs[i].data.size() != pb.data.size()) {
        std::ostringstream o;
        o << "Cannot copy param " << i << " weights from layer '" << lname << "'; shape mismatch.  Source param shape is";
        for (auto d : blobs[i].shape) o << " " << d;
        o << "; target param shape is";
        for (auto d : pb.shape) o << " " << d;
        throw std::runtime_error(o.str());
      }
      pb.data = blobs[i].data;
    }
    target->params_dirty = true;
  }
}

void Net::save(const std::string& path) const {
  if (params_dev_newer_) const_cast<Net*>(this)->sync_params_to_host();  // training updated the device arena
  std::string out;
  put_bytes(out, 1, name_);
  for (const auto& L : layers_) {
    std::string lm;
    put_bytes(lm, 1, L.name);
    put_bytes(lm, 2, L.type);
    for (int b : L.bottoms) put_bytes(lm, 3, tensors_[b].name);
    for (int t : L.tops) put_bytes(lm, 4, tensors_[t].name);
    for (const auto& pb : L.params) {
      std::string bm;
      std::string packed(reinterpret_cast<const char*>(pb.data.data()), pb.data.size() * 4);
      put_bytes(bm, 5, packed);
      std::string dims, shp;
      for (int d : pb.shape) put_varint(dims, (uint64_t)d);
      put_bytes(shp, 1, dims);
      put_bytes(bm, 7, shp);
      put_bytes(lm, 7, bm);
    }
    put_bytes(out, 100, lm);
  }
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("Could not open " + path + " for writing");
  f.write(out.data(), (std::streamsize)out.size());
}

// ---- SolverState (caffe.proto:217-222): iter = 1, learned_net = 2, history = 3 (repeated BlobProto), current_step = 4 ----
void write_solverstate(const std::string& path, int iter, const std::string& learned_net, int current_step,
                       const std::vector<std::pair<std::vector<int>, std::vector<float>>>& history) {
  std::string out;
  put_key(out, 1, 0);
  put_varint(out, (uint64_t)iter);
  put_bytes(out, 2, learned_net);
  for (const auto& h : history) {
    std::string bm;
    std::string packed(reinterpret_cast<const char*>(h.second.data()), h.second.size() * 4);
    put_bytes(bm, 5, packed);
    std::string dims, shp;
    for (int d : h.first) put_varint(dims, (uint64_t)d);
    put_bytes(shp, 1, dims);
    put_bytes(bm, 7, shp);
    put_bytes(out, 3, bm);
  }
  put_key(out, 4, 0);
  put_varint(out, (uint64_t)current_step);
  std::ofstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("Could not open " + path + " for writing");
  f.write(out.data(), (std::streamsize)out.size());
}

void read_solverstate(const std::string& path, int* iter, std::string* learned_net, int* current_step,
                      std::vector<std::vector<float>>* history) {
  std::ifstream f(path, std::ios::binary);
  if (!f) throw std::runtime_error("Could not open " + path);
  std::string buf((std::istreambuf_iterator<char>(f)), std::istreambuf_iterator<char>());
  Reader r{reinterpret_cast<const uint8_t*>(buf.data()), reinterpret_cast<const uint8_t*>(buf.data()) + buf.size()};
  *iter = 0;
  *current_step = 0;
  learned_net->clear();
  history->clear();
  while (r.ok()) {
    const uint64_t key = r.varint();
    const int field = (int)(key >> 3), wire = (int)(key & 7);
    if (field == 1 && wire == 0) *iter = (int)r.varint();
    else if (field == 4 && wire == 0) *current_step = (int)r.varint();
    else if (field == 2 && wire == 2) {
      Reader s2 = r.sub();
      learned_net->assign(reinterpret_cast<const char*>(s2.p), (size_t)(s2.end - s2.p));
    } else if (field == 3 && wire == 2) {
      history->push_back(read_blob(r.sub()).data);
    } else {
      r.skip(wire);
    }
  }
}

}  // namespace eco
