#include "onnx_reader.h"

#include <cstdio>
#include <cstdint>
#include <cstring>
#include <stdexcept>

namespace pb200 {
namespace {

struct Span {
  const uint8_t* p;
  const uint8_t* end;
};

uint64_t varint(Span& s) {
  uint64_t v = 0;
  int shift = 0;
  while (true) {
    if (s.p >= s.end || shift > 63) throw std::runtime_error("onnx: truncated varint");
    uint8_t b = *s.p++;
    v |= uint64_t(b & 0x7F) << shift;
    if (!(b & 0x80)) return v;
    shift += 7;
  }
}

struct Field {
  uint32_t no;
  uint32_t wt;
  uint64_t value;  // wt 0
  Span sub;        // wt 2 (payload) / wt 1,5 (fixed bytes)
};

bool next(Span& s, Field& f) {
  if (s.p >= s.end) return false;
  uint64_t key = varint(s);
  f.no = uint32_t(key >> 3);
  f.wt = uint32_t(key & 7);
  switch (f.wt) {
    case 0: f.value = varint(s); break;
    case 1:
      if (s.end - s.p < 8) throw std::runtime_error("onnx: truncated fixed64");
      f.sub = {s.p, s.p + 8};
      s.p += 8;
      break;
    case 2: {
      uint64_t n = varint(s);
      if (uint64_t(s.end - s.p) < n) throw std::runtime_error("onnx: truncated length-delimited field");
      f.sub = {s.p, s.p + n};
      s.p += n;
      break;
    }
    case 5:
      if (s.end - s.p < 4) throw std::runtime_error("onnx: truncated fixed32");
      f.sub = {s.p, s.p + 4};
      s.p += 4;
      break;
    default: throw std::runtime_error("onnx: unsupported wire type");
  }
  return true;
}

std::string str(const Span& s) { return std::string(reinterpret_cast<const char*>(s.p), s.end - s.p); }

void parse_tensor(Span s, OnnxTensor& t) {
  Field f;
  std::vector<int64_t> i64;
  while (next(s, f)) {
    switch (f.no) {
      case 1:
        if (f.wt == 0) t.dims.push_back(int64_t(f.value));
        else { Span p = f.sub; while (p.p < p.end) t.dims.push_back(int64_t(varint(p))); }
        break;
      case 2: t.dtype = int(f.value); break;
      case 8: t.name = str(f.sub); break;
      case 9: t.raw = f.sub.p; t.raw_bytes = size_t(f.sub.end - f.sub.p); break;
      case 4:  // float_data (packed or not)
        for (const uint8_t* p = f.sub.p; p + 4 <= f.sub.end; p += 4) {
          float v;
          std::memcpy(&v, p, 4);
          t.owned.push_back(v);
        }
        break;
      default: break;
    }
  }
  // A truncated or hostile file must fail here, not as an out-of-bounds read in the packer (the reference gets this
  // validation from onnxruntime's model loader).
  int64_t n = 1;
  for (int64_t d : t.dims) {
    if (d < 0 || d > (int64_t(1) << 31)) throw std::runtime_error("onnx: tensor '" + t.name + "' has a bad dimension");
    n *= d;
    if (n > (int64_t(1) << 36)) throw std::runtime_error("onnx: tensor '" + t.name + "' is implausibly large");
  }
  if (t.dtype == 1 && t.raw == nullptr && t.owned.empty() && n != 0)
    throw std::runtime_error("onnx: float tensor '" + t.name + "' has no data");
  if (t.dtype == 1 && t.raw != nullptr && int64_t(t.raw_bytes) != n * 4)
    throw std::runtime_error("onnx: tensor '" + t.name + "' raw_data size does not match dims");
  if (t.dtype == 1 && t.raw == nullptr && int64_t(t.owned.size()) != n)
    throw std::runtime_error("onnx: tensor '" + t.name + "' float_data size does not match dims");
  if (t.dtype == 1 && t.raw != nullptr && (reinterpret_cast<uintptr_t>(t.raw) & 3u) != 0) {
    // raw_data sits at an arbitrary byte offset of the file: give f32() an aligned copy
    t.owned.resize(size_t(n));
    std::memcpy(t.owned.data(), t.raw, size_t(n) * 4);
    t.raw = nullptr;
    t.raw_bytes = 0;
  }
}

void parse_attr(Span s, OnnxNode& n) {
  Field f;
  std::string name;
  std::vector<int64_t> ints;
  while (next(s, f)) {
    if (f.no == 1) name = str(f.sub);
    else if (f.no == 3 && f.wt == 0) ints.push_back(int64_t(f.value));
    else if (f.no == 8) {
      if (f.wt == 0) ints.push_back(int64_t(f.value));
      else { Span p = f.sub; while (p.p < p.end) ints.push_back(int64_t(varint(p))); }
    }
  }
  if (!ints.empty()) n.ints[name] = std::move(ints);
}

void parse_node(Span s, OnnxNode& n) {
  Field f;
  while (next(s, f)) {
    switch (f.no) {
      case 1: n.inputs.push_back(str(f.sub)); break;
      case 2: n.outputs.push_back(str(f.sub)); break;
      case 3: n.name = str(f.sub); break;
      case 4: n.op_type = str(f.sub); break;
      case 5: parse_attr(f.sub, n); break;
      default: break;  // doc_string (6) etc.
    }
  }
}

std::string value_info_name(Span s) {
  Field f;
  while (next(s, f))
    if (f.no == 1 && f.wt == 2) return str(f.sub);
  return "";
}

void parse_graph(Span s, OnnxModel& m) {
  Field f;
  while (next(s, f)) {
    if (f.wt != 2) continue;
    if (f.no == 1) {
      m.nodes.emplace_back();
      parse_node(f.sub, m.nodes.back());
    } else if (f.no == 5) {
      m.initializers.emplace_back();
      parse_tensor(f.sub, m.initializers.back());
    } else if (f.no == 11) {
      m.inputs.push_back(value_info_name(f.sub));
    } else if (f.no == 12) {
      m.outputs.push_back(value_info_name(f.sub));
    }
  }
}

}  // namespace

void load_onnx(const std::string& path, OnnxModel& m) {
  FILE* fp = std::fopen(path.c_str(), "rb");
  if (!fp) throw std::runtime_error("cannot open voice model '" + path + "'");
  std::fseek(fp, 0, SEEK_END);
  long n = std::ftell(fp);
  std::fseek(fp, 0, SEEK_SET);
  if (n <= 0) {
    std::fclose(fp);
    throw std::runtime_error("voice model '" + path + "' is empty");
  }
  m.bytes.resize(size_t(n));
  size_t got = std::fread(m.bytes.data(), 1, size_t(n), fp);
  std::fclose(fp);
  if (got != size_t(n)) throw std::runtime_error("short read on '" + path + "'");
  Span s{m.bytes.data(), m.bytes.data() + m.bytes.size()};
  Field f;
  bool have_graph = false;
  while (next(s, f)) {
    if (f.no == 1 && f.wt == 0) m.ir_version = int64_t(f.value);
    else if (f.no == 2 && f.wt == 2) m.producer = str(f.sub);
    else if (f.no == 7 && f.wt == 2) { parse_graph(f.sub, m); have_graph = true; }
    else if (f.no == 8 && f.wt == 2) {
      Span o = f.sub;
      Field g;
      while (next(o, g))
        if (g.no == 2 && g.wt == 0) m.opset = int64_t(g.value);
    }
  }
  if (!have_graph) throw std::runtime_error("'" + path + "' holds no ONNX graph");
  for (size_t i = 0; i < m.initializers.size(); ++i) m.init_index[m.initializers[i].name] = i;
}

}  // namespace pb200
