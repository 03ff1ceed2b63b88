// Minimal ONNX protobuf wire reader for piper voice files (no protobuf / onnx dependency).
//
// Reads exactly what the engine needs from the graph the reference exporter writes
// (/root/reference/src/python/piper_train/export_onnx.py:88-101): the initializers
// (TensorProto: dims=1, data_type=2, float_data=4, int64_data=7, name=8, raw_data=9) and the
// nodes (NodeProto: input=1, output=2, name=3, op_type=4, attribute=5) with integer attributes
// (AttributeProto: name=1, i=3, ints=8).  The file bytes are kept alive by OnnxModel so tensors
// can point straight into raw_data without a copy.
#pragma once
#include <cstdint>
#include <map>
#include <string>
#include <unordered_map>
#include <vector>

namespace pb200 {

struct OnnxTensor {
  std::string name;
  std::vector<int64_t> dims;
  int dtype = 0;                 // 1 = float32, 7 = int64
  const uint8_t* raw = nullptr;  // little-endian payload inside OnnxModel::bytes (or owned)
  size_t raw_bytes = 0;
  std::vector<float> owned;      // when the file used float_data instead of raw_data
  int64_t numel() const {
    int64_t n = 1;
    for (auto d : dims) n *= d;
    return n;
  }
  const float* f32() const { return owned.empty() ? reinterpret_cast<const float*>(raw) : owned.data(); }
};

struct OnnxNode {
  std::string op_type, name;
  std::vector<std::string> inputs, outputs;
  std::map<std::string, std::vector<int64_t>> ints;
  int64_t attr(const char* key, int64_t dflt, size_t idx = 0) const {
    auto it = ints.find(key);
    if (it == ints.end() || it->second.size() <= idx) return dflt;
    return it->second[idx];
  }
};

struct OnnxModel {
  std::vector<uint8_t> bytes;
  std::string producer;
  int64_t ir_version = 0, opset = 0;
  std::vector<OnnxNode> nodes;
  std::vector<OnnxTensor> initializers;
  std::unordered_map<std::string, size_t> init_index;
  std::vector<std::string> inputs, outputs;
  const OnnxTensor* find(const std::string& name) const {
    auto it = init_index.find(name);
    return it == init_index.end() ? nullptr : &initializers[it->second];
  }
};

// Throws std::runtime_error on malformed input.
void load_onnx(const std::string& path, OnnxModel& out);

}  // namespace pb200
