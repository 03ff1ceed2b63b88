// Small recursive-descent JSON reader: just enough for piper's `<voice>.onnx.json`
// (objects, arrays, strings with \uXXXX escapes, numbers, true/false/null).  Header-only.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace minijson {

struct Value {
  enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;  // insertion order kept

  bool contains(const std::string& k) const { return find(k) != nullptr; }
  const Value* find(const std::string& k) const {
    if (kind != Object) return nullptr;
    for (const auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  const Value& at(const std::string& k) const {
    const Value* v = find(k);
    if (!v) throw std::runtime_error("json: missing key '" + k + "'");
    return *v;
  }
  double number() const {
    if (kind != Number) throw std::runtime_error("json: not a number");
    return num;
  }
  const std::string& string() const {
    if (kind != String) throw std::runtime_error("json: not a string");
    return str;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {}
  Value parse() {
    Value v = value();
    ws();
    if (i_ != s_.size()) fail("trailing characters");
    return v;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m + " at offset " + std::to_string(i_)); }
  void ws() { while (i_ < s_.size() && (s_[i_] == ' ' || s_[i_] == '\n' || s_[i_] == '\t' || s_[i_] == '\r')) ++i_; }
  char peek() { ws(); if (i_ >= s_.size()) fail("unexpected end"); return s_[i_]; }
  void expect(char c) { if (peek() != c) fail("unexpected character"); ++i_; }
  static void utf8(std::string& out, uint32_t cp) {
    if (cp < 0x80) out += char(cp);
    else if (cp < 0x800) { out += char(0xC0 | (cp >> 6)); out += char(0x80 | (cp & 0x3F)); }
    else if (cp < 0x10000) { out += char(0xE0 | (cp >> 12)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
    else { out += char(0xF0 | (cp >> 18)); out += char(0x80 | ((cp >> 12) & 0x3F)); out += char(0x80 | ((cp >> 6) & 0x3F)); out += char(0x80 | (cp & 0x3F)); }
  }
  uint32_t hex4() {
    if (i_ + 4 > s_.size()) fail("bad \\u escape");
    uint32_t v = 0;
    for (int k = 0; k < 4; ++k) {
      char c = s_[i_++];
      v <<= 4;
      if (c >= '0' && c <= '9') v |= uint32_t(c - '0');
      else if (c >= 'a' && c <= 'f') v |= uint32_t(c - 'a' + 10);
      else if (c >= 'A' && c <= 'F') v |= uint32_t(c - 'A' + 10);
      else fail("bad hex digit");
    }
    return v;
  }
  std::string string_lit() {
    expect('"');
    std::string out;
    while (true) {
      if (i_ >= s_.size()) fail("unterminated string");
      char c = s_[i_++];
      if (c == '"') return out;
      if (c != '\\') { out += c; continue; }
      if (i_ >= s_.size()) fail("bad escape");
      char e = s_[i_++];
      switch (e) {
        case '"': out += '"'; break;
        case '\\': out += '\\'; break;
        case '/': out += '/'; break;
        case 'b': out += '\b'; break;
        case 'f': out += '\f'; break;
        case 'n': out += '\n'; break;
        case 'r': out += '\r'; break;
        case 't': out += '\t'; break;
        case 'u': {
          uint32_t cp = hex4();
          if (cp >= 0xD800 && cp < 0xDC00 && i_ + 6 <= s_.size() && s_[i_] == '\\' && s_[i_ + 1] == 'u') {
            i_ += 2;
            uint32_t lo = hex4();
            cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
          }
          utf8(out, cp);
          break;
        }
        default: fail("unknown escape");
      }
    }
  }
  Value value() {
    char c = peek();
    Value v;
    if (c == '{') {
      ++i_;
      v.kind = Value::Object;
      if (peek() == '}') { ++i_; return v; }
      while (true) {
        std::string k = string_lit();
        expect(':');
        v.obj.emplace_back(std::move(k), value());
        char d = peek();
        ++i_;
        if (d == '}') return v;
        if (d != ',') fail("expected , or }");
        ws();
      }
    }
    if (c == '[') {
      ++i_;
      v.kind = Value::Array;
      if (peek() == ']') { ++i_; return v; }
      while (true) {
        v.arr.push_back(value());
        char d = peek();
        ++i_;
        if (d == ']') return v;
        if (d != ',') fail("expected , or ]");
      }
    }
    if (c == '"') { v.kind = Value::String; v.str = string_lit(); return v; }
    if (s_.compare(i_, 4, "true") == 0) { i_ += 4; v.kind = Value::Bool; v.b = true; return v; }
    if (s_.compare(i_, 5, "false") == 0) { i_ += 5; v.kind = Value::Bool; return v; }
    if (s_.compare(i_, 4, "null") == 0) { i_ += 4; return v; }
    char* end = nullptr;
    v.num = std::strtod(s_.c_str() + i_, &end);
    if (end == s_.c_str() + i_) fail("unexpected token");
    i_ = size_t(end - s_.c_str());
    v.kind = Value::Number;
    return v;
  }
};

inline Value parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace minijson
