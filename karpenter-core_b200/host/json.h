// Minimal JSON DOM (objects, arrays, strings, numbers, bools, null) used to load
// test problems into kmodel::Problem. Not a general-purpose library.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace kjson {

struct Value {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;  // Str, and the raw text of Num
  std::vector<Value> arr;
  std::vector<std::pair<std::string, Value>> obj;

  bool is_null() const { return kind == Null; }
  const Value* get(const std::string& k) const {
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  bool has(const std::string& k) const {
    const Value* v = get(k);
    return v && !v->is_null();
  }
  std::string s(const std::string& k, const std::string& def = "") const {
    const Value* v = get(k);
    if (!v || v->is_null()) return def;
    return v->str;
  }
  double d(const std::string& k, double def = 0) const {
    const Value* v = get(k);
    if (!v || v->is_null()) return def;
    return v->num;
  }
  int64_t i(const std::string& k, int64_t def = 0) const {
    const Value* v = get(k);
    if (!v || v->is_null()) return def;
    return (int64_t)v->num;
  }
  bool boolean(const std::string& k, bool def = false) const {
    const Value* v = get(k);
    if (!v || v->is_null()) return def;
    return v->b;
  }
};

class Parser {
 public:
  explicit Parser(const char* s) : p_(s) {}
  Value parse() {
    Value v = value();
    ws();
    if (*p_) fail("trailing characters");
    return v;
  }

 private:
  const char* p_;
  [[noreturn]] void fail(const char* m) { throw std::runtime_error(std::string("json: ") + m); }
  void ws() {
    while (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r') ++p_;
  }
  Value value() {
    ws();
    Value v;
    switch (*p_) {
      case '{': {
        v.kind = Value::Obj;
        ++p_;
        ws();
        if (*p_ == '}') { ++p_; return v; }
        for (;;) {
          ws();
          std::string k = string();
          ws();
          if (*p_ != ':') fail("expected ':'");
          ++p_;
          v.obj.emplace_back(std::move(k), value());
          ws();
          if (*p_ == ',') { ++p_; continue; }
          if (*p_ == '}') { ++p_; return v; }
          fail("expected ',' or '}'");
        }
      }
      case '[': {
        v.kind = Value::Arr;
        ++p_;
        ws();
        if (*p_ == ']') { ++p_; return v; }
        for (;;) {
          v.arr.push_back(value());
          ws();
          if (*p_ == ',') { ++p_; continue; }
          if (*p_ == ']') { ++p_; return v; }
          fail("expected ',' or ']'");
        }
      }
      case '"':
        v.kind = Value::Str;
        v.str = string();
        return v;
      case 't':
        if (std::string(p_, 4) != "true") fail("bad literal");
        p_ += 4; v.kind = Value::Bool; v.b = true; return v;
      case 'f':
        if (std::string(p_, 5) != "false") fail("bad literal");
        p_ += 5; v.kind = Value::Bool; v.b = false; return v;
      case 'n':
        if (std::string(p_, 4) != "null") fail("bad literal");
        p_ += 4; return v;
      default: {
        char* end = nullptr;
        v.num = std::strtod(p_, &end);
        if (end == p_) fail("unexpected character");
        v.kind = Value::Num;
        v.str.assign(p_, (size_t)(end - p_));
        p_ = end;
        return v;
      }
    }
  }
  std::string string() {
    if (*p_ != '"') fail("expected string");
    ++p_;
    std::string out;
    while (*p_ && *p_ != '"') {
      if (*p_ == '\\') {
        ++p_;
        switch (*p_) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case 'b': out += '\b'; break;
          case 'f': out += '\f'; break;
          case 'u': {
            unsigned cp = (unsigned)std::strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: out += *p_;
        }
        ++p_;
      } else {
        out += *p_++;
      }
    }
    if (*p_ != '"') fail("unterminated string");
    ++p_;
    return out;
  }
};

}  // namespace kjson
