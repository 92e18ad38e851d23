// Minimal JSON DOM parser (objects, arrays, strings, numbers, true/false/null).
// Used for speech_tokenizer/config.json (Cfg.swift:361-383, 574-582 keys) and the
// safetensors header.  Throws std::runtime_error on malformed input.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace q3 {

struct Json {
  enum Kind { Null, Bool, Num, Str, Arr, Obj } kind = Null;
  bool b = false;
  double num = 0;
  std::string str;
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;  // insertion order kept

  const Json* find(const std::string& k) const {
    if (kind != Obj) return nullptr;
    for (auto& kv : obj)
      if (kv.first == k) return &kv.second;
    return nullptr;
  }
  bool is_null() const { return kind == Null; }
};

class JsonParser {
 public:
  JsonParser(const char* p, size_t n) : p_(p), end_(p + n) {}
  Json parse() {
    Json v = value();
    ws();
    if (p_ != end_) fail("trailing characters");
    return v;
  }

 private:
  const char* p_;
  const char* end_;
  [[noreturn]] void fail(const char* why) { throw std::runtime_error(std::string("json: ") + why); }
  void ws() {
    while (p_ < end_ && (*p_ == ' ' || *p_ == '\n' || *p_ == '\t' || *p_ == '\r')) ++p_;
  }
  bool lit(const char* s) {
    size_t n = strlen_(s);
    if ((size_t)(end_ - p_) >= n && std::equal(s, s + n, p_)) {
      p_ += n;
      return true;
    }
    return false;
  }
  static size_t strlen_(const char* s) {
    size_t n = 0;
    while (s[n]) ++n;
    return n;
  }
  Json value() {
    ws();
    if (p_ >= end_) fail("unexpected end");
    Json v;
    char c = *p_;
    if (c == '{') {
      ++p_;
      v.kind = Json::Obj;
      ws();
      if (p_ < end_ && *p_ == '}') {
        ++p_;
        return v;
      }
      for (;;) {
        ws();
        if (p_ >= end_ || *p_ != '"') fail("expected key");
        std::string k = string_();
        ws();
        if (p_ >= end_ || *p_ != ':') fail("expected ':'");
        ++p_;
        v.obj.emplace_back(std::move(k), value());
        ws();
        if (p_ < end_ && *p_ == ',') {
          ++p_;
          continue;
        }
        if (p_ < end_ && *p_ == '}') {
          ++p_;
          break;
        }
        fail("expected ',' or '}'");
      }
    } else if (c == '[') {
      ++p_;
      v.kind = Json::Arr;
      ws();
      if (p_ < end_ && *p_ == ']') {
        ++p_;
        return v;
      }
      for (;;) {
        v.arr.push_back(value());
        ws();
        if (p_ < end_ && *p_ == ',') {
          ++p_;
          continue;
        }
        if (p_ < end_ && *p_ == ']') {
          ++p_;
          break;
        }
        fail("expected ',' or ']'");
      }
    } else if (c == '"') {
      v.kind = Json::Str;
      v.str = string_();
    } else if (lit("true")) {
      v.kind = Json::Bool;
      v.b = true;
    } else if (lit("false")) {
      v.kind = Json::Bool;
    } else if (lit("null")) {
      v.kind = Json::Null;
    } else {
      char* e = nullptr;
      std::string tmp(p_, (size_t)std::min<ptrdiff_t>(end_ - p_, 64));
      v.num = std::strtod(tmp.c_str(), &e);
      if (e == tmp.c_str()) fail("bad token");
      p_ += (e - tmp.c_str());
      v.kind = Json::Num;
    }
    return v;
  }
  std::string string_() {
    ++p_;  // opening quote
    std::string s;
    while (p_ < end_ && *p_ != '"') {
      if (*p_ == '\\') {
        ++p_;
        if (p_ >= end_) fail("bad escape");
        switch (*p_) {
          case 'n': s += '\n'; break;
          case 't': s += '\t'; break;
          case 'r': s += '\r'; break;
          case 'b': s += '\b'; break;
          case 'f': s += '\f'; break;
          case 'u': {
            if (end_ - p_ < 5) fail("bad \\u");
            unsigned cp = (unsigned)std::strtoul(std::string(p_ + 1, 4).c_str(), nullptr, 16);
            p_ += 4;
            if (cp < 0x80) s += (char)cp;
            else if (cp < 0x800) { s += (char)(0xC0 | (cp >> 6)); s += (char)(0x80 | (cp & 0x3F)); }
            else { s += (char)(0xE0 | (cp >> 12)); s += (char)(0x80 | ((cp >> 6) & 0x3F)); s += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: s += *p_;
        }
        ++p_;
      } else {
        s += *p_++;
      }
    }
    if (p_ >= end_) fail("unterminated string");
    ++p_;
    return s;
  }
};

}  // namespace q3
