// json.hpp — a small JSON DOM (parse + serialise) for the host side of the hot path.
// Objects keep insertion order (Go's encoding/json emits struct fields in declaration order; we mirror that by
// building objects in the same order).
#pragma once
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace cch {

struct Json {
  enum Type { Null, Bool, Number, String, Array, Object } type = Null;
  bool b = false;
  std::string s;          // String: the value; Number: the literal text as written
  std::vector<Json> arr;
  std::vector<std::pair<std::string, Json>> obj;

  Json() {}
  static Json null() { return Json(); }
  static Json boolean(bool v) { Json j; j.type = Bool; j.b = v; return j; }
  static Json number(long long v) { Json j; j.type = Number; j.s = std::to_string(v); return j; }
  static Json number_text(const std::string &t) { Json j; j.type = Number; j.s = t; return j; }
  static Json string(const std::string &v) { Json j; j.type = String; j.s = v; return j; }
  static Json array() { Json j; j.type = Array; return j; }
  static Json object() { Json j; j.type = Object; return j; }

  bool is_null() const { return type == Null; }
  bool is_object() const { return type == Object; }
  bool is_array() const { return type == Array; }
  bool is_string() const { return type == String; }

  const Json *find(const std::string &k) const {
    if (type != Object) return nullptr;
    for (auto &kv : obj) if (kv.first == k) return &kv.second;
    return nullptr;
  }
  // path lookup: get("spec","nodeName")
  const Json &at(const std::string &k) const {
    static const Json none;
    const Json *p = find(k);
    return p ? *p : none;
  }
  Json &set(const std::string &k, Json v) {
    for (auto &kv : obj) if (kv.first == k) { kv.second = std::move(v); return kv.second; }
    type = Object;
    obj.emplace_back(k, std::move(v));
    return obj.back().second;
  }
  void push(Json v) { type = Array; arr.push_back(std::move(v)); }
  std::string str(const std::string &dflt = "") const { return type == String ? s : dflt; }
  long long i64(long long dflt = 0) const {
    if (type == Number) return strtoll(s.c_str(), nullptr, 10);
    return dflt;
  }
  bool truthy() const { return type == Bool && b; }
  size_t size() const { return type == Array ? arr.size() : (type == Object ? obj.size() : 0); }
};

class JsonParser {
 public:
  explicit JsonParser(std::string_view t) : t_(t) {}   // a view: items of a list are parsed in place, without copying their text
  Json parse() {
    Json v = value();
    ws();
    if (p_ != t_.size()) fail("trailing characters");
    return v;
  }

 private:
  std::string_view t_;
  size_t p_ = 0;
  int depth_ = 0;                    // nesting of the value being parsed; capped so that a hostile document cannot overflow the stack
  static constexpr int kMaxDepth = 512;
  struct Nest { JsonParser &p; explicit Nest(JsonParser &q) : p(q) { if (++p.depth_ > kMaxDepth) p.fail("nesting too deep"); } ~Nest() { p.depth_--; } };
  [[noreturn]] void fail(const char *m) { throw std::runtime_error(std::string("json: ") + m + " at offset " + std::to_string(p_)); }
  void ws() { while (p_ < t_.size() && (t_[p_] == ' ' || t_[p_] == '\n' || t_[p_] == '\t' || t_[p_] == '\r')) p_++; }
  Json value() {
    ws();
    if (p_ >= t_.size()) fail("unexpected end");
    char c = t_[p_];
    if (c == '{') return object();
    if (c == '[') return array();
    if (c == '"') return Json::string(string());
    if (c == 't') { expect("true"); return Json::boolean(true); }
    if (c == 'f') { expect("false"); return Json::boolean(false); }
    if (c == 'n') { expect("null"); return Json::null(); }
    return number();
  }
  void expect(const char *w) {
    size_t n = strlen(w);
    if (t_.compare(p_, n, w) != 0) fail("bad literal");
    p_ += n;
  }
  Json number() {
    size_t b = p_;
    if (p_ < t_.size() && (t_[p_] == '-' || t_[p_] == '+')) p_++;
    while (p_ < t_.size() && (isdigit((unsigned char)t_[p_]) || t_[p_] == '.' || t_[p_] == 'e' || t_[p_] == 'E' || t_[p_] == '-' || t_[p_] == '+')) p_++;
    if (b == p_) fail("bad value");
    return Json::number_text(std::string(t_.substr(b, p_ - b)));
  }
  std::string string() {
    std::string out;
    p_++;  // opening quote
    while (true) {
      // fast path: copy the run up to the next quote or backslash in one go
      const char *b = t_.data() + p_, *e = t_.data() + t_.size(), *q = b;
      while (q < e && *q != '"' && *q != '\\') q++;
      if (q > b) { out.append(b, (size_t)(q - b)); p_ += (size_t)(q - b); }
      if (p_ >= t_.size()) fail("unterminated string");
      char c = t_[p_++];
      if (c == '"') break;
      if (c == '\\') {
        if (p_ >= t_.size()) fail("bad escape");
        char e = t_[p_++];
        switch (e) {
          case '"': out += '"'; break; case '\\': out += '\\'; break; case '/': out += '/'; break;
          case 'b': out += '\b'; break; case 'f': out += '\f'; break; case 'n': out += '\n'; break;
          case 'r': out += '\r'; break; case 't': out += '\t'; break;
          case 'u': {
            if (p_ + 4 > t_.size()) fail("bad \\u");
            unsigned cp = (unsigned)strtoul(std::string(t_.substr(p_, 4)).c_str(), nullptr, 16);
            p_ += 4;
            if (cp >= 0xD800 && cp <= 0xDBFF && p_ + 6 <= t_.size() && t_[p_] == '\\' && t_[p_ + 1] == 'u') {
              unsigned lo = (unsigned)strtoul(std::string(t_.substr(p_ + 2, 4)).c_str(), nullptr, 16);
              p_ += 6;
              cp = 0x10000 + ((cp - 0xD800) << 10) + (lo - 0xDC00);
            }
            if (cp < 0x80) out += (char)cp;
            else if (cp < 0x800) { out += (char)(0xC0 | (cp >> 6)); out += (char)(0x80 | (cp & 0x3F)); }
            else if (cp < 0x10000) { out += (char)(0xE0 | (cp >> 12)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            else { out += (char)(0xF0 | (cp >> 18)); out += (char)(0x80 | ((cp >> 12) & 0x3F)); out += (char)(0x80 | ((cp >> 6) & 0x3F)); out += (char)(0x80 | (cp & 0x3F)); }
            break;
          }
          default: fail("bad escape");
        }
      } else out += c;
    }
    return out;
  }
  Json array() {
    Nest nest(*this);
    Json a = Json::array();
    p_++;
    ws();
    if (p_ < t_.size() && t_[p_] == ']') { p_++; return a; }
    while (true) {
      a.arr.push_back(value());
      ws();
      if (p_ >= t_.size()) fail("unterminated array");
      if (t_[p_] == ',') { p_++; continue; }
      if (t_[p_] == ']') { p_++; break; }
      fail("expected , or ]");
    }
    return a;
  }
  Json object() {
    Nest nest(*this);
    Json o = Json::object();
    p_++;
    ws();
    if (p_ < t_.size() && t_[p_] == '}') { p_++; return o; }
    while (true) {
      ws();
      if (p_ >= t_.size() || t_[p_] != '"') fail("expected key");
      std::string k = string();
      ws();
      if (p_ >= t_.size() || t_[p_] != ':') fail("expected :");
      p_++;
      o.obj.emplace_back(std::move(k), value());
      ws();
      if (p_ >= t_.size()) fail("unterminated object");
      if (t_[p_] == ',') { p_++; continue; }
      if (t_[p_] == '}') { p_++; break; }
      fail("expected , or }");
    }
    return o;
  }
};

inline Json parse_json(std::string_view text) { return JsonParser(text).parse(); }

// Go's encoding/json escapes <, >, & and U+2028/2029 as well (HTML-safe by default: json.Marshal).
inline void json_escape(const std::string &s, std::string &out) {
  out += '"';
  for (size_t i = 0; i < s.size(); i++) {
    unsigned char c = (unsigned char)s[i];
    switch (c) {
      case '"': out += "\\\""; break;
      case '\\': out += "\\\\"; break;
      case '\n': out += "\\n"; break;
      case '\r': out += "\\r"; break;
      case '\t': out += "\\t"; break;
      case '<': out += "\\u003c"; break;
      case '>': out += "\\u003e"; break;
      case '&': out += "\\u0026"; break;
      default:
        if (c < 0x20) { char b[8]; snprintf(b, sizeof(b), "\\u%04x", c); out += b; }
        else out += (char)c;
    }
  }
  out += '"';
}

inline void json_dump(const Json &j, std::string &out) {
  switch (j.type) {
    case Json::Null: out += "null"; break;
    case Json::Bool: out += j.b ? "true" : "false"; break;
    case Json::Number: out += j.s; break;
    case Json::String: json_escape(j.s, out); break;
    case Json::Array:
      out += '[';
      for (size_t i = 0; i < j.arr.size(); i++) { if (i) out += ','; json_dump(j.arr[i], out); }
      out += ']';
      break;
    case Json::Object:
      out += '{';
      for (size_t i = 0; i < j.obj.size(); i++) {
        if (i) out += ',';
        json_escape(j.obj[i].first, out);
        out += ':';
        json_dump(j.obj[i].second, out);
      }
      out += '}';
      break;
  }
}

inline std::string json_dump(const Json &j) { std::string o; json_dump(j, o); return o; }

}  // namespace cch
