// prototxt.hpp -- protobuf text-format reader for caffe_3d net definitions (no protoc / libprotobuf
// in this image).  Reads what ReadProtoFromTextFile (caffe_3d/src/caffe/util/io.cpp) would read for
// the schema in caffe_3d/src/caffe/proto/caffe.proto; field semantics (defaults, repeated-ness) are
// applied by the consumers in net.cpp, the parser keeps every field as an ordered list.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace eco {
namespace pt {

struct Msg;
struct Value {
  std::string text;            // scalar token (string literal contents, number, enum identifier)
  bool is_string = false;
  std::shared_ptr<Msg> msg;    // set for message-typed fields
};

struct Msg {
  std::vector<std::pair<std::string, Value>> fields;  // in file order

  std::vector<const Value*> all(const std::string& name) const {
    std::vector<const Value*> out;
    for (auto& f : fields)
      if (f.first == name) out.push_back(&f.second);
    return out;
  }
  bool has(const std::string& name) const {
    for (auto& f : fields)
      if (f.first == name) return true;
    return false;
  }
  const Value* last(const std::string& name) const {
    const Value* v = nullptr;
    for (auto& f : fields)
      if (f.first == name) v = &f.second;
    return v;
  }
  const Msg* msg(const std::string& name) const {
    const Value* v = last(name);
    return (v && v->msg) ? v->msg.get() : nullptr;
  }
  std::string str(const std::string& name, const std::string& def = "") const {
    const Value* v = last(name);
    return v ? v->text : def;
  }
  double num(const std::string& name, double def) const {
    const Value* v = last(name);
    return v ? std::strtod(v->text.c_str(), nullptr) : def;
  }
  long integer(const std::string& name, long def) const {
    const Value* v = last(name);
    return v ? std::strtol(v->text.c_str(), nullptr, 0) : def;
  }
  bool boolean(const std::string& name, bool def) const {
    const Value* v = last(name);
    if (!v) return def;
    return v->text == "true" || v->text == "True" || v->text == "1";
  }
  std::vector<long> integers(const std::string& name) const {
    std::vector<long> out;
    for (auto* v : all(name)) out.push_back(std::strtol(v->text.c_str(), nullptr, 0));
    return out;
  }
  std::vector<std::string> strs(const std::string& name) const {
    std::vector<std::string> out;
    for (auto* v : all(name)) out.push_back(v->text);
    return out;
  }
};

class Parser {
 public:
  explicit Parser(const std::string& s) : s_(s) {}
  std::shared_ptr<Msg> parse() {
    auto m = std::make_shared<Msg>();
    parse_fields(*m, '\0');
    return m;
  }

 private:
  const std::string& s_;
  size_t i_ = 0;

  [[noreturn]] void fail(const std::string& what) const {
    size_t line = 1;
    for (size_t k = 0; k < i_ && k < s_.size(); ++k)
      if (s_[k] == '\n') ++line;
    throw std::runtime_error("prototxt parse error at line " + std::to_string(line) + ": " + what);
  }
  void skip_ws() {
    while (i_ < s_.size()) {
      const char c = s_[i_];
      if (c == '#') {
        while (i_ < s_.size() && s_[i_] != '\n') ++i_;
      } else if (c == ' ' || c == '\t' || c == '\n' || c == '\r') {
        ++i_;
      } else {
        break;
      }
    }
  }
  static bool is_ident(char c) {
    return (c >= 'a' && c <= 'z') || (c >= 'A' && c <= 'Z') || (c >= '0' && c <= '9') || c == '_' || c == '.' ||
           c == '-' || c == '+';
  }
  std::string ident() {
    const size_t b = i_;
    while (i_ < s_.size() && is_ident(s_[i_])) ++i_;
    if (b == i_) fail(std::string("unexpected character '") + (i_ < s_.size() ? s_[i_] : '?') + "'");
    return s_.substr(b, i_ - b);
  }
  std::string quoted() {
    const char q = s_[i_++];
    std::string out;
    while (i_ < s_.size() && s_[i_] != q) {
      char c = s_[i_++];
      if (c == '\\' && i_ < s_.size()) {
        const char e = s_[i_++];
        switch (e) {
          case 'n': c = '\n'; break;
          case 't': c = '\t'; break;
          case 'r': c = '\r'; break;
          case '0': c = '\0'; break;
          default: c = e; break;
        }
      }
      out.push_back(c);
    }
    if (i_ >= s_.size()) fail("unterminated string");
    ++i_;
    return out;
  }
  Value scalar() {
    Value v;
    skip_ws();
    if (i_ < s_.size() && (s_[i_] == '"' || s_[i_] == '\'')) {
      v.is_string = true;
      v.text = quoted();
      skip_ws();
      while (i_ < s_.size() && (s_[i_] == '"' || s_[i_] == '\'')) {  // adjacent literals concatenate
        v.text += quoted();
        skip_ws();
      }
    } else {
      v.text = ident();
    }
    return v;
  }
  void parse_fields(Msg& m, char close) {
    for (;;) {
      skip_ws();
      if (i_ >= s_.size()) {
        if (close != '\0') fail("unterminated message");
        return;
      }
      const char c = s_[i_];
      if (c == close && close != '\0') {
        ++i_;
        return;
      }
      if (c == ',' || c == ';') {
        ++i_;
        continue;
      }
      const std::string name = ident();
      skip_ws();
      if (i_ < s_.size() && s_[i_] == ':') {
        ++i_;
        skip_ws();
      }
      if (i_ >= s_.size()) fail("field '" + name + "' has no value");
      if (s_[i_] == '{' || s_[i_] == '<') {
        const char cl = s_[i_] == '{' ? '}' : '>';
        ++i_;
        Value v;
        v.msg = std::make_shared<Msg>();
        parse_fields(*v.msg, cl);
        m.fields.emplace_back(name, std::move(v));
      } else if (s_[i_] == '[') {
        ++i_;
        for (;;) {
          skip_ws();
          if (i_ >= s_.size()) fail("unterminated list");
          if (s_[i_] == ']') {
            ++i_;
            break;
          }
          if (s_[i_] == ',') {
            ++i_;
            continue;
          }
          if (s_[i_] == '{' || s_[i_] == '<') {
            const char cl = s_[i_] == '{' ? '}' : '>';
            ++i_;
            Value v;
            v.msg = std::make_shared<Msg>();
            parse_fields(*v.msg, cl);
            m.fields.emplace_back(name, std::move(v));
          } else {
            m.fields.emplace_back(name, scalar());
          }
        }
      } else {
        m.fields.emplace_back(name, scalar());
      }
    }
  }
};

inline std::shared_ptr<Msg> parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace pt
}  // namespace eco
