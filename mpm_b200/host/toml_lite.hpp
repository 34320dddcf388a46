// The subset of TOML the reference's scene files use (scenes/*.toml, read by src/main.cu:31-66
// through cpptoml): comments, [[array-of-tables]] headers, and key = value pairs whose values are
// strings, integers, floats, booleans or (possibly multi-line) arrays of those.  Plain [table]
// headers and dotted keys are rejected with an error rather than misread.
#pragma once
#include <cctype>
#include <cstdlib>
#include <fstream>
#include <map>
#include <sstream>
#include <stdexcept>
#include <string>
#include <vector>

namespace mpmh {

struct TomlValue {
  enum Kind { STRING, NUMBER, BOOL, ARRAY } kind = NUMBER;
  std::string s;
  double d = 0.0;
  bool b = false;
  std::vector<TomlValue> a;
};

struct TomlTable {
  std::map<std::string, TomlValue> kv;
  // the accessors main.cu uses: get_as<T>(key).value_or(default), get_array_of<double>(key)
  double number_or(const std::string& k, double dflt) const {
    auto it = kv.find(k);
    return (it != kv.end() && it->second.kind == TomlValue::NUMBER) ? it->second.d : dflt;
  }
  std::string string_or(const std::string& k, const std::string& dflt) const {
    auto it = kv.find(k);
    return (it != kv.end() && it->second.kind == TomlValue::STRING) ? it->second.s : dflt;
  }
  bool numbers(const std::string& k, std::vector<double>& out) const {
    auto it = kv.find(k);
    if (it == kv.end() || it->second.kind != TomlValue::ARRAY) return false;
    out.clear();
    for (const TomlValue& v : it->second.a) {
      if (v.kind != TomlValue::NUMBER) return false;
      out.push_back(v.d);
    }
    return true;
  }
};

struct TomlDoc {
  TomlTable root;
  std::map<std::string, std::vector<TomlTable>> arrays;  // [[name]] in file order
  const std::vector<TomlTable>& table_array(const std::string& name) const {
    static const std::vector<TomlTable> empty;
    auto it = arrays.find(name);
    return it == arrays.end() ? empty : it->second;
  }
};

class TomlParser {
 public:
  explicit TomlParser(const std::string& text) : t_(text) {}

  TomlDoc parse() {
    TomlDoc doc;
    TomlTable* cur = &doc.root;
    while (true) {
      skip_ws_nl();
      if (eof()) break;
      if (peek() == '[') {
        if (t_.compare(p_, 2, "[[") != 0) fail("plain [table] headers are not supported");
        p_ += 2;
        skip_ws();
        const std::string name = key();
        skip_ws();
        if (t_.compare(p_, 2, "]]") != 0) fail("expected ]]");
        p_ += 2;
        end_of_line();
        doc.arrays[name].emplace_back();
        cur = &doc.arrays[name].back();
        continue;
      }
      const std::string k = key();
      skip_ws();
      if (eof() || peek() != '=') fail("expected '=' after key '" + k + "'");
      ++p_;
      skip_ws();
      cur->kv[k] = value();
      end_of_line();
    }
    return doc;
  }

 private:
  const std::string& t_;
  size_t p_ = 0;
  int line() const {
    int n = 1;
    for (size_t i = 0; i < p_ && i < t_.size(); ++i) n += t_[i] == '\n';
    return n;
  }
  [[noreturn]] void fail(const std::string& what) const {
    throw std::runtime_error("TOML line " + std::to_string(line()) + ": " + what);
  }
  bool eof() const { return p_ >= t_.size(); }
  char peek() const { return t_[p_]; }
  void skip_ws() {
    while (!eof() && (peek() == ' ' || peek() == '\t')) ++p_;
  }
  void skip_comment() {
    if (!eof() && peek() == '#')
      while (!eof() && peek() != '\n') ++p_;
  }
  void skip_ws_nl() {
    while (!eof()) {
      const char c = peek();
      if (c == ' ' || c == '\t' || c == '\n' || c == '\r') ++p_;
      else if (c == '#') skip_comment();
      else break;
    }
  }
  void end_of_line() {
    skip_ws();
    skip_comment();
    if (!eof() && peek() == '\r') ++p_;
    if (!eof() && peek() != '\n') fail("unexpected characters after value");
  }
  std::string key() {
    if (!eof() && (peek() == '"' || peek() == '\'')) return string_value();
    const size_t b = p_;
    while (!eof() && (std::isalnum((unsigned char)peek()) || peek() == '_' || peek() == '-')) ++p_;
    if (p_ == b) fail("expected a key");
    if (!eof() && peek() == '.') fail("dotted keys are not supported");
    return t_.substr(b, p_ - b);
  }
  std::string string_value() {
    const char q = peek();
    ++p_;
    std::string out;
    while (true) {
      if (eof() || peek() == '\n') fail("unterminated string");
      char c = t_[p_++];
      if (c == q) break;
      if (q == '"' && c == '\\') {
        if (eof()) fail("bad escape");
        const char e = t_[p_++];
        switch (e) {
          case 'n': out += '\n'; break;
          case 't': out += '\t'; break;
          case 'r': out += '\r'; break;
          case '"': out += '"'; break;
          case '\\': out += '\\'; break;
          default: fail("unsupported escape");
        }
      } else {
        out += c;
      }
    }
    return out;
  }
  TomlValue value() {
    TomlValue v;
    if (eof()) fail("expected a value");
    const char c = peek();
    if (c == '"' || c == '\'') {
      v.kind = TomlValue::STRING;
      v.s = string_value();
    } else if (c == '[') {
      v.kind = TomlValue::ARRAY;
      ++p_;
      while (true) {
        skip_ws_nl();
        if (eof()) fail("unterminated array");
        if (peek() == ']') {
          ++p_;
          break;
        }
        v.a.push_back(value());
        skip_ws_nl();
        if (!eof() && peek() == ',') ++p_;
        else if (!eof() && peek() != ']') fail("expected ',' or ']' in array");
      }
    } else if (t_.compare(p_, 4, "true") == 0) {
      v.kind = TomlValue::BOOL;
      v.b = true;
      p_ += 4;
    } else if (t_.compare(p_, 5, "false") == 0) {
      v.kind = TomlValue::BOOL;
      p_ += 5;
    } else {
      std::string num;
      while (!eof() && (std::isalnum((unsigned char)peek()) || peek() == '+' || peek() == '-' || peek() == '.' || peek() == '_')) {
        if (peek() != '_') num += peek();
        ++p_;
      }
      if (num.empty()) fail("expected a value");
      char* end = nullptr;
      v.kind = TomlValue::NUMBER;
      if (num == "inf" || num == "+inf") v.d = 1.0 / 0.0;
      else if (num == "-inf") v.d = -1.0 / 0.0;
      else {
        v.d = std::strtod(num.c_str(), &end);
        if (!end || *end) fail("bad number '" + num + "'");
      }
    }
    return v;
  }
};

inline TomlDoc toml_parse_file(const std::string& path) {
  std::ifstream f(path);
  if (!f) throw std::runtime_error("cannot open scene file '" + path + "'");
  std::stringstream ss;
  ss << f.rdbuf();
  const std::string text = ss.str();
  return TomlParser(text).parse();
}

}  // namespace mpmh
