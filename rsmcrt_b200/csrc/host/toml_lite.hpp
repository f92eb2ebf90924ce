// toml_lite.hpp — the subset of TOML v1.0 that the reference's res/*.toml files use (the reference parses
// with toml-f, fpm.toml:8, which is not vendored).  Supported: comments, [table], [[array-of-tables]],
// bare/quoted keys, basic and literal strings, integers, floats (exponent, underscores, inf/nan), booleans,
// (multi-line) arrays of scalars or arrays.  Not supported (and not used by any shipped config): dotted keys,
// inline tables, multi-line strings, dates.
#pragma once
#include <cctype>
#include <cmath>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <vector>

namespace toml_lite {

struct Value;
using Table = std::map<std::string, Value>;

struct Value {
    enum Kind { NONE, STRING, INT, FLOAT, BOOL, ARRAY, TABLE, TABLE_ARRAY } kind = NONE;
    std::string s;
    long long i = 0;
    double f = 0;
    bool b = false;
    std::vector<Value> arr;                     // ARRAY
    std::shared_ptr<Table> tbl;                 // TABLE
    std::vector<std::shared_ptr<Table>> tarr;   // TABLE_ARRAY
    int line = 0;
    bool is_number() const { return kind == INT || kind == FLOAT; }
    double as_double() const { return kind == INT ? (double)i : f; }
};

struct ParseError : std::runtime_error {
    int line;
    ParseError(const std::string& m, int l) : std::runtime_error("toml line " + std::to_string(l) + ": " + m), line(l) {}
};

class Parser {
  public:
    explicit Parser(const std::string& text) : t_(text) {}
    Table parse() {
        Table root;
        Table* cur = &root;
        while (true) {
            skip_ws_nl();
            if (eof()) break;
            if (peek() == '[') {
                ++p_;
                bool arr = false;
                if (peek() == '[') { arr = true; ++p_; }
                skip_ws();
                std::string name = parse_key();
                skip_ws();
                expect(']');
                if (arr) expect(']');
                end_of_line();
                Value& v = root[name];
                if (arr) {
                    if (v.kind == Value::NONE) v.kind = Value::TABLE_ARRAY;
                    if (v.kind != Value::TABLE_ARRAY) throw ParseError("'" + name + "' redefined", line_);
                    v.tarr.push_back(std::make_shared<Table>());
                    cur = v.tarr.back().get();
                } else {
                    if (v.kind != Value::NONE) throw ParseError("table '" + name + "' defined twice", line_);
                    v.kind = Value::TABLE;
                    v.tbl = std::make_shared<Table>();
                    cur = v.tbl.get();
                }
                continue;
            }
            std::string key = parse_key();
            skip_ws();
            expect('=');
            skip_ws();
            Value val = parse_value();
            val.line = line_;
            end_of_line();
            if (cur->count(key)) throw ParseError("key '" + key + "' defined twice", line_);
            (*cur)[key] = val;
        }
        return root;
    }

  private:
    const std::string& t_;
    size_t p_ = 0;
    int line_ = 1;
    bool eof() const { return p_ >= t_.size(); }
    char peek() const { return eof() ? '\0' : t_[p_]; }
    void expect(char c) {
        if (peek() != c) throw ParseError(std::string("expected '") + c + "'", line_);
        ++p_;
    }
    void skip_ws() {
        while (!eof() && (t_[p_] == ' ' || t_[p_] == '\t')) ++p_;
    }
    void skip_comment() {
        if (peek() == '#')
            while (!eof() && t_[p_] != '\n') ++p_;
    }
    void skip_ws_nl() {
        while (!eof()) {
            char c = t_[p_];
            if (c == ' ' || c == '\t' || c == '\r') ++p_;
            else if (c == '\n') { ++p_; ++line_; }
            else if (c == '#') skip_comment();
            else break;
        }
    }
    void end_of_line() {
        skip_ws();
        skip_comment();
        if (peek() == '\r') ++p_;
        if (eof()) return;
        if (peek() != '\n') throw ParseError("unexpected trailing characters", line_);
        ++p_;
        ++line_;
    }
    std::string parse_key() {
        if (peek() == '"' || peek() == '\'') return parse_string();
        size_t s = p_;
        while (!eof() && (std::isalnum((unsigned char)t_[p_]) || t_[p_] == '_' || t_[p_] == '-')) ++p_;
        if (p_ == s) throw ParseError("expected a key", line_);
        return t_.substr(s, p_ - s);
    }
    std::string parse_string() {
        char q = t_[p_++];
        std::string out;
        while (true) {
            if (eof() || t_[p_] == '\n') throw ParseError("unterminated string", line_);
            char c = t_[p_++];
            if (c == q) break;
            if (q == '"' && c == '\\') {
                char e = t_[p_++];
                switch (e) {
                    case 'n': out += '\n'; break;
                    case 't': out += '\t'; break;
                    case '\\': out += '\\'; break;
                    case '"': out += '"'; break;
                    default: out += e;
                }
            } else
                out += c;
        }
        return out;
    }
    Value parse_value() {
        Value v;
        char c = peek();
        if (c == '"' || c == '\'') {
            v.kind = Value::STRING;
            v.s = parse_string();
            return v;
        }
        if (c == '[') {
            ++p_;
            v.kind = Value::ARRAY;
            while (true) {
                skip_ws_nl();
                if (peek() == ']') { ++p_; break; }
                v.arr.push_back(parse_value());
                skip_ws_nl();
                if (peek() == ',') { ++p_; continue; }
                skip_ws_nl();
                if (peek() == ']') { ++p_; break; }
                throw ParseError("expected ',' or ']' in array", line_);
            }
            return v;
        }
        size_t s = p_;
        while (!eof() && !std::isspace((unsigned char)t_[p_]) && t_[p_] != ',' && t_[p_] != ']' && t_[p_] != '#') ++p_;
        std::string tok = t_.substr(s, p_ - s);
        if (tok == "true" || tok == "false") {
            v.kind = Value::BOOL;
            v.b = tok == "true";
            return v;
        }
        std::string clean;
        for (char ch : tok)
            if (ch != '_') clean += ch;
        if (clean.empty()) throw ParseError("expected a value", line_);
        bool is_float = clean.find_first_of(".eE") != std::string::npos || clean.find("inf") != std::string::npos ||
                        clean.find("nan") != std::string::npos;
        char* end = nullptr;
        if (is_float) {
            v.kind = Value::FLOAT;
            v.f = std::strtod(clean.c_str(), &end);
        } else {
            v.kind = Value::INT;
            v.i = std::strtoll(clean.c_str(), &end, 10);
        }
        if (!end || *end != '\0') throw ParseError("cannot parse value '" + tok + "'", line_);
        return v;
    }
};

inline Table parse(const std::string& text) { return Parser(text).parse(); }

}  // namespace toml_lite
