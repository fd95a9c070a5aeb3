// Minimal JSON reader (objects, arrays, strings, numbers, booleans, null) for the run configuration files.
// The reference uses nlohmann/json (fetched by CPM, CMakeLists.txt:39-43); the configs only need this subset.
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace qkdldpc {

class Json {
public:
    enum class Kind { Null, Bool, Number, String, Array, Object };
    Kind kind = Kind::Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::vector<std::pair<std::string, Json>> obj;   // insertion order kept

    static Json parse(const std::string &text) {
        size_t p = 0;
        Json j = value(text, p);
        ws(text, p);
        if (p != text.size()) throw std::runtime_error("JSON: trailing characters at offset " + std::to_string(p));
        return j;
    }
    bool contains(const std::string &key) const {
        if (kind != Kind::Object) return false;
        for (auto &kv : obj) if (kv.first == key) return true;
        return false;
    }
    const Json &at(const std::string &key) const {
        if (kind != Kind::Object) throw std::runtime_error("JSON: not an object while reading key '" + key + "'");
        for (auto &kv : obj) if (kv.first == key) return kv.second;
        throw std::runtime_error("JSON: key '" + key + "' not found");
    }
    const std::vector<Json> &items() const {
        if (kind != Kind::Array) throw std::runtime_error("JSON: array expected");
        return arr;
    }
    double as_double() const {
        if (kind != Kind::Number) throw std::runtime_error("JSON: number expected");
        return num;
    }
    size_t as_size() const {
        if (kind != Kind::Number || num < 0) throw std::runtime_error("JSON: non-negative integer expected");
        return static_cast<size_t>(num);
    }
    bool as_bool() const {
        if (kind != Kind::Bool) throw std::runtime_error("JSON: boolean expected");
        return b;
    }
    bool empty() const { return kind == Kind::Null || (kind == Kind::Object && obj.empty()) || (kind == Kind::Array && arr.empty()); }

private:
    static void ws(const std::string &s, size_t &p) {
        if (p == 0 && s.size() >= 3 && (unsigned char)s[0] == 0xEF && (unsigned char)s[1] == 0xBB && (unsigned char)s[2] == 0xBF) p = 3;
        while (p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) ++p;
    }
    static Json value(const std::string &s, size_t &p) {
        ws(s, p);
        if (p >= s.size()) throw std::runtime_error("JSON: unexpected end of input");
        Json j;
        const char c = s[p];
        if (c == '{') {
            j.kind = Kind::Object;
            ++p;
            ws(s, p);
            if (p < s.size() && s[p] == '}') { ++p; return j; }
            while (true) {
                ws(s, p);
                Json k = string_(s, p);
                ws(s, p);
                if (p >= s.size() || s[p] != ':') throw std::runtime_error("JSON: ':' expected at offset " + std::to_string(p));
                ++p;
                j.obj.emplace_back(k.str, value(s, p));
                ws(s, p);
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == '}') { ++p; return j; }
                throw std::runtime_error("JSON: ',' or '}' expected at offset " + std::to_string(p));
            }
        }
        if (c == '[') {
            j.kind = Kind::Array;
            ++p;
            ws(s, p);
            if (p < s.size() && s[p] == ']') { ++p; return j; }
            while (true) {
                j.arr.push_back(value(s, p));
                ws(s, p);
                if (p < s.size() && s[p] == ',') { ++p; continue; }
                if (p < s.size() && s[p] == ']') { ++p; return j; }
                throw std::runtime_error("JSON: ',' or ']' expected at offset " + std::to_string(p));
            }
        }
        if (c == '"') return string_(s, p);
        if (s.compare(p, 4, "true") == 0) { j.kind = Kind::Bool; j.b = true; p += 4; return j; }
        if (s.compare(p, 5, "false") == 0) { j.kind = Kind::Bool; j.b = false; p += 5; return j; }
        if (s.compare(p, 4, "null") == 0) { p += 4; return j; }
        char *end = nullptr;
        j.num = std::strtod(s.c_str() + p, &end);
        if (end == s.c_str() + p) throw std::runtime_error("JSON: unexpected character at offset " + std::to_string(p));
        j.kind = Kind::Number;
        p = static_cast<size_t>(end - s.c_str());
        return j;
    }
    static Json string_(const std::string &s, size_t &p) {
        if (p >= s.size() || s[p] != '"') throw std::runtime_error("JSON: string expected at offset " + std::to_string(p));
        Json j;
        j.kind = Kind::String;
        ++p;
        while (p < s.size() && s[p] != '"') {
            if (s[p] == '\\' && p + 1 < s.size()) {
                const char e = s[p + 1];
                j.str += (e == 'n') ? '\n' : (e == 't') ? '\t' : (e == 'r') ? '\r' : e;
                p += 2;
            } else {
                j.str += s[p++];
            }
        }
        if (p >= s.size()) throw std::runtime_error("JSON: unterminated string");
        ++p;
        return j;
    }
};

}  // namespace qkdldpc
