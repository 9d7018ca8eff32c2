// Minimal JSON reader for job.json and scene files (the reference reads them with nlohmann json:
// /root/reference/include/job.h:18-68, src/scene_parser.cpp:140-200).  Only what those files use:
// objects, arrays, strings, numbers, booleans, null.  Accessing a missing key yields a null value,
// mirroring nlohmann's operator[] on a mutable object.
#pragma once

#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <utility>
#include <vector>

namespace pathed {

class Json {
public:
    enum class Type { Null, Bool, Number, String, Array, Object };

    Json() : m_type(Type::Null) {}

    static Json parse(const std::string &text)
    {
        size_t pos = 0;
        Json value = parseValue(text, pos);
        skipSpace(text, pos);
        if (pos != text.size()) { throw std::runtime_error("json: trailing characters"); }
        return value;
    }

    Type type() const { return m_type; }
    bool isNull() const { return m_type == Type::Null; }
    bool isObject() const { return m_type == Type::Object; }
    bool isArray() const { return m_type == Type::Array; }
    bool isString() const { return m_type == Type::String; }
    bool isNumber() const { return m_type == Type::Number; }
    bool isBool() const { return m_type == Type::Bool; }

    // missing key / wrong container -> shared null, like nlohmann's operator[] creating a null member
    const Json &operator[](const std::string &key) const
    {
        static const Json null;
        if (m_type != Type::Object) { return null; }
        for (const auto &member : *m_members) { if (member.first == key) { return member.second; } }
        return null;
    }
    const Json &operator[](size_t index) const
    {
        static const Json null;
        if (m_type != Type::Array || index >= m_items->size()) { return null; }
        return (*m_items)[index];
    }
    size_t size() const
    {
        if (m_type == Type::Array) { return m_items->size(); }
        if (m_type == Type::Object) { return m_members->size(); }
        return 0;
    }
    const std::vector<Json> &items() const
    {
        static const std::vector<Json> empty;
        return m_type == Type::Array ? *m_items : empty;
    }
    const std::vector<std::pair<std::string, Json>> &members() const
    {
        static const std::vector<std::pair<std::string, Json>> empty;
        return m_type == Type::Object ? *m_members : empty;
    }

    // typed getters throw std::runtime_error on a type mismatch (nlohmann: type_error)
    const std::string &asString() const
    {
        if (m_type != Type::String) { throw std::runtime_error("json: type must be string"); }
        return m_string;
    }
    double asNumber() const
    {
        if (m_type != Type::Number) { throw std::runtime_error("json: type must be number"); }
        return m_number;
    }
    int asInt() const { return (int)asNumber(); }
    bool asBool() const
    {
        if (m_type != Type::Bool) { throw std::runtime_error("json: type must be boolean"); }
        return m_bool;
    }

    std::string dump(int indent = 0, int depth = 0) const
    {
        const std::string pad(indent * (depth + 1), ' '), padEnd(indent * depth, ' ');
        const char *nl = indent ? "\n" : "";
        switch (m_type) {
        case Type::Null: return "null";
        case Type::Bool: return m_bool ? "true" : "false";
        case Type::Number: {
            char buffer[64];
            if (m_number == (long long)m_number) { snprintf(buffer, sizeof(buffer), "%lld", (long long)m_number); }
            else { snprintf(buffer, sizeof(buffer), "%.17g", m_number); }
            return buffer;
        }
        case Type::String: return quote(m_string);
        case Type::Array: {
            std::string out = "[";
            for (size_t i = 0; i < m_items->size(); i++) {
                out += (i ? "," : "") + std::string(nl) + pad + (*m_items)[i].dump(indent, depth + 1);
            }
            return out + (m_items->empty() ? "" : nl + padEnd) + "]";
        }
        case Type::Object: {
            std::string out = "{";
            for (size_t i = 0; i < m_members->size(); i++) {
                out += (i ? "," : "") + std::string(nl) + pad + quote((*m_members)[i].first) + ": " +
                       (*m_members)[i].second.dump(indent, depth + 1);
            }
            return out + (m_members->empty() ? "" : nl + padEnd) + "}";
        }
        }
        return "";
    }

private:
    Type m_type;
    bool m_bool = false;
    double m_number = 0.0;
    std::string m_string;
    std::shared_ptr<std::vector<Json>> m_items;
    std::shared_ptr<std::vector<std::pair<std::string, Json>>> m_members;

    static std::string quote(const std::string &s)
    {
        std::string out = "\"";
        for (char ch : s) {
            if (ch == '"' || ch == '\\') { out += '\\'; out += ch; }
            else if (ch == '\n') { out += "\\n"; }
            else if (ch == '\t') { out += "\\t"; }
            else { out += ch; }
        }
        return out + "\"";
    }
    static void skipSpace(const std::string &t, size_t &p)
    {
        while (p < t.size() && (t[p] == ' ' || t[p] == '\n' || t[p] == '\t' || t[p] == '\r')) { p++; }
    }
    static Json parseValue(const std::string &t, size_t &p)
    {
        skipSpace(t, p);
        if (p >= t.size()) { throw std::runtime_error("json: unexpected end of input"); }
        Json v;
        const char ch = t[p];
        if (ch == '{') {
            v.m_type = Type::Object;
            v.m_members = std::make_shared<std::vector<std::pair<std::string, Json>>>();
            p++; skipSpace(t, p);
            if (p < t.size() && t[p] == '}') { p++; return v; }
            for (;;) {
                skipSpace(t, p);
                std::string key = parseString(t, p);
                skipSpace(t, p);
                if (p >= t.size() || t[p] != ':') { throw std::runtime_error("json: expected ':'"); }
                p++;
                Json member = parseValue(t, p);
                bool replaced = false;
                for (auto &existing : *v.m_members) {
                    if (existing.first == key) { existing.second = member; replaced = true; }
                }
                if (!replaced) { v.m_members->emplace_back(key, member); }
                skipSpace(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == '}') { p++; break; }
                throw std::runtime_error("json: expected ',' or '}'");
            }
        } else if (ch == '[') {
            v.m_type = Type::Array;
            v.m_items = std::make_shared<std::vector<Json>>();
            p++; skipSpace(t, p);
            if (p < t.size() && t[p] == ']') { p++; return v; }
            for (;;) {
                v.m_items->push_back(parseValue(t, p));
                skipSpace(t, p);
                if (p < t.size() && t[p] == ',') { p++; continue; }
                if (p < t.size() && t[p] == ']') { p++; break; }
                throw std::runtime_error("json: expected ',' or ']'");
            }
        } else if (ch == '"') {
            v.m_type = Type::String;
            v.m_string = parseString(t, p);
        } else if (t.compare(p, 4, "true") == 0) {
            v.m_type = Type::Bool; v.m_bool = true; p += 4;
        } else if (t.compare(p, 5, "false") == 0) {
            v.m_type = Type::Bool; v.m_bool = false; p += 5;
        } else if (t.compare(p, 4, "null") == 0) {
            p += 4;
        } else {
            char *end = nullptr;
            v.m_number = std::strtod(t.c_str() + p, &end);
            if (end == t.c_str() + p) { throw std::runtime_error("json: invalid value"); }
            v.m_type = Type::Number;
            p = (size_t)(end - t.c_str());
        }
        return v;
    }
    static std::string parseString(const std::string &t, size_t &p)
    {
        if (p >= t.size() || t[p] != '"') { throw std::runtime_error("json: expected string"); }
        p++;
        std::string out;
        while (p < t.size() && t[p] != '"') {
            if (t[p] == '\\' && p + 1 < t.size()) {
                p++;
                switch (t[p]) {
                case 'n': out += '\n'; break;
                case 't': out += '\t'; break;
                case 'r': out += '\r'; break;
                case 'b': out += '\b'; break;
                case 'f': out += '\f'; break;
                case 'u': {
                    const unsigned code = (unsigned)std::strtoul(t.substr(p + 1, 4).c_str(), nullptr, 16);
                    if (code < 0x80) { out += (char)code; }
                    else if (code < 0x800) { out += (char)(0xC0 | (code >> 6)); out += (char)(0x80 | (code & 0x3F)); }
                    else { out += (char)(0xE0 | (code >> 12)); out += (char)(0x80 | ((code >> 6) & 0x3F)); out += (char)(0x80 | (code & 0x3F)); }
                    p += 4;
                    break;
                }
                default: out += t[p];
                }
                p++;
            } else {
                out += t[p++];
            }
        }
        if (p >= t.size()) { throw std::runtime_error("json: unterminated string"); }
        p++;
        return out;
    }
};

} // namespace pathed
