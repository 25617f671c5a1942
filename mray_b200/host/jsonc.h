// jsonc.h — a small JSON-with-comments reader for the MRay scene format (Docs/markdown/scene/mrayScene.md): objects, arrays,
// numbers, strings, booleans, null, // line and /* block */ comments, trailing commas tolerated. Values keep their source order.
#pragma once
#include <cstdint>
#include <cstdlib>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <string_view>
#include <utility>
#include <vector>

namespace jsonc
{

struct Value
{
    enum Kind { Null, Bool, Number, String, Array, Object } kind = Null;
    bool b = false;
    double num = 0.0;
    bool isInteger = false;
    std::string str;
    std::vector<Value> arr;
    std::vector<std::pair<std::string, Value>> obj;

    bool IsArray() const { return kind == Array; }
    bool IsObject() const { return kind == Object; }
    bool IsNumber() const { return kind == Number; }
    bool IsString() const { return kind == String; }
    size_t Size() const { return kind == Array ? arr.size() : kind == Object ? obj.size() : 1; }
    const Value* Find(std::string_view key) const
    {
        if(kind != Object) return nullptr;
        for(const auto& kv : obj) if(kv.first == key) return &kv.second;
        return nullptr;
    }
    const Value& At(std::string_view key) const
    {
        const Value* v = Find(key);
        if(!v) throw std::runtime_error("json: key \"" + std::string(key) + "\" not found");
        return *v;
    }
    const Value& operator[](size_t i) const
    {
        if(kind != Array || i >= arr.size()) throw std::runtime_error("json: array index out of range");
        return arr[i];
    }
    double AsNumber() const { if(kind != Number) throw std::runtime_error("json: number expected"); return num; }
    uint32_t AsU32() const { if(kind != Number) throw std::runtime_error("json: integer expected"); return uint32_t(num); }
    bool AsBool() const { if(kind != Bool) throw std::runtime_error("json: boolean expected"); return b; }
    const std::string& AsString() const { if(kind != String) throw std::runtime_error("json: string expected"); return str; }
};

class Parser
{
    std::string_view s;
    size_t p = 0;

    [[noreturn]] void Fail(const char* what) const
    {
        size_t line = 1;
        for(size_t i = 0; i < p && i < s.size(); i++) if(s[i] == '\n') line++;
        throw std::runtime_error(std::string("json: ") + what + " at line " + std::to_string(line));
    }
    void SkipSpace()
    {
        for(;;)
        {
            while(p < s.size() && (s[p] == ' ' || s[p] == '\t' || s[p] == '\n' || s[p] == '\r')) p++;
            if(p + 1 < s.size() && s[p] == '/' && s[p + 1] == '/') { while(p < s.size() && s[p] != '\n') p++; continue; }
            if(p + 1 < s.size() && s[p] == '/' && s[p + 1] == '*')
            {
                p += 2;
                while(p + 1 < s.size() && !(s[p] == '*' && s[p + 1] == '/')) p++;
                if(p + 1 >= s.size()) Fail("unterminated comment");
                p += 2; continue;
            }
            break;
        }
    }
    Value ParseString()
    {
        Value v; v.kind = Value::String;
        p++;   // opening quote
        while(p < s.size() && s[p] != '"')
        {
            char c = s[p++];
            if(c == '\\')
            {
                if(p >= s.size()) Fail("bad escape");
                char e = s[p++];
                switch(e)
                {
                    case 'n': v.str += '\n'; break; case 't': v.str += '\t'; break; case 'r': v.str += '\r'; break;
                    case 'b': v.str += '\b'; break; case 'f': v.str += '\f'; break;
                    case 'u':
                    {
                        if(p + 4 > s.size()) Fail("bad \\u escape");
                        unsigned cp = unsigned(std::strtoul(std::string(s.substr(p, 4)).c_str(), nullptr, 16)); p += 4;
                        if(cp < 0x80) v.str += char(cp);
                        else if(cp < 0x800) { v.str += char(0xC0 | (cp >> 6)); v.str += char(0x80 | (cp & 0x3F)); }
                        else { v.str += char(0xE0 | (cp >> 12)); v.str += char(0x80 | ((cp >> 6) & 0x3F)); v.str += char(0x80 | (cp & 0x3F)); }
                        break;
                    }
                    default: v.str += e; break;   // \" \\ \/
                }
            }
            else v.str += c;
        }
        if(p >= s.size()) Fail("unterminated string");
        p++;
        return v;
    }
    Value ParseNumber()
    {
        size_t b = p;
        if(p < s.size() && (s[p] == '-' || s[p] == '+')) p++;
        bool integer = true;
        while(p < s.size() && ((s[p] >= '0' && s[p] <= '9') || s[p] == '.' || s[p] == 'e' || s[p] == 'E' || s[p] == '-' || s[p] == '+'))
        {
            if(s[p] == '.' || s[p] == 'e' || s[p] == 'E') integer = false;
            p++;
        }
        if(b == p) Fail("value expected");
        Value v; v.kind = Value::Number; v.isInteger = integer;
        v.num = std::strtod(std::string(s.substr(b, p - b)).c_str(), nullptr);
        return v;
    }
    Value ParseValue()
    {
        SkipSpace();
        if(p >= s.size()) Fail("unexpected end");
        char c = s[p];
        if(c == '{')
        {
            Value v; v.kind = Value::Object; p++;
            for(;;)
            {
                SkipSpace();
                if(p < s.size() && s[p] == '}') { p++; break; }
                if(p >= s.size() || s[p] != '"') Fail("key expected");
                std::string key = ParseString().str;
                SkipSpace();
                if(p >= s.size() || s[p] != ':') Fail("':' expected");
                p++;
                v.obj.emplace_back(std::move(key), ParseValue());
                SkipSpace();
                if(p < s.size() && s[p] == ',') { p++; continue; }
                if(p < s.size() && s[p] == '}') { p++; break; }
                Fail("',' or '}' expected");
            }
            return v;
        }
        if(c == '[')
        {
            Value v; v.kind = Value::Array; p++;
            for(;;)
            {
                SkipSpace();
                if(p < s.size() && s[p] == ']') { p++; break; }
                v.arr.push_back(ParseValue());
                SkipSpace();
                if(p < s.size() && s[p] == ',') { p++; continue; }
                if(p < s.size() && s[p] == ']') { p++; break; }
                Fail("',' or ']' expected");
            }
            return v;
        }
        if(c == '"') return ParseString();
        if(s.compare(p, 4, "true") == 0) { p += 4; Value v; v.kind = Value::Bool; v.b = true; return v; }
        if(s.compare(p, 5, "false") == 0) { p += 5; Value v; v.kind = Value::Bool; v.b = false; return v; }
        if(s.compare(p, 4, "null") == 0) { p += 4; return Value{}; }
        return ParseNumber();
    }

    public:
    explicit Parser(std::string_view text) : s(text) {}
    Value Parse()
    {
        Value v = ParseValue();
        SkipSpace();
        if(p != s.size()) Fail("trailing characters");
        return v;
    }
};

inline Value Parse(std::string_view text) { return Parser(text).Parse(); }

} // namespace jsonc
