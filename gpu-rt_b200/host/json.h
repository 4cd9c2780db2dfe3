/* json.h — minimal recursive-descent JSON reader for the glTF front end (no third-party deps). */
#pragma once
#include <cstdlib>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace gpurt {

struct Json {
    enum Type { Null, Bool, Num, Str, Arr, Obj } type = Null;
    bool b = false;
    double num = 0;
    std::string str;
    std::vector<Json> arr;
    std::map<std::string, Json> obj;

    bool has(const std::string& k) const { return type == Obj && obj.count(k); }
    const Json& operator[](const std::string& k) const {
        static const Json null;
        auto it = obj.find(k);
        return (type == Obj && it != obj.end()) ? it->second : null;
    }
    const Json& operator[](size_t i) const {
        static const Json null;
        return (type == Arr && i < arr.size()) ? arr[i] : null;
    }
    size_t size() const { return type == Arr ? arr.size() : 0; }
    double number(double def) const { return type == Num ? num : def; }
    /* a number outside int's range (1e30, NaN) is not an index: the default, not an undefined conversion */
    int integer(int def) const { return type == Num && num >= -2147483648.0 && num <= 2147483647.0 ? (int)num : def; }
    std::string string(const std::string& def = "") const { return type == Str ? str : def; }

    static bool parse(const std::string& text, Json& out, std::string& err) {
        Parser p{text.data(), text.data() + text.size()};
        if(!p.value(out)) {
            err = "JSON parse error near byte " + std::to_string(p.cur - text.data());
            return false;
        }
        return true;
    }

private:
    struct Parser {
        const char* cur;
        const char* end;
        int depth = 0; /* nesting of the value being read: glTF needs about ten levels; a file of 100,000 '[' must not
                          recurse until the stack ends */
        static constexpr int kMaxDepth = 200;
        void ws() {
            while(cur < end && (*cur == ' ' || *cur == '\n' || *cur == '\t' || *cur == '\r')) cur++;
        }
        bool lit(const char* s) {
            const char* c = cur;
            while(*s) {
                if(c >= end || *c != *s) return false;
                c++, s++;
            }
            cur = c;
            return true;
        }
        static void utf8(std::string& s, unsigned cp) {
            if(cp < 0x80) s += (char)cp;
            else if(cp < 0x800) s += (char)(0xC0 | (cp >> 6)), s += (char)(0x80 | (cp & 0x3F));
            else
                s += (char)(0xE0 | (cp >> 12)), s += (char)(0x80 | ((cp >> 6) & 0x3F)),
                    s += (char)(0x80 | (cp & 0x3F));
        }
        bool string(std::string& s) {
            if(cur >= end || *cur != '"') return false;
            cur++;
            while(cur < end && *cur != '"') {
                if(*cur == '\\') {
                    if(++cur >= end) return false;
                    switch(*cur) {
                    case 'n': s += '\n'; break;
                    case 't': s += '\t'; break;
                    case 'r': s += '\r'; break;
                    case 'b': s += '\b'; break;
                    case 'f': s += '\f'; break;
                    case 'u': {
                        if(end - cur < 5) return false;
                        unsigned cp = (unsigned)strtoul(std::string(cur + 1, cur + 5).c_str(), nullptr, 16);
                        utf8(s, cp);
                        cur += 4;
                    } break;
                    default: s += *cur;
                    }
                    cur++;
                } else
                    s += *cur++;
            }
            if(cur >= end) return false;
            cur++;
            return true;
        }
        bool value(Json& v) {
            if(depth >= kMaxDepth) return false;
            depth++;
            bool ok = value_body(v);
            depth--;
            return ok;
        }
        bool value_body(Json& v) {
            ws();
            if(cur >= end) return false;
            char c = *cur;
            if(c == '{') {
                v.type = Obj;
                cur++;
                ws();
                if(cur < end && *cur == '}') return cur++, true;
                for(;;) {
                    ws();
                    std::string k;
                    if(!string(k)) return false;
                    ws();
                    if(cur >= end || *cur++ != ':') return false;
                    if(!value(v.obj[k])) return false;
                    ws();
                    if(cur >= end) return false;
                    if(*cur == ',') { cur++; continue; }
                    if(*cur == '}') return cur++, true;
                    return false;
                }
            }
            if(c == '[') {
                v.type = Arr;
                cur++;
                ws();
                if(cur < end && *cur == ']') return cur++, true;
                for(;;) {
                    v.arr.emplace_back();
                    if(!value(v.arr.back())) return false;
                    ws();
                    if(cur >= end) return false;
                    if(*cur == ',') { cur++; continue; }
                    if(*cur == ']') return cur++, true;
                    return false;
                }
            }
            if(c == '"') return v.type = Str, string(v.str);
            if(lit("true")) return v.type = Bool, v.b = true, true;
            if(lit("false")) return v.type = Bool, v.b = false, true;
            if(lit("null")) return v.type = Null, true;
            char* e = nullptr;
            v.num = strtod(cur, &e);
            if(e == cur) return false;
            cur = e;
            v.type = Num;
            return true;
        }
    };
};

} // namespace gpurt
