// txh_geojson.cpp -- single-pass scanner of NHD flowline GeoJSON (host only).
//
// The reference's load_nhd_geojson (tx_fast_hydrology/muskingum.py:877-917) json.load()s the file and walks the
// features in Python list comprehensions: COMID, toCOMID, Shape_Length per feature, then an id -> index join.  At
// CONUS size (2.7M features, ~GBs of geometry) that is minutes of interpreter time before the first routing step.
// This scanner walks the JSON once, keeps only the three attributes (geometry is skipped, never materialised) and
// does the id -> index join with one sort.  Structure handled: {"features": [ {"attributes": {...}, "geometry":
// {...}}, ... ]} with keys in any order, numbers as integers or floats, a missing / null / non-numeric toCOMID
// (= outlet, muskingum.py:897-902).
#include <algorithm>
#include <cerrno>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <numeric>
#include <string>
#include <vector>

#include "../../include/txh.h"

int txh_set_error_(int code, const char* msg);      // txh_capi.cu: sets the calling thread's txh_last_error()

namespace {

struct Scanner {
    const char* p;
    const char* end;
    std::string err;

    void ws() { while (p < end && (*p == ' ' || *p == '\n' || *p == '\t' || *p == '\r')) ++p; }
    bool fail(const char* what) { if (err.empty()) err = std::string("GeoJSON: ") + what; return false; }
    bool expect(char c) { ws(); if (p < end && *p == c) { ++p; return true; } return fail("unexpected character"); }
    // string token; out (optional) receives the raw bytes between the quotes (escapes left as they are)
    bool string(std::string* out)
    {
        ws();
        if (p >= end || *p != '"') return fail("string expected");
        const char* s = ++p;
        while (p < end && *p != '"') p += (*p == '\\' && p + 1 < end) ? 2 : 1;
        if (p >= end) return fail("unterminated string");
        if (out) out->assign(s, p - s);
        ++p;
        return true;
    }
    bool skip_value()
    {
        ws();
        if (p >= end) return fail("value expected");
        if (*p == '"') return string(nullptr);
        if (*p == '{' || *p == '[') {
            int depth = 0;
            while (p < end) {
                const char c = *p;
                if (c == '"') { if (!string(nullptr)) return false; continue; }
                if (c == '{' || c == '[') ++depth;
                else if (c == '}' || c == ']') { if (--depth == 0) { ++p; return true; } }
                ++p;
            }
            return fail("unterminated container");
        }
        while (p < end && *p != ',' && *p != '}' && *p != ']') ++p;          // number / true / false / null
        return true;
    }
    // number -> (ok, value); anything else is skipped and reported as "absent"
    bool number(bool* present, double* v, long long* iv, bool* is_int)
    {
        ws();
        *present = false;
        if (p < end && (*p == '-' || (*p >= '0' && *p <= '9'))) {
            char* e1 = nullptr; char* e2 = nullptr;
            errno = 0;
            const long long i = strtoll(p, &e1, 10);
            const double d = strtod(p, &e2);
            if (e2 == p) return fail("bad number");
            *present = true; *v = d;
            *is_int = (e1 == e2) && errno == 0;
            *iv = *is_int ? i : (long long)d;
            p = e2;
            return true;
        }
        return skip_value();
    }
};

// what the sizing call (capacity = 0) parsed, kept for the call that fetches it: the file is read once
struct Parsed {
    std::string path;
    std::vector<int64_t> ids, to;
    std::vector<uint8_t> has_to;
    std::vector<double> len;
};
thread_local Parsed g_parsed;

}  // namespace

extern "C" int txh_scan_nhd_geojson(const char* path, int64_t capacity, int64_t* comid, int64_t* endnodes, double* shape_length,
                                    int64_t* count)
{
    if (!path || !count) return txh_set_error_(TXH_E_INVALID, "null argument");
    std::vector<int64_t> ids, to;
    std::vector<uint8_t> has_to;
    std::vector<double> len;
    bool found = false;
    if (!g_parsed.path.empty() && g_parsed.path == path && capacity >= (int64_t)g_parsed.ids.size() && comid && endnodes &&
        shape_length) {
        ids.swap(g_parsed.ids); to.swap(g_parsed.to); has_to.swap(g_parsed.has_to); len.swap(g_parsed.len);
        g_parsed.path.clear();
        found = true;
    } else {
    g_parsed = Parsed();
    FILE* fp = fopen(path, "rb");
    if (!fp) return txh_set_error_(TXH_E_INVALID, (std::string("cannot open ") + path).c_str());
    std::vector<char> buf;
    fseek(fp, 0, SEEK_END);
    const long sz = ftell(fp);
    fseek(fp, 0, SEEK_SET);
    buf.resize(sz > 0 ? (size_t)sz + 1 : 1);
    const size_t got = sz > 0 ? fread(buf.data(), 1, (size_t)sz, fp) : 0;
    fclose(fp);
    buf[got] = 0;
    Scanner sc{buf.data(), buf.data() + got, {}};
    std::string key;
    if (!sc.expect('{')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
    for (;;) {                                                        // top-level members
        sc.ws();
        if (sc.p < sc.end && *sc.p == '}') break;
        if (!sc.string(&key) || !sc.expect(':')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
        if (key != "features") { if (!sc.skip_value()) return txh_set_error_(TXH_E_INVALID, sc.err.c_str()); }
        else {
            found = true;
            if (!sc.expect('[')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
            sc.ws();
            if (sc.p < sc.end && *sc.p == ']') ++sc.p;
            else for (;;) {                                           // features
                if (!sc.expect('{')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                long long id = 0, t = 0; double sl = 0.0;
                bool has_id = false, has_t = false, has_len = false;
                sc.ws();
                if (sc.p < sc.end && *sc.p == '}') ++sc.p;
                else for (;;) {                                       // members of a feature
                    if (!sc.string(&key) || !sc.expect(':')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                    if (key != "attributes") { if (!sc.skip_value()) return txh_set_error_(TXH_E_INVALID, sc.err.c_str()); }
                    else {
                        if (!sc.expect('{')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                        sc.ws();
                        if (sc.p < sc.end && *sc.p == '}') ++sc.p;
                        else for (;;) {
                            if (!sc.string(&key) || !sc.expect(':')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                            bool pr = false, isint = false; double dv = 0.0; long long iv = 0;
                            if (key == "COMID") { if (!sc.number(&pr, &dv, &iv, &isint)) return txh_set_error_(TXH_E_INVALID, sc.err.c_str()); if (pr) { id = iv; has_id = true; } }
                            else if (key == "toCOMID") { if (!sc.number(&pr, &dv, &iv, &isint)) return txh_set_error_(TXH_E_INVALID, sc.err.c_str()); if (pr) { t = iv; has_t = true; } }
                            else if (key == "Shape_Length") { if (!sc.number(&pr, &dv, &iv, &isint)) return txh_set_error_(TXH_E_INVALID, sc.err.c_str()); if (pr) { sl = dv; has_len = true; } }
                            else if (!sc.skip_value()) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                            sc.ws();
                            if (sc.p < sc.end && *sc.p == ',') { ++sc.p; continue; }
                            if (!sc.expect('}')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                            break;
                        }
                    }
                    sc.ws();
                    if (sc.p < sc.end && *sc.p == ',') { ++sc.p; continue; }
                    if (!sc.expect('}')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                    break;
                }
                if (!has_id) return txh_set_error_(TXH_E_INVALID, "GeoJSON: a feature has no numeric COMID");
                if (!has_len) return txh_set_error_(TXH_E_INVALID, "GeoJSON: a feature has no numeric Shape_Length");
                ids.push_back(id); to.push_back(t); has_to.push_back(has_t); len.push_back(sl);
                sc.ws();
                if (sc.p < sc.end && *sc.p == ',') { ++sc.p; continue; }
                if (!sc.expect(']')) return txh_set_error_(TXH_E_INVALID, sc.err.c_str());
                break;
            }
        }
        sc.ws();
        if (sc.p < sc.end && *sc.p == ',') { ++sc.p; continue; }
        break;
    }
    }
    if (!found) return txh_set_error_(TXH_E_INVALID, "GeoJSON: no \"features\" member");
    const int64_t n = (int64_t)ids.size();
    *count = n;
    if (capacity < n || !comid || !endnodes || !shape_length) {                   // sizing call: keep the parse
        g_parsed.path = path;
        g_parsed.ids.swap(ids); g_parsed.to.swap(to); g_parsed.has_to.swap(has_to); g_parsed.len.swap(len);
        return TXH_OK;
    }
    // id -> index join: first occurrence of an id wins (as a lookup in a unique pandas index does)
    std::vector<int64_t> order(n);
    std::iota(order.begin(), order.end(), 0);
    std::stable_sort(order.begin(), order.end(), [&](int64_t a, int64_t b) { return ids[a] < ids[b]; });
    for (int64_t i = 0; i < n; ++i) {
        comid[i] = ids[i]; shape_length[i] = len[i];
        int64_t e = i;                                                // no downstream feature: an outlet (self-loop)
        if (has_to[i]) {
            auto it = std::lower_bound(order.begin(), order.end(), to[i], [&](int64_t a, long long v) { return ids[a] < v; });
            if (it != order.end() && ids[*it] == to[i]) e = *it;
        }
        endnodes[i] = e;
    }
    return TXH_OK;
}
