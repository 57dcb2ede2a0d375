// Host-side ingest of the C ABI (include/impop_b200.h, "Ingest"): GFA v1 text of one window -> bit-packed
// haplotype x node presence matrix, node lengths, path names (and optional multiset coverage counts).
//
// Replaces the text hand-off in front of the similarity tool: the reference extracts a window graph and lets
// `odgi similarity -i tmp.gfa` walk its paths (run_pica2_odgi.sh:60-96); `impg similarity -r REGION` does the
// same from alignments (run_h-fst.sh:65-67).  One matrix row per path line (P or W), as `odgi similarity` with
// no grouping flags makes one group per path; node k = k-th S line.  Plain C++ on the host: no CUDA in here.
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>

#include <string>
#include <vector>

#include "../../include/impop_b200.h"

namespace {

struct Line {
    const char *p, *e;   // [p, e): one line without its terminator
};

// Next line of [cur, end); false at the end.  Strips a trailing '\r'.
inline bool next_line(const char *&cur, const char *end, Line &ln) {
    if (cur >= end) return false;
    const char *nl = (const char *)memchr(cur, '\n', (size_t)(end - cur));
    const char *e = nl ? nl : end;
    ln.p = cur;
    ln.e = (e > cur && e[-1] == '\r') ? e - 1 : e;
    cur = nl ? nl + 1 : end;
    return true;
}

// k-th tab-separated field of a line (0-based); false if the line has fewer fields.
inline bool field(const Line &ln, int k, const char *&fp, const char *&fe) {
    const char *p = ln.p;
    for (int i = 0; i < k; ++i) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(ln.e - p));
        if (!t) return false;
        p = t + 1;
    }
    const char *t = (const char *)memchr(p, '\t', (size_t)(ln.e - p));
    fp = p;
    fe = t ? t : ln.e;
    return true;
}

inline uint64_t hash_bytes(const char *p, size_t n) {   // FNV-1a, finalised
    uint64_t h = 1469598103934665603ull;
    for (size_t i = 0; i < n; ++i) { h ^= (unsigned char)p[i]; h *= 1099511628211ull; }
    h ^= h >> 29; h *= 0xBF58476D1CE4E5B9ull; h ^= h >> 32;
    return h;
}

// Open-addressing map segment name -> node index (names are slices of the caller's text).
struct SegMap {
    struct Slot { const char *p; uint32_t len; int32_t idx; };
    std::vector<Slot> slots;
    uint64_t mask = 0;
    void init(size_t n) {
        size_t cap = 16;
        while (cap < 2 * n + 2) cap <<= 1;
        slots.assign(cap, Slot{nullptr, 0u, -1});
        mask = cap - 1;
    }
    bool insert(const char *p, size_t len, int32_t idx) {       // false: duplicate name
        uint64_t h = hash_bytes(p, len) & mask;
        while (slots[h].p) {
            if (slots[h].len == len && memcmp(slots[h].p, p, len) == 0) return false;
            h = (h + 1) & mask;
        }
        slots[h] = Slot{p, (uint32_t)len, idx};
        return true;
    }
    int32_t find(const char *p, size_t len) const {
        uint64_t h = hash_bytes(p, len) & mask;
        while (slots[h].p) {
            if (slots[h].len == len && memcmp(slots[h].p, p, len) == 0) return slots[h].idx;
            h = (h + 1) & mask;
        }
        return -1;
    }
};

// Segment name -> node index.  Graph builders number their segments, so when EVERY name is a canonical decimal
// integer (no sign, no leading zero, at most 9 digits) and the numbers are dense enough, a step is resolved by
// parsing its digits and indexing an array; any other file goes through the hash map.  A step token that is not a
// canonical integer cannot equal any name of such a file, so it is reported as undefined either way.
struct SegIndex {
    SegMap map;
    std::vector<int32_t> direct;
    uint32_t lo = 0xFFFFFFFFu, hi = 0;
    bool numeric = true, use_direct = false;
    static inline bool canonical(const char *p, size_t len, uint32_t &v) {
        if (len == 0 || len > 9 || (len > 1 && p[0] == '0')) return false;
        uint32_t x = 0;
        for (size_t i = 0; i < len; ++i) {
            const unsigned d = (unsigned)(p[i] - '0');
            if (d > 9u) return false;
            x = x * 10u + d;
        }
        v = x;
        return true;
    }
    void init(size_t n) { map.init(n); }
    bool insert(const char *p, size_t len, int32_t idx) {
        uint32_t v;
        if (numeric && canonical(p, len, v)) { lo = v < lo ? v : lo; hi = v > hi ? v : hi; } else numeric = false;
        return map.insert(p, len, idx);
    }
    void finish(size_t n) {                                      // after the last insert
        if (!numeric || n == 0 || (uint64_t)(hi - lo) + 1 > 8ull * n + 1024) return;
        direct.assign((size_t)(hi - lo) + 1, -1);
        for (const SegMap::Slot &sl : map.slots)
            if (sl.p) { uint32_t v = 0; canonical(sl.p, sl.len, v); direct[v - lo] = sl.idx; }
        use_direct = true;
    }
    inline int32_t find(const char *p, size_t len) const {
        if (use_direct) {
            uint32_t v;
            if (!canonical(p, len, v) || v < lo || v > hi) return -1;
            return direct[v - lo];
        }
        return map.find(p, len);
    }
};

// Length of a segment: its sequence, or the LN:i: tag when the sequence is '*'.
inline bool segment_length(const Line &ln, uint64_t &len) {
    const char *sp, *se;
    if (!field(ln, 2, sp, se)) return false;
    if (!(se - sp == 1 && *sp == '*')) { len = (uint64_t)(se - sp); return true; }
    for (int k = 3;; ++k) {
        const char *tp, *te;
        if (!field(ln, k, tp, te)) break;
        if (te - tp > 5 && memcmp(tp, "LN:i:", 5) == 0) {
            uint64_t v = 0;
            for (const char *q = tp + 5; q < te; ++q) {
                if (*q < '0' || *q > '9') return false;
                v = v * 10 + (uint64_t)(*q - '0');
                if (v > 0xFFFFFFFFull) return false;
            }
            len = v;
            return true;
        }
    }
    len = 0;      // '*' without LN: length unknown, counts as 0 (contributes nothing to any statistic)
    return true;
}

// Calls f(name_begin, name_end) for every step of a P line's segment list ("11+,12-,...") or a W line's walk
// (">11<12..."); false on a malformed step.
template <typename F>
inline bool for_each_step(char kind, const char *p, const char *e, F f) {
    if (kind == 'P') {
        if (e - p == 1 && *p == '*') return true;
        while (p < e) {
            const char *se = p;
            while (se < e && *se != ',') ++se;                  // steps are a few bytes long: a plain loop beats memchr
            if (se - p < 2 || (se[-1] != '+' && se[-1] != '-')) return false;
            if (!f(p, se - 1)) return false;
            p = se < e ? se + 1 : e;
        }
        return true;
    }
    if (e - p == 1 && *p == '*') return true;
    while (p < e) {
        if (*p != '>' && *p != '<') return false;
        const char *q = p + 1;
        while (q < e && *q != '>' && *q != '<') ++q;
        if (q == p + 1) return false;
        if (!f(p + 1, q)) return false;
        p = q;
    }
    return true;
}

// The row name of a path line.  P: the path name as it stands.  W: sample#hap#seqid[:start-end] (PanSN).
inline bool path_name(char kind, const Line &ln, std::string &out) {
    const char *a, *b;
    out.clear();
    if (kind == 'P') {
        if (!field(ln, 1, a, b)) return false;
        out.assign(a, b);
        return true;
    }
    for (int k = 1; k <= 3; ++k) {
        if (!field(ln, k, a, b)) return false;
        if (k > 1) out.push_back('#');
        out.append(a, b);
    }
    const char *s0, *s1, *e0, *e1;
    if (!field(ln, 4, s0, s1) || !field(ln, 5, e0, e1)) return false;
    if (!(s1 - s0 == 1 && *s0 == '*') && !(e1 - e0 == 1 && *e0 == '*')) {
        out.push_back(':'); out.append(s0, s1); out.push_back('-'); out.append(e0, e1);
    }
    return true;
}

inline int steps_field(char kind) { return kind == 'P' ? 2 : 6; }

inline uint64_t text_hash(const char *text, int64_t bytes) {
    uint64_t h0 = 0x9E3779B97F4A7C15ull ^ (uint64_t)bytes, h1 = 0xC2B2AE3D27D4EB4Full, h2 = 0x165667B19E3779F9ull, h3 = 0x27D4EB2F165667C5ull;
    int64_t i = 0;
    for (; i + 32 <= bytes; i += 32) {                           // four independent lanes: the multiplies overlap
        uint64_t w[4];
        memcpy(w, text + i, 32);
        h0 = (h0 ^ w[0]) * 0xD6E8FEB86659FD93ull; h0 ^= h0 >> 29;
        h1 = (h1 ^ w[1]) * 0xD6E8FEB86659FD93ull; h1 ^= h1 >> 29;
        h2 = (h2 ^ w[2]) * 0xD6E8FEB86659FD93ull; h2 ^= h2 >> 29;
        h3 = (h3 ^ w[3]) * 0xD6E8FEB86659FD93ull; h3 ^= h3 >> 29;
    }
    uint64_t h = h0 ^ (h1 * 0x9E3779B97F4A7C15ull) ^ (h2 * 0xBF58476D1CE4E5B9ull) ^ (h3 * 0x94D049BB133111EBull);
    for (; i < bytes; ++i) h = (h ^ (unsigned char)text[i]) * 0x100000001B3ull;
    return h ^ (h >> 31);
}

// What impop_gfa_scan found, kept for the impop_gfa_fill that follows on the same thread with the same text (pointer,
// length and hash of the text must agree), so that the fill does not walk every step a third time.
struct GfaMemo {
    const char *text = nullptr;
    int64_t bytes = -1;
    uint64_t hash = 0;
    impop_gfa_info_t info;
};
thread_local GfaMemo g_gfa_memo;

}  // namespace

extern "C" {

int impop_gfa_scan(const char *text, int64_t bytes, impop_gfa_info_t *info) {
    if (!text || bytes < 0 || !info) return IMPOP_ERR_ARG;
    impop_gfa_info_t out = {0, 0, 0, 0, 0};
    g_gfa_memo.text = nullptr; g_gfa_memo.bytes = -1;
    const char *cur = text, *end = text + bytes;
    Line ln;
    std::string name;
    int64_t lineno = 0;
    while (next_line(cur, end, ln)) {
        ++lineno;
        if (ln.e - ln.p < 2 || ln.p[1] != '\t') continue;
        const char kind = ln.p[0];
        if (kind == 'S') {
            ++out.segments;
        } else if (kind == 'P' || kind == 'W') {
            if (!path_name(kind, ln, name)) { out.error_line = lineno; *info = out; return IMPOP_ERR_ARG; }
            const char *sp, *se;
            if (!field(ln, steps_field(kind), sp, se)) { out.error_line = lineno; *info = out; return IMPOP_ERR_ARG; }
            // steps by their separators (a loop the compiler vectorises); malformed steps are reported by impop_gfa_fill,
            // which walks them anyway
            int64_t steps = 0;
            if (!(se - sp == 1 && *sp == '*')) {
                if (kind == 'P') { steps = se > sp ? 1 : 0; for (const char *q = sp; q < se; ++q) steps += *q == ','; }
                else for (const char *q = sp; q < se; ++q) steps += (*q == '>') | (*q == '<');
            }
            ++out.paths;
            out.name_bytes += (int64_t)name.size() + 1;
            out.steps += steps;
        }
    }
    *info = out;
    g_gfa_memo.text = text; g_gfa_memo.bytes = bytes; g_gfa_memo.hash = text_hash(text, bytes); g_gfa_memo.info = out;
    return IMPOP_OK;
}

int impop_gfa_fill(const char *text, int64_t bytes, int32_t pitch_words, uint32_t *x_bits_host, uint32_t *node_len_host,
                   uint16_t *counts_host, char *names_host, int64_t *name_off_host, int64_t *error_line, int32_t *revisits) {
    if (error_line) *error_line = 0;
    if (revisits) *revisits = 0;
    bool revisit = false;
    if (!text || bytes < 0 || pitch_words < 0 || pitch_words % 4 != 0) return IMPOP_ERR_ARG;
    impop_gfa_info_t info;
    if (g_gfa_memo.text == text && g_gfa_memo.bytes == bytes && g_gfa_memo.hash == text_hash(text, bytes)) {
        info = g_gfa_memo.info;                                   // scanned (and validated) just before
        g_gfa_memo.text = nullptr; g_gfa_memo.bytes = -1;
    } else {
        int rc = impop_gfa_scan(text, bytes, &info);
        g_gfa_memo.text = nullptr; g_gfa_memo.bytes = -1;
        if (rc != IMPOP_OK) { if (error_line) *error_line = info.error_line; return rc; }
    }
    if ((int64_t)pitch_words * 32 < info.segments) return IMPOP_ERR_ARG;
    if ((info.paths && (!x_bits_host || !names_host || !name_off_host)) || (info.segments && !node_len_host)) return IMPOP_ERR_ARG;
    if (info.segments > 0x7FFFFFFFll || info.paths > 0x7FFFFFFFll) return IMPOP_ERR_RANGE;

    // pass 1: segments in file order
    SegIndex map;
    map.init((size_t)info.segments);
    const char *cur = text, *end = text + bytes;
    Line ln;
    int64_t lineno = 0;
    int32_t seg = 0;
    while (next_line(cur, end, ln)) {
        ++lineno;
        if (ln.e - ln.p < 2 || ln.p[1] != '\t' || ln.p[0] != 'S') continue;
        const char *np, *ne;
        uint64_t len = 0;
        if (!field(ln, 1, np, ne) || ne == np || !segment_length(ln, len) || len > 0xFFFFFFFFull ||
            !map.insert(np, (size_t)(ne - np), seg)) {
            if (error_line) *error_line = lineno;
            return IMPOP_ERR_ARG;
        }
        node_len_host[seg++] = (uint32_t)len;
    }
    map.finish((size_t)info.segments);
    // pass 2: path lines in file order
    if (info.paths) memset(x_bits_host, 0, sizeof(uint32_t) * (size_t)info.paths * (size_t)pitch_words);
    if (counts_host && info.paths) memset(counts_host, 0, sizeof(uint16_t) * (size_t)info.paths * (size_t)info.segments);
    cur = text; lineno = 0;
    int64_t row = 0, off = 0;
    std::string name;
    while (next_line(cur, end, ln)) {
        ++lineno;
        if (ln.e - ln.p < 2 || ln.p[1] != '\t') continue;
        const char kind = ln.p[0];
        if (kind != 'P' && kind != 'W') continue;
        path_name(kind, ln, name);
        name_off_host[row] = off;
        memcpy(names_host + off, name.c_str(), name.size() + 1);
        off += (int64_t)name.size() + 1;
        uint32_t *xr = x_bits_host + (size_t)row * (size_t)pitch_words;
        uint16_t *cr = counts_host ? counts_host + (size_t)row * (size_t)info.segments : nullptr;
        const char *sp, *se;
        if (!field(ln, steps_field(kind), sp, se)) { if (error_line) *error_line = lineno; return IMPOP_ERR_ARG; }
        // P line over numbered segments (what graph builders write): digits, sign and comma in ONE pass over the bytes, the
        // node by array index.  Anything unusual -- a token that is not a canonical number, a missing sign -- and the line is
        // done again by the general tokenizer below, which decides what is an error.
        bool done = false;
        if (kind == 'P' && map.use_direct && !(se - sp == 1 && *sp == '*')) {
            const char *p = sp;
            const uint32_t lo = map.lo, span = map.hi - map.lo;
            const int32_t *direct = map.direct.data();
            done = true;
            const bool revisit_before = revisit;
            int32_t cw = -1;                                          // presence word being filled (steps mostly ascend: one
            uint32_t cbits = 0u;                                      // load and one store per word instead of a read-modify-write per step)
            while (p < se) {
                uint32_t v = 0;
                const char *q = p;
                unsigned d;
                while (q < se && (d = (unsigned)(*q - '0')) <= 9u) { v = v * 10u + d; ++q; }
                const size_t nd = (size_t)(q - p);
                if (nd == 0 || nd > 9 || (nd > 1 && *p == '0') || q >= se || (*q != '+' && *q != '-') || (q + 1 < se && q[1] != ',')) { done = false; break; }
                const uint32_t rel = v - lo;
                const int32_t k = rel <= span ? direct[rel] : -1;
                if (k < 0) { done = false; break; }
                if ((k >> 5) != cw) {
                    if (cw >= 0) xr[cw] = cbits;
                    cw = k >> 5;
                    cbits = xr[cw];
                }
                revisit |= (cbits >> (k & 31)) & 1u;                  // the path has been here before
                cbits |= 1u << (k & 31);
                if (cr && cr[k] != 0xFFFFu) ++cr[k];
                p = q + 2;                                            // past the sign and the comma (or the end)
            }
            if (done && cw >= 0) xr[cw] = cbits;
            if (!done) {                                              // start the row over
                revisit = revisit_before;
                memset(xr, 0, sizeof(uint32_t) * (size_t)pitch_words);
                if (cr) memset(cr, 0, sizeof(uint16_t) * (size_t)info.segments);
            }
        }
        const bool ok = done || for_each_step(kind, sp, se, [&](const char *a, const char *b) {
            const int32_t k = map.find(a, (size_t)(b - a));
            if (k < 0) return false;                      // step over a segment the file does not define
            revisit |= (xr[k >> 5] >> (k & 31)) & 1u;
            xr[k >> 5] |= 1u << (k & 31);
            if (cr && cr[k] != 0xFFFFu) ++cr[k];
            return true;
        });
        if (!ok) { if (error_line) *error_line = lineno; return IMPOP_ERR_ARG; }
        ++row;
    }
    if (info.paths) name_off_host[row] = off;
    if (revisits) *revisits = revisit ? 1 : 0;
    return IMPOP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// All-pairs similarity table (the TSV `odgi similarity` / `impg similarity` print and pica2.py:6-58,
// h-fst.py:84-119 parse with csv.DictReader -- 47-92 % of those scripts' run time at 466 haplotypes, SURVEY 8 a-1 /
// a-4) -> sorted names + dense identity matrix (NaN = pair absent), last row of a repeated pair wins (pica2.py:44).
// Only MACHINE-CLEAN text is taken here: ASCII, no quote characters, every row as wide as the header needs, numbers
// in plain decimal / exponent form or nan / inf.  Anything else reports status 1 and the caller uses the general
// reader, so that the reference's behaviour on odd input (csv quoting, Python's float() grammar, its error
// messages) is reproduced by the code that mirrors it line by line.
// ------------------------------------------------------------------------------------------------------------
namespace {

struct TsvCols { int a = -1, b = -1, v = -1, need = 0; };

inline bool tsv_header(const Line &ln, TsvCols &c) {
    int k = 0, na = 0, nb = 0, nv = 0;
    const char *p = ln.p;
    while (true) {
        const char *t = (const char *)memchr(p, '\t', (size_t)(ln.e - p));
        const char *fe = t ? t : ln.e;
        const size_t len = (size_t)(fe - p);
        if (len == 7 && memcmp(p, "group.a", 7) == 0) { c.a = k; ++na; }
        else if (len == 7 && memcmp(p, "group.b", 7) == 0) { c.b = k; ++nb; }
        else if (len == 18 && memcmp(p, "estimated.identity", 18) == 0) { c.v = k; ++nv; }
        ++k;
        if (!t) break;
        p = t + 1;
    }
    c.need = 1 + (c.a > c.b ? (c.a > c.v ? c.a : c.v) : (c.b > c.v ? c.b : c.v));
    return na == 1 && nb == 1 && nv == 1;
}

// Plain decimal / exponent number or [+-]nan / inf / infinity, nothing else (no spaces, no hex, no underscores).
inline bool tsv_number(const char *p, const char *e, double &out) {
    if (p == e || e - p > 64) return false;
    const char *q = p;
    if (*q == '+' || *q == '-') ++q;
    if (q == e) return false;
    auto ieq = [&](const char *w) {
        size_t n = strlen(w);
        if ((size_t)(e - q) != n) return false;
        for (size_t i = 0; i < n; ++i) if ((q[i] | 0x20) != w[i]) return false;
        return true;
    };
    if (!(ieq("nan") || ieq("inf") || ieq("infinity"))) {
        int digits = 0;
        while (q < e && *q >= '0' && *q <= '9') { ++q; ++digits; }
        if (q < e && *q == '.') { ++q; while (q < e && *q >= '0' && *q <= '9') { ++q; ++digits; } }
        if (digits == 0) return false;
        if (q < e && (*q == 'e' || *q == 'E')) {
            ++q;
            if (q < e && (*q == '+' || *q == '-')) ++q;
            int ed = 0;
            while (q < e && *q >= '0' && *q <= '9') { ++q; ++ed; }
            if (ed == 0) return false;
        }
        if (q != e) return false;
    }
    char buf[72];
    memcpy(buf, p, (size_t)(e - p));
    buf[e - p] = 0;
    char *end = nullptr;
    out = strtod(buf, &end);                 // glibc: correctly rounded, as CPython's float()
    return end == buf + (e - p);
}

struct NameMap {                              // name -> first-seen index
    SegMap map;
    std::vector<Line> names;
    void init(size_t cap) { map.init(cap); }
    int32_t get(const char *p, const char *e) {
        int32_t k = map.find(p, (size_t)(e - p));
        if (k >= 0) return k;
        if (2 * (names.size() + 2) > map.slots.size()) {          // grow: rehash everything
            SegMap bigger;
            bigger.init(map.slots.size());
            for (size_t i = 0; i < names.size(); ++i) bigger.insert(names[i].p, (size_t)(names[i].e - names[i].p), (int32_t)i);
            map = bigger;
        }
        k = (int32_t)names.size();
        map.insert(p, (size_t)(e - p), k);
        names.push_back(Line{p, e});
        return k;
    }
};

// One pass over a clean table: calls row(ia, ib, value) per data row.  Returns 0 clean, 1 needs the general reader.
template <typename F>
int tsv_walk(const char *text, int64_t bytes, NameMap &nm, int64_t &rows, F row) {
    for (int64_t i = 0; i < bytes; ++i) {
        const unsigned char ch = (unsigned char)text[i];
        if (ch == '"' || ch >= 0x80 || ch == 0) return 1;
    }
    const char *cur = text, *end = text + bytes;
    Line ln;
    if (!next_line(cur, end, ln)) return 1;                       // empty file: the general reader words the error
    TsvCols c;
    if (!tsv_header(ln, c)) return 1;
    rows = 0;
    while (next_line(cur, end, ln)) {
        if (ln.e == ln.p) continue;                               // csv.DictReader skips blank lines
        if (memchr(ln.p, '\r', (size_t)(ln.e - ln.p))) return 1;
        const char *ap, *ae, *bp, *be, *vp, *ve, *xp, *xe;
        if (!field(ln, c.need - 1, xp, xe)) return 1;             // row narrower than the header needs
        field(ln, c.a, ap, ae); field(ln, c.b, bp, be); field(ln, c.v, vp, ve);
        double v;
        if (!tsv_number(vp, ve, v)) return 1;
        ++rows;
        row(nm.get(ap, ae), nm.get(bp, be), v);
    }
    return 0;
}

struct TsvRow { int32_t a, b; double v; };

// impop_tsv_scan keeps what it parsed for the impop_tsv_fill that follows on the same thread with the same text, so the
// table is walked (names hashed, numbers converted) once, not twice.  Per thread, no shared state; the fill trusts the
// memo only if pointer, length and a hash of the whole text agree, and parses again otherwise.
struct TsvMemo {
    const char *text = nullptr;
    int64_t bytes = -1;
    uint64_t hash = 0;
    std::vector<TsvRow> rows;
    std::vector<Line> names;                  // first-seen order, pointing into `text`
};
thread_local TsvMemo g_tsv_memo;

}  // namespace

extern "C" {

int impop_tsv_scan(const char *text, int64_t bytes, impop_tsv_info_t *info) {
    if (!text || bytes < 0 || !info) return IMPOP_ERR_ARG;
    impop_tsv_info_t out = {0, 0, 0, 0, 0};
    NameMap nm;
    nm.init(1024);
    TsvMemo &memo = g_tsv_memo;
    memo.text = nullptr; memo.bytes = -1; memo.rows.clear(); memo.names.clear();
    int64_t rows = 0;
    out.status = tsv_walk(text, bytes, nm, rows, [&](int32_t a, int32_t b, double v) { memo.rows.push_back(TsvRow{a, b, v}); });
    if (out.status == 0) {
        out.rows = rows;
        out.names = (int64_t)nm.names.size();
        for (const Line &s : nm.names) out.name_bytes += (int64_t)(s.e - s.p) + 1;
        memo.names = nm.names;
        memo.text = text; memo.bytes = bytes; memo.hash = text_hash(text, bytes);
    } else {
        memo.rows.clear();
    }
    *info = out;
    return IMPOP_OK;
}

int impop_tsv_fill(const char *text, int64_t bytes, double *matrix_host, char *names_host, int64_t *name_off_host) {
    if (!text || bytes < 0 || !name_off_host) return IMPOP_ERR_ARG;
    NameMap nm;
    std::vector<TsvRow> rows;
    TsvMemo &memo = g_tsv_memo;
    if (memo.text == text && memo.bytes == bytes && memo.hash == text_hash(text, bytes)) {
        rows.swap(memo.rows);                                     // parsed by the impop_tsv_scan just before
        nm.names.swap(memo.names);
        memo.text = nullptr; memo.bytes = -1;
    } else {
        nm.init(1024);
        int64_t count = 0;
        if (tsv_walk(text, bytes, nm, count, [&](int32_t a, int32_t b, double v) { rows.push_back(TsvRow{a, b, v}); }) != 0)
            return IMPOP_ERR_ARG;
    }
    const size_t n = nm.names.size();
    if (n && (!matrix_host || !names_host)) return IMPOP_ERR_ARG;
    // names in byte order (= Python's str order on ASCII), first-seen index -> sorted rank
    std::vector<int32_t> order(n), rank(n);
    for (size_t i = 0; i < n; ++i) order[i] = (int32_t)i;
    std::sort(order.begin(), order.end(), [&](int32_t x, int32_t y) {
        const Line &a = nm.names[x], &b = nm.names[y];
        const size_t la = (size_t)(a.e - a.p), lb = (size_t)(b.e - b.p);
        const int c = memcmp(a.p, b.p, la < lb ? la : lb);
        return c != 0 ? c < 0 : la < lb;
    });
    int64_t off = 0;
    for (size_t r = 0; r < n; ++r) {
        rank[order[r]] = (int32_t)r;
        const Line &s = nm.names[order[r]];
        name_off_host[r] = off;
        memcpy(names_host + off, s.p, (size_t)(s.e - s.p));
        names_host[off + (s.e - s.p)] = 0;
        off += (int64_t)(s.e - s.p) + 1;
    }
    name_off_host[n] = off;
    const double nan = __builtin_nan("");
    for (size_t i = 0; i < n * n; ++i) matrix_host[i] = nan;
    for (const TsvRow &r : rows) {                                // file order: the last row of a pair stays
        const size_t i = (size_t)rank[r.a], j = (size_t)rank[r.b];
        matrix_host[i * n + j] = r.v;
        matrix_host[j * n + i] = r.v;
    }
    return IMPOP_OK;
}

}  // extern "C"

// ------------------------------------------------------------------------------------------------------------
// Column compaction of a batch of windows (ingest, once per window; exact).
//   * nodes visited by EVERY haplotype of the window (the backbone of a pangenome window graph: a third of the nodes
//     of an HPRC-shaped window) add the same constant to every intersection I_ij and path length A_i: they are
//     merged into ONE node whose length is the sum of theirs (flags == 0) or into the window constant C (affine form);
//   * nodes visited by no haplotype, and nodes of length 0, contribute nothing and are dropped;
//   * the remaining nodes are ordered by length (ties: original order), so that the presence words of unit-length
//     nodes (SNP alleles) come first: prep_rows then needs one popcount per 32 such nodes and row.
// IMPOP_COMPACT_PAIRS (affine form, include/impop_b200.h):
//   * nodes with IDENTICAL presence columns (variants in perfect linkage) are merged into one column of their summed
//     length that counts for all of them in S;
//   * two columns r, a that are COMPLEMENTARY over the window's rows (x_r + x_a = 1 for every haplotype: the two
//     branches of a bi-allelic bubble, the commonest shape in a variation graph) become ONE column: with x_r = 1 - x_a,
//         len_r x_ri x_rj + len_a x_ai x_aj = len_r - len_r x_ai - len_r x_aj + (len_r + len_a) x_ai x_aj ,
//     i.e. the kept column a gets the weight len_r + len_a, len_r goes into the window constant C and len_r x_ai into the
//     row term R_i:  I_ij = sum_k len'_k x'_ik x'_jk + C - R_i - R_j  (exact integers), A_i = I_ii.
// IMPOP_COMPACT_REPLICATE: a weight >= 255 is spread over ceil(w / 254) copies of its column (byte weights, no separate
//     heavy chunk on the device) when the copies fit into the padding of the window's 128-column chunks.
// I, A, U, the segregating-node count and hence every statistic are unchanged.  The similarity tools do the equivalent
// implicitly: they walk paths, not matrix columns.  Fewer columns = fewer operand bytes to expand and multiply, and
// fewer bytes to upload.
// ------------------------------------------------------------------------------------------------------------
#include <mutex>
#include <thread>
#include <unordered_map>

namespace {

struct CompactCol {
    int32_t src;        // original column whose presence bits the output column carries
    uint32_t w;         // weight (node length; merged: sum)
    uint8_t mult;       // nodes of positive length the column stands for in S (0: replica)
};

struct CompactPlan {
    std::vector<CompactCol> cols;                          // output columns, output order
    std::vector<std::pair<int32_t, uint32_t>> radj;        // (kept column, removed length): R_i += length when the row carries the column
    uint64_t c = 0;                  // window constant C
    uint64_t total = 0;              // sum of the lengths of all visited nodes (must stay below 2^31)
    int32_t m_out = 0;
    int64_t site_runs = 0;           // runs of segregating nodes between nodes every row visits (original node order)
    bool bad = false;                // a multiplicity or weight does not fit
};

inline uint64_t mix64(uint64_t z) {
    z += 0x9e3779b97f4a7c15ull;
    z = (z ^ (z >> 30)) * 0xbf58476d1ce4e5b9ull;
    z = (z ^ (z >> 27)) * 0x94d049bb133111ebull;
    return z ^ (z >> 31);
}

// 32 x 32 bit transpose, least significant bit first: out[c] bit r = in[r] bit c.
inline void transpose32(uint32_t A[32]) {
    uint32_t m = 0x0000FFFFu;
    for (int j = 16; j != 0; j >>= 1, m ^= (m << j))
        for (int k = 0; k < 32; k = (k + j + 1) & ~j) {
            const uint32_t t = ((A[k] >> j) ^ A[k + j]) & m;
            A[k] ^= t << j;
            A[k + j] ^= t;
        }
}

// The window column by column: column k = nb words, bit (i & 31) of word i >> 5 = haplotype i carries node k.
// (The compaction compares, merges and gathers COLUMNS; the caller's matrix is row-major.)
struct ColMajor {
    int nb = 0;
    uint32_t last_mask = 0xffffffffu;           // valid rows of the last word
    std::vector<uint32_t> w;
    const uint32_t *col(int32_t k) const { return w.data() + (size_t)k * nb; }
    uint32_t mask(int i) const { return i == nb - 1 ? last_mask : 0xffffffffu; }
};

void to_columns(int32_t n, int32_t m, int32_t pitch, const uint32_t *x, ColMajor &cm) {
    const int words = (m + 31) / 32;
    cm.nb = (n + 31) / 32;
    cm.last_mask = (n & 31) ? ((1u << (n & 31)) - 1u) : 0xffffffffu;
    cm.w.assign((size_t)words * 32 * (size_t)cm.nb, 0u);
    uint32_t A[32];
    for (int rb = 0; rb < cm.nb; ++rb) {
        const int r0 = rb * 32, nr = std::min(32, n - r0);
        for (int wc = 0; wc < words; ++wc) {
            for (int r = 0; r < nr; ++r) A[r] = x[(size_t)(r0 + r) * pitch + wc];
            for (int r = nr; r < 32; ++r) A[r] = 0u;
            transpose32(A);
            for (int c = 0; c < 32; ++c) cm.w[((size_t)wc * 32 + c) * cm.nb + rb] = A[c];
        }
    }
}

void compact_plan(int32_t n, int32_t m, int32_t pitch, const uint32_t *x, const uint32_t *len, uint32_t flags, CompactPlan &pl,
                  ColMajor &cm) {
    to_columns(n, m, pitch, x, cm);
    const int nb = cm.nb;
    pl.cols.clear(); pl.radj.clear();
    pl.c = 0; pl.total = 0; pl.site_runs = 0; pl.bad = false;
    uint64_t const_len = 0;
    std::vector<int32_t> var;                              // variable columns of positive length, original order
    bool in_run = false;
    for (int32_t k = 0; k < m; ++k) {
        if (len[k] == 0u) continue;
        const uint32_t *c = cm.col(k);
        bool any = false, all = n > 0;
        for (int i = 0; i < nb; ++i) { any |= c[i] != 0u; all &= c[i] == cm.mask(i); }
        if (all) { const_len += len[k]; pl.total += len[k]; in_run = false; }
        else if (any) {
            var.push_back(k);
            pl.total += len[k];
            if (!in_run) { ++pl.site_runs; in_run = true; }
        }
    }
    const bool affine = (flags & IMPOP_COMPACT_PAIRS) != 0u;
    if (!affine) {
        for (int32_t k : var) pl.cols.push_back(CompactCol{k, len[k], 1});
    } else {
        // column signatures: a hash of the column's words, and of its complement's; candidates are verified word by word
        auto sig = [&](const uint32_t *c, bool complement) {
            uint64_t h = 0x243f6a8885a308d3ull;
            for (int i = 0; i < nb; ++i) h = mix64(h ^ (uint64_t)(complement ? (~c[i] & cm.mask(i)) : c[i]));
            return h;
        };
        auto related = [&](int32_t a, int32_t b, bool complement) {
            const uint32_t *ca = cm.col(a), *cb = cm.col(b);
            for (int i = 0; i < nb; ++i)
                if ((ca[i] ^ cb[i]) != (complement ? cm.mask(i) : 0u)) return false;
            return true;
        };
        struct Group { int32_t rep; uint64_t len; uint32_t cnt; bool used; };
        std::vector<Group> groups;
        std::unordered_map<uint64_t, int32_t> by_hash;
        by_hash.reserve(var.size() * 2);
        for (int32_t k : var) {
            const uint64_t h = sig(cm.col(k), false);
            auto it = by_hash.find(h);
            if (it != by_hash.end()) {
                Group &g = groups[it->second];
                const bool same = related(g.rep, k, false);
                if (same && g.cnt < 127u) { g.len += len[k]; ++g.cnt; continue; }
                groups.push_back(Group{k, len[k], 1u, false});
                if (same) it->second = (int32_t)groups.size() - 1;          // the group is full: further copies join its successor
                continue;                                                   // (same signature, different column: stays alone)
            }
            by_hash.emplace(h, (int32_t)groups.size());
            groups.push_back(Group{k, len[k], 1u, false});
        }
        pl.c = const_len;
        for (size_t gi = 0; gi < groups.size(); ++gi) {
            Group &g = groups[gi];
            if (g.used) continue;
            g.used = true;
            auto it = by_hash.find(sig(cm.col(g.rep), true));
            if (it != by_hash.end() && (size_t)it->second != gi) {
                Group &o = groups[it->second];
                if (!o.used && related(g.rep, o.rep, true)) {
                    o.used = true;
                    const uint64_t wsum = g.len + o.len;
                    if (wsum > 0xffffffffull) { pl.bad = true; return; }
                    pl.cols.push_back(CompactCol{g.rep, (uint32_t)wsum, (uint8_t)(g.cnt + o.cnt)});
                    pl.c += o.len;
                    pl.radj.emplace_back(g.rep, (uint32_t)o.len);
                    continue;
                }
            }
            if (g.len > 0xffffffffull) { pl.bad = true; return; }
            pl.cols.push_back(CompactCol{g.rep, (uint32_t)g.len, (uint8_t)g.cnt});
        }
    }
    std::stable_sort(pl.cols.begin(), pl.cols.end(), [](const CompactCol &a, const CompactCol &b) { return a.w < b.w; });
    if (!affine) {
        if (const_len > 0) {
            if (const_len > 0xffffffffull) { pl.bad = true; return; }
            pl.cols.push_back(CompactCol{-1, (uint32_t)const_len, 1});          // src -1: the all-ones column
        }
    } else if (flags & IMPOP_COMPACT_REPLICATE) {
        // weights >= 255 as copies of the column with byte weights <= 254, where that does not cost more 128-column
        // chunks than the device's heavy columns would (one entry per 255 * 255 of weight, padded to a chunk of their own)
        const int64_t base = (int64_t)pl.cols.size();
        int64_t extra_all = 0, entries = 0;
        size_t first_heavy = pl.cols.size();
        for (size_t j = 0; j < pl.cols.size(); ++j)
            if (pl.cols[j].w >= 255u) {
                if (first_heavy == pl.cols.size()) first_heavy = j;
                extra_all += (int64_t)((pl.cols[j].w + 253u) / 254u) - 1;
                entries += (int64_t)((pl.cols[j].w / 255u + 254u) / 255u);
            }
        if (extra_all > 0) {
            auto chunks = [](int64_t cols) { return (cols + 127) / 128; };
            const int64_t dev_chunks = chunks(base) + chunks(entries);
            int64_t budget = chunks(base + extra_all) <= dev_chunks ? extra_all : chunks(std::max<int64_t>(base, 1)) * 128 - base;
            std::vector<CompactCol> out(pl.cols.begin(), pl.cols.begin() + (ptrdiff_t)first_heavy), kept_heavy;
            for (size_t j = first_heavy; j < pl.cols.size(); ++j) {        // ascending weight: the cheapest splits first
                const CompactCol cc = pl.cols[j];
                const int64_t copies = (int64_t)((cc.w + 253u) / 254u);
                if (copies - 1 > budget) { kept_heavy.push_back(cc); continue; }
                budget -= copies - 1;
                uint32_t rest = cc.w;
                for (int64_t t = 0; t < copies; ++t) {
                    const uint32_t part = (uint32_t)((rest + (uint32_t)(copies - t) - 1u) / (uint32_t)(copies - t));   // even split, each <= 254
                    out.push_back(CompactCol{cc.src, part, (uint8_t)(t == 0 ? cc.mult : 0)});
                    rest -= part;
                }
            }
            std::stable_sort(out.begin(), out.end(), [](const CompactCol &a, const CompactCol &b) { return a.w < b.w; });
            out.insert(out.end(), kept_heavy.begin(), kept_heavy.end());
            pl.cols.swap(out);
        }
    }
    pl.m_out = (int32_t)pl.cols.size();
}

template <typename F>
void for_windows(int32_t windows, int32_t threads, F fn) {
    threads = std::max(1, std::min(threads, windows));
    if (threads == 1) { for (int32_t w = 0; w < windows; ++w) fn(w); return; }
    std::vector<std::thread> pool;
    for (int t = 0; t < threads; ++t)
        pool.emplace_back([=] { for (int32_t w = t; w < windows; w += threads) fn(w); });
    for (auto &th : pool) th.join();
}

// The plans of the last impop_compact_scan, kept for the impop_compact_fill that follows on the same arrays
// (the pair search is the expensive half of the compaction: it runs once).
struct PlanCache {
    std::mutex mu;
    const void *x = nullptr, *len = nullptr;
    int32_t windows = -1;
    uint32_t flags = 0;
    uint64_t key = 0;                // content of the descriptor arrays, of every node length and of a sample of the presence words
    std::vector<CompactPlan> plans;
} g_plan_cache;

// Guard of the plan cache: pointers and sizes alone would accept another batch that happens to live at the same addresses.
uint64_t batch_key(int32_t windows, const int32_t *n, const int32_t *m, const int32_t *pitch_words, const int64_t *x_off,
                   const int64_t *len_off, const uint32_t *x_bits, const uint32_t *node_len) {
    uint64_t h = text_hash((const char *)n, 4 * (int64_t)windows) ^ mix64(text_hash((const char *)m, 4 * (int64_t)windows)) ^
                 mix64(text_hash((const char *)pitch_words, 4 * (int64_t)windows) + 1) ^ mix64(text_hash((const char *)x_off, 8 * (int64_t)windows) + 2) ^
                 mix64(text_hash((const char *)len_off, 8 * (int64_t)windows) + 3);
    for (int32_t w = 0; w < windows; ++w) {
        h = mix64(h ^ text_hash((const char *)(node_len + len_off[w]), 4 * (int64_t)m[w]));
        if (n[w] > 0 && pitch_words[w] > 0) {                       // first and last row of the window
            h = mix64(h ^ text_hash((const char *)(x_bits + x_off[w]), 4 * (int64_t)pitch_words[w]));
            h = mix64(h ^ text_hash((const char *)(x_bits + x_off[w] + (int64_t)(n[w] - 1) * pitch_words[w]), 4 * (int64_t)pitch_words[w]));
        }
    }
    return h;
}

}  // namespace

extern "C" {

int impop_compact_scan(int32_t windows, const int32_t *n, const int32_t *m, const int32_t *pitch_words, const int64_t *x_off,
                       const int64_t *len_off, const uint32_t *x_bits, const uint32_t *node_len, int32_t threads, uint32_t flags,
                       int32_t *m_out, int64_t *site_runs_out) {
    if (windows < 0 || (windows > 0 && (!n || !m || !pitch_words || !x_off || !len_off || !m_out))) return IMPOP_ERR_ARG;
    if ((flags & IMPOP_COMPACT_REPLICATE) && !(flags & IMPOP_COMPACT_PAIRS)) return IMPOP_ERR_ARG;
    for (int32_t w = 0; w < windows; ++w)
        if (n[w] < 0 || m[w] < 0 || (int64_t)pitch_words[w] * 32 < m[w] || (m[w] > 0 && n[w] > 0 && (!x_bits || !node_len)))
            return IMPOP_ERR_ARG;
    std::vector<CompactPlan> plans((size_t)windows);
    for_windows(windows, threads, [&](int32_t w) {
        CompactPlan &pl = plans[w];
        ColMajor cm;
        compact_plan(n[w], m[w], pitch_words[w], x_bits + x_off[w], node_len + len_off[w], flags, pl, cm);
        m_out[w] = pl.m_out;
        if (site_runs_out) site_runs_out[w] = pl.site_runs;
    });
    bool bad = false;
    for (const CompactPlan &pl : plans) bad |= pl.bad;
    {
        std::lock_guard<std::mutex> lk(g_plan_cache.mu);
        g_plan_cache.x = x_bits; g_plan_cache.len = node_len; g_plan_cache.windows = windows; g_plan_cache.flags = flags;
        g_plan_cache.key = batch_key(windows, n, m, pitch_words, x_off, len_off, x_bits, node_len);
        g_plan_cache.plans.swap(plans);
    }
    return bad ? IMPOP_ERR_RANGE : IMPOP_OK;
}

int impop_compact_fill(int32_t windows, const int32_t *n, const int32_t *m, const int32_t *pitch_words, const int64_t *x_off,
                       const int64_t *len_off, const uint32_t *x_bits, const uint32_t *node_len, int32_t threads, uint32_t flags,
                       const int32_t *out_pitch_words, const int64_t *out_x_off, const int64_t *out_len_off,
                       uint32_t *x_out, uint32_t *len_out, const int64_t *out_row_off, int32_t *row_adj_out,
                       int64_t *win_const_out, uint8_t *col_mult_out) {
    if (windows < 0 || (windows > 0 && (!n || !m || !pitch_words || !x_off || !len_off || !out_pitch_words || !out_x_off ||
                                        !out_len_off || !x_out || !len_out)))
        return IMPOP_ERR_ARG;
    const bool affine = (flags & IMPOP_COMPACT_PAIRS) != 0u;
    if ((flags & IMPOP_COMPACT_REPLICATE) && !affine) return IMPOP_ERR_ARG;
    if (affine && windows > 0 && (!out_row_off || !row_adj_out || !win_const_out || !col_mult_out)) return IMPOP_ERR_ARG;
    std::vector<CompactPlan> cached;
    {
        std::lock_guard<std::mutex> lk(g_plan_cache.mu);
        if (g_plan_cache.x == x_bits && g_plan_cache.len == node_len && g_plan_cache.windows == windows &&
            g_plan_cache.flags == flags && (int32_t)g_plan_cache.plans.size() == windows &&
            g_plan_cache.key == batch_key(windows, n, m, pitch_words, x_off, len_off, x_bits, node_len)) {
            cached.swap(g_plan_cache.plans);
            g_plan_cache.windows = -1;
        }
    }
    std::vector<int> bad((size_t)std::max(windows, 1), 0);
    for_windows(windows, threads, [&](int32_t w) {
        CompactPlan local;
        ColMajor cm;
        const uint32_t *x = x_bits + x_off[w], *len = node_len + len_off[w];
        if (cached.empty()) compact_plan(n[w], m[w], pitch_words[w], x, len, flags, local, cm);
        else to_columns(n[w], m[w], pitch_words[w], x, cm);
        const CompactPlan &pl = cached.empty() ? local : cached[w];
        const int32_t op = out_pitch_words[w];
        if (pl.bad || (int64_t)op * 32 < pl.m_out || (affine && pl.total >= (1ull << 31))) { bad[w] = 1; return; }
        uint32_t *lo = len_out + out_len_off[w];
        const int32_t nv = pl.m_out, nn = n[w];
        for (int32_t j = 0; j < nv; ++j) lo[j] = pl.cols[j].w;
        if (col_mult_out) {
            uint8_t *mo = col_mult_out + out_len_off[w];
            for (int32_t j = 0; j < nv; ++j) mo[j] = pl.cols[j].mult;
        }
        if (win_const_out) win_const_out[w] = (int64_t)pl.c;
        // output rows: 32 output columns x 32 rows at a time, gathered as column words and transposed back
        uint32_t *xo = x_out + out_x_off[w];
        uint32_t A[32];
        for (int rb = 0; rb < cm.nb; ++rb) {
            const int r0 = rb * 32, nr = std::min(32, nn - r0);
            for (int32_t ow = 0; ow < op; ++ow) {
                bool nonzero = false;
                for (int c = 0; c < 32; ++c) {
                    const int32_t j = ow * 32 + c;
                    uint32_t v = 0u;
                    if (j < nv) v = pl.cols[j].src < 0 ? cm.mask(rb) : cm.col(pl.cols[j].src)[rb];
                    A[c] = v;
                    nonzero |= v != 0u;
                }
                if (nonzero) transpose32(A);
                for (int r = 0; r < nr; ++r) xo[(size_t)(r0 + r) * op + ow] = A[r];
            }
        }
        if (row_adj_out) {
            int32_t *ra = row_adj_out + out_row_off[w];
            for (int32_t i = 0; i < nn; ++i) ra[i] = 0;
            for (const auto &pr : pl.radj) {
                const uint32_t *c = cm.col(pr.first);
                for (int rb = 0; rb < cm.nb; ++rb)
                    for (uint32_t b = c[rb]; b; b &= b - 1) ra[rb * 32 + __builtin_ctz(b)] += (int32_t)pr.second;
            }
        }
    });
    for (int32_t w = 0; w < windows; ++w)
        if (bad[w]) return IMPOP_ERR_RANGE;
    return IMPOP_OK;
}

}  // extern "C"
