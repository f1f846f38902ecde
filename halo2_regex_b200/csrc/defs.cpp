// Definition loaders and the dense-table packer (host side, no CUDA).
#include "defs.hpp"

#include <algorithm>
#include <cstdarg>
#include <cstdio>
#include <cstring>

#include "../../include/b2r.h"

namespace b2r {

static thread_local char g_err[512] = "";
void set_error(const char* fmt, ...) {
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof g_err, fmt, ap);
    va_end(ap);
}
const char* get_error() { return g_err; }

int read_file(const char* path, std::string& out) {
    FILE* f = fopen(path, "rb");
    if (!f) { set_error("cannot open %s", path); return B2R_ERR_IO; }
    char buf[1 << 16];
    size_t n;
    out.clear();
    while ((n = fread(buf, 1, sizeof buf, f)) > 0) out.append(buf, n);
    fclose(f);
    return B2R_OK;
}

// ---- line / token scanner ------------------------------------------------------------------------------------
// Mirrors BufRead::lines() + str::split_whitespace() + str::parse::<u64>() (reference src/defs.rs:84-92, 218-226):
// lines end at '\n' (one preceding '\r' dropped), no line after a trailing '\n'; separators are Unicode White_Space;
// a token is an optional '+' followed by ASCII digits that fit u64; anything else (incl. invalid UTF-8) is an error.
namespace {

// decodes one UTF-8 scalar at p (< end); returns its length or 0 if malformed
int utf8_next(const unsigned char* p, const unsigned char* end, uint32_t& cp) {
    unsigned c = *p;
    if (c < 0x80) { cp = c; return 1; }
    int n = (c >= 0xF0) ? 4 : (c >= 0xE0) ? 3 : (c >= 0xC2) ? 2 : 0;
    if (!n || c > 0xF4 || end - p < n) return 0;
    cp = c & (0xFF >> (n + 1));
    for (int i = 1; i < n; i++) {
        if ((p[i] & 0xC0) != 0x80) return 0;
        cp = (cp << 6) | (p[i] & 0x3F);
    }
    if ((n == 3 && (cp < 0x800 || (cp >= 0xD800 && cp <= 0xDFFF))) || (n == 4 && (cp < 0x10000 || cp > 0x10FFFF))) return 0;
    return n;
}
bool is_white_space(uint32_t cp) {
    return (cp >= 9 && cp <= 13) || cp == 0x20 || cp == 0x85 || cp == 0xA0 || cp == 0x1680 || (cp >= 0x2000 && cp <= 0x200A) ||
           cp == 0x2028 || cp == 0x2029 || cp == 0x202F || cp == 0x205F || cp == 0x3000;
}

struct LineScanner {
    const unsigned char *cur, *end;
    uint64_t line_idx = 0;
    std::vector<uint64_t> tok;
    LineScanner(const char* t, size_t n) : cur((const unsigned char*)t), end((const unsigned char*)t + n) {}
    // 1 = a line was read into tok, 0 = end of text, -1 = malformed line
    int next() {
        if (cur >= end) return 0;
        const unsigned char* nl = (const unsigned char*)memchr(cur, '\n', end - cur);
        const unsigned char* stop = nl ? nl : end;
        const unsigned char* next_line = nl ? nl + 1 : end;
        if (nl && stop > cur && stop[-1] == '\r') stop--;
        tok.clear();
        bool in_tok = false, have_digit = false, bad = false;
        uint64_t val = 0;
        auto finish = [&]() {
            if (in_tok) { if (!have_digit) bad = true; tok.push_back(val); }
            in_tok = false; have_digit = false; val = 0;
        };
        for (const unsigned char* p = cur; p < stop;) {
            uint32_t cp;
            int n = utf8_next(p, stop, cp);
            if (!n) { bad = true; break; }
            if (is_white_space(cp)) finish();
            else if (!in_tok && cp == '+') in_tok = true;
            else if (cp >= '0' && cp <= '9') {
                unsigned d = cp - '0';
                if (val > (UINT64_MAX - d) / 10) bad = true;
                val = val * 10 + d;
                in_tok = have_digit = true;
            } else bad = true;
            p += n;
        }
        finish();
        cur = next_line;
        return bad ? -1 : 1;
    }
};

}  // namespace

int parse_allstr(const char* text, size_t len, AllstrDef& out, uint64_t* err_line) {
    out = AllstrDef();
    LineScanner sc(text, len);
    for (int r; (r = sc.next()) != 0; sc.line_idx++) {
        const auto& e = sc.tok;
        const uint64_t idx = sc.line_idx;
        const size_t need = idx <= 2 ? 1 : 3;  // elements[0] / elements[2] index panics on a short line
        if (r < 0 || e.size() < need) {
            if (err_line) *err_line = idx;
            set_error("allstr definition: malformed line %llu", (unsigned long long)idx);
            return B2R_ERR_PARSE;
        }
        if (idx == 0) out.first_state_val = e[0];
        else if (idx == 1) out.accepted_state_val = e[0];
        else if (idx == 2) out.largest_state_val = e[0];
        else out.state_lookup[{e[0], (uint8_t)e[2]}] = {idx, e[1]};  // `elements[2] as u8`; insert overwrites
    }
    return B2R_OK;
}

int parse_substr(const char* text, size_t len, SubstrDef& out, uint64_t* err_line) {
    out = SubstrDef();
    LineScanner sc(text, len);
    for (int r; (r = sc.next()) != 0; sc.line_idx++) {
        const auto& e = sc.tok;
        const uint64_t idx = sc.line_idx;
        const size_t need = idx <= 2 ? 1 : (idx <= 4 ? 0 : 2);
        if (r < 0 || e.size() < need) {
            if (err_line) *err_line = idx;
            set_error("substr definition: malformed line %llu", (unsigned long long)idx);
            return B2R_ERR_PARSE;
        }
        switch (idx) {
            case 0: out.max_length = e[0]; break;
            case 1: out.min_position = e[0]; break;
            case 2: out.max_position = e[0]; break;
            case 3: out.start_states = e; break;
            case 4: out.end_states = e; break;
            default: out.valid_state_transitions.insert({e[0], e[1]});
        }
    }
    return B2R_OK;
}

std::vector<Transition> AllstrDef::in_table_order() const {
    std::vector<Transition> v;
    v.reserve(state_lookup.size());
    for (const auto& kv : state_lookup) v.push_back({kv.first.second, kv.first.first, kv.second.second, kv.second.first});
    std::sort(v.begin(), v.end(), [](const Transition& a, const Transition& b) { return a.line_idx < b.line_idx; });
    return v;
}

// ---- packer ----------------------------------------------------------------------------------------------------
int pack_def(const AllstrDef& a, const std::vector<const SubstrDef*>& substrs, uint32_t substr_id_offset, PackedDef& out) {
    out = PackedDef();
    if (a.largest_state_val + 1 > 65535) {
        set_error("largest_state_val %llu: more than 65535 states are not supported", (unsigned long long)a.largest_state_val);
        return B2R_ERR_UNSUPPORTED;
    }
    const uint32_t S = (uint32_t)a.largest_state_val + 1;
    if (a.first_state_val >= S) {
        set_error("first_state_val %llu exceeds largest_state_val", (unsigned long long)a.first_state_val);
        return B2R_ERR_UNSUPPORTED;
    }
    if ((uint64_t)substr_id_offset + substrs.size() > 256) {
        set_error("substr ids above 255 are not supported");
        return B2R_ERR_UNSUPPORTED;
    }
    out.num_states = S;
    out.first_state = (uint32_t)a.first_state_val;
    out.accepted_state = a.accepted_state_val < S ? (uint32_t)a.accepted_state_val : 0xFFFFFFFFu;
    out.state_width = S <= 255 ? 1 : 2;
    out.substr_id_offset = substr_id_offset;
    out.num_substrs = (uint32_t)substrs.size();

    // dense [256][S] entries, default invalid
    std::vector<uint32_t> dense(256 * (size_t)S, ENT_INVALID);
    const std::vector<Transition> order = a.in_table_order();
    out.rows.push_back({0, S, S, 0});  // src/table.rs:101
    out.row_bin.push_back(0xFFFFFFFFu);
    for (const Transition& t : order) {
        if (t.cur >= S || t.next >= S) {
            set_error("state id %llu exceeds largest_state_val %llu (the dummy state would collide with a real one)",
                      (unsigned long long)std::max(t.cur, t.next), (unsigned long long)a.largest_state_val);
            return B2R_ERR_UNSUPPORTED;
        }
        uint32_t sid = 0, flags = 0;
        for (size_t k = 0; k < substrs.size(); k++) {  // first match wins (src/lib.rs:831-840, src/table.rs:111-120)
            if (substrs[k]->valid_state_transitions.count({t.cur, t.next})) {
                sid = substr_id_offset + (uint32_t)k;
                const auto& ss = substrs[k]->start_states;
                const auto& es = substrs[k]->end_states;
                if (std::find(ss.begin(), ss.end(), t.cur) != ss.end()) flags |= ENT_IS_START;
                if (std::find(es.begin(), es.end(), t.next) != es.end()) flags |= ENT_IS_END;
                break;
            }
        }
        dense[(size_t)t.ch * S + t.cur] = (uint32_t)t.next | (sid << ENT_SID_SHIFT) | flags;
        if (sid && t.cur < 64) out.hot_states[t.cur >> 5] |= 1u << (t.cur & 31);
        out.rows.push_back({t.ch, t.cur, t.next, sid});
        out.row_bin.push_back((uint32_t)t.ch * S + (uint32_t)t.cur);
    }

    if (S < 64) out.hot_states[S >> 5] |= 1u << (S & 31);   // the walk's trap state (after an invalid transition) must be looked at

    // byte equivalence classes: bytes with identical [S] columns share a class
    out.byte_class.assign(256, 0);
    std::map<std::vector<uint32_t>, uint32_t> seen;
    for (int c = 0; c < 256; c++) {
        std::vector<uint32_t> col(dense.begin() + (size_t)c * S, dense.begin() + (size_t)(c + 1) * S);
        auto it = seen.find(col);
        if (it == seen.end()) {
            it = seen.emplace(col, (uint32_t)seen.size()).first;
            out.trans.insert(out.trans.end(), col.begin(), col.end());
        }
        out.byte_class[c] = (uint8_t)it->second;
    }
    out.num_classes = (uint32_t)seen.size();

    // endpoint table (src/table.rs:126-196) and its counter bins
    out.erows.push_back({0, S, S});
    out.erow_start_bin.push_back(0xFFFFFFFFu);
    out.erow_end_bin.push_back(0xFFFFFFFFu);
    for (size_t k = 0; k < substrs.size(); k++) {
        const uint64_t id = substr_id_offset + k;
        for (uint64_t s : substrs[k]->start_states) {
            bool first = true;
            for (const EndpointRow& r : out.erows) if (r.sid == id && r.start == s && r.end == S) first = false;
            out.erows.push_back({id, s, S});
            out.erow_start_bin.push_back(first && s < S ? (uint32_t)(k * S + s) : 0xFFFFFFFFu);
            out.erow_end_bin.push_back(0xFFFFFFFFu);
        }
        for (uint64_t e : substrs[k]->end_states) {
            bool first = true;
            for (const EndpointRow& r : out.erows) if (r.sid == id && r.start == S && r.end == e) first = false;
            out.erows.push_back({id, S, e});
            out.erow_start_bin.push_back(0xFFFFFFFFu);
            out.erow_end_bin.push_back(first && e < S ? (uint32_t)(k * S + e) : 0xFFFFFFFFu);
        }
    }
    return B2R_OK;
}

uint32_t padded_states(uint32_t num_states) {
    uint32_t p = 1;
    while (p < num_states + 1) p <<= 1;
    return p;
}

void build_walk_table(const PackedDef& def, std::vector<uint32_t>& out) {
    const uint32_t S = def.num_states, P = padded_states(S);
    out.assign((size_t)def.num_classes * P, (S << 16) | 1u);
    for (uint32_t k = 0; k < def.num_classes; k++)
        for (uint32_t s = 0; s < S; s++) {
            const uint32_t e = def.trans[(size_t)k * S + s];
            if (e & ENT_INVALID) continue;
            out[(size_t)k * P + s] = ((e & ENT_NEXT_MASK) << 16) | ((e & ENT_SID_MASK) ? 1u : 0u);
        }
}

}  // namespace b2r
