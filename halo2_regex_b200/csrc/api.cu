// extern "C" boundary (include/b2r.h): handles, table queries, the batch entry points.
#include <cuda_runtime.h>

#include <cstdio>
#include <cstdlib>
#include <algorithm>
#include <cstring>
#include <memory>
#include <new>
#include <string>
#include <vector>

#include "../../include/b2r.h"
#include "defs.hpp"
#include "kernels.cuh"
#include "long.cuh"

using namespace b2r;

struct b2r_allstr { AllstrDef def; };
struct b2r_substr { SubstrDef def; };

#define CUDA_TRY(expr)                                                                   \
    do {                                                                                 \
        cudaError_t _e = (expr);                                                         \
        if (_e != cudaSuccess) {                                                         \
            set_error("%s failed: %s", #expr, cudaGetErrorString(_e));                   \
            return B2R_ERR_CUDA;                                                         \
        }                                                                                \
    } while (0)

namespace {

struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int reserve(size_t n) {
        if (n <= cap) return B2R_OK;
        if (p) cudaFree(p);
        p = nullptr; cap = 0;
        CUDA_TRY(cudaMalloc(&p, n));
        cap = n;
        return B2R_OK;
    }
    void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

struct DevDef {
    uint8_t* byte_class = nullptr;
    uint32_t* trans = nullptr;
    uint32_t* hot = nullptr;     // walk table [C][P] (walk.cuh)
    uint32_t *row_bin = nullptr, *erow_start_bin = nullptr, *erow_end_bin = nullptr;
    unsigned long long *hist = nullptr, *ep_start = nullptr, *ep_end = nullptr;  // inside cfg->scratch
};

}  // namespace

struct b2r_config {
    int device = -1;             // -1: host-only handle (table queries), no matching
    uint64_t max_chars = 0;
    uint32_t n_defs = 0;
    PackedDef packed[B2R_MAX_DEFS];
    DevDef dev[B2R_MAX_DEFS];
    void* tables = nullptr;      // one allocation holding every constant table
    void* scratch = nullptr;     // BatchCounters + hist + endpoint counters, zeroed per batch
    size_t scratch_bytes = 0;
    b2r_batch_status* d_batch_status = nullptr;
    int force_table_mode = -1;        // testing hooks: B2R_TABLE_MODE=repl|plain|global, B2R_HIST_MODE=smem|global
    int force_hist_mode = -1;
    DevBuf ws_fmask;                  // granule flags (walk -> emit)
    DevBuf ws_long;                   // long-string path: chunk offsets, transition-vector tree, entry states, flag summary
    DevBuf ws_states[B2R_MAX_DEFS];   // state column of a def the caller did not ask for (emit reads it)
    WalkParams last = {};
    bool have_last = false;
    uint32_t last_launches = 0;
    bool timing = false;
    cudaEvent_t ev[4] = {nullptr, nullptr, nullptr, nullptr};   // before walk, after walk, after emit, after finalize
    // staging for the host-pointer entry point
    DevBuf ws_bytes, ws_offsets, ws_cols;
    cudaStream_t host_stream = nullptr;
    cudaStream_t in_stream = nullptr, out_stream = nullptr;   // host entry point: H2D / D2H copies overlapping the kernels
    static constexpr int MAX_SLICES = 8;
    cudaEvent_t ev_in[MAX_SLICES] = {}, ev_done[MAX_SLICES] = {};
    BatchCounters* h_slices = nullptr;    // pinned: the counters of every slice of a host batch
    cudaEvent_t ev_fork = nullptr;
};

namespace {

struct DeviceGuard {
    int prev = -1;
    bool ok = true;
    explicit DeviceGuard(int dev) {
        if (cudaGetDevice(&prev) != cudaSuccess) { ok = false; return; }
        if (prev != dev && cudaSetDevice(dev) != cudaSuccess) ok = false;
    }
    ~DeviceGuard() { if (prev >= 0) cudaSetDevice(prev); }
};

size_t align_up(size_t x, size_t a) { return (x + a - 1) / a * a; }

int upload_tables(b2r_config* c) {
    size_t total = 0;
    for (uint32_t d = 0; d < c->n_defs; d++) {
        const PackedDef& pd = c->packed[d];
        total += align_up(256, 256) + align_up(pd.trans.size() * 4, 256) + align_up(pd.row_bin.size() * 4, 256) +
                 2 * align_up(pd.erows.size() * 4, 256) + align_up((size_t)pd.num_classes * padded_states(pd.num_states) * 4, 256);
    }
    CUDA_TRY(cudaMalloc(&c->tables, total));
    unsigned char* base = (unsigned char*)c->tables;
    size_t off = 0;
    auto put = [&](const void* src, size_t n) -> void* {
        void* dst = base + off;
        cudaMemcpy(dst, src, n, cudaMemcpyHostToDevice);
        off += align_up(n, 256);
        return dst;
    };
    size_t scratch = align_up(sizeof(BatchCounters), 256);
    for (uint32_t d = 0; d < c->n_defs; d++) {
        const PackedDef& pd = c->packed[d];
        c->dev[d].byte_class = (uint8_t*)put(pd.byte_class.data(), 256);
        c->dev[d].trans = (uint32_t*)put(pd.trans.data(), pd.trans.size() * 4);
        {
            std::vector<uint32_t> hot;
            build_walk_table(pd, hot);
            c->dev[d].hot = (uint32_t*)put(hot.data(), hot.size() * 4);
        }
        c->dev[d].row_bin = (uint32_t*)put(pd.row_bin.data(), pd.row_bin.size() * 4);
        c->dev[d].erow_start_bin = (uint32_t*)put(pd.erow_start_bin.data(), pd.erows.size() * 4);
        c->dev[d].erow_end_bin = (uint32_t*)put(pd.erow_end_bin.data(), pd.erows.size() * 4);
        scratch += align_up((size_t)256 * pd.num_states * 8, 256) + 2 * align_up((size_t)std::max<uint32_t>(pd.num_substrs, 1) * pd.num_states * 8, 256);
    }
    CUDA_TRY(cudaGetLastError());
    CUDA_TRY(cudaMalloc(&c->scratch, scratch));
    c->scratch_bytes = scratch;
    unsigned char* sb = (unsigned char*)c->scratch;
    size_t so = align_up(sizeof(BatchCounters), 256);
    for (uint32_t d = 0; d < c->n_defs; d++) {
        const PackedDef& pd = c->packed[d];
        c->dev[d].hist = (unsigned long long*)(sb + so); so += align_up((size_t)256 * pd.num_states * 8, 256);
        const size_t ep = align_up((size_t)std::max<uint32_t>(pd.num_substrs, 1) * pd.num_states * 8, 256);
        c->dev[d].ep_start = (unsigned long long*)(sb + so); so += ep;
        c->dev[d].ep_end = (unsigned long long*)(sb + so); so += ep;
    }
    CUDA_TRY(cudaMalloc((void**)&c->d_batch_status, sizeof(b2r_batch_status)));
    if (const char* tm = getenv("B2R_TABLE_MODE"))
        c->force_table_mode = !strcmp(tm, "repl") ? (int)TABLE_REPL : !strcmp(tm, "plain") ? (int)TABLE_PLAIN : !strcmp(tm, "plain16") ? (int)TABLE_PLAIN16 : !strcmp(tm, "global") ? (int)TABLE_GLOBAL : -1;
    if (const char* hm = getenv("B2R_HIST_MODE"))
        c->force_hist_mode = !strcmp(hm, "smem") ? (int)HIST_SMEM : !strcmp(hm, "global") ? (int)HIST_GLOBAL : -1;
    return B2R_OK;
}

bool aligned16(const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15) == 0; }

int check_outputs(const b2r_config* c, const b2r_outputs* o, bool check_ptrs, uint64_t M = 0) {
    if (!M) M = c->max_chars;
    if (o->row_pitch < M || o->row_pitch % 16) { set_error("row_pitch %llu must be >= max_chars_size and a multiple of 16", (unsigned long long)o->row_pitch); return B2R_ERR_ALIGNMENT; }
    if (o->bitmap_pitch < (M + 7) / 8 || o->bitmap_pitch % 4) { set_error("bitmap_pitch %llu must be >= ceil(M/8) and a multiple of 4", (unsigned long long)o->bitmap_pitch); return B2R_ERR_ALIGNMENT; }
    if (!check_ptrs) return B2R_OK;
    bool ok = aligned16(o->masked_chars) && aligned16(o->masked_substr_ids);
    for (uint32_t d = 0; d < c->n_defs; d++)
        ok = ok && aligned16(o->states[d]) && aligned16(o->substr_ids[d]) && aligned16(o->start_enable[d]) && aligned16(o->end_enable[d]);
    if (!ok) { set_error("output columns must be 16-byte aligned"); return B2R_ERR_ALIGNMENT; }
    return B2R_OK;
}

void fill_walk_params(const b2r_config* c, WalkParams& p, const uint8_t* d_bytes, const uint64_t* d_offsets, uint64_t n, uint64_t total_bytes,
                      const b2r_outputs* o, uint64_t max_chars) {
    memset(&p, 0, sizeof p);
    p.bytes = d_bytes; p.offsets = d_offsets; p.n_strings = n; p.total_bytes = total_bytes;
    p.row_pitch = o->row_pitch; p.bitmap_pitch = o->bitmap_pitch;
    p.max_chars = (uint32_t)max_chars; p.n_defs = c->n_defs;
    for (uint32_t d = 0; d < c->n_defs; d++) {
        const PackedDef& pd = c->packed[d];
        DefDev& dd = p.def[d];
        dd.byte_class = c->dev[d].byte_class; dd.trans = c->dev[d].trans; dd.hot = c->dev[d].hot;
        dd.padded_states = padded_states(pd.num_states);
        dd.hot_states[0] = pd.hot_states[0]; dd.hot_states[1] = pd.hot_states[1];
        dd.num_states = pd.num_states; dd.num_classes = pd.num_classes; dd.first_state = pd.first_state;
        dd.accepted_state = pd.accepted_state; dd.sid_offset = pd.substr_id_offset; dd.num_substrs = pd.num_substrs;
        dd.hist = c->dev[d].hist; dd.ep_start = c->dev[d].ep_start; dd.ep_end = c->dev[d].ep_end;
        dd.states = o->states[d]; dd.substr_ids = o->substr_ids[d]; dd.start_enable = o->start_enable[d]; dd.end_enable = o->end_enable[d];
    }
    p.masked_chars = o->masked_chars; p.masked_substr_ids = o->masked_substr_ids;
    p.status = o->status; p.records = o->records; p.compact_bytes = o->compact_bytes;
    p.max_records = o->records ? o->max_records : 0; p.compact_pitch = o->compact_bytes ? o->compact_pitch : 0;
    p.counters = (BatchCounters*)c->scratch;
    p.n_tiles = (uint32_t)((n + 31) / 32);
    { const char* dbg = getenv("B2R_DEBUG"); p.debug = dbg ? (uint32_t)atoi(dbg) : 0u; }
    { const char* f = getenv("B2R_SPREAD_FILL"); p.spread_fill = (f && f[0] == '0') ? 0u : 1u; }   // testing hook
    { const char* f = getenv("B2R_FUSE"); p.fuse = (f && f[0] == '0') ? 0u : 1u; }   // testing hook: B2R_FUSE=0 runs emit_kernel as its own launch
    p.fm_words = (uint32_t)((((max_chars - 1) + 15) / 16 + 31) / 32);
    uint64_t ep = 0;
    for (uint32_t d = 0; d < c->n_defs; d++) ep += 2ull * c->packed[d].num_substrs * c->packed[d].num_states * 4ull;
    ep = (ep + 15) & ~15ull;
    p.ep_smem_bytes = ep <= 8192 ? (uint32_t)ep : 0u;   // many substrs x many states: count with global atomics instead
    uint64_t et = 0;
    for (uint32_t d = 0; d < c->n_defs; d++) et += (uint64_t)c->packed[d].num_classes * c->packed[d].num_states * 4ull + 256ull;
    p.emit_smem_tables = et + p.ep_smem_bytes <= 40 * 1024 ? 1u : 0u;
}

int enqueue_finalize(b2r_config* c, const b2r_outputs* o, uint64_t n, uint64_t max_chars, cudaStream_t st) {
    FinalizeParams f;
    memset(&f, 0, sizeof f);
    f.n_defs = c->n_defs; f.accumulate = (o->flags & B2R_OUT_ACCUMULATE_MULT) ? 1 : 0;
    f.n_rows_total = n * max_chars; f.counters = (const BatchCounters*)c->scratch;
    bool any = false;
    for (uint32_t d = 0; d < c->n_defs; d++) {
        auto& fd = f.def[d];
        fd.hist = c->dev[d].hist; fd.row_bin = c->dev[d].row_bin; fd.n_rows = (uint32_t)c->packed[d].rows.size();
        fd.ep_start = c->dev[d].ep_start; fd.ep_end = c->dev[d].ep_end;
        fd.erow_start_bin = c->dev[d].erow_start_bin; fd.erow_end_bin = c->dev[d].erow_end_bin; fd.n_erows = (uint32_t)c->packed[d].erows.size();
        fd.mult = (unsigned long long*)o->mult[d]; fd.endpoint_mult = (unsigned long long*)o->endpoint_mult[d];
        any = any || fd.mult || fd.endpoint_mult;
    }
    if (!any) return B2R_OK;
    c->last_launches++;
    return launch_finalize(f, st);
}

}  // namespace

extern "C" {

const char* b2r_last_error(void) { return get_error(); }
const char* b2r_version(void) { return "b2r 0.1 (sm_100a)"; }

// ---- AllstrRegexDef ------------------------------------------------------------------------------------------------
int b2r_allstr_parse(const char* text, size_t len, b2r_allstr** out, uint64_t* err_line) {
    if (!out || (!text && len)) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    std::unique_ptr<b2r_allstr> a(new (std::nothrow) b2r_allstr);
    int rc = parse_allstr(text, len, a->def, err_line);
    if (rc) return rc;
    *out = a.release();
    return B2R_OK;
}
int b2r_allstr_read_from_text(const char* path, b2r_allstr** out, uint64_t* err_line) {
    std::string s;
    int rc = read_file(path, s);
    if (rc) return rc;
    return b2r_allstr_parse(s.data(), s.size(), out, err_line);
}
void b2r_allstr_free(b2r_allstr* a) { delete a; }
uint64_t b2r_allstr_first_state_val(const b2r_allstr* a) { return a->def.first_state_val; }
uint64_t b2r_allstr_accepted_state_val(const b2r_allstr* a) { return a->def.accepted_state_val; }
uint64_t b2r_allstr_largest_state_val(const b2r_allstr* a) { return a->def.largest_state_val; }
uint64_t b2r_allstr_num_transitions(const b2r_allstr* a) { return a->def.state_lookup.size(); }
int b2r_allstr_lookup(const b2r_allstr* a, uint8_t ch, uint64_t state, uint64_t* line_idx, uint64_t* next) {
    auto it = a->def.state_lookup.find({state, ch});
    if (it == a->def.state_lookup.end()) return 0;
    if (line_idx) *line_idx = it->second.first;
    if (next) *next = it->second.second;
    return 1;
}
int b2r_allstr_entries(const b2r_allstr* a, uint64_t* out4, uint64_t capacity_rows) {
    const auto v = a->def.in_table_order();
    if (capacity_rows < v.size()) { set_error("capacity too small"); return B2R_ERR_INVALID_ARG; }
    for (size_t i = 0; i < v.size(); i++) { out4[4 * i] = v[i].ch; out4[4 * i + 1] = v[i].cur; out4[4 * i + 2] = v[i].next; out4[4 * i + 3] = v[i].line_idx; }
    return B2R_OK;
}

// ---- SubstrRegexDef ------------------------------------------------------------------------------------------------
int b2r_substr_parse(const char* text, size_t len, b2r_substr** out, uint64_t* err_line) {
    if (!out || (!text && len)) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    std::unique_ptr<b2r_substr> s(new (std::nothrow) b2r_substr);
    int rc = parse_substr(text, len, s->def, err_line);
    if (rc) return rc;
    *out = s.release();
    return B2R_OK;
}
int b2r_substr_read_from_text(const char* path, b2r_substr** out, uint64_t* err_line) {
    std::string s;
    int rc = read_file(path, s);
    if (rc) return rc;
    return b2r_substr_parse(s.data(), s.size(), out, err_line);
}
int b2r_substr_new(uint64_t max_length, uint64_t min_position, uint64_t max_position, const uint64_t* pairs, uint64_t n_pairs,
                   const uint64_t* start_states, uint64_t n_start, const uint64_t* end_states, uint64_t n_end, b2r_substr** out) {
    if (!out) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    std::unique_ptr<b2r_substr> s(new (std::nothrow) b2r_substr);
    s->def.max_length = max_length; s->def.min_position = min_position; s->def.max_position = max_position;
    for (uint64_t i = 0; i < n_pairs; i++) s->def.valid_state_transitions.insert({pairs[2 * i], pairs[2 * i + 1]});
    s->def.start_states.assign(start_states, start_states + n_start);
    s->def.end_states.assign(end_states, end_states + n_end);
    *out = s.release();
    return B2R_OK;
}
void b2r_substr_free(b2r_substr* s) { delete s; }
uint64_t b2r_substr_max_length(const b2r_substr* s) { return s->def.max_length; }
uint64_t b2r_substr_min_position(const b2r_substr* s) { return s->def.min_position; }
uint64_t b2r_substr_max_position(const b2r_substr* s) { return s->def.max_position; }
uint64_t b2r_substr_num_transitions(const b2r_substr* s) { return s->def.valid_state_transitions.size(); }
uint64_t b2r_substr_num_start_states(const b2r_substr* s) { return s->def.start_states.size(); }
uint64_t b2r_substr_num_end_states(const b2r_substr* s) { return s->def.end_states.size(); }
int b2r_substr_transitions(const b2r_substr* s, uint64_t* out2, uint64_t capacity) {
    if (capacity < s->def.valid_state_transitions.size()) { set_error("capacity too small"); return B2R_ERR_INVALID_ARG; }
    size_t i = 0;
    for (const auto& pr : s->def.valid_state_transitions) { out2[2 * i] = pr.first; out2[2 * i + 1] = pr.second; i++; }
    return B2R_OK;
}
int b2r_substr_start_states(const b2r_substr* s, uint64_t* out, uint64_t capacity) {
    if (capacity < s->def.start_states.size()) { set_error("capacity too small"); return B2R_ERR_INVALID_ARG; }
    std::copy(s->def.start_states.begin(), s->def.start_states.end(), out);
    return B2R_OK;
}
int b2r_substr_end_states(const b2r_substr* s, uint64_t* out, uint64_t capacity) {
    if (capacity < s->def.end_states.size()) { set_error("capacity too small"); return B2R_ERR_INVALID_ARG; }
    std::copy(s->def.end_states.begin(), s->def.end_states.end(), out);
    return B2R_OK;
}
int b2r_substr_contains(const b2r_substr* s, uint64_t cur, uint64_t next) { return s->def.valid_state_transitions.count({cur, next}) ? 1 : 0; }

// ---- RegexVerifyConfig -----------------------------------------------------------------------------------------------
int b2r_config_new(const b2r_allstr* const* allstr, const b2r_substr* const* const* substrs, const uint32_t* n_substrs, uint32_t n_defs,
                   uint64_t max_chars_size, int device, b2r_config** out) {
    if (!allstr || !n_substrs || !out || n_defs == 0) { set_error("null / empty argument"); return B2R_ERR_INVALID_ARG; }
    if (n_defs > B2R_MAX_DEFS) { set_error("%u regex defs: at most %d are supported", n_defs, B2R_MAX_DEFS); return B2R_ERR_UNSUPPORTED; }
    if (max_chars_size == 0 || max_chars_size > 0xFFFFFFF0ull) { set_error("max_chars_size out of range"); return B2R_ERR_INVALID_ARG; }
    std::unique_ptr<b2r_config> c(new (std::nothrow) b2r_config);
    c->n_defs = n_defs; c->max_chars = max_chars_size; c->device = device;
    uint32_t offset = 1;  // src/lib.rs:780, 827
    uint64_t max_sum = 0;
    for (uint32_t d = 0; d < n_defs; d++) {
        std::vector<const SubstrDef*> subs;
        for (uint32_t k = 0; k < n_substrs[d]; k++) subs.push_back(&substrs[d][k]->def);
        int rc = pack_def(allstr[d]->def, subs, offset, c->packed[d]);
        if (rc) return rc;
        if (n_substrs[d]) max_sum += offset + n_substrs[d] - 1;
        offset += n_substrs[d];  // src/table.rs:197
    }
    // one storage width for every state column of the config: 2 bytes as soon as one def has a dummy state > 255
    bool any_wide = false;
    for (uint32_t d = 0; d < n_defs; d++) any_wide = any_wide || c->packed[d].state_width == 2;
    if (any_wide) for (uint32_t d = 0; d < n_defs; d++) c->packed[d].state_width = 2;
    if (max_sum > 255) { set_error("sum of the largest substr ids over defs is %llu > 255", (unsigned long long)max_sum); return B2R_ERR_UNSUPPORTED; }
    if (device >= 0) {
        int n_dev = 0;
        if (cudaGetDeviceCount(&n_dev) != cudaSuccess || device >= n_dev) {
            set_error("CUDA device %d is not available (%d devices); this library has no CPU fallback", device, n_dev);
            return B2R_ERR_CUDA;
        }
        cudaDeviceProp prop;
        CUDA_TRY(cudaGetDeviceProperties(&prop, device));
        if (prop.major != 10) { set_error("device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major, prop.minor); return B2R_ERR_CUDA; }
        DeviceGuard g(device);
        if (!g.ok) { set_error("cudaSetDevice(%d) failed", device); return B2R_ERR_CUDA; }
        int rc = upload_tables(c.get());
        if (rc) return rc;
        CUDA_TRY(cudaStreamCreateWithFlags(&c->host_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->in_stream, cudaStreamNonBlocking));
        CUDA_TRY(cudaStreamCreateWithFlags(&c->out_stream, cudaStreamNonBlocking));
        for (auto& e : c->ev_in) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        for (auto& e : c->ev_done) CUDA_TRY(cudaEventCreateWithFlags(&e, cudaEventDisableTiming));
        CUDA_TRY(cudaMallocHost((void**)&c->h_slices, sizeof(BatchCounters) * b2r_config::MAX_SLICES));
        CUDA_TRY(cudaEventCreateWithFlags(&c->ev_fork, cudaEventDisableTiming));
        for (auto& e : c->ev) CUDA_TRY(cudaEventCreate(&e));
    }
    *out = c.release();
    return B2R_OK;
}

void b2r_config_free(b2r_config* c) {
    if (!c) return;
    if (c->device >= 0) {
        DeviceGuard g(c->device);
        cudaFree(c->tables); cudaFree(c->scratch); cudaFree(c->d_batch_status);
        c->ws_fmask.release(); c->ws_long.release();
        for (auto& b : c->ws_states) b.release();
        c->ws_bytes.release(); c->ws_offsets.release(); c->ws_cols.release();
        if (c->host_stream) cudaStreamDestroy(c->host_stream);
        if (c->in_stream) cudaStreamDestroy(c->in_stream);
        if (c->out_stream) cudaStreamDestroy(c->out_stream);
        for (auto& e : c->ev_in) if (e) cudaEventDestroy(e);
        for (auto& e : c->ev_done) if (e) cudaEventDestroy(e);
        if (c->h_slices) cudaFreeHost(c->h_slices);
        if (c->ev_fork) cudaEventDestroy(c->ev_fork);
        for (auto& e : c->ev) if (e) cudaEventDestroy(e);
    }
    delete c;
}
uint32_t b2r_config_num_defs(const b2r_config* c) { return c->n_defs; }
uint64_t b2r_config_max_chars_size(const b2r_config* c) { return c->max_chars; }
int b2r_config_device(const b2r_config* c) { return c->device; }
uint32_t b2r_config_state_width(const b2r_config* c, uint32_t d) { return d < c->n_defs ? c->packed[d].state_width : 0; }
uint64_t b2r_config_dummy_state(const b2r_config* c, uint32_t d) { return d < c->n_defs ? c->packed[d].num_states : 0; }
uint32_t b2r_config_substr_id_offset(const b2r_config* c, uint32_t d) { return d < c->n_defs ? c->packed[d].substr_id_offset : 0; }
uint32_t b2r_config_num_byte_classes(const b2r_config* c, uint32_t d) { return d < c->n_defs ? c->packed[d].num_classes : 0; }
uint64_t b2r_config_recommended_row_pitch(const b2r_config* c) { return align_up(c->max_chars, 32); }
uint64_t b2r_config_recommended_bitmap_pitch(const b2r_config* c) { return align_up((c->max_chars + 7) / 8, 32); }

uint64_t b2r_table_num_rows(const b2r_config* c, uint32_t d) { return d < c->n_defs ? c->packed[d].rows.size() : 0; }
int b2r_table_rows(const b2r_config* c, uint32_t d, uint64_t* out4, uint64_t capacity_rows) {
    if (d >= c->n_defs || capacity_rows < c->packed[d].rows.size()) { set_error("bad def index / capacity"); return B2R_ERR_INVALID_ARG; }
    const auto& rows = c->packed[d].rows;
    for (size_t i = 0; i < rows.size(); i++) { out4[4 * i] = rows[i].ch; out4[4 * i + 1] = rows[i].cur; out4[4 * i + 2] = rows[i].next; out4[4 * i + 3] = rows[i].sid; }
    return B2R_OK;
}
uint64_t b2r_endpoint_num_rows(const b2r_config* c, uint32_t d) { return d < c->n_defs ? c->packed[d].erows.size() : 0; }
int b2r_endpoint_rows(const b2r_config* c, uint32_t d, uint64_t* out3, uint64_t capacity_rows) {
    if (d >= c->n_defs || capacity_rows < c->packed[d].erows.size()) { set_error("bad def index / capacity"); return B2R_ERR_INVALID_ARG; }
    const auto& rows = c->packed[d].erows;
    for (size_t i = 0; i < rows.size(); i++) { out3[3 * i] = rows[i].sid; out3[3 * i + 1] = rows[i].start; out3[3 * i + 2] = rows[i].end; }
    return B2R_OK;
}

// ---- the hot call ------------------------------------------------------------------------------------------------------
static int match_batch_impl(b2r_config* c, const uint8_t* d_bytes, const uint64_t* d_offsets, uint64_t n, uint64_t total_bytes,
                            const b2r_outputs* o, uint64_t max_chars, cudaStream_t st) {
    if (c->device < 0) { set_error("this handle was created without a device (device = -1): no CPU fallback exists"); return B2R_ERR_CUDA; }
    if (!o || (n && (!d_bytes && total_bytes)) || (n && !d_offsets)) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    if (!aligned16(d_bytes)) { set_error("bytes must be 16-byte aligned"); return B2R_ERR_ALIGNMENT; }
    int rc = check_outputs(c, o, true);
    if (rc) return rc;
    DeviceGuard g(c->device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", c->device); return B2R_ERR_CUDA; }
    c->last_launches = 0;
    CUDA_TRY(cudaMemsetAsync(c->scratch, 0, c->scratch_bytes, st));
    CUDA_TRY(cudaMemsetAsync(c->scratch, 0xFF, sizeof(unsigned long long), st));  // BatchCounters::first_bad = none
    WalkParams& p = c->last;
    fill_walk_params(c, p, d_bytes, d_offsets, n, total_bytes, o, max_chars);
    c->have_last = true;
    bool wide = false;
    for (uint32_t d = 0; d < c->n_defs; d++) wide = wide || c->packed[d].state_width == 2;
    for (uint32_t d = 0; d < c->n_defs; d++)
    if (n) {
        // walk -> emit hand-over: granule flags, and a scratch state column for every def the caller does not want
        if ((rc = c->ws_fmask.reserve(std::max<size_t>((size_t)p.fm_words * n * 4, 16)))) return rc;
        p.fmask = (uint32_t*)c->ws_fmask.p;
        for (uint32_t d = 0; d < c->n_defs; d++) {
            if (p.def[d].states) continue;
            if ((rc = c->ws_states[d].reserve((size_t)n * o->row_pitch * (wide ? 2 : 1)))) return rc;
            p.def[d].states = c->ws_states[d].p;
        }
        if ((rc = plan_walk(p, wide, c->force_table_mode, c->force_hist_mode))) return rc;
    }
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[0], st));
    if (n) {
        if ((rc = launch_walk(p, wide, st, nullptr))) return rc;
        c->last_launches++;
    }
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[1], st));
    if (n && !p.fuse) {
        if ((rc = launch_emit(p, wide, st, nullptr))) return rc;
        c->last_launches++;
    }
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[2], st));
    rc = enqueue_finalize(c, o, n, max_chars, st);
    if (rc) return rc;
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[3], st));
    return B2R_OK;
}

int b2r_match_batch(b2r_config* c, const uint8_t* d_bytes, const uint64_t* d_offsets, uint64_t n, uint64_t total_bytes,
                    const b2r_outputs* d_out, void* cuda_stream) {
    if (!c) { set_error("null config"); return B2R_ERR_INVALID_ARG; }
    return match_batch_impl(c, d_bytes, d_offsets, n, total_bytes, d_out, c->max_chars, (cudaStream_t)cuda_stream);
}

int b2r_batch_result(b2r_config* c, void* cuda_stream, b2r_batch_status* out) {
    if (!c || c->device < 0 || !c->have_last) { set_error("no batch has been enqueued on this handle"); return B2R_ERR_INVALID_ARG; }
    DeviceGuard g(c->device);
    cudaStream_t st = (cudaStream_t)cuda_stream;
    CUDA_TRY(cudaStreamSynchronize(st));
    BatchCounters h;
    CUDA_TRY(cudaMemcpy(&h, c->scratch, sizeof h, cudaMemcpyDeviceToHost));
    b2r_batch_status r;
    memset(&r, 0, sizeof r);
    r.n_overlap_lo = (uint32_t)h.n_overlap;
    if (h.first_bad != ~0ull) {
        int rc = launch_diagnose(c->last, h.first_bad, c->d_batch_status, st);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(&r, c->d_batch_status, sizeof r, cudaMemcpyDeviceToHost));
        if (r.code == B2R_ERR_INVALID_TRANSITION)
            set_error("The transition from %u by %u is invalid! (string %llu, position %u, def %u)", r.state, (unsigned)r.byte,
                      (unsigned long long)r.string_idx, r.pos, (unsigned)r.def);
        else if (r.code == B2R_ERR_TOO_LONG)
            set_error("string %llu is longer than max_chars_size-1", (unsigned long long)r.string_idx);
    }
    if (out) *out = r;
    return r.code;
}

// ---- host-pointer entry point ----------------------------------------------------------------------------------------
int b2r_match_batch_host(b2r_config* c, const uint8_t* h_bytes, const uint64_t* h_offsets, uint64_t n, const b2r_outputs* ho,
                         b2r_batch_status* result) {
    if (!c || !ho || (n && !h_offsets)) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    if (c->device < 0) { set_error("this handle was created without a device (device = -1): no CPU fallback exists"); return B2R_ERR_CUDA; }
    int rc = check_outputs(c, ho, false);  // same pitches are used on the device; host pointer alignment is irrelevant
    if (rc) return rc;
    DeviceGuard g(c->device);
    cudaStream_t st = c->host_stream;
    const uint64_t total = n ? h_offsets[n] : 0;
    const uint64_t base = n ? h_offsets[0] : 0;
    if (total < base) { set_error("offsets must be non-decreasing"); return B2R_ERR_INVALID_ARG; }
    const uint64_t nbytes = total - base;
    if ((rc = c->ws_bytes.reserve(align_up(nbytes + 16, 256)))) return rc;
    if ((rc = c->ws_offsets.reserve((n + 1) * 8))) return rc;
    // device columns, same layout as the host ones
    const uint64_t rp = ho->row_pitch, bp = ho->bitmap_pitch;
    size_t need = 0;
    auto slot = [&](size_t bytes) { size_t o = need; need += align_up(bytes, 256); return o; };
    // stride: bytes per string (0: not per string).  pinned: page-locked destination, the copy is asynchronous; a copy into
    // pageable memory blocks the calling thread until everything queued before it on its stream is done, so those are
    // issued after the last slice instead of inside the pipeline (they would serialise the H2D of slice i+1 behind the
    // D2H of slice i: measured 110 ms instead of 92 ms per 2^20-string batch with three small pageable columns).
    struct Copy { size_t off; void* host; size_t bytes; size_t stride; bool pinned; };
    auto is_pinned = [](const void* h) {
        cudaPointerAttributes a;
        if (cudaPointerGetAttributes(&a, h) != cudaSuccess) { cudaGetLastError(); return false; }
        return a.type == cudaMemoryTypeHost;
    };
    std::vector<Copy> copies;
    b2r_outputs dout = *ho;
    size_t off_states[B2R_MAX_DEFS], off_sid[B2R_MAX_DEFS], off_se[B2R_MAX_DEFS], off_ee[B2R_MAX_DEFS], off_mult[B2R_MAX_DEFS], off_em[B2R_MAX_DEFS];
    for (uint32_t d = 0; d < c->n_defs; d++) {
        const size_t w = c->packed[d].state_width;
        off_states[d] = ho->states[d] ? slot(n * rp * w) : 0;
        off_sid[d] = ho->substr_ids[d] ? slot(n * rp) : 0;
        off_se[d] = ho->start_enable[d] ? slot(n * bp) : 0;
        off_ee[d] = ho->end_enable[d] ? slot(n * bp) : 0;
        off_mult[d] = ho->mult[d] ? slot(c->packed[d].rows.size() * 8) : 0;
        off_em[d] = ho->endpoint_mult[d] ? slot(c->packed[d].erows.size() * 16) : 0;
    }
    const size_t off_mc = ho->masked_chars ? slot(n * rp) : 0, off_ms = ho->masked_substr_ids ? slot(n * rp) : 0;
    const size_t off_st = ho->status ? slot(n * sizeof(b2r_string_status)) : 0;
    const size_t off_rec = ho->records ? slot(n * (size_t)ho->max_records * sizeof(b2r_substr_record)) : 0;
    const size_t off_cb = ho->compact_bytes ? slot(n * (size_t)ho->compact_pitch) : 0;
    if ((rc = c->ws_cols.reserve(need + 256))) return rc;
    unsigned char* cb = (unsigned char*)c->ws_cols.p;
    auto bind = [&](void* host, size_t off, size_t bytes, size_t stride = 0) -> void* {
        if (!host) return nullptr;
        copies.push_back({off, host, bytes, stride, is_pinned(host)});
        return cb + off;
    };
    const bool acc = (ho->flags & B2R_OUT_ACCUMULATE_MULT) != 0;
    for (uint32_t d = 0; d < c->n_defs; d++) {
        const size_t w = c->packed[d].state_width;
        dout.states[d] = bind(ho->states[d], off_states[d], n * rp * w, rp * w);
        dout.substr_ids[d] = (uint8_t*)bind(ho->substr_ids[d], off_sid[d], n * rp, rp);
        dout.start_enable[d] = (uint8_t*)bind(ho->start_enable[d], off_se[d], n * bp, bp);
        dout.end_enable[d] = (uint8_t*)bind(ho->end_enable[d], off_ee[d], n * bp, bp);
        dout.mult[d] = (uint64_t*)bind(ho->mult[d], off_mult[d], c->packed[d].rows.size() * 8);
        dout.endpoint_mult[d] = (uint64_t*)bind(ho->endpoint_mult[d], off_em[d], c->packed[d].erows.size() * 16);
        if (acc) {
            if (ho->mult[d]) CUDA_TRY(cudaMemcpyAsync(dout.mult[d], ho->mult[d], c->packed[d].rows.size() * 8, cudaMemcpyHostToDevice, st));
            if (ho->endpoint_mult[d]) CUDA_TRY(cudaMemcpyAsync(dout.endpoint_mult[d], ho->endpoint_mult[d], c->packed[d].erows.size() * 16, cudaMemcpyHostToDevice, st));
        }
    }
    dout.masked_chars = (uint8_t*)bind(ho->masked_chars, off_mc, n * rp, rp);
    dout.masked_substr_ids = (uint8_t*)bind(ho->masked_substr_ids, off_ms, n * rp, rp);
    dout.status = (b2r_string_status*)bind(ho->status, off_st, n * sizeof(b2r_string_status), sizeof(b2r_string_status));
    dout.records = (b2r_substr_record*)bind(ho->records, off_rec, n * (size_t)ho->max_records * sizeof(b2r_substr_record), (size_t)ho->max_records * sizeof(b2r_substr_record));
    dout.compact_bytes = (uint8_t*)bind(ho->compact_bytes, off_cb, n * (size_t)ho->compact_pitch, (size_t)ho->compact_pitch);

    // The batch is cut into slices of strings: the H2D copy of slice i+1 and the D2H copy of slice i-1 run on their own
    // streams while the kernels of slice i run (PCIe is full duplex; the copies are the end-to-end bottleneck).
    // inputs: keep the caller's offsets (the kernel adds them to the base pointer, so shift the base instead):
    // d_bytes + offsets[j] must address string j: d_bytes = ws + (base & 15) - base  (16-byte aligned by construction)
    unsigned char* const d_in = (unsigned char*)c->ws_bytes.p + (base & 15);
    const uint8_t* d_bytes = d_in - base;
    const uint64_t* d_offsets = (const uint64_t*)c->ws_offsets.p;
    int n_slices = n >= 16384 ? b2r_config::MAX_SLICES : 1;
    { const char* e = getenv("B2R_SLICES"); if (e && atoi(e) >= 1 && atoi(e) <= b2r_config::MAX_SLICES && n >= 16384) n_slices = atoi(e); }   // testing hook
    const bool trace = getenv("B2R_TRACE_HOST") != nullptr;               // timing aid: where the copies sit on the time line
    cudaEvent_t tev[4] = {};
    if (trace) for (auto& e : tev) CUDA_TRY(cudaEventCreate(&e));
    if (trace) CUDA_TRY(cudaEventRecord(tev[0], st));
    CUDA_TRY(cudaEventRecord(c->ev_fork, st));                           // accumulate uploads / earlier work on the compute stream
    CUDA_TRY(cudaStreamWaitEvent(c->in_stream, c->ev_fork, 0));
    CUDA_TRY(cudaStreamWaitEvent(c->out_stream, c->ev_fork, 0));
    if (n) CUDA_TRY(cudaMemcpyAsync(c->ws_offsets.p, h_offsets, (n + 1) * 8, cudaMemcpyHostToDevice, c->in_stream));
    std::vector<WalkParams> slice_params(n_slices);
    std::vector<uint64_t> slice_lo(n_slices);
    for (int i = 0; i < n_slices; i++) {
        const uint64_t lo = n * (uint64_t)i / n_slices, hi = n * (uint64_t)(i + 1) / n_slices, ni = hi - lo;
        slice_lo[i] = lo;
        const uint64_t b0 = n ? h_offsets[lo] : 0, b1 = n ? h_offsets[hi] : 0;
        if (b1 < b0) { set_error("offsets must be non-decreasing"); return B2R_ERR_INVALID_ARG; }
        if (b1 > b0) CUDA_TRY(cudaMemcpyAsync(d_in + (b0 - base), h_bytes + b0, b1 - b0, cudaMemcpyHostToDevice, c->in_stream));
        CUDA_TRY(cudaEventRecord(c->ev_in[i], c->in_stream));
        CUDA_TRY(cudaStreamWaitEvent(st, c->ev_in[i], 0));
        b2r_outputs ds = dout;                                           // this slice's rows of every column
        for (uint32_t d = 0; d < c->n_defs; d++) {
            const size_t w = c->packed[d].state_width;
            if (ds.states[d]) ds.states[d] = (unsigned char*)ds.states[d] + lo * rp * w;
            if (ds.substr_ids[d]) ds.substr_ids[d] += lo * rp;
            if (ds.start_enable[d]) ds.start_enable[d] += lo * bp;
            if (ds.end_enable[d]) ds.end_enable[d] += lo * bp;
        }
        if (ds.masked_chars) ds.masked_chars += lo * rp;
        if (ds.masked_substr_ids) ds.masked_substr_ids += lo * rp;
        if (ds.status) ds.status += lo;
        if (ds.records) ds.records += lo * (size_t)ho->max_records;
        if (ds.compact_bytes) ds.compact_bytes += lo * (size_t)ho->compact_pitch;
        if (i > 0) ds.flags |= B2R_OUT_ACCUMULATE_MULT;                  // the multiplicities of the slices add up
        rc = match_batch_impl(c, d_bytes, d_offsets + lo, ni, total, &ds, c->max_chars, st);
        if (rc) return rc;
        slice_params[i] = c->last;
        CUDA_TRY(cudaMemcpyAsync(c->h_slices + i, c->scratch, sizeof(BatchCounters), cudaMemcpyDeviceToHost, st));
        CUDA_TRY(cudaEventRecord(c->ev_done[i], st));
        CUDA_TRY(cudaStreamWaitEvent(c->out_stream, c->ev_done[i], 0));
        if (trace && i == 0) CUDA_TRY(cudaEventRecord(tev[2], c->out_stream));
        if (trace && i == n_slices - 1) CUDA_TRY(cudaEventRecord(tev[1], c->in_stream));
        for (const Copy& cp : copies)
            if (cp.stride && cp.pinned && ni) CUDA_TRY(cudaMemcpyAsync((unsigned char*)cp.host + lo * cp.stride, cb + cp.off + lo * cp.stride, ni * cp.stride, cudaMemcpyDeviceToHost, c->out_stream));
    }
    for (const Copy& cp : copies)
        if (cp.stride && !cp.pinned && n) CUDA_TRY(cudaMemcpyAsync(cp.host, cb + cp.off, n * cp.stride, cudaMemcpyDeviceToHost, c->out_stream));
    for (const Copy& cp : copies)
        if (!cp.stride && cp.bytes) CUDA_TRY(cudaMemcpyAsync(cp.host, cb + cp.off, cp.bytes, cudaMemcpyDeviceToHost, st));
    if (trace) CUDA_TRY(cudaEventRecord(tev[3], c->out_stream));
    CUDA_TRY(cudaStreamSynchronize(st));
    CUDA_TRY(cudaStreamSynchronize(c->out_stream));
    if (trace) {
        float h2d_end = 0, d2h_begin = 0, d2h_end = 0;
        cudaEventElapsedTime(&h2d_end, tev[0], tev[1]); cudaEventElapsedTime(&d2h_begin, tev[0], tev[2]); cudaEventElapsedTime(&d2h_end, tev[0], tev[3]);
        fprintf(stderr, "[b2r] host call, %d slices: last H2D done at %.2f ms, first D2H starts at %.2f ms, last D2H done at %.2f ms\n", n_slices, h2d_end, d2h_begin, d2h_end);
        for (auto& e : tev) cudaEventDestroy(e);
    }

    // the batch result: the lowest failing string over all slices (reference: the first panic), overlaps summed
    b2r_batch_status r;
    memset(&r, 0, sizeof r);
    uint64_t n_overlap = 0;
    for (int i = 0; i < n_slices; i++) n_overlap += c->h_slices[i].n_overlap;
    for (int i = 0; i < n_slices; i++) {
        if (c->h_slices[i].first_bad == ~0ull) continue;
        rc = launch_diagnose(slice_params[i], c->h_slices[i].first_bad, c->d_batch_status, st);
        if (rc) return rc;
        CUDA_TRY(cudaStreamSynchronize(st));
        CUDA_TRY(cudaMemcpy(&r, c->d_batch_status, sizeof r, cudaMemcpyDeviceToHost));
        r.string_idx += slice_lo[i];
        if (r.code == B2R_ERR_INVALID_TRANSITION)
            set_error("The transition from %u by %u is invalid! (string %llu, position %u, def %u)", r.state, (unsigned)r.byte,
                      (unsigned long long)r.string_idx, r.pos, (unsigned)r.def);
        else if (r.code == B2R_ERR_TOO_LONG)
            set_error("string %llu is longer than max_chars_size-1", (unsigned long long)r.string_idx);
        break;
    }
    r.n_overlap_lo = (uint32_t)n_overlap;
    if (result) *result = r;
    return r.code;
}

int b2r_match_substrs(b2r_config* c, const uint8_t* characters, uint64_t len, const b2r_outputs* h_out, b2r_batch_status* result) {
    if (!c || !h_out) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    const uint64_t offsets[2] = {0, len};
    return b2r_match_batch_host(c, characters, offsets, 1, h_out, result);
}

int b2r_match_long(b2r_config* c, const uint8_t* d_bytes, uint64_t len, const b2r_outputs* o, void* cuda_stream) {
    if (!c || !o || (!d_bytes && len)) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    if (c->device < 0) { set_error("this handle was created without a device (device = -1): no CPU fallback exists"); return B2R_ERR_CUDA; }
    if (len + 1 > 0xFFFFFFF0ull) { set_error("string of %llu bytes: at most 2^32 - 17 rows are supported", (unsigned long long)len); return B2R_ERR_UNSUPPORTED; }
    if (!aligned16(d_bytes)) { set_error("bytes must be 16-byte aligned"); return B2R_ERR_ALIGNMENT; }
    const uint64_t M = len + 1;
    int rc = check_outputs(c, o, true, M);
    if (rc) return rc;
    DeviceGuard g(c->device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", c->device); return B2R_ERR_CUDA; }
    cudaStream_t st = (cudaStream_t)cuda_stream;
    bool wide = false;
    for (uint32_t d = 0; d < c->n_defs; d++) wide = wide || c->packed[d].state_width == 2;
    c->last_launches = 0;
    CUDA_TRY(cudaMemsetAsync(c->scratch, 0, c->scratch_bytes, st));
    CUDA_TRY(cudaMemsetAsync(c->scratch, 0xFF, sizeof(unsigned long long), st));  // BatchCounters::first_bad = none

    // ---- workspace: chunk offsets (+ the {0, len} pair of the whole string), flag words, summary, tree of transition vectors
    const uint32_t n_chunks = (uint32_t)std::max<uint64_t>(1, (len + LONG_CHUNK - 1) / LONG_CHUNK);
    const uint32_t chunk_fm_words = (LONG_CHUNK / 16 + 31) / 32;
    const size_t nodes = long_level_nodes(n_chunks);
    size_t need = 0;
    auto slot = [&](size_t bytes) { size_t off = need; need += align_up(bytes, 256); return off; };
    const size_t off_offsets = slot((size_t)(n_chunks + 3) * 8), off_fmask = slot((size_t)chunk_fm_words * n_chunks * 4), off_summary = slot((size_t)((n_chunks + 31) / 32) * 4);
    size_t off_maps[B2R_MAX_DEFS], off_entry[B2R_MAX_DEFS], off_states[B2R_MAX_DEFS];
    for (uint32_t d = 0; d < c->n_defs; d++) {
        off_maps[d] = slot(nodes * (c->packed[d].num_states + 1) * 2);
        off_entry[d] = slot(nodes * 2);
        off_states[d] = o->states[d] ? 0 : slot(align_up(M, 16) * (wide ? 2 : 1) + 64);
    }
    uint32_t max_s1 = 0;
    for (uint32_t d = 0; d < c->n_defs; d++) max_s1 = std::max(max_s1, c->packed[d].num_states + 1);
    const size_t off_uniq = slot((size_t)n_chunks * 4 * 2), off_nuniq = slot(n_chunks), off_which = slot((size_t)n_chunks * max_s1);
    if ((rc = c->ws_long.reserve(need))) return rc;
    unsigned char* ws = (unsigned char*)c->ws_long.p;
    uint64_t* d_offsets = (uint64_t*)(ws + off_offsets);
    const uint64_t whole[2] = {0, len};
    CUDA_TRY(cudaMemcpyAsync(d_offsets + n_chunks + 1, whole, sizeof whole, cudaMemcpyHostToDevice, st));

    // ---- 1 + 2: transition vectors of the chunks, composition tree, entry state of every chunk ------------------------------
    LongParams lp;
    memset(&lp, 0, sizeof lp);
    lp.bytes = d_bytes; lp.len = len; lp.n_chunks = n_chunks; lp.n_defs = c->n_defs; lp.offsets = d_offsets;
    lp.uniq = (uint16_t*)(ws + off_uniq); lp.n_uniq = ws + off_nuniq; lp.which = ws + off_which;
    for (uint32_t d = 0; d < c->n_defs; d++) {
        lp.def[d].byte_class = c->dev[d].byte_class; lp.def[d].trans = c->dev[d].trans;
        lp.def[d].num_states = c->packed[d].num_states; lp.def[d].first_state = c->packed[d].first_state;
        lp.def[d].num_classes = c->packed[d].num_classes;
        lp.def[d].maps = (uint16_t*)(ws + off_maps[d]); lp.def[d].entry = (uint16_t*)(ws + off_entry[d]);
    }
    // the sparse columns are zeroed on a side stream while the (latency-bound) prefix stages and the walk run; joined before
    // the emit stage
    {
        cudaStream_t zs = c->in_stream;
        CUDA_TRY(cudaEventRecord(c->ev_fork, st));
        CUDA_TRY(cudaStreamWaitEvent(zs, c->ev_fork, 0));
        auto zero = [&](void* ptr, size_t bytes) { return ptr ? cudaMemsetAsync(ptr, 0, bytes, zs) : cudaSuccess; };
        const size_t col_bytes = align_up(M, 16), bm_bytes = align_up((M + 7) / 8, 4);
        for (uint32_t d = 0; d < c->n_defs; d++) {
            CUDA_TRY(zero(o->substr_ids[d], col_bytes));
            CUDA_TRY(zero(o->start_enable[d], bm_bytes));
            CUDA_TRY(zero(o->end_enable[d], bm_bytes));
        }
        CUDA_TRY(zero(o->masked_chars, col_bytes));
        CUDA_TRY(zero(o->masked_substr_ids, col_bytes));
        CUDA_TRY(cudaEventRecord(c->ev_done[0], zs));
    }
    if ((rc = launch_long_prepare(lp, st, &c->last_launches))) return rc;

    // ---- 3: the walk, chunk = string, one contiguous state row ------------------------------------------------------------
    b2r_outputs seg = *o;
    seg.row_pitch = LONG_CHUNK;
    WalkParams pw;
    fill_walk_params(c, pw, d_bytes, d_offsets, n_chunks, len, &seg, LONG_CHUNK + 1);
    pw.fuse = 0; pw.prefilled = 0; pw.segment_mode = 1;
    pw.fmask = (uint32_t*)(ws + off_fmask);
    for (uint32_t d = 0; d < c->n_defs; d++) {
        if (!pw.def[d].states) pw.def[d].states = ws + off_states[d];
        pw.def[d].init_states = lp.def[d].entry;                          // level 0 of the tree
    }
    if ((rc = plan_walk(pw, wide, c->force_table_mode, c->force_hist_mode))) return rc;
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[0], st));
    if ((rc = launch_walk(pw, wide, st, nullptr))) return rc;
    c->last_launches++;
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[1], st));

    // ---- 4: the ordered emit stage of the one string (sparse columns zeroed by now) --------------------------------------------
    CUDA_TRY(cudaStreamWaitEvent(st, c->ev_done[0], 0));
    WalkParams& pe = c->last;
    fill_walk_params(c, pe, d_bytes, d_offsets + n_chunks + 1, 1, len, o, M);
    pe.fuse = 0; pe.prefilled = 1;
    pe.fmask = pw.fmask;
    for (uint32_t d = 0; d < c->n_defs; d++) pe.def[d].states = pw.def[d].states;
    c->have_last = true;
    if ((rc = launch_long_emit(pe, wide, pw.fmask, (uint32_t*)(ws + off_summary), n_chunks, chunk_fm_words, st, &c->last_launches))) return rc;
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[2], st));
    rc = enqueue_finalize(c, o, 1, M, st);
    if (rc) return rc;
    if (c->timing) CUDA_TRY(cudaEventRecord(c->ev[3], st));
    return B2R_OK;
}

// Host-pointer variant of the long-string path: the string goes up in one copy, the columns come back in one copy each.
int b2r_match_long_host(b2r_config* c, const uint8_t* h_bytes, uint64_t len, const b2r_outputs* ho, b2r_batch_status* result) {
    if (!c || !ho || (!h_bytes && len)) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    if (c->device < 0) { set_error("this handle was created without a device (device = -1): no CPU fallback exists"); return B2R_ERR_CUDA; }
    const uint64_t M = len + 1;
    int rc = check_outputs(c, ho, false, M);
    if (rc) return rc;
    DeviceGuard g(c->device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", c->device); return B2R_ERR_CUDA; }
    cudaStream_t st = c->host_stream;
    if ((rc = c->ws_bytes.reserve(align_up(len + 16, 256)))) return rc;
    const uint64_t rp = ho->row_pitch, bp = ho->bitmap_pitch;
    size_t need = 0;
    auto slot = [&](size_t bytes) { size_t o = need; need += align_up(bytes, 256); return o; };
    struct Copy { size_t off; void* host; size_t bytes; };
    std::vector<Copy> copies;
    std::vector<size_t> offs;
    b2r_outputs dout = *ho;
    auto want = [&](void* host, size_t bytes) -> size_t {
        if (!host) { offs.push_back(0); return 0; }
        const size_t o = slot(bytes);
        copies.push_back({o, host, bytes});
        offs.push_back(o);
        return o;
    };
    // first pass: sizes; second pass (after the reserve): device pointers
    for (uint32_t d = 0; d < c->n_defs; d++) {
        want(ho->states[d], rp * c->packed[d].state_width); want(ho->substr_ids[d], rp); want(ho->start_enable[d], bp); want(ho->end_enable[d], bp);
        want(ho->mult[d], c->packed[d].rows.size() * 8); want(ho->endpoint_mult[d], c->packed[d].erows.size() * 16);
    }
    want(ho->masked_chars, rp); want(ho->masked_substr_ids, rp); want(ho->status, sizeof(b2r_string_status));
    want(ho->records, (size_t)ho->max_records * sizeof(b2r_substr_record)); want(ho->compact_bytes, (size_t)ho->compact_pitch);
    if ((rc = c->ws_cols.reserve(need + 256))) return rc;
    unsigned char* cb = (unsigned char*)c->ws_cols.p;
    size_t k = 0;
    auto dev = [&](void* host) -> void* { const size_t o = offs[k++]; return host ? cb + o : nullptr; };
    for (uint32_t d = 0; d < c->n_defs; d++) {
        dout.states[d] = dev(ho->states[d]); dout.substr_ids[d] = (uint8_t*)dev(ho->substr_ids[d]);
        dout.start_enable[d] = (uint8_t*)dev(ho->start_enable[d]); dout.end_enable[d] = (uint8_t*)dev(ho->end_enable[d]);
        dout.mult[d] = (uint64_t*)dev(ho->mult[d]); dout.endpoint_mult[d] = (uint64_t*)dev(ho->endpoint_mult[d]);
        if (ho->flags & B2R_OUT_ACCUMULATE_MULT) {
            if (ho->mult[d]) CUDA_TRY(cudaMemcpyAsync(dout.mult[d], ho->mult[d], c->packed[d].rows.size() * 8, cudaMemcpyHostToDevice, st));
            if (ho->endpoint_mult[d]) CUDA_TRY(cudaMemcpyAsync(dout.endpoint_mult[d], ho->endpoint_mult[d], c->packed[d].erows.size() * 16, cudaMemcpyHostToDevice, st));
        }
    }
    dout.masked_chars = (uint8_t*)dev(ho->masked_chars); dout.masked_substr_ids = (uint8_t*)dev(ho->masked_substr_ids);
    dout.status = (b2r_string_status*)dev(ho->status); dout.records = (b2r_substr_record*)dev(ho->records);
    dout.compact_bytes = (uint8_t*)dev(ho->compact_bytes);
    if (len) CUDA_TRY(cudaMemcpyAsync(c->ws_bytes.p, h_bytes, len, cudaMemcpyHostToDevice, st));
    if ((rc = b2r_match_long(c, (const uint8_t*)c->ws_bytes.p, len, &dout, st))) return rc;
    for (const Copy& cp : copies) CUDA_TRY(cudaMemcpyAsync(cp.host, cb + cp.off, cp.bytes, cudaMemcpyDeviceToHost, st));
    return b2r_batch_result(c, st, result);
}

uint32_t b2r_last_launch_count(const b2r_config* c) { return c ? c->last_launches : 0; }
int b2r_config_set_timing(b2r_config* c, int enable) { if (!c) return B2R_ERR_INVALID_ARG; c->timing = enable != 0; return B2R_OK; }
int b2r_last_kernel_ms(b2r_config* c, float* walk_ms, float* total_ms) {
    if (!c || !c->timing || c->device < 0) { set_error("timing is not enabled on this handle"); return B2R_ERR_INVALID_ARG; }
    DeviceGuard g(c->device);
    CUDA_TRY(cudaEventSynchronize(c->ev[3]));
    if (walk_ms) CUDA_TRY(cudaEventElapsedTime(walk_ms, c->ev[0], c->ev[1]));
    if (total_ms) CUDA_TRY(cudaEventElapsedTime(total_ms, c->ev[0], c->ev[3]));
    return B2R_OK;
}
int b2r_last_stage_ms(b2r_config* c, float* ms3) {
    if (!c || !c->timing || c->device < 0 || !ms3) { set_error("timing is not enabled on this handle"); return B2R_ERR_INVALID_ARG; }
    DeviceGuard g(c->device);
    CUDA_TRY(cudaEventSynchronize(c->ev[3]));
    for (int i = 0; i < 3; i++) CUDA_TRY(cudaEventElapsedTime(ms3 + i, c->ev[i], c->ev[i + 1]));
    return B2R_OK;
}
int b2r_last_plan(const b2r_config* c, uint32_t* table_mode, uint32_t* hist_mode) {
    if (!c || !c->have_last) { set_error("no batch has been enqueued on this handle"); return B2R_ERR_INVALID_ARG; }
    if (table_mode) *table_mode = c->last.table_mode;
    if (hist_mode) *hist_mode = c->last.hist_mode;
    return B2R_OK;
}

}  // extern "C"
