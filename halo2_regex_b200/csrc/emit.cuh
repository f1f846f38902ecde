// emit_kernel — stage 2 of the batch path: everything the reference derives from the state sequence.
//
// One WARP per string.  Inputs: the string, its state columns and its granule flags (walk.cuh).  Outputs, written as whole
// rows (zeros and values merged before the store, so DRAM sees exactly the algorithmic bytes):
//   per-def substr ids (src/lib.rs:825-845), start_enable / end_enable bitmaps (:482-513), the endpoint-lookup
//   multiplicities (:235-284), masked_chars / masked_substr_ids (:740-764), substring records + compact bytes, the
//   accept flag (:427-457) and the status record.
//
// Step 1: every 16-byte vector of every column whose granule is NOT flagged is zero (a row with a non-zero value has a
//         non-zero substr id in some def, and start/end flags require one): coalesced 16-byte zero stores.
// Step 2: flagged 32-row windows, in order, lane = row: the row's packed entries are looked up again from
//         (byte, state) — one lookup replaces the reference's HashSet probes and `contains` scans — and give substr id,
//         is_start, is_end(next row).  Ballots turn them into bitmap words; shuffles give the neighbour rows.
// Masks.  With b_1 < b_2 < ... the rows where the id sum changes AND is_start_sum|is_end_sum is set, the forward scan
//         (src/lib.rs:598-645) sets start_mask at b_k when is_start_sum[b_k] and resets it when only is_end_sum[b_k]; the
//         backward scan (:663-714) sets end_mask for the rows before b_{k+1} when is_end_sum[b_{k+1}] and resets it when
//         only is_start_sum[b_{k+1}].  Hence mask = 1 exactly on [b_k, b_{k+1}) with is_start at b_k and is_end at b_{k+1}.
//         The boundaries are streamed in order; a closing boundary makes the warp overwrite the masked rows of
//         [b_k, b_{k+1}) (they may extend over unflagged granules) — same L2 lines it has just zeroed.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "defs.hpp"
#include "kernels.cuh"

namespace b2r {

constexpr uint32_t NO_POS = 0xFFFFFFFFu;
constexpr int EMIT_THREADS = 256;

__device__ __forceinline__ uint32_t ent_sid(uint32_t e) { return (e >> ENT_SID_SHIFT) & 0xFFu; }

// Failure details of string j in the reference's order: derive_states walks def 0 over the whole string first, then
// def 1, ... (src/lib.rs:806-821), so the panic belongs to the LOWEST def index that fails, at its first failing byte.
static __device__ __noinline__ b2r_batch_status diagnose_string(const WalkParams& p, uint64_t j) {
    b2r_batch_status r = {};
    r.string_idx = j;
    const uint64_t off = p.offsets[j], end = p.offsets[j + 1];
    if (end < off || end - off > (uint64_t)(p.max_chars - 1)) {
        r.code = B2R_ERR_TOO_LONG; r.pos = NO_POS;
        return r;
    }
    for (uint32_t d = 0; d < p.n_defs && r.code == 0; d++) {
        uint32_t s = p.def[d].first_state;
        for (uint64_t i = off; i < end; i++) {
            const uint32_t c = p.bytes[i];
            const uint32_t e = p.def[d].trans[(uint32_t)p.def[d].byte_class[c] * p.def[d].num_states + s];
            if (e & ENT_INVALID) {
                r.code = B2R_ERR_INVALID_TRANSITION; r.pos = (uint32_t)(i - off); r.state = s; r.byte = (uint8_t)c; r.def = (uint8_t)d;
                break;
            }
            s = e & ENT_NEXT_MASK;
        }
    }
    return r;
}

// the reference panics (src/lib.rs:817) / the string does not fit: mark the string, remember the lowest failing index
static __device__ __noinline__ void kill_string(const WalkParams& p, uint64_t idx) {
    atomicMin(&p.counters->first_bad, (unsigned long long)idx);
    if (p.status) {
        const b2r_batch_status r = diagnose_string(p, idx);
        b2r_string_status st = {};
        st.flags = (r.code == B2R_ERR_TOO_LONG) ? B2R_ST_TOO_LONG : B2R_ST_INVALID_TRANSITION;
        st.err_pos = (r.code == B2R_ERR_TOO_LONG) ? NO_POS : r.pos; st.err_state = r.state; st.err_byte = r.byte; st.err_def = r.def;
        p.status[idx] = st;
    }
}

// per-CTA view of the lookup tables (shared memory when they fit, else global) and the endpoint counters
template <int D>
struct EmitTables {
    const uint8_t* cls[D];
    const uint32_t* trans[D];
    uint32_t* ep_s[D];          // shared-memory endpoint counters: [0,K*S) start lookups, [K*S,2*K*S) end lookups; null = global
};

// per-warp running sums published once at the end of the kernel
struct EmitTotals {
    unsigned long long pad_rows = 0;
    uint32_t n_ok = 0, n_overlap = 0;
};

// rows [a,b) of string j are masked: start_mask = end_mask = 1 (src/lib.rs:740-764), b <= len.  Writes masked_chars,
// masked_substr_ids, the compact bytes and one record per maximal run of a constant id sum.  Warp-cooperative; every
// argument is warp-uniform.  Returns the updated (n_rec, n_cmp).
template <int D, typename ST>
static __device__ __noinline__ uint2 emit_segment(const WalkParams& p, const EmitTables<D> tb, uint64_t j, const uint8_t* src, uint32_t a, uint32_t b,
                                                  uint32_t n_rec, uint32_t n_cmp) {
    const int lane = threadIdx.x & 31;
    uint8_t* const mc = p.masked_chars ? p.masked_chars + j * p.row_pitch : nullptr;
    uint8_t* const ms = p.masked_substr_ids ? p.masked_substr_ids + j * p.row_pitch : nullptr;
    uint8_t* const cb = p.compact_bytes ? p.compact_bytes + j * (uint64_t)p.compact_pitch : nullptr;
    auto record = [&](uint32_t start, uint32_t len, uint32_t sid, uint32_t coff) {
        if (lane == 0 && p.records && n_rec < p.max_records) {
            b2r_substr_record r; r.start = start; r.len = len; r.substr_id = sid; r.compact_off = coff;
            p.records[j * p.max_records + n_rec] = r;
        }
        n_rec++;
    };
    uint32_t run_start = a, run_sum = 0, last_sum = 0;
    for (uint32_t base = a; base < b; base += 32) {
        const uint32_t i = base + lane;
        const bool in = i < b;
        uint32_t c = 0, sum = 0;
        if (in) {
            c = src[i];
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t S = p.def[d].num_states;
                const uint32_t s = (uint32_t) reinterpret_cast<const ST*>(p.def[d].states)[j * p.row_pitch + i];
                if (s < S) sum += ent_sid(tb.trans[d][(uint32_t)tb.cls[d][c] * S + s]);
            }
            if (mc) mc[i] = (uint8_t)c;
            if (ms) ms[i] = (uint8_t)sum;
            const uint32_t k = n_cmp + (i - a);
            if (cb && k < p.compact_pitch) cb[k] = (uint8_t)c;
        }
        // records: maximal runs of a constant id sum
        uint32_t before = __shfl_up_sync(0xffffffffu, sum, 1);
        if (lane == 0) before = last_sum;
        const bool starts = in && (i == a || sum != before);
        uint32_t sm = __ballot_sync(0xffffffffu, starts);
        while (sm) {
            const int l = __ffs((int)sm) - 1;
            sm &= sm - 1;
            const uint32_t pos = base + l;
            const uint32_t s_here = __shfl_sync(0xffffffffu, sum, l);
            if (pos != a) record(run_start, pos - run_start, run_sum, n_cmp + (run_start - a));
            run_start = pos; run_sum = s_here;
        }
        last_sum = __shfl_sync(0xffffffffu, sum, 31);
    }
    record(run_start, b - run_start, run_sum, n_cmp + (run_start - a));
    return make_uint2(n_rec, n_cmp + (b - a));
}

template <int D, typename ST>
struct StringEmitter {
    const WalkParams& p;
    const EmitTables<D>& tb;
    const int lane;
    // the string
    uint64_t j;
    const uint8_t* src;
    uint32_t L;
    // streaming state (warp-uniform)
    bool prev_valid, prev_is;   // last boundary: is_start_sum set there
    uint32_t prev_pos;
    bool carry_valid;           // rows up to carry_pos-1 have been examined; carry_s / carry_ie belong to row carry_pos-1
    uint32_t carry_pos, carry_s, carry_ie;
    uint32_t n_rec, n_cmp;
    bool overlap;

    __device__ __forceinline__ StringEmitter(const WalkParams& p_, const EmitTables<D>& tb_, int lane_) : p(p_), tb(tb_), lane(lane_) {}

    __device__ __forceinline__ uint32_t state_at(int d, uint32_t row) const {
        return (uint32_t) reinterpret_cast<const ST*>(p.def[d].states)[j * p.row_pitch + row];
    }
    // packed entry of row `row` (< L): the transition taken from state s on byte c
    __device__ __forceinline__ uint32_t entry(int d, uint32_t c, uint32_t s) const {
        const uint32_t S = p.def[d].num_states;
        if (s >= S) return ENT_INVALID;                                  // the walk parked this def in the trap state
        return tb.trans[d][(uint32_t)tb.cls[d][c] * S + s];
    }

    // a row where the id sum changes and is_start_sum | is_end_sum is set
    __device__ __forceinline__ void boundary(uint32_t pos, bool b_is, bool b_ie) {
        if (prev_valid && prev_is && b_ie) {
            const uint2 r = emit_segment<D, ST>(p, tb, j, src, prev_pos, pos, n_rec, n_cmp);
            n_rec = r.x; n_cmp = r.y;
        }
        prev_valid = true; prev_pos = pos; prev_is = b_is;
    }

    // the rows after carry_pos-1 are not examined (their granule is not flagged: id sum 0, no is_start); row carry_pos can
    // still be a boundary: the id sum drops to 0 there and is_end_sum[carry_pos] comes from row carry_pos-1
    __device__ __forceinline__ void close_gap() {
        if (carry_valid && carry_s != 0 && carry_ie != 0) {
            if (carry_ie > 1) overlap = true;
            boundary(carry_pos, false, true);
        }
        carry_valid = false; carry_s = 0; carry_ie = 0;
    }

    // rows [32t, 32t+32); halves: bit h set = granule 2t+h is flagged (only those bytes are written here).  Returns false
    // when the string hit an invalid transition.
    __device__ __forceinline__ bool window(uint32_t t, uint32_t halves) {
        const uint32_t i = 32 * t + lane;
        if (carry_valid && carry_pos != 32 * t) close_gap();
        const bool is_char = i < L;
        uint32_t c = 0;
        if (is_char) c = src[i];
        uint32_t sum = 0, is_sum = 0, ie_next = 0, inval = 0;
        uint32_t e[D], s[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            e[d] = 0; s[d] = 0;
            if (is_char) {
                s[d] = state_at(d, i);
                e[d] = entry(d, c, s[d]);
                inval |= e[d] & ENT_INVALID;
                sum += ent_sid(e[d]);
                is_sum += (e[d] >> 24) & 1u;
                ie_next += (e[d] >> 25) & 1u;
            }
        }
        if (__any_sync(0xffffffffu, inval != 0)) return false;
        uint32_t sum_before = __shfl_up_sync(0xffffffffu, sum, 1);
        uint32_t ie_here = __shfl_up_sync(0xffffffffu, ie_next, 1);
        if (lane == 0) { sum_before = carry_s; ie_here = carry_ie; }     // zero unless the previous window was examined too
        const bool bnd = (sum != sum_before) && (is_sum | ie_here);
        if (__any_sync(0xffffffffu, is_sum > 1 || ie_here > 1)) overlap = true;

        // ---- per-def columns --------------------------------------------------------------------------------------
        const bool mine = (halves >> (lane >> 4)) & 1u;                  // my half of the window is flagged (else zero-filled in step 1)
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t S = p.def[d].num_states;
            const uint32_t sid = ent_sid(e[d]);
            if (mine && p.def[d].substr_ids && i < p.max_chars) p.def[d].substr_ids[j * p.row_pitch + i] = (uint8_t)sid;
            const bool st = (e[d] & ENT_IS_START) != 0, en = (e[d] & ENT_IS_END) != 0;
            const uint32_t sw = __ballot_sync(0xffffffffu, st), ew = __ballot_sync(0xffffffffu, en);
            if (lane == 0) {                                             // the whole bitmap word belongs to this window
                if (p.def[d].start_enable) *reinterpret_cast<uint32_t*>(p.def[d].start_enable + j * p.bitmap_pitch + 4 * t) = sw;
                if (p.def[d].end_enable) *reinterpret_cast<uint32_t*>(p.def[d].end_enable + j * p.bitmap_pitch + 4 * t) = ew;
            }
            if (st) {                                                    // endpoint lookup, src/lib.rs:235-258
                const uint32_t bin = (sid - p.def[d].sid_offset) * S + s[d];
                if (tb.ep_s[d]) atomicAdd(tb.ep_s[d] + bin, 1u); else atomicAdd(p.def[d].ep_start + bin, 1ull);
            }
            if (en) {                                                    // src/lib.rs:260-284
                const uint32_t bin = (sid - p.def[d].sid_offset) * S + (e[d] & ENT_NEXT_MASK);
                if (tb.ep_s[d]) atomicAdd(tb.ep_s[d] + p.def[d].num_substrs * S + bin, 1u); else atomicAdd(p.def[d].ep_end + bin, 1ull);
            }
        }
        if (mine && i < p.max_chars) {                                   // masked rows are overwritten when their segment closes
            if (p.masked_chars) p.masked_chars[j * p.row_pitch + i] = 0;
            if (p.masked_substr_ids) p.masked_substr_ids[j * p.row_pitch + i] = 0;
        }
        __syncwarp();

        // ---- boundaries, in order ----------------------------------------------------------------------------------
        uint32_t bm = __ballot_sync(0xffffffffu, bnd);
        const uint32_t ism = __ballot_sync(0xffffffffu, is_sum != 0), iem = __ballot_sync(0xffffffffu, ie_here != 0);
        while (bm) {
            const int l = __ffs((int)bm) - 1;
            bm &= bm - 1;
            boundary(32 * t + l, (ism >> l) & 1u, (iem >> l) & 1u);
        }
        carry_valid = true; carry_pos = 32 * t + 32;
        carry_s = __shfl_sync(0xffffffffu, sum, 31);
        carry_ie = __shfl_sync(0xffffffffu, ie_next, 31);
        return true;
    }

    __device__ __forceinline__ void run(uint64_t j_, EmitTotals& tot) {
        j = j_;
        const uint64_t N = p.n_strings;
        const uint32_t M = p.max_chars;
        const uint64_t off = p.offsets[j], end = p.offsets[j + 1];
        if (end < off || end - off > (uint64_t)(M - 1)) {                // len must be <= M-1
            if (lane == 0) kill_string(p, j);
            return;
        }
        src = p.bytes + off;
        L = (uint32_t)(end - off);
        const uint64_t rp = p.row_pitch, bp = p.bitmap_pitch;

        // ---- step 1: zero every vector whose granule is not flagged -------------------------------------------------
        const uint32_t nvec = (M + 15) / 16;
        const uint4 z = make_uint4(0, 0, 0, 0);
        for (uint32_t k = 0; k * 32 < nvec; k++) {
            const uint32_t v = k * 32 + lane;
            const uint32_t fw = k < p.fm_words ? __ldg(p.fmask + (size_t)k * N + j) : 0u;
            if (v < nvec && !((fw >> lane) & 1u)) {
                const uint64_t o = j * rp + 16ull * v;
#pragma unroll
                for (int d = 0; d < D; d++)
                    if (p.def[d].substr_ids) *reinterpret_cast<uint4*>(p.def[d].substr_ids + o) = z;
                if (p.masked_chars) *reinterpret_cast<uint4*>(p.masked_chars + o) = z;
                if (p.masked_substr_ids) *reinterpret_cast<uint4*>(p.masked_substr_ids + o) = z;
            }
        }
        const uint32_t nbw = (M + 31) / 32;                              // bitmap words: word t = rows [32t, 32t+32) = granules 2t, 2t+1
        for (uint32_t k = 0; k * 32 < nbw; k++) {
            const uint32_t t = k * 32 + lane;
            const uint32_t wi = (2 * t) >> 5;
            const uint32_t fw = (t < nbw && wi < p.fm_words) ? __ldg(p.fmask + (size_t)wi * N + j) : 0u;
            if (t < nbw && !((fw >> ((2 * t) & 31)) & 3u)) {
#pragma unroll
                for (int d = 0; d < D; d++) {
                    if (p.def[d].start_enable) *reinterpret_cast<uint32_t*>(p.def[d].start_enable + j * bp + 4 * t) = 0u;
                    if (p.def[d].end_enable) *reinterpret_cast<uint32_t*>(p.def[d].end_enable + j * bp + 4 * t) = 0u;
                }
            }
        }
        __syncwarp();

        // ---- step 2: flagged windows, in order ----------------------------------------------------------------------
        prev_valid = false; prev_is = false; prev_pos = 0;
        carry_valid = false; carry_pos = 0; carry_s = 0; carry_ie = 0;
        n_rec = 0; n_cmp = 0; overlap = false;
        for (uint32_t w = 0; w < p.fm_words; w++) {
            uint32_t bits = __ldg(p.fmask + (size_t)w * N + j);
            while (bits) {
                const uint32_t g = (uint32_t)__ffs((int)bits) - 1u;
                const uint32_t lo = g & ~1u;
                const uint32_t halves = (bits >> lo) & 3u;
                bits &= ~(3u << lo);
                if (!window((w * 32 + g) >> 1, halves)) {
                    if (lane == 0) kill_string(p, j);
                    return;
                }
            }
        }
        close_gap();

        // ---- accept rule (src/lib.rs:427-457) and the status record ----------------------------------------------------
        uint32_t flags = 0;
#pragma unroll
        for (int d = 0; d < D; d++)
            if (state_at(d, L) == p.def[d].accepted_state) flags |= B2R_ST_ACCEPTED(d);
        if (overlap) flags |= B2R_ST_OVERLAP;
        if (p.records && n_rec > p.max_records) flags |= B2R_ST_RECORDS_TRUNCATED;
        if (p.compact_bytes && n_cmp > p.compact_pitch) flags |= B2R_ST_COMPACT_TRUNCATED;
        if (p.status && lane == 0) {
            b2r_string_status st = {};
            st.flags = flags; st.err_pos = NO_POS; st.n_records = n_rec; st.n_compact = n_cmp;
            p.status[j] = st;
        }
        tot.pad_rows += M - L; tot.n_ok++; tot.n_overlap += overlap ? 1u : 0u;
    }
};

template <int D, typename ST>
__global__ void __launch_bounds__(EMIT_THREADS) emit_kernel(const __grid_constant__ WalkParams p) {
    extern __shared__ __align__(16) unsigned char esmem[];
    const int lane = threadIdx.x & 31;
    EmitTables<D> tb;
    // shared memory: endpoint counters, then (optionally) the lookup tables
    uint32_t off = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        tb.ep_s[d] = p.ep_smem_bytes ? reinterpret_cast<uint32_t*>(esmem + off) : nullptr;
        if (p.ep_smem_bytes) off += 2u * p.def[d].num_substrs * p.def[d].num_states * 4u;
    }
    off = (off + 15u) & ~15u;
    for (uint32_t i = threadIdx.x; i < off / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(esmem)[i] = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        if (p.emit_smem_tables) {
            const uint32_t n = p.def[d].num_classes * p.def[d].num_states;
            uint32_t* const t = reinterpret_cast<uint32_t*>(esmem + off);
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) t[i] = p.def[d].trans[i];
            off += n * 4;
            uint8_t* const c = esmem + off;
            for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) c[i] = p.def[d].byte_class[i];
            off += 256;
            tb.trans[d] = t; tb.cls[d] = c;
        } else {
            tb.trans[d] = p.def[d].trans; tb.cls[d] = p.def[d].byte_class;
        }
    }
    __syncthreads();

    const uint64_t warps_total = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t gw = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    StringEmitter<D, ST> em(p, tb, lane);
    EmitTotals tot;
    for (uint64_t j = gw; j < p.n_strings; j += warps_total) em.run(j, tot);

    if (lane == 0) {
        if (tot.pad_rows) atomicAdd(&p.counters->pad_rows, tot.pad_rows);
        if (tot.n_overlap) atomicAdd(&p.counters->n_overlap, (unsigned long long)tot.n_overlap);
        if (tot.n_ok) atomicAdd(&p.counters->n_ok_strings, (unsigned long long)tot.n_ok);
    }
    __syncthreads();
#pragma unroll
    for (int d = 0; d < D; d++) {
        if (!tb.ep_s[d]) continue;
        const uint32_t ks = p.def[d].num_substrs * p.def[d].num_states;
        for (uint32_t i = threadIdx.x; i < 2 * ks; i += blockDim.x) {
            const uint32_t v = tb.ep_s[d][i];
            if (v) atomicAdd((i < ks ? p.def[d].ep_start + i : p.def[d].ep_end + (i - ks)), (unsigned long long)v);
        }
    }
}

}  // namespace b2r
