// Lane-private "rare row" machinery shared by the walk kernels (sm_100a).
//
// A row is RARE when one of its packed entries changes the per-def substr id, carries is_start / is_end, or is an
// invalid transition.  Everything the reference derives beyond the state column hangs off those rows:
//   per-def substr ids (src/lib.rs:825-845)           -> run fills of the (zero-initialised) substr_ids column
//   is_start / is_end (src/lib.rs:847-888)            -> bits of the start_enable / end_enable bitmaps (:482-513) and the
//                                                        endpoint-lookup multiplicities (:235-284)
//   start_mask / end_mask scans (src/lib.rs:598-714)  -> evaluated in closed form, see `boundary` below
//   masked outputs (src/lib.rs:740-764)               -> fills of masked_chars / masked_substr_ids, substring records
// The hot loops only RECORD rare rows in a small per-lane queue (local memory); `drain` replays them in order.  In a
// uniform-length batch all 32 lanes of a warp drain at the same time, so the replay runs in lockstep instead of
// serialising the warp once per event.
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "defs.hpp"
#include "kernels.cuh"

namespace b2r {

constexpr uint32_t NO_POS = 0xFFFFFFFFu;
constexpr int QCAP = 8;                 // queued rare rows per lane before an early drain

// entry field accessors common to both encodings (defs.hpp): substr id in bits 16..23, flags in bits 24..26
__device__ __forceinline__ uint32_t ent_sid(uint32_t e) { return (e >> 16) & 0xFFu; }

template <int D>
struct Event {
    uint32_t pos, c;
    uint32_t e[D];      // entries of the row
    uint32_t s[D];      // states AT the row (before the transition)
    uint32_t nx[D];     // states after the transition
};

// Table policy: how finalize_segment re-walks a stretch (only for multi-run segments).
struct ClassTables {    // class-compressed tables: entry = trans[class[c]*S + s], next state in bits 0..15
    const uint8_t* cls;
    const uint32_t* trans;
    uint32_t S;
    __device__ __forceinline__ uint32_t lookup(uint32_t c, uint32_t s) const { return trans[(uint32_t)cls[c] * S + s]; }
    __device__ __forceinline__ static uint32_t next(uint32_t e) { return e & 0xFFFFu; }
};
struct DirectTables {   // direct [256][65] table in shared memory: entry at c*260 + s*4, next state in bits 8..15
    const unsigned char* tab;
    __device__ __forceinline__ uint32_t lookup(uint32_t c, uint32_t s) const { return *reinterpret_cast<const uint32_t*>(tab + c * 260 + s * 4); }
    __device__ __forceinline__ static uint32_t next(uint32_t e) { return (e >> 8) & 0xFFu; }
};

template <int D, typename TB>
struct Cold {
    uint64_t idx;           // string index
    const uint8_t* src;     // first byte of the string
    uint32_t len;
    uint32_t run_sid[D];    // current per-def substr id
    uint32_t run_start[D];  // first row of the current per-def substr-id run
    uint32_t seg_state[D];  // states at seg_start (to re-walk a multi-run segment)
    uint32_t sum_run;       // current id sum over defs
    uint32_t ie_pos, ie_val;// is_end sum that applies to boundary ie_pos
    uint32_t seg_sum;       // id sum at seg_start
    int32_t seg_start;      // first row of the pending (start-masked, not yet end-resolved) segment, -1 if none
    uint32_t seg_multi;     // the pending segment contains an unflagged id change
    uint32_t n_rec, n_cmp, flags;
    uint32_t bm_idx[2][D];  // word index of the bitmap word being accumulated (start_enable / end_enable), NO_POS = none
    uint32_t bm_val[2][D];
    uint32_t nq;
    Event<D> q[QCAP];
    TB tb[D];
    uint32_t* ep_s[D];      // shared-memory endpoint counters of def d: [0,K*S) start lookups, [K*S,2*K*S) end lookups; null = global

    __device__ __forceinline__ void init(uint64_t idx_, const uint8_t* src_, uint32_t len_) {
        idx = idx_; src = src_; len = len_;
        sum_run = 0; ie_pos = NO_POS; ie_val = 0; seg_sum = 0; seg_start = -1; seg_multi = 0;
        n_rec = 0; n_cmp = 0; flags = 0; nq = 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            run_sid[d] = 0; run_start[d] = 0; seg_state[d] = 0;
            bm_idx[0][d] = bm_idx[1][d] = NO_POS; bm_val[0][d] = bm_val[1][d] = 0;
        }
    }
};

__device__ __forceinline__ void fill_bytes(uint8_t* row, uint32_t a, uint32_t b, uint32_t v) {
    uint32_t i = a;
    while (i < b && (i & 15u)) row[i++] = (uint8_t)v;
    const uint32_t v4 = v * 0x01010101u;
    const uint4 vv = make_uint4(v4, v4, v4, v4);
    for (; i + 16 <= b; i += 16) *reinterpret_cast<uint4*>(row + i) = vv;   // rows are 16-byte aligned
    while (i < b) row[i++] = (uint8_t)v;
}

template <int D, typename TB>
__device__ __forceinline__ void emit_record(const WalkParams& p, Cold<D, TB>& k, uint32_t start, uint32_t len, uint32_t sid, uint32_t coff) {
    if (p.records && k.n_rec < p.max_records) {
        b2r_substr_record r; r.start = start; r.len = len; r.substr_id = sid; r.compact_off = coff;
        p.records[k.idx * p.max_records + k.n_rec] = r;
    }
    k.n_rec++;
}
template <int D, typename TB>
__device__ __forceinline__ void put_masked(const WalkParams& p, Cold<D, TB>& k, uint32_t i, uint32_t c) {
    if (p.masked_chars) p.masked_chars[k.idx * p.row_pitch + i] = (uint8_t)c;
    if (p.compact_bytes && k.n_cmp < p.compact_pitch) p.compact_bytes[k.idx * p.compact_pitch + k.n_cmp] = (uint8_t)c;
    k.n_cmp++;
}

// rows [a,b) are masked: start_mask = end_mask = 1 (src/lib.rs:740-764)
template <int D, typename TB>
__device__ __noinline__ void finalize_segment(const WalkParams& p, Cold<D, TB>& k, uint32_t a, uint32_t b) {
    if (!k.seg_multi) {
        emit_record(p, k, a, b - a, k.seg_sum, k.n_cmp);
        for (uint32_t i = a; i < b; i++) put_masked(p, k, i, k.src[i]);
        if (p.masked_substr_ids) fill_bytes(p.masked_substr_ids + k.idx * p.row_pitch, a, b, k.seg_sum);
    } else {  // the id sum changes inside the segment without a flag: re-walk it from the saved states
        uint32_t st[D];
#pragma unroll
        for (int d = 0; d < D; d++) st[d] = k.seg_state[d];
        uint32_t run_a = a, run_sum = 0, run_coff = k.n_cmp;
        for (uint32_t i = a; i < b; i++) {
            const uint32_t c = k.src[i];
            uint32_t sum = 0;
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t e = k.tb[d].lookup(c, st[d]);
                sum += ent_sid(e);
                st[d] = TB::next(e);
            }
            if (i == a) run_sum = sum;
            else if (sum != run_sum) { emit_record(p, k, run_a, i - run_a, run_sum, run_coff); run_a = i; run_sum = sum; run_coff = k.n_cmp; }
            put_masked(p, k, i, c);
            if (p.masked_substr_ids) p.masked_substr_ids[k.idx * p.row_pitch + i] = (uint8_t)sum;
        }
        emit_record(p, k, run_a, b - run_a, run_sum, run_coff);
    }
}

// Boundary `pos` where the id sum changes from k.sum_run to new_sum.
// Closed form of the two scans: the forward scan (src/lib.rs:613-642, idx = pos) sets start_mask when is_start_sum[pos]
// is set and resets it when only is_end_sum[pos] is; the backward scan (src/lib.rs:678-710, M-idx = pos) sets end_mask
// for the rows BEFORE pos when is_end_sum[pos] is set and resets it when only is_start_sum[pos] is.  Hence
// mask = start_mask & end_mask is 1 exactly on [b_k, b_{k+1}) for consecutive flagged boundaries b_k < b_{k+1} with
// is_start at b_k and is_end at b_{k+1}.  s[] = states at row pos.
template <int D, typename TB>
__device__ __forceinline__ void boundary(const WalkParams& p, Cold<D, TB>& k, uint32_t pos, uint32_t new_sum, uint32_t is_sum, uint32_t ie_sum, const uint32_t* s) {
    if (is_sum | ie_sum) {
        if (k.seg_start >= 0 && ie_sum) finalize_segment<D, TB>(p, k, (uint32_t)k.seg_start, pos);
        if (is_sum) {
            k.seg_start = (int32_t)pos; k.seg_sum = new_sum; k.seg_multi = 0;
#pragma unroll
            for (int d = 0; d < D; d++) k.seg_state[d] = s[d];
        } else k.seg_start = -1;
    } else if (k.seg_start >= 0) k.seg_multi = 1;
}

// the bitmaps are zero-initialised and a 32-bit word belongs to one row, so whole words are stored without a read
template <int D, typename TB>
__device__ __forceinline__ void bitmap_set(Cold<D, TB>& k, int which, int d, uint8_t* bitmap, uint64_t pitch, uint32_t pos) {
    if (!bitmap) return;
    const uint32_t wi = pos >> 5;
    if (k.bm_idx[which][d] != wi) {
        if (k.bm_idx[which][d] != NO_POS) reinterpret_cast<uint32_t*>(bitmap + k.idx * pitch)[k.bm_idx[which][d]] = k.bm_val[which][d];
        k.bm_idx[which][d] = wi; k.bm_val[which][d] = 0;
    }
    k.bm_val[which][d] |= 1u << (pos & 31);
}
template <int D, typename TB>
__device__ __forceinline__ void bitmap_flush(Cold<D, TB>& k, int which, int d, uint8_t* bitmap, uint64_t pitch) {
    if (bitmap && k.bm_idx[which][d] != NO_POS) reinterpret_cast<uint32_t*>(bitmap + k.idx * pitch)[k.bm_idx[which][d]] = k.bm_val[which][d];
    k.bm_idx[which][d] = NO_POS;
}

// one rare row, in order
template <int D, typename TB>
__device__ __forceinline__ void process_event(const WalkParams& p, Cold<D, TB>& k, const Event<D>& ev) {
    const uint32_t pos = ev.pos;
    uint32_t new_sum = 0, is_sum = 0, ie_next = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const uint32_t e = ev.e[d];
        const uint32_t sid = ent_sid(e);
        const uint32_t S = p.def[d].num_states;
        new_sum += sid;
        is_sum += (e >> 24) & 1u;
        ie_next += (e >> 25) & 1u;
        if (sid != k.run_sid[d]) {                   // per-def substr id run ends (src/lib.rs:825-845)
            if (k.run_sid[d] && p.def[d].substr_ids) fill_bytes(p.def[d].substr_ids + k.idx * p.row_pitch, k.run_start[d], pos, k.run_sid[d]);
            k.run_start[d] = pos; k.run_sid[d] = sid;
        }
        if (e & ENT_IS_START) {                      // start_enable, src/lib.rs:482-493; endpoint lookup :235-258
            bitmap_set(k, 0, d, p.def[d].start_enable, p.bitmap_pitch, pos);
            const uint32_t bin = (sid - p.def[d].sid_offset) * S + ev.s[d];
            if (k.ep_s[d]) atomicAdd(k.ep_s[d] + bin, 1u); else atomicAdd(p.def[d].ep_start + bin, 1ull);
        }
        if (e & ENT_IS_END) {                        // end_enable, src/lib.rs:501-513; endpoint lookup :260-284
            bitmap_set(k, 1, d, p.def[d].end_enable, p.bitmap_pitch, pos);
            const uint32_t bin = (sid - p.def[d].sid_offset) * S + ev.nx[d];
            if (k.ep_s[d]) atomicAdd(k.ep_s[d] + p.def[d].num_substrs * S + bin, 1u); else atomicAdd(p.def[d].ep_end + bin, 1ull);
        }
    }
    const uint32_t ie_here = (k.ie_pos == pos) ? k.ie_val : 0;
    if (is_sum > 1 || ie_here > 1) k.flags |= B2R_ST_OVERLAP;
    if (new_sum != k.sum_run) boundary<D, TB>(p, k, pos, new_sum, is_sum, ie_here, ev.s);
    k.sum_run = new_sum;
    k.ie_pos = pos + 1; k.ie_val = ie_next;
}

template <int D, typename TB>
__device__ __noinline__ void drain(const WalkParams& p, Cold<D, TB>& k) {
    const uint32_t n = k.nq;
    for (uint32_t i = 0; i < n; i++) process_event<D, TB>(p, k, k.q[i]);
    k.nq = 0;
}

// row `len`: the final-state row (src/lib.rs:404-418), last boundary, accept rule (src/lib.rs:427-457).  s[] = final states.
template <int D, typename TB>
__device__ __noinline__ void finish_string(const WalkParams& p, Cold<D, TB>& k, const uint32_t* s) {
    drain<D, TB>(p, k);
    const uint32_t L = k.len;
    const uint32_t ie_here = (k.ie_pos == L) ? k.ie_val : 0;
    if (ie_here > 1) k.flags |= B2R_ST_OVERLAP;
#pragma unroll
    for (int d = 0; d < D; d++) {
        if (k.run_sid[d] && p.def[d].substr_ids) fill_bytes(p.def[d].substr_ids + k.idx * p.row_pitch, k.run_start[d], L, k.run_sid[d]);
        if (s[d] == p.def[d].accepted_state) k.flags |= B2R_ST_ACCEPTED(d);
        bitmap_flush(k, 0, d, p.def[d].start_enable, p.bitmap_pitch);
        bitmap_flush(k, 1, d, p.def[d].end_enable, p.bitmap_pitch);
    }
    if (k.sum_run != 0) boundary<D, TB>(p, k, L, 0, 0, ie_here, s);
    if (p.status) {
        b2r_string_status st = {};
        st.flags = k.flags; st.err_pos = NO_POS;
        if (p.records && k.n_rec > p.max_records) st.flags |= B2R_ST_RECORDS_TRUNCATED;
        if (p.compact_bytes && k.n_cmp > p.compact_pitch) st.flags |= B2R_ST_COMPACT_TRUNCATED;
        st.n_records = k.n_rec; st.n_compact = k.n_cmp;
        p.status[k.idx] = st;
    }
}

// ---- CTA-level shared-memory counters (endpoint lookups, padded rows, overlaps): one global atomic per CTA and bin ----
struct CtaCounters {
    unsigned long long pad_rows, n_overlap, n_ok;
    unsigned long long reserved;
};

template <int D>
__device__ __forceinline__ uint32_t ep_smem_layout(const WalkParams& p, unsigned char* base, uint32_t* (&ep)[D]) {
    uint32_t off = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        ep[d] = p.ep_smem_bytes ? reinterpret_cast<uint32_t*>(base + off) : nullptr;
        off += 2u * p.def[d].num_substrs * p.def[d].num_states * 4u;
    }
    return off;
}

// called by every thread of the CTA before the tile loop (followed by __syncthreads) ...
template <int D>
__device__ __forceinline__ void cta_counters_init(const WalkParams& p, unsigned char* ep_base, CtaCounters* cc) {
    for (uint32_t i = threadIdx.x; i < p.ep_smem_bytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(ep_base)[i] = 0;
    if (threadIdx.x == 0) { cc->pad_rows = 0; cc->n_overlap = 0; cc->n_ok = 0; cc->reserved = 0; }
}
// ... once per tile by every lane of the warp ...
__device__ __forceinline__ void cta_counters_tile(CtaCounters* cc, bool ok, uint32_t pad, bool overlap) {
    const uint32_t pad_sum = __reduce_add_sync(0xffffffffu, ok ? pad : 0u);
    const uint32_t ov = __popc(__ballot_sync(0xffffffffu, ok && overlap));
    const uint32_t okc = __popc(__ballot_sync(0xffffffffu, ok));
    if ((threadIdx.x & 31) == 0) {
        if (pad_sum) atomicAdd(&cc->pad_rows, (unsigned long long)pad_sum);
        if (ov) atomicAdd(&cc->n_overlap, (unsigned long long)ov);
        if (okc) atomicAdd(&cc->n_ok, (unsigned long long)okc);
    }
}
// ... and after the tile loop (after a __syncthreads) to publish
template <int D>
__device__ __forceinline__ void cta_counters_flush(const WalkParams& p, uint32_t* const (&ep)[D], const CtaCounters* cc) {
#pragma unroll
    for (int d = 0; d < D; d++) {
        if (!ep[d]) continue;
        const uint32_t ks = p.def[d].num_substrs * p.def[d].num_states;
        for (uint32_t i = threadIdx.x; i < 2 * ks; i += blockDim.x) {
            const uint32_t v = ep[d][i];
            if (v) atomicAdd((i < ks ? p.def[d].ep_start + i : p.def[d].ep_end + (i - ks)), (unsigned long long)v);
        }
    }
    if (threadIdx.x == 0) {
        if (cc->pad_rows) atomicAdd(&p.counters->pad_rows, cc->pad_rows);
        if (cc->n_overlap) atomicAdd(&p.counters->n_overlap, cc->n_overlap);
        if (cc->n_ok) atomicAdd(&p.counters->n_ok_strings, cc->n_ok);
    }
}

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

// the reference panics (src/lib.rs:817): mark the string, remember the lowest failing index of the batch
template <int D, typename TB>
__device__ __noinline__ void kill_string(const WalkParams& p, Cold<D, TB>& k) {
    atomicMin(&p.counters->first_bad, (unsigned long long)k.idx);
    if (p.status) {
        const b2r_batch_status r = diagnose_string(p, k.idx);
        b2r_string_status st = {};
        st.flags = (r.code == B2R_ERR_TOO_LONG) ? B2R_ST_TOO_LONG : B2R_ST_INVALID_TRANSITION;
        st.err_pos = (r.code == B2R_ERR_TOO_LONG) ? NO_POS : r.pos; st.err_state = r.state; st.err_byte = r.byte; st.err_def = r.def;
        p.status[k.idx] = st;
    }
}

}  // namespace b2r
