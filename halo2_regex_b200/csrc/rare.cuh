// Lane-private "rare row" machinery shared by the walk kernels (sm_100a).
//
// A row is RARE when one of its packed entries changes the per-def substr id, carries is_start / is_end, or is an
// invalid transition.  Everything the reference derives beyond the state column hangs off those rows:
//   per-def substr ids (src/lib.rs:825-845)           -> run fills of the (zero-initialised) substr_ids column
//   is_start / is_end (src/lib.rs:847-888)            -> bits of the start_enable / end_enable bitmaps (:482-513) and the
//                                                        endpoint-lookup multiplicities (:235-284)
//   start_mask / end_mask scans (src/lib.rs:598-714)  -> evaluated in closed form, see `boundary` below
//   masked outputs (src/lib.rs:740-764)               -> fills of masked_chars / masked_substr_ids, substring records
// The state carried from one rare row to the next is a handful of words per lane ("cold" state).  The direct kernel keeps
// it in shared memory as a struct of arrays (stride 32 words: conflict-free, and no local-memory traffic — with ~200 KB of
// the SM configured as shared memory the L1 is too small to hold 512 threads' stacks); the generic kernel keeps it in a
// small local array (stride 1).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "defs.hpp"
#include "kernels.cuh"

namespace b2r {

constexpr uint32_t NO_POS = 0xFFFFFFFFu;

// entry field accessors common to both encodings (defs.hpp): substr id in bits 16..23, flags in bits 24..26
__device__ __forceinline__ uint32_t ent_sid(uint32_t e) { return (e >> 16) & 0xFFu; }

__device__ __forceinline__ uint32_t lds8(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u8 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}

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

// ---- cold state: field f of this lane lives at base[f * STRIDE] ---------------------------------------------------------
enum ColdField : int {
    CF_SUM_RUN = 0,   // current id sum over defs
    CF_IE,            // is_end sum for boundary `pos`, packed as pos << 3 | sum (NO_POS = none)
    CF_SEG_SUM,       // id sum at seg_start
    CF_SEG_START,     // first row of the pending (start-masked, not yet end-resolved) segment << 1 | multi-run flag; NO_POS = none
    CF_NREC,
    CF_NCMP,
    CF_FLAGS,
    CF_NQ,            // queued rare rows
    CF_PER_DEF        // then per def: run_start[d], seg_state[d], run_sid[d]
};
__host__ __device__ constexpr int cold_fields(int D) { return CF_PER_DEF + 3 * D; }

// Queue of rare rows awaiting their (heavy) processing.  It lives in a per-lane slice of a global scratch buffer that
// stays L2-resident: pushes are fire-and-forget stores, the replay at the end of the string reads it back in lockstep.
// Word w of this lane's slice is at qbase[w * 32] (the 32 lanes of a warp share 128-byte lines).
constexpr int QCAP = 8;
__host__ __device__ constexpr int queue_words(int D) { return QCAP * (1 + 2 * D); }   // per event: pos, then (entry, state) per def

template <int D, int STRIDE>
struct Cold {
    uint32_t* base;
    __device__ __forceinline__ uint32_t& f(int field) const { return base[field * STRIDE]; }
    __device__ __forceinline__ uint32_t& run_start(int d) const { return base[(CF_PER_DEF + 3 * d) * STRIDE]; }
    __device__ __forceinline__ uint32_t& seg_state(int d) const { return base[(CF_PER_DEF + 3 * d + 1) * STRIDE]; }
    __device__ __forceinline__ uint32_t& run_sid(int d) const { return base[(CF_PER_DEF + 3 * d + 2) * STRIDE]; }
    __device__ __forceinline__ void init() const {
        f(CF_SUM_RUN) = 0; f(CF_IE) = NO_POS; f(CF_SEG_SUM) = 0; f(CF_SEG_START) = NO_POS; f(CF_NREC) = 0; f(CF_NCMP) = 0; f(CF_FLAGS) = 0;
        f(CF_NQ) = 0;
#pragma unroll
        for (int d = 0; d < D; d++) { run_start(d) = 0; seg_state(d) = 0; run_sid(d) = 0; }
    }
};

// per-string context the rare path needs (built from registers at the call site)
template <int D, typename TB>
struct RowCtx {
    uint64_t idx;           // string index
    const uint8_t* src;     // first byte of the string (global)
    uint32_t len;
    uint32_t tile_pos;      // rows >= tile_pos are also available in shared memory at tile_s + (row - tile_pos); NO_POS = none
    uint32_t tile_s;
    TB tb[D];
    uint32_t* ep_s[D];      // shared-memory endpoint counters of def d: [0,K*S) start lookups, [K*S,2*K*S) end lookups; null = global
    uint32_t* qbase;        // this lane's slice of the global event queue (word w at qbase[w * 32])
    __device__ __forceinline__ uint32_t char_at(uint32_t i) const {
        return (i >= tile_pos) ? lds8(tile_s + (i - tile_pos)) : (uint32_t)src[i];
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

template <int D, int ST, typename TB>
__device__ __forceinline__ void emit_record(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x, uint32_t start, uint32_t len, uint32_t sid, uint32_t coff) {
    const uint32_t n = k.f(CF_NREC);
    if (p.records && n < p.max_records) {
        b2r_substr_record r; r.start = start; r.len = len; r.substr_id = sid; r.compact_off = coff;
        p.records[x.idx * p.max_records + n] = r;
    }
    k.f(CF_NREC) = n + 1;
}

// rows [a,b) are masked: start_mask = end_mask = 1 (src/lib.rs:740-764)
template <int D, int ST, typename TB>
__device__ __noinline__ void finalize_segment(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x, uint32_t a, uint32_t b, uint32_t multi) {
    uint32_t n_cmp = k.f(CF_NCMP);
    uint8_t* mc = p.masked_chars ? p.masked_chars + x.idx * p.row_pitch : nullptr;
    uint8_t* cb = p.compact_bytes ? p.compact_bytes + x.idx * p.compact_pitch : nullptr;
    if (!multi) {
        const uint32_t sum = k.f(CF_SEG_SUM);
        emit_record(p, k, x, a, b - a, sum, n_cmp);
        for (uint32_t i = a; i < b; i++) {
            const uint32_t c = x.char_at(i);
            if (mc) mc[i] = (uint8_t)c;
            if (cb && n_cmp < p.compact_pitch) cb[n_cmp] = (uint8_t)c;
            n_cmp++;
        }
        if (p.masked_substr_ids) fill_bytes(p.masked_substr_ids + x.idx * p.row_pitch, a, b, sum);
    } else {  // the id sum changes inside the segment without a flag: re-walk it from the saved states
        uint32_t st[D];
#pragma unroll
        for (int d = 0; d < D; d++) st[d] = k.seg_state(d);
        uint32_t run_a = a, run_sum = 0, run_coff = n_cmp;
        for (uint32_t i = a; i < b; i++) {
            const uint32_t c = x.char_at(i);
            uint32_t sum = 0;
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t e = x.tb[d].lookup(c, st[d]);
                sum += ent_sid(e);
                st[d] = TB::next(e);
            }
            if (i == a) run_sum = sum;
            else if (sum != run_sum) { emit_record(p, k, x, run_a, i - run_a, run_sum, run_coff); run_a = i; run_sum = sum; run_coff = n_cmp; }
            if (mc) mc[i] = (uint8_t)c;
            if (cb && n_cmp < p.compact_pitch) cb[n_cmp] = (uint8_t)c;
            n_cmp++;
            if (p.masked_substr_ids) p.masked_substr_ids[x.idx * p.row_pitch + i] = (uint8_t)sum;
        }
        emit_record(p, k, x, run_a, b - run_a, run_sum, run_coff);
    }
    k.f(CF_NCMP) = n_cmp;
}

// Boundary `pos` where the id sum changes to new_sum.
// Closed form of the two scans: the forward scan (src/lib.rs:613-642, idx = pos) sets start_mask when is_start_sum[pos]
// is set and resets it when only is_end_sum[pos] is; the backward scan (src/lib.rs:678-710, M-idx = pos) sets end_mask
// for the rows BEFORE pos when is_end_sum[pos] is set and resets it when only is_start_sum[pos] is.  Hence
// mask = start_mask & end_mask is 1 exactly on [b_k, b_{k+1}) for consecutive flagged boundaries b_k < b_{k+1} with
// is_start at b_k and is_end at b_{k+1}.  s[] = states at row pos.
template <int D, int ST, typename TB>
__device__ __forceinline__ void boundary(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x, uint32_t pos, uint32_t new_sum, uint32_t is_sum,
                                         uint32_t ie_sum, const uint32_t* s) {
    const uint32_t seg = k.f(CF_SEG_START);
    if (is_sum | ie_sum) {
        if (seg != NO_POS && ie_sum) finalize_segment<D, ST, TB>(p, k, x, seg >> 1, pos, seg & 1u);
        if (is_sum) {
            k.f(CF_SEG_START) = pos << 1; k.f(CF_SEG_SUM) = new_sum;
#pragma unroll
            for (int d = 0; d < D; d++) k.seg_state(d) = s[d];
        } else k.f(CF_SEG_START) = NO_POS;
    } else if (seg != NO_POS) k.f(CF_SEG_START) = seg | 1u;
}

// the bitmaps are zero-initialised before any rare row is processed: set single bits with a fire-and-forget atomic OR
__device__ __forceinline__ void bitmap_set(uint8_t* bitmap_row, uint32_t pos) {
    atomicOr(reinterpret_cast<unsigned int*>(bitmap_row) + (pos >> 5), 1u << (pos & 31));
}

// One rare row, in order.  e[] = entries of the row, s[] = states at the row, nx[] = states after it.
template <int D, int ST, typename TB>
__device__ __forceinline__ void process_row(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x, uint32_t pos, const uint32_t* e, const uint32_t* s,
                                            const uint32_t* nx) {
    uint32_t new_sum = 0, is_sum = 0, ie_next = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const uint32_t sid = ent_sid(e[d]);
        const uint32_t S = p.def[d].num_states;
        new_sum += sid;
        is_sum += (e[d] >> 24) & 1u;
        ie_next += (e[d] >> 25) & 1u;
        const uint32_t run = k.run_sid(d);
        if (sid != run) {                            // per-def substr id run ends (src/lib.rs:825-845)
            if (run && p.def[d].substr_ids) fill_bytes(p.def[d].substr_ids + x.idx * p.row_pitch, k.run_start(d), pos, run);
            k.run_start(d) = pos; k.run_sid(d) = sid;
        }
        if (e[d] & ENT_IS_START) {                   // start_enable, src/lib.rs:482-493; endpoint lookup :235-258
            if (p.def[d].start_enable) bitmap_set(p.def[d].start_enable + x.idx * p.bitmap_pitch, pos);
            const uint32_t bin = (sid - p.def[d].sid_offset) * S + s[d];
            if (x.ep_s[d]) atomicAdd(x.ep_s[d] + bin, 1u); else atomicAdd(p.def[d].ep_start + bin, 1ull);
        }
        if (e[d] & ENT_IS_END) {                     // end_enable, src/lib.rs:501-513; endpoint lookup :260-284
            if (p.def[d].end_enable) bitmap_set(p.def[d].end_enable + x.idx * p.bitmap_pitch, pos);
            const uint32_t bin = (sid - p.def[d].sid_offset) * S + nx[d];
            if (x.ep_s[d]) atomicAdd(x.ep_s[d] + p.def[d].num_substrs * S + bin, 1u); else atomicAdd(p.def[d].ep_end + bin, 1ull);
        }
    }
    const uint32_t ie = k.f(CF_IE);
    const uint32_t ie_here = (ie != NO_POS && (ie >> 3) == pos) ? (ie & 7u) : 0u;
    if (is_sum > 1 || ie_here > 1) k.f(CF_FLAGS) |= B2R_ST_OVERLAP;
    if (new_sum != k.f(CF_SUM_RUN)) boundary<D, ST, TB>(p, k, x, pos, new_sum, is_sum, ie_here, s);
    k.f(CF_SUM_RUN) = new_sum;
    k.f(CF_IE) = ((pos + 1) << 3) | ie_next;
}

// queue a rare row: e[] = entries, s[] = states at the row
template <int D, int ST, typename TB>
__device__ __forceinline__ void drain(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x);

template <int D, int ST, typename TB>
__device__ __forceinline__ void push_row(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x, uint32_t pos, const uint32_t* e, const uint32_t* s) {
    uint32_t n = k.f(CF_NQ);
    if (n == QCAP) { drain<D, ST, TB>(p, k, x); n = 0; }    // rare: more than QCAP rare rows before the end of the string
    uint32_t* q = x.qbase + n * (1 + 2 * D) * 32;
    __stcg(q, pos);
#pragma unroll
    for (int d = 0; d < D; d++) { __stcg(q + (1 + 2 * d) * 32, e[d]); __stcg(q + (2 + 2 * d) * 32, s[d]); }
    k.f(CF_NQ) = n + 1;
}

// replay the queued rows in order (the heavy part: fills, bitmap bits, endpoint counters, mask segments)
template <int D, int ST, typename TB>
__device__ __forceinline__ void drain(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x) {
    const uint32_t n = k.f(CF_NQ);
    for (uint32_t i = 0; i < n; i++) {
        const uint32_t* q = x.qbase + i * (1 + 2 * D) * 32;
        const uint32_t pos = __ldcg(q);
        uint32_t e[D], s[D], nx[D];
#pragma unroll
        for (int d = 0; d < D; d++) { e[d] = __ldcg(q + (1 + 2 * d) * 32); s[d] = __ldcg(q + (2 + 2 * d) * 32); nx[d] = TB::next(e[d]); }
        process_row<D, ST, TB>(p, k, x, pos, e, s, nx);
    }
    k.f(CF_NQ) = 0;
}

// row `len`: the final-state row (src/lib.rs:404-418), last boundary, accept rule (src/lib.rs:427-457).  s[] = final states.
template <int D, int ST, typename TB>
__device__ __noinline__ void finish_string(const WalkParams& p, const Cold<D, ST>& k, const RowCtx<D, TB>& x, const uint32_t* s) {
    drain<D, ST, TB>(p, k, x);
    const uint32_t L = x.len;
    const uint32_t ie = k.f(CF_IE);
    const uint32_t ie_here = (ie != NO_POS && (ie >> 3) == L) ? (ie & 7u) : 0u;
    uint32_t flags = k.f(CF_FLAGS);
    if (ie_here > 1) flags |= B2R_ST_OVERLAP;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const uint32_t run = k.run_sid(d);
        if (run && p.def[d].substr_ids) fill_bytes(p.def[d].substr_ids + x.idx * p.row_pitch, k.run_start(d), L, run);
        if (s[d] == p.def[d].accepted_state) flags |= B2R_ST_ACCEPTED(d);
    }
    k.f(CF_FLAGS) = flags;
    if (k.f(CF_SUM_RUN) != 0) boundary<D, ST, TB>(p, k, x, L, 0, 0, ie_here, s);
    if (p.status) {
        b2r_string_status st = {};
        st.flags = flags; st.err_pos = NO_POS;
        const uint32_t n_rec = k.f(CF_NREC), n_cmp = k.f(CF_NCMP);
        if (p.records && n_rec > p.max_records) st.flags |= B2R_ST_RECORDS_TRUNCATED;
        if (p.compact_bytes && n_cmp > p.compact_pitch) st.flags |= B2R_ST_COMPACT_TRUNCATED;
        st.n_records = n_rec; st.n_compact = n_cmp;
        p.status[x.idx] = st;
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

}  // namespace b2r
