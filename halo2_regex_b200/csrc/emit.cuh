// emit_kernel — stage 2 of the batch path: everything the reference derives from the state sequence.
//
// One WARP per tile of 32 consecutive strings.  Inputs: the strings, their state columns and granule flags (walk.cuh).
// Outputs (every row is written within microseconds by one warp, so zeros and values merge in L2 and DRAM sees the
// algorithmic bytes once):
//   per-def substr ids (src/lib.rs:825-845), start_enable / end_enable bitmaps (:482-513), the endpoint-lookup
//   multiplicities (:235-284), masked_chars / masked_substr_ids (:740-764), substring records + compact bytes, the
//   accept flag (:427-457) and the status record.
//
// Phase A (lane = string): offsets, granule flags, final states of the tile's 32 strings (coalesced).
// Phase B: the tile's rows of every sparse column are contiguous in memory: blanket zero-fill with 16-byte stores.
// Phase C (lane = string): every lane scans the flagged granules of ITS string, in order.  A row's packed entries are
//         looked up again from (byte, state) — one lookup replaces the reference's HashSet probes and `contains` scans —
//         and give substr id, is_start, is_end(next row).  Non-zero values overwrite the zeros of phase B in L2.
// Phase D (lane = string): accept flags and status records.
// Masks.  With b_1 < b_2 < ... the rows where the id sum changes AND is_start_sum|is_end_sum is set, the forward scan
//         (src/lib.rs:598-645) sets start_mask at b_k when is_start_sum[b_k] and resets it when only is_end_sum[b_k]; the
//         backward scan (:663-714) sets end_mask for the rows before b_{k+1} when is_end_sum[b_{k+1}] and resets it when
//         only is_start_sum[b_{k+1}].  Hence mask = 1 exactly on [b_k, b_{k+1}) with is_start at b_k and is_end at b_{k+1}.
//         The boundaries are streamed in order; a closing boundary makes the lane overwrite the masked rows of
//         [b_k, b_{k+1}) (they may extend over unflagged granules) — same L2 lines the warp has just zeroed.
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
// masked_substr_ids, the compact bytes and one record per maximal run of a constant id sum.  Lane-private (one thread,
// its own string).  Returns the updated (n_rec, n_cmp).
template <int D, typename ST>
static __device__ __noinline__ uint2 emit_segment(const WalkParams& p, const EmitTables<D> tb, uint64_t j, const uint8_t* src, uint32_t a, uint32_t b,
                                                  uint32_t n_rec, uint32_t n_cmp) {
    uint8_t* const mc = p.masked_chars ? p.masked_chars + j * p.row_pitch : nullptr;
    uint8_t* const ms = p.masked_substr_ids ? p.masked_substr_ids + j * p.row_pitch : nullptr;
    uint8_t* const cb = p.compact_bytes ? p.compact_bytes + j * (uint64_t)p.compact_pitch : nullptr;
    auto record = [&](uint32_t start, uint32_t len, uint32_t sid, uint32_t coff) {
        if (p.records && n_rec < p.max_records) {
            b2r_substr_record r; r.start = start; r.len = len; r.substr_id = sid; r.compact_off = coff;
            p.records[j * p.max_records + n_rec] = r;
        }
        n_rec++;
    };
    uint32_t run_start = a, run_sum = 0;
    for (uint32_t i = a; i < b; i++) {
        const uint32_t c = src[i];
        uint32_t sum = 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t S = p.def[d].num_states;
            const uint32_t s = (uint32_t) reinterpret_cast<const ST*>(p.def[d].states)[j * p.row_pitch + i];
            if (s < S) sum += ent_sid(tb.trans[d][(uint32_t)tb.cls[d][c] * S + s]);
        }
        if (mc && !(p.debug & 16)) mc[i] = (uint8_t)c;
        if (ms && sum && !(p.debug & 16)) ms[i] = (uint8_t)sum;
        const uint32_t k = n_cmp + (i - a);
        if (cb && k < p.compact_pitch) cb[k] = (uint8_t)c;
        if (i == a) run_sum = sum;
        else if (sum != run_sum) {                                       // records: maximal runs of a constant id sum
            record(run_start, i - run_start, run_sum, n_cmp + (run_start - a));
            run_start = i; run_sum = sum;
        }
    }
    record(run_start, b - run_start, run_sum, n_cmp + (run_start - a));
    return make_uint2(n_rec, n_cmp + (b - a));
}

// Lane-private scan of ONE string's flagged granules, in order (phase C).  Streams the boundaries of the mask algebra.
template <int D, typename ST>
struct LaneScan {
    const WalkParams& p;
    const EmitTables<D>& tb;
    uint64_t j;
    const uint8_t* src;
    uint32_t L;
    bool prev_valid, prev_is;   // last boundary: is_start_sum set there
    uint32_t prev_pos;
    uint32_t next_row;          // last examined row + 1; NO_POS = none yet / gap closed
    uint32_t carry_s, carry_ie; // id sum of row next_row-1, is_end_sum[next_row] (0 after a gap)
    uint32_t n_rec, n_cmp;
    bool overlap, invalid;
    uint32_t bw_t, sw[D], ew[D];   // bitmap words being assembled: rows [32*bw_t, 32*bw_t+32)

    __device__ __forceinline__ LaneScan(const WalkParams& p_, const EmitTables<D>& tb_, uint64_t j_, const uint8_t* src_, uint32_t L_)
        : p(p_), tb(tb_), j(j_), src(src_), L(L_) {
        prev_valid = false; prev_is = false; prev_pos = 0;
        next_row = NO_POS; carry_s = 0; carry_ie = 0;
        n_rec = 0; n_cmp = 0; overlap = false; invalid = false;
        bw_t = NO_POS;
#pragma unroll
        for (int d = 0; d < D; d++) { sw[d] = 0; ew[d] = 0; }
    }

    // a row where the id sum changes and is_start_sum | is_end_sum is set
    __device__ __forceinline__ void boundary(uint32_t pos, bool b_is, bool b_ie) {
        if (prev_valid && prev_is && b_ie) {
            const uint2 r = emit_segment<D, ST>(p, tb, j, src, prev_pos, pos, n_rec, n_cmp);
            n_rec = r.x; n_cmp = r.y;
        }
        prev_valid = true; prev_pos = pos; prev_is = b_is;
    }
    // the rows from next_row on are not examined (their granule is not flagged: id sum 0, no is_start) or lie past the end
    // of the string; row next_row can still be a boundary: the id sum drops to 0 there and is_end_sum[next_row] comes from
    // the row before
    __device__ __forceinline__ void close_gap() {
        if (next_row != NO_POS && carry_s != 0 && carry_ie != 0) {
            if (carry_ie > 1) overlap = true;
            boundary(next_row, false, true);
        }
        next_row = NO_POS; carry_s = 0; carry_ie = 0;
    }
    __device__ __forceinline__ void flush_bitmap_words() {
        if (bw_t == NO_POS || (p.debug & 16)) return;
#pragma unroll
        for (int d = 0; d < D; d++) {
            if (sw[d] && p.def[d].start_enable) *reinterpret_cast<uint32_t*>(p.def[d].start_enable + j * p.bitmap_pitch + 4 * bw_t) = sw[d];
            if (ew[d] && p.def[d].end_enable) *reinterpret_cast<uint32_t*>(p.def[d].end_enable + j * p.bitmap_pitch + 4 * bw_t) = ew[d];
            sw[d] = 0; ew[d] = 0;
        }
    }
    // row i < L with byte c and states s[].  Rows with an id sum of 0 in every def may be skipped by the caller: they are
    // treated like the rows of an unflagged granule.
    __device__ __forceinline__ void row(uint32_t i, uint32_t c, const uint32_t* s) {
        if (next_row != i) close_gap();
        uint32_t sum = 0, is_sum = 0, ie_next = 0;
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t S = p.def[d].num_states;
            const uint32_t e = s[d] < S ? tb.trans[d][(uint32_t)tb.cls[d][c] * S + s[d]] : ENT_INVALID;   // trap state: the walk parked this def
            if (e & ENT_INVALID) { invalid = true; continue; }
            const uint32_t sid = ent_sid(e);
            if (!sid) continue;
            sum += sid;
            if (p.def[d].substr_ids && !(p.debug & 16)) p.def[d].substr_ids[j * p.row_pitch + i] = (uint8_t)sid;
            if (e & ENT_IS_START) {                                      // start_enable, endpoint lookup src/lib.rs:235-258
                is_sum++;
                sw[d] |= 1u << (i & 31);
                const uint32_t bin = (sid - p.def[d].sid_offset) * S + s[d];
                if (tb.ep_s[d]) atomicAdd(tb.ep_s[d] + bin, 1u); else atomicAdd(p.def[d].ep_start + bin, 1ull);
            }
            if (e & ENT_IS_END) {                                        // end_enable, endpoint lookup src/lib.rs:260-284
                ie_next++;
                ew[d] |= 1u << (i & 31);
                const uint32_t bin = (sid - p.def[d].sid_offset) * S + (e & ENT_NEXT_MASK);
                if (tb.ep_s[d]) atomicAdd(tb.ep_s[d] + p.def[d].num_substrs * S + bin, 1u); else atomicAdd(p.def[d].ep_end + bin, 1ull);
            }
        }
        if (is_sum > 1 || carry_ie > 1) overlap = true;
        if (sum != carry_s && (is_sum | carry_ie)) boundary(i, is_sum != 0, carry_ie != 0);
        carry_s = sum; carry_ie = ie_next; next_row = i + 1;
    }
    // granule g: rows [16g, 16g+16)
    __device__ __forceinline__ void granule(uint32_t g) {
        const uint32_t base = 16 * g;
        if ((g >> 1) != bw_t) { flush_bitmap_words(); bw_t = g >> 1; }
        // 16 bytes of the string from an arbitrary address: two aligned 16-byte loads, shifted into place.  The second
        // one is only touched when the bytes needed reach into it.
        const uintptr_t addr = reinterpret_cast<uintptr_t>(src + base);
        const uint32_t sh = (uint32_t)(addr & 15);
        const uint32_t n = L - base < 16 ? L - base : 16;              // rows of this granule that are characters (>= 1)
        const uint4* q = reinterpret_cast<const uint4*>(addr - sh);
        const uint4 v0 = __ldg(q);
        uint4 v1 = make_uint4(0, 0, 0, 0);
        if (sh + n > 16) v1 = __ldg(q + 1);
        uint32_t x[8] = {v0.x, v0.y, v0.z, v0.w, v1.x, v1.y, v1.z, v1.w};
        uint32_t y[7], z[5], w[4];
#pragma unroll
        for (int k = 0; k < 7; k++) y[k] = (sh & 4) ? x[k + 1] : x[k];
#pragma unroll
        for (int k = 0; k < 5; k++) z[k] = (sh & 8) ? y[k + 2] : y[k];
#pragma unroll
        for (int k = 0; k < 4; k++) w[k] = __funnelshift_r(z[k], z[k + 1], (sh & 3) * 8);
        uint32_t sv[D][4 * sizeof(ST)];
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint4* sp = reinterpret_cast<const uint4*>(reinterpret_cast<const ST*>(p.def[d].states) + j * p.row_pitch + base);
            const uint4 a = __ldg(sp);
            sv[d][0] = a.x; sv[d][1] = a.y; sv[d][2] = a.z; sv[d][3] = a.w;
            if (sizeof(ST) == 2) { const uint4 b2 = __ldg(sp + 1); sv[d][4] = b2.x; sv[d][5] = b2.y; sv[d][6] = b2.z; sv[d][7] = b2.w; }
        }
        // pass 1, unrolled, all lanes in step: which rows carry a substr id (or an invalid transition)?
        uint32_t hot = 0;
#pragma unroll
        for (int r = 0; r < 16; r++) {
            const uint32_t c = (w[r >> 2] >> (8 * (r & 3))) & 0xFFu;
#pragma unroll
            for (int d = 0; d < D; d++) {
                const uint32_t S = p.def[d].num_states;
                const uint32_t st = sizeof(ST) == 1 ? (sv[d][r >> 2] >> (8 * (r & 3))) & 0xFFu : (sv[d][r >> 1] >> (16 * (r & 1))) & 0xFFFFu;
                const uint32_t e = st < S ? tb.trans[d][(uint32_t)tb.cls[d][c] * S + st] : ENT_INVALID;
                if (e & (ENT_SID_MASK | ENT_INVALID)) hot |= 1u << r;
            }
        }
        hot &= (1u << n) - 1u;                                           // rows past the end of the string are not characters
        // pass 2, one loop body shared by all lanes: each lane handles ITS next hot row
        while (hot) {
            const uint32_t r = (uint32_t)__ffs((int)hot) - 1u;
            hot &= hot - 1;
            const uint32_t wq = r < 8 ? (r < 4 ? w[0] : w[1]) : (r < 12 ? w[2] : w[3]);
            const uint32_t c = (wq >> (8 * (r & 3))) & 0xFFu;
            uint32_t st[D];
#pragma unroll
            for (int d = 0; d < D; d++) {
                if (sizeof(ST) == 1) {
                    const uint32_t q = r < 8 ? (r < 4 ? sv[d][0] : sv[d][1]) : (r < 12 ? sv[d][2] : sv[d][3]);
                    st[d] = (q >> (8 * (r & 3))) & 0xFFu;
                } else {
                    const uint32_t lo = r < 4 ? (r < 2 ? sv[d][0] : sv[d][1]) : (r < 6 ? sv[d][2] : sv[d][3]);
                    const uint32_t hi = r < 12 ? (r < 10 ? sv[d][4 * (sizeof(ST) - 1)] : sv[d][4 * (sizeof(ST) - 1) + 1]) : (r < 14 ? sv[d][4 * (sizeof(ST) - 1) + 2] : sv[d][4 * (sizeof(ST) - 1) + 3]);
                    st[d] = ((r < 8 ? lo : hi) >> (16 * (r & 1))) & 0xFFFFu;
                }
            }
            row(base + r, c, st);
        }
    }
    __device__ __forceinline__ void finish() { close_gap(); flush_bitmap_words(); }
};

// Warp-cooperative emitter for one tile of 32 consecutive strings.
template <int D, typename ST>
struct TileEmitter {
    const WalkParams& p;
    const EmitTables<D>& tb;
    const int lane;

    __device__ __forceinline__ TileEmitter(const WalkParams& p_, const EmitTables<D>& tb_, int lane_) : p(p_), tb(tb_), lane(lane_) {}

    // One tile.  Phase A (lane = string): offsets, granule flags, final states.  Phase B: the tile's rows of every sparse
    // column are contiguous — blanket zero-fill with 16-byte stores.  Phase C (lane = string): every lane scans the flagged
    // granules of its own string.  Phase D (lane = string): accept rule (src/lib.rs:427-457) and the status records.
    __device__ __forceinline__ void run_tile(uint64_t tile, EmitTotals& tot) {
        const uint64_t N = p.n_strings;
        const uint32_t M = p.max_chars;
        const uint64_t rp = p.row_pitch, bp = p.bitmap_pitch;
        const uint64_t tile_base = tile * 32;
        const uint64_t jl = tile_base + lane;
        const bool valid = jl < N;
        const uint64_t rows_here = N - tile_base < 32 ? N - tile_base : 32;

        // ---- A ----------------------------------------------------------------------------------------------------
        uint64_t off = 0, end = 0;
        if (valid) { off = p.offsets[jl]; end = p.offsets[jl + 1]; }
        const bool too_long = valid && (end < off || end - off > (uint64_t)(M - 1));   // len must be <= M-1
        const bool live = valid && !too_long;
        const uint32_t Ll = live ? (uint32_t)(end - off) : 0u;
        uint32_t fw0 = 0, fw1 = 0;
        if (live && p.fm_words > 0) fw0 = __ldg(p.fmask + jl);
        if (live && p.fm_words > 1) fw1 = __ldg(p.fmask + N + jl);
        uint32_t fin[D];
#pragma unroll
        for (int d = 0; d < D; d++) fin[d] = (live && !(p.debug & 4)) ? (uint32_t) reinterpret_cast<const ST*>(p.def[d].states)[jl * rp + Ll] : 0xFFFFFFFEu;

        // ---- B ----------------------------------------------------------------------------------------------------
        auto zero_region = [&](uint8_t* base, uint64_t bytes) {   // base 16-byte aligned, bytes a multiple of 4
            if (!base || (p.debug & 1)) return;
            const uint64_t nv = bytes / 16;
            const uint4 z = make_uint4(0, 0, 0, 0);
            for (uint64_t v = lane; v < nv; v += 32) reinterpret_cast<uint4*>(base)[v] = z;
            for (uint64_t o = nv * 16 + (uint64_t)lane * 4; o < bytes; o += 128) *reinterpret_cast<uint32_t*>(base + o) = 0u;
        };
        auto phase_b = [&]() {
#pragma unroll
        for (int d = 0; d < D; d++) {
            zero_region(p.def[d].substr_ids ? p.def[d].substr_ids + tile_base * rp : nullptr, rows_here * rp);
            zero_region(p.def[d].start_enable ? p.def[d].start_enable + tile_base * bp : nullptr, rows_here * bp);
            zero_region(p.def[d].end_enable ? p.def[d].end_enable + tile_base * bp : nullptr, rows_here * bp);
        }
        zero_region(p.masked_chars ? p.masked_chars + tile_base * rp : nullptr, rows_here * rp);
        zero_region(p.masked_substr_ids ? p.masked_substr_ids + tile_base * rp : nullptr, rows_here * rp);
            __syncwarp();                                                // the zeros are ordered before the values of phase C
        };
        if (!(p.debug & 32)) phase_b();

        // ---- C ----------------------------------------------------------------------------------------------------
        uint32_t r_nrec = 0, r_ncmp = 0, r_flags = 0;
        if (live && !(p.debug & 2) && (p.fm_words > 2 || (fw0 | fw1) != 0)) {
            LaneScan<D, ST> sc(p, tb, jl, p.bytes + off, Ll);
            for (uint32_t w = 0; w < p.fm_words && !sc.invalid; w++) {
                uint32_t bits = w == 0 ? fw0 : w == 1 ? fw1 : __ldg(p.fmask + (size_t)w * N + jl);
                while (bits && !sc.invalid) {
                    const uint32_t g = (uint32_t)__ffs((int)bits) - 1u;
                    bits &= bits - 1;
                    sc.granule(w * 32 + g);
                }
            }
            if (sc.invalid) r_flags = B2R_ST_INVALID_TRANSITION;
            else {
                sc.finish();
                r_nrec = sc.n_rec; r_ncmp = sc.n_cmp;
                if (sc.overlap) r_flags = B2R_ST_OVERLAP;
            }
        }

        if (p.debug & 32) phase_b();
        // ---- D ----------------------------------------------------------------------------------------------------
        if (valid && !(p.debug & 8)) {
            if (too_long || (r_flags & B2R_ST_INVALID_TRANSITION)) kill_string(p, jl);
            else {
                uint32_t flags = r_flags;
#pragma unroll
                for (int d = 0; d < D; d++)
                    if (fin[d] == p.def[d].accepted_state) flags |= B2R_ST_ACCEPTED(d);
                if (p.records && r_nrec > p.max_records) flags |= B2R_ST_RECORDS_TRUNCATED;
                if (p.compact_bytes && r_ncmp > p.compact_pitch) flags |= B2R_ST_COMPACT_TRUNCATED;
                if (p.status) {
                    b2r_string_status st = {};
                    st.flags = flags; st.err_pos = NO_POS; st.n_records = r_nrec; st.n_compact = r_ncmp;
                    p.status[jl] = st;
                }
                tot.pad_rows += M - Ll; tot.n_ok++; tot.n_overlap += (r_flags & B2R_ST_OVERLAP) ? 1u : 0u;
            }
        }
        __syncwarp();
    }
};

// shared memory of the emitter: endpoint counters, then (optionally) the lookup tables.  Every thread of the CTA calls
// this, followed by a __syncthreads().
template <int D>
__device__ __forceinline__ void emit_tables_init(const WalkParams& p, unsigned char* esmem, EmitTables<D>& tb) {
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
}

// publish the per-lane totals and the CTA's endpoint counters; every thread of the CTA calls this
template <int D>
__device__ __forceinline__ void emit_publish(const WalkParams& p, const EmitTables<D>& tb, const EmitTotals& tot) {
    unsigned long long psum = tot.pad_rows;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) psum += __shfl_xor_sync(0xffffffffu, psum, o);
    const uint32_t nok = __reduce_add_sync(0xffffffffu, tot.n_ok), nov = __reduce_add_sync(0xffffffffu, tot.n_overlap);
    if ((threadIdx.x & 31) == 0) {
        if (psum) atomicAdd(&p.counters->pad_rows, psum);
        if (nov) atomicAdd(&p.counters->n_overlap, (unsigned long long)nov);
        if (nok) atomicAdd(&p.counters->n_ok_strings, (unsigned long long)nok);
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

template <int D, typename ST>
__global__ void __launch_bounds__(EMIT_THREADS) emit_kernel(const __grid_constant__ WalkParams p) {
    extern __shared__ __align__(16) unsigned char esmem[];
    const int lane = threadIdx.x & 31;
    EmitTables<D> tb;
    emit_tables_init<D>(p, esmem, tb);
    __syncthreads();

    const uint64_t warps_total = (uint64_t)gridDim.x * (blockDim.x >> 5);
    const uint64_t gw = (uint64_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    TileEmitter<D, ST> em(p, tb, lane);
    EmitTotals tot;
    for (uint64_t tile = gw; tile < p.n_tiles; tile += warps_total) em.run_tile(tile, tot);
    emit_publish<D>(p, tb, tot);
}

}  // namespace b2r
