// walk_kernel — the hot kernel of the DFA witness-generation path (sm_100a).
//
// One LANE per string, one WARP per tile of 32 strings:
//   * the input bytes of the tile's 32 strings are staged chunk by chunk (CH positions) into a padded shared-memory
//     tile with 16-byte coalesced global loads, so that each lane then reads ITS string with conflict-free LDS.128;
//   * each lane walks its DFA(s) sequentially: one packed-entry lookup per byte per def in the shared-memory
//     class/transition tables (entry = next state | substr id | is_start | is_end | invalid, see defs.hpp), which replaces
//     derive_states + derive_substr_ids + derive_is_start_end of the reference (src/lib.rs:804-888);
//   * the state column is transposed back through shared memory and stored with 16-byte coalesced stores; the sparse
//     columns (substr ids, start/end enable bitmaps, masked chars / ids) are zero-filled with coalesced stores and
//     patched by the owning lane only where they are non-zero (out-of-line "rare row" path);
//   * the start_mask/end_mask scans of the reference (src/lib.rs:598-714) are evaluated in closed form: events happen
//     only at boundaries where the id sum changes and is_start/is_end is set; mask = 1 exactly on the stretch between
//     an event with is_start and the NEXT event when that one has is_end (DESIGN.md "mask algebra");
//   * lookup-input multiplicities are accumulated per (byte, state) in a shared-memory histogram and flushed with
//     global 64-bit atomics (table row r <-> bin via PackedDef::row_bin).
#pragma once
#include <cuda_runtime.h>

#include <cstdint>

#include "defs.hpp"
#include "kernels.cuh"

namespace b2r {

constexpr int CH = 64;              // positions per staged chunk
constexpr int IN_PITCH = CH + 16;   // +1 vector for unaligned strings; 5 x 16 B keeps LDS.128 conflict-free
constexpr uint32_t NO_POS = 0xFFFFFFFFu;

template <typename ST>
struct StTile {
    static constexpr int ROW_BYTES = CH * (int)sizeof(ST);
    static constexpr int PITCH = ROW_BYTES + 16;        // odd multiple of 16 B
    static constexpr int VPR = ROW_BYTES / 16;          // vectors per row
};

// lane-private state that only the rare path touches (lives in local memory; its address is taken)
template <int D>
struct Cold {
    uint64_t idx;           // string index
    const uint8_t* src;     // first byte of the string
    uint32_t len;
    uint32_t run_start[D];  // first row of the current per-def substr-id run
    uint32_t seg_state[D];  // states at seg_start (to re-walk a multi-run segment)
    uint32_t sum_run;       // current id sum over defs
    uint32_t ie_pos, ie_val;// is_end sum that applies to boundary ie_pos
    uint32_t seg_sum;       // id sum at seg_start
    int32_t seg_start;      // first row of the pending (start-masked, not yet end-resolved) segment, -1 if none
    uint32_t seg_multi;     // the pending segment contains an unflagged id change
    uint32_t n_rec, n_cmp, flags;
    uint32_t dead;          // invalid transition seen / string skipped
    const uint8_t* cls[D];  // tables (generic pointers: shared or global)
    const uint32_t* trans[D];
};

__device__ __forceinline__ void fill_bytes(uint8_t* row, uint32_t a, uint32_t b, uint32_t v) {
    uint32_t i = a;
    while (i < b && (i & 15u)) row[i++] = (uint8_t)v;
    const uint32_t v4 = v * 0x01010101u;
    const uint4 vv = make_uint4(v4, v4, v4, v4);
    for (; i + 16 <= b; i += 16) *reinterpret_cast<uint4*>(row + i) = vv;   // rows are 16-byte aligned
    while (i < b) row[i++] = (uint8_t)v;
}

template <int D>
__device__ __forceinline__ void emit_record(const WalkParams& p, Cold<D>& k, uint32_t start, uint32_t len, uint32_t sid, uint32_t coff) {
    if (p.records && k.n_rec < p.max_records) {
        b2r_substr_record r; r.start = start; r.len = len; r.substr_id = sid; r.compact_off = coff;
        p.records[k.idx * p.max_records + k.n_rec] = r;
    }
    k.n_rec++;
}
template <int D>
__device__ __forceinline__ void put_masked(const WalkParams& p, Cold<D>& k, uint32_t i, uint32_t c) {
    if (p.masked_chars) p.masked_chars[k.idx * p.row_pitch + i] = (uint8_t)c;
    if (p.compact_bytes && k.n_cmp < p.compact_pitch) p.compact_bytes[k.idx * p.compact_pitch + k.n_cmp] = (uint8_t)c;
    k.n_cmp++;
}

// rows [a,b) are masked: start_mask = end_mask = 1 (src/lib.rs:740-764)
template <int D>
__device__ __noinline__ void finalize_segment(const WalkParams& p, Cold<D>& k, uint32_t a, uint32_t b) {
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
                const uint32_t e = k.trans[d][(uint32_t)k.cls[d][c] * p.def[d].num_states + st[d]];
                sum += (e & ENT_SID_MASK) >> ENT_SID_SHIFT;
                st[d] = e & ENT_NEXT_MASK;
            }
            if (i == a) run_sum = sum;
            else if (sum != run_sum) { emit_record(p, k, run_a, i - run_a, run_sum, run_coff); run_a = i; run_sum = sum; run_coff = k.n_cmp; }
            put_masked(p, k, i, c);
            if (p.masked_substr_ids) p.masked_substr_ids[k.idx * p.row_pitch + i] = (uint8_t)sum;
        }
        emit_record(p, k, run_a, b - run_a, run_sum, run_coff);
    }
}

// boundary `pos` where the id sum changes from k.sum_run to new_sum (forward event: src/lib.rs:613-642 at idx = pos;
// backward event: src/lib.rs:678-710 at M-idx = pos).  s[] = states at row pos.
template <int D>
__device__ __forceinline__ void boundary(const WalkParams& p, Cold<D>& k, uint32_t pos, uint32_t new_sum, uint32_t is_sum, uint32_t ie_sum, const uint32_t* s) {
    if (is_sum | ie_sum) {
        if (k.seg_start >= 0 && ie_sum) finalize_segment<D>(p, k, (uint32_t)k.seg_start, pos);
        if (is_sum) {
            k.seg_start = (int32_t)pos; k.seg_sum = new_sum; k.seg_multi = 0;
#pragma unroll
            for (int d = 0; d < D; d++) k.seg_state[d] = s[d];
        } else k.seg_start = -1;
    } else if (k.seg_start >= 0) k.seg_multi = 1;
}

__device__ __forceinline__ void set_bit(uint8_t* bitmap_row, uint32_t pos) {
    uint32_t* w = reinterpret_cast<uint32_t*>(bitmap_row) + (pos >> 5);
    __stcg(w, __ldcg(w) | (1u << (pos & 31)));
}

// Failure details of string j in the reference's order: derive_states walks def 0 over the whole string first, then
// def 1, ... (src/lib.rs:806-821), so the panic belongs to the LOWEST def index that fails, at its first failing byte.
static __device__ __noinline__ b2r_batch_status diagnose_string(const WalkParams& p, uint64_t j) {
    b2r_batch_status r = {};
    r.string_idx = j;
    const uint64_t off = p.offsets[j], end = p.offsets[j + 1];
    if (end < off || end - off > (uint64_t)(p.max_chars - 1)) {
        r.code = B2R_ERR_TOO_LONG;
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

// a row whose entries carry a substr id change, a flag or an invalid transition.
// e[] = entries of row pos, s[] = states at row pos, expect[] (in/out) = current per-def id in entry position.
template <int D>
__device__ __noinline__ void rare_row(const WalkParams& p, Cold<D>& k, uint32_t pos, uint32_t c, const uint32_t* e, const uint32_t* s, uint32_t* expect) {
    uint32_t invalid = 0;
#pragma unroll
    for (int d = 0; d < D; d++) invalid |= e[d] & ENT_INVALID;
    if (invalid && !k.dead) {   // the reference panics (src/lib.rs:817)
        k.dead = 1;
        atomicMin(&p.counters->first_bad, (unsigned long long)k.idx);
        if (p.status) {
            const b2r_batch_status r = diagnose_string(p, k.idx);
            b2r_string_status st = {};
            st.flags = B2R_ST_INVALID_TRANSITION; st.err_pos = r.pos; st.err_state = r.state; st.err_byte = r.byte; st.err_def = r.def;
            p.status[k.idx] = st;
        }
    }
    if (k.dead) return;
    uint32_t new_sum = 0, is_sum = 0, ie_next = 0;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const uint32_t sid = (e[d] & ENT_SID_MASK) >> ENT_SID_SHIFT;
        const uint32_t S = p.def[d].num_states;
        new_sum += sid;
        is_sum += (e[d] >> 24) & 1u;
        ie_next += (e[d] >> 25) & 1u;
        if ((e[d] & ENT_SID_MASK) != expect[d]) {   // per-def substr id run ends (src/lib.rs:825-845)
            const uint32_t old = expect[d] >> ENT_SID_SHIFT;
            if (old && p.def[d].substr_ids) fill_bytes(p.def[d].substr_ids + k.idx * p.row_pitch, k.run_start[d], pos, old);
            k.run_start[d] = pos; expect[d] = e[d] & ENT_SID_MASK;
        }
        if (e[d] & ENT_IS_START) {                   // start_enable, src/lib.rs:482-493; endpoint lookup :235-258
            if (p.def[d].start_enable) set_bit(p.def[d].start_enable + k.idx * p.bitmap_pitch, pos);
            atomicAdd(p.def[d].ep_start + (sid - p.def[d].sid_offset) * S + s[d], 1ull);
        }
        if (e[d] & ENT_IS_END) {                     // end_enable, src/lib.rs:501-513; endpoint lookup :260-284
            if (p.def[d].end_enable) set_bit(p.def[d].end_enable + k.idx * p.bitmap_pitch, pos);
            atomicAdd(p.def[d].ep_end + (sid - p.def[d].sid_offset) * S + (e[d] & ENT_NEXT_MASK), 1ull);
        }
    }
    const uint32_t ie_here = (k.ie_pos == pos) ? k.ie_val : 0;
    if (is_sum > 1 || ie_here > 1) k.flags |= B2R_ST_OVERLAP;
    if (new_sum != k.sum_run) boundary<D>(p, k, pos, new_sum, is_sum, ie_here, s);
    k.sum_run = new_sum;
    k.ie_pos = pos + 1; k.ie_val = ie_next;
}

// row `len`: the final-state row (src/lib.rs:404-418), last boundary, accept rule (src/lib.rs:427-457)
template <int D>
__device__ __noinline__ void finish_string(const WalkParams& p, Cold<D>& k, const uint32_t* s, const uint32_t* expect) {
    const uint32_t L = k.len;
    const uint32_t ie_here = (k.ie_pos == L) ? k.ie_val : 0;
    if (ie_here > 1) k.flags |= B2R_ST_OVERLAP;
#pragma unroll
    for (int d = 0; d < D; d++) {
        const uint32_t old = expect[d] >> ENT_SID_SHIFT;
        if (old && p.def[d].substr_ids) fill_bytes(p.def[d].substr_ids + k.idx * p.row_pitch, k.run_start[d], L, old);
        if (s[d] == p.def[d].accepted_state) k.flags |= B2R_ST_ACCEPTED(d);
    }
    if (k.sum_run != 0) boundary<D>(p, k, L, 0, 0, ie_here, s);
    if (p.status) {
        b2r_string_status st = {};
        st.flags = k.flags; st.err_pos = NO_POS;
        if (p.records && k.n_rec > p.max_records) st.flags |= B2R_ST_RECORDS_TRUNCATED;
        if (p.compact_bytes && k.n_cmp > p.compact_pitch) st.flags |= B2R_ST_COMPACT_TRUNCATED;
        st.n_records = k.n_rec; st.n_compact = k.n_cmp;
        p.status[k.idx] = st;
    }
}

template <int D, typename ST, bool TBL_SMEM, bool HIST_SMEM, int WARPS>
__global__ void __launch_bounds__(WARPS * 32) walk_kernel(const __grid_constant__ WalkParams p) {
    extern __shared__ __align__(16) unsigned char smem[];
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // ---- shared memory carve-up ------------------------------------------------------------------------------
    unsigned char* sp = smem;
    const uint8_t* cls_t[D];
    const uint32_t* trans_t[D];
    uint32_t* hist_t[D];
    uint32_t S_[D];
#pragma unroll
    for (int d = 0; d < D; d++) {
        S_[d] = p.def[d].num_states;
        if (TBL_SMEM) {
            uint32_t* tr = reinterpret_cast<uint32_t*>(sp);
            const uint32_t n = p.def[d].num_classes * S_[d];
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) tr[i] = p.def[d].trans[i];
            sp += (size_t)n * 4;
            uint8_t* cl = sp;
            for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) cl[i] = p.def[d].byte_class[i];
            sp += 256;
            trans_t[d] = tr; cls_t[d] = cl;
        } else {
            trans_t[d] = p.def[d].trans; cls_t[d] = p.def[d].byte_class;
        }
        if (HIST_SMEM) {
            uint32_t* h = reinterpret_cast<uint32_t*>(sp);
            const uint32_t n = 256u * S_[d];
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) h[i] = 0;
            sp += (size_t)n * 4;
            hist_t[d] = h;
        } else hist_t[d] = nullptr;
    }
    sp = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(sp) + 15) & ~uintptr_t(15));
    unsigned char* in_tile = sp + (size_t)warp * (32 * IN_PITCH + D * 32 * StTile<ST>::PITCH);
    unsigned char* st_tile = in_tile + 32 * IN_PITCH;
    __syncthreads();

    const uint32_t M = p.max_chars;
    const uint32_t Mpad = (M + 15u) & ~15u;                 // rows written (row_pitch >= Mpad by contract)
    const uint32_t n_chunks = (Mpad + CH - 1) / CH;
    const uint64_t rp = p.row_pitch;
    const bool want_hist = p.want_hist != 0;

    for (uint32_t tile = blockIdx.x * WARPS + warp; tile < p.n_tiles; tile += gridDim.x * WARPS) {
        const uint64_t tile_base = (uint64_t)tile * 32;
        const uint64_t idx = tile_base + lane;
        const bool valid = idx < p.n_strings;
        uint64_t off = 0, end = 0;
        if (valid) { off = p.offsets[idx]; end = p.offsets[idx + 1]; }
        Cold<D> k;
        k.idx = idx; k.sum_run = 0; k.ie_pos = NO_POS; k.ie_val = 0; k.seg_sum = 0; k.seg_start = -1; k.seg_multi = 0;
        k.n_rec = 0; k.n_cmp = 0; k.flags = 0; k.dead = valid ? 0u : 1u;
        uint32_t s[D], expect[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            s[d] = p.def[d].first_state; expect[d] = 0; k.run_start[d] = 0; k.seg_state[d] = 0;
            k.cls[d] = cls_t[d]; k.trans[d] = trans_t[d];
        }
        if (valid && (end < off || end - off > (uint64_t)(M - 1))) {     // len must be <= M-1
            k.dead = 1; end = off;
            atomicMin(&p.counters->first_bad, (unsigned long long)idx);
            if (p.status) {
                b2r_string_status st = {};
                st.flags = B2R_ST_TOO_LONG; st.err_pos = NO_POS;
                p.status[idx] = st;
            }
        }
        const uint32_t L = (uint32_t)(end - off);
        k.len = L; k.src = p.bytes + off;
        bool dead = k.dead != 0;   // register mirror (k lives in local memory: its address escapes to the rare path)
        const uint64_t abase = off & ~uint64_t(15);
        const uint32_t shift = (uint32_t)(off & 15);
        const bool any_shift = __any_sync(0xffffffffu, shift != 0);

        // zero the tile's bitmap rows (32 adjacent rows are contiguous)
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint64_t nrows = (p.n_strings - tile_base < 32) ? (p.n_strings - tile_base) : 32;
            const uint64_t words = nrows * p.bitmap_pitch / 4;
            if (p.def[d].start_enable) {
                uint32_t* w = reinterpret_cast<uint32_t*>(p.def[d].start_enable + tile_base * p.bitmap_pitch);
                for (uint64_t i = lane; i < words; i += 32) w[i] = 0;
            }
            if (p.def[d].end_enable) {
                uint32_t* w = reinterpret_cast<uint32_t*>(p.def[d].end_enable + tile_base * p.bitmap_pitch);
                for (uint64_t i = lane; i < words; i += 32) w[i] = 0;
            }
        }

        // ---- chunk loop ------------------------------------------------------------------------------------------
#pragma unroll 1
        for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
            const uint32_t cbase = chunk * CH;
            // (1) stage the input chunk: rows of 4 (aligned tile) or 5 vectors starting at the 16-byte aligned string base
#pragma unroll
            for (int i = 0; i < 5; i++) {
                const int v = lane + 32 * i;
                const int row = any_shift ? v / 5 : v >> 2;
                const int kk = any_shift ? v - row * 5 : v & 3;
                const uint64_t rbase = __shfl_sync(0xffffffffu, abase, row & 31);
                const uint64_t rend = __shfl_sync(0xffffffffu, end, row & 31);
                if (row < 32) {
                    const uint64_t g = rbase + cbase + (uint32_t)kk * 16;
                    uint4 val = make_uint4(0, 0, 0, 0);
                    if (g < rend) val = *reinterpret_cast<const uint4*>(p.bytes + g);
                    *reinterpret_cast<uint4*>(in_tile + row * IN_PITCH + kk * 16) = val;
                }
            }
            // (2) zero-fill the sparse byte columns of this chunk (coalesced, straight from registers)
            {
                const uint4 z = make_uint4(0, 0, 0, 0);
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int v = lane + 32 * i;
                    const int row = v >> 2, kk = v & 3;
                    const uint32_t col = cbase + kk * 16;
                    if (tile_base + row < p.n_strings && col < Mpad) {
                        const uint64_t o = (tile_base + row) * rp + col;
                        if (p.masked_chars) *reinterpret_cast<uint4*>(p.masked_chars + o) = z;
                        if (p.masked_substr_ids) *reinterpret_cast<uint4*>(p.masked_substr_ids + o) = z;
#pragma unroll
                        for (int d = 0; d < D; d++)
                            if (p.def[d].substr_ids) *reinterpret_cast<uint4*>(p.def[d].substr_ids + o) = z;
                    }
                }
            }
            __syncwarp();

            // (3) walk: this lane's string, positions [cbase, cbase+CH)
            const unsigned char* my_in = in_tile + lane * IN_PITCH;
#pragma unroll 1
            for (int g = 0; g < CH / 16; g++) {
                const uint32_t gbase = cbase + g * 16;
                uint32_t w[4];
                if (!any_shift) {
                    const uint4 t = *reinterpret_cast<const uint4*>(my_in + g * 16);
                    w[0] = t.x; w[1] = t.y; w[2] = t.z; w[3] = t.w;
                } else {
                    const uint32_t* q = reinterpret_cast<const uint32_t*>(my_in + g * 16 + (shift & ~3u));
                    const uint32_t sh = (shift & 3u) * 8;
                    const uint32_t x0 = q[0], x1 = q[1], x2 = q[2], x3 = q[3], x4 = q[4];
                    w[0] = __funnelshift_r(x0, x1, sh); w[1] = __funnelshift_r(x1, x2, sh);
                    w[2] = __funnelshift_r(x2, x3, sh); w[3] = __funnelshift_r(x3, x4, sh);
                }
                ST pk[D][16];
                if (gbase + 16 <= L && !dead) {
                    // all 16 rows are real characters
#pragma unroll
                    for (int b = 0; b < 16; b++) {
                        const uint32_t c = (w[b >> 2] >> ((b & 3) * 8)) & 0xFFu;
                        uint32_t e[D];
                        uint32_t rare = 0;
#pragma unroll
                        for (int d = 0; d < D; d++) {
                            if (TBL_SMEM) e[d] = trans_t[d][(uint32_t)cls_t[d][c] * S_[d] + s[d]];
                            else e[d] = __ldg(trans_t[d] + (uint32_t)__ldg(cls_t[d] + c) * S_[d] + s[d]);
                            if (want_hist) {
                                if (HIST_SMEM) atomicAdd(hist_t[d] + c * S_[d] + s[d], 1u);
                                else atomicAdd(p.def[d].hist + c * S_[d] + s[d], 1ull);
                            }
                            pk[d][b] = (ST)s[d];
                            rare |= (e[d] ^ expect[d]) & ENT_RARE_MASK;
                        }
                        if (rare) {   // only copies escape, so s/e/expect stay in registers
                            uint32_t te[D], ts[D], tx[D];
#pragma unroll
                            for (int d = 0; d < D; d++) { te[d] = e[d]; ts[d] = s[d]; tx[d] = expect[d]; }
                            rare_row<D>(p, k, gbase + b, c, te, ts, tx);
#pragma unroll
                            for (int d = 0; d < D; d++) expect[d] = tx[d];
                            dead = k.dead != 0;
                            if (dead) break;
                        }
#pragma unroll
                        for (int d = 0; d < D; d++) s[d] = e[d] & ENT_NEXT_MASK;
                    }
                } else {
                    // ragged end: characters, then the final-state row, then dummy rows
#pragma unroll 1
                    for (int b = 0; b < 16; b++) {
                        const uint32_t pos = gbase + b;
                        const uint32_t c = (w[b >> 2] >> ((b & 3) * 8)) & 0xFFu;
                        if (pos < L && !dead) {
                            uint32_t e[D];
                            uint32_t rare = 0;
#pragma unroll
                            for (int d = 0; d < D; d++) {
                                if (TBL_SMEM) e[d] = trans_t[d][(uint32_t)cls_t[d][c] * S_[d] + s[d]];
                                else e[d] = __ldg(trans_t[d] + (uint32_t)__ldg(cls_t[d] + c) * S_[d] + s[d]);
                                if (want_hist) {
                                    if (HIST_SMEM) atomicAdd(hist_t[d] + c * S_[d] + s[d], 1u);
                                    else atomicAdd(p.def[d].hist + c * S_[d] + s[d], 1ull);
                                }
                                rare |= (e[d] ^ expect[d]) & ENT_RARE_MASK;
                            }
                            // state byte goes straight to the tile (dynamic b: no register array indexing)
#pragma unroll
                            for (int d = 0; d < D; d++)
                                reinterpret_cast<ST*>(st_tile + (d * 32 + lane) * StTile<ST>::PITCH)[g * 16 + b] = (ST)s[d];
                            if (rare) {
                                uint32_t te[D], ts[D], tx[D];
#pragma unroll
                                for (int d = 0; d < D; d++) { te[d] = e[d]; ts[d] = s[d]; tx[d] = expect[d]; }
                                rare_row<D>(p, k, pos, c, te, ts, tx);
#pragma unroll
                                for (int d = 0; d < D; d++) expect[d] = tx[d];
                                dead = k.dead != 0;
                            }
                            if (!dead) {
#pragma unroll
                                for (int d = 0; d < D; d++) s[d] = e[d] & ENT_NEXT_MASK;
                            }
                        } else {
#pragma unroll
                            for (int d = 0; d < D; d++)
                                reinterpret_cast<ST*>(st_tile + (d * 32 + lane) * StTile<ST>::PITCH)[g * 16 + b] = (ST)((pos <= L) ? s[d] : S_[d]);  // final state, then dummy
                            if (pos == L && !dead) {
                                uint32_t ts[D], tx[D];
#pragma unroll
                                for (int d = 0; d < D; d++) { ts[d] = s[d]; tx[d] = expect[d]; }
                                finish_string<D>(p, k, ts, tx);
                            }
                        }
                    }
                    continue;   // tile bytes already written
                }
#pragma unroll
                for (int d = 0; d < D; d++) {
                    unsigned char* dst = st_tile + (d * 32 + lane) * StTile<ST>::PITCH + g * 16 * (int)sizeof(ST);
                    if (sizeof(ST) == 1) {
                        uint32_t r[4];
#pragma unroll
                        for (int q = 0; q < 4; q++)
                            r[q] = (uint32_t)pk[d][4 * q] | ((uint32_t)pk[d][4 * q + 1] << 8) | ((uint32_t)pk[d][4 * q + 2] << 16) | ((uint32_t)pk[d][4 * q + 3] << 24);
                        *reinterpret_cast<uint4*>(dst) = make_uint4(r[0], r[1], r[2], r[3]);
                    } else {
                        uint32_t r[8];
#pragma unroll
                        for (int q = 0; q < 8; q++) r[q] = (uint32_t)pk[d][2 * q] | ((uint32_t)pk[d][2 * q + 1] << 16);
                        *reinterpret_cast<uint4*>(dst) = make_uint4(r[0], r[1], r[2], r[3]);
                        *reinterpret_cast<uint4*>(dst + 16) = make_uint4(r[4], r[5], r[6], r[7]);
                    }
                }
            }
            __syncwarp();

            // (4) store the state tile (coalesced 16-byte vectors)
#pragma unroll
            for (int d = 0; d < D; d++) {
                if (!p.def[d].states) continue;
                constexpr int VPR = StTile<ST>::VPR;
#pragma unroll
                for (int i = 0; i < VPR; i++) {
                    const int v = lane + 32 * i;
                    const int row = v / VPR, kk = v % VPR;
                    const uint32_t col = cbase + kk * (16 / (int)sizeof(ST));
                    if (tile_base + row < p.n_strings && col < Mpad) {
                        const uint4 val = *reinterpret_cast<const uint4*>(st_tile + (d * 32 + row) * StTile<ST>::PITCH + kk * 16);
                        *reinterpret_cast<uint4*>(reinterpret_cast<unsigned char*>(p.def[d].states) + ((tile_base + row) * rp + col) * sizeof(ST)) = val;
                    }
                }
            }
            __syncwarp();
        }

        // per-tile counters: rows with enable = 0 all look up table row 0 (src/lib.rs:218-232 with enable = 0)
        {
            const bool ok = valid && !dead;
            const uint32_t pad_sum = __reduce_add_sync(0xffffffffu, ok ? (M - L) : 0u);
            const uint32_t ov = __popc(__ballot_sync(0xffffffffu, ok && (k.flags & B2R_ST_OVERLAP) != 0));
            const uint32_t okc = __popc(__ballot_sync(0xffffffffu, ok));
            if (lane == 0) {
                if (pad_sum) atomicAdd(&p.counters->pad_rows, (unsigned long long)pad_sum);
                if (ov) atomicAdd(&p.counters->n_overlap, (unsigned long long)ov);
                if (okc) atomicAdd(&p.counters->n_ok_strings, (unsigned long long)okc);
            }
        }
    }

    // ---- flush the multiplicity bins -----------------------------------------------------------------------------
    if (HIST_SMEM) {
        __syncthreads();
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t n = 256u * S_[d];
            for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
                const uint32_t v = hist_t[d][i];
                if (v) atomicAdd(p.def[d].hist + i, (unsigned long long)v);
            }
        }
    }
}

template <int D, typename ST, bool TS, bool HS, int WARPS>
static int launch_one(const WalkParams& p, size_t smem, int grid, cudaStream_t st) {
    auto kern = walk_kernel<D, ST, TS, HS, WARPS>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2R_ERR_CUDA; }
    kern<<<grid, WARPS * 32, smem, st>>>(p);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("walk_kernel launch: %s", cudaGetErrorString(e)); return B2R_ERR_CUDA; }
    return B2R_OK;
}

constexpr int WALK_WARPS = 8;

// one instantiation set per number of defs (walk_inst.cu is compiled once per D, in parallel)
template <int D>
int launch_walk_d(const WalkParams& p, bool wide, bool ts, bool hs, size_t smem, int grid, cudaStream_t st);

}  // namespace b2r
