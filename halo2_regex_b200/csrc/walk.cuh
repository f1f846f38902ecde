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
#include "rare.cuh"

namespace b2r {

constexpr int CH = 64;              // positions per staged chunk
constexpr int IN_PITCH = CH + 16;   // +1 vector for unaligned strings; 5 x 16 B keeps LDS.128 conflict-free

template <typename ST>
struct StTile {
    static constexpr int ROW_BYTES = CH * (int)sizeof(ST);
    static constexpr int PITCH = ROW_BYTES + 16;        // odd multiple of 16 B
    static constexpr int VPR = ROW_BYTES / 16;          // vectors per row
};

// one rare row handled immediately, out of line (class-compressed entry encoding: next state in bits 0..15)
template <int D>
__device__ __noinline__ void rare_row_generic(const WalkParams& p, const Cold<D, 1>& k, const RowCtx<D, ClassTables>& x, uint32_t pos, const uint32_t* e,
                                              const uint32_t* s, uint32_t* expect) {
    uint32_t nx[D], run_sid[D];
#pragma unroll
    for (int d = 0; d < D; d++) { nx[d] = e[d] & ENT_NEXT_MASK; run_sid[d] = expect[d] >> 16; }
    (void)nx; (void)run_sid;
    push_row<D, 1, ClassTables>(p, k, x, pos, e, s);
#pragma unroll
    for (int d = 0; d < D; d++) expect[d] = e[d] & ENT_SID_MASK;
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
    CtaCounters* const cc = reinterpret_cast<CtaCounters*>(sp);
    unsigned char* const ep_base = sp + sizeof(CtaCounters);
    uint32_t* ep_s[D];
    ep_smem_layout<D>(p, ep_base, ep_s);
    cta_counters_init<D>(p, ep_base, cc);
    sp = ep_base + ((p.ep_smem_bytes + 15u) & ~15u);
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
        uint32_t cold_store[cold_fields(D)];                              // cold state in local memory (stride 1)
        const Cold<D, 1> k{cold_store};
        bool dead = !valid;
        if (valid && (end < off || end - off > (uint64_t)(M - 1))) {     // len must be <= M-1
            dead = true; end = off;
            kill_string(p, idx);
        }
        const uint32_t L = (uint32_t)(end - off);
        k.init();
        RowCtx<D, ClassTables> x;
        x.idx = idx; x.src = p.bytes + off; x.len = L; x.tile_pos = NO_POS; x.tile_s = 0;
        x.qbase = p.queue + ((size_t)(blockIdx.x * WARPS + warp) * queue_words(D)) * 32 + lane;
        uint32_t s[D], expect[D];
#pragma unroll
        for (int d = 0; d < D; d++) {
            s[d] = p.def[d].first_state; expect[d] = 0;
            x.tb[d].cls = cls_t[d]; x.tb[d].trans = trans_t[d]; x.tb[d].S = S_[d]; x.ep_s[d] = ep_s[d];
        }
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
                        if (rare) {   // only copies escape to the out-of-line call: s/e/expect stay in registers
                            uint32_t inval = 0;
#pragma unroll
                            for (int d = 0; d < D; d++) inval |= e[d] & ENT_INVALID;
                            if (inval) { if (!dead) { dead = true; kill_string(p, idx); } }
                            else if (!dead) {
                                uint32_t te[D], ts[D], tx[D];
#pragma unroll
                                for (int d = 0; d < D; d++) { te[d] = e[d]; ts[d] = s[d]; tx[d] = expect[d]; }
                                rare_row_generic<D>(p, k, x, gbase + b, te, ts, tx);
#pragma unroll
                                for (int d = 0; d < D; d++) expect[d] = tx[d];
                            }
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
                                uint32_t inval = 0;
#pragma unroll
                                for (int d = 0; d < D; d++) inval |= e[d] & ENT_INVALID;
                                if (inval) { dead = true; kill_string(p, idx); }
                                else {
                                    uint32_t te[D], ts[D], tx[D];
#pragma unroll
                                    for (int d = 0; d < D; d++) { te[d] = e[d]; ts[d] = s[d]; tx[d] = expect[d]; }
                                    rare_row_generic<D>(p, k, x, pos, te, ts, tx);
#pragma unroll
                                    for (int d = 0; d < D; d++) expect[d] = tx[d];
                                }
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
                                uint32_t ts[D];
#pragma unroll
                                for (int d = 0; d < D; d++) ts[d] = s[d];
                                finish_string<D, 1, ClassTables>(p, k, x, ts);
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
        cta_counters_tile(cc, valid && !dead, M - L, (cold_store[CF_FLAGS] & B2R_ST_OVERLAP) != 0);
    }

    __syncthreads();
    cta_counters_flush<D>(p, ep_s, cc);

    // ---- flush the multiplicity bins -----------------------------------------------------------------------------
    if (HIST_SMEM) {
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
