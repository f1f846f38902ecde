// walk_direct_kernel — the hot kernel for definitions with at most 64 states per def (all shipped DFAs: 13..29 states).
//
// One LANE per string, one WARP per tile of 32 strings, persistent CTAs (one per SM).
//
// Shared memory per CTA
//   table  D x 64 KiB   direct next-state table: entry (u32) of (byte c, state s) at c*260 + s*4 (row stride 65 words: bank = (c+s) mod 32), so that ONE byte-permute
//                       builds the address from the previous entry (byte0 = s<<2) and the input word (byte1 = c):
//                         entry = [ next<<2 | next<<8 | substr_id<<16 | flags<<24 ]   (flags: is_start, is_end, invalid)
//                       this is the "dense 256 x S next-state table staged into shared memory" of the north star, padded to 64
//                       states per byte;
//   hist   D x 64 KiB   multiplicity bins, same (c,s) addressing (+HIST_OFF); when S <= 32 the bins live in the unused upper
//                       half of each 256-byte table row instead (HIST_OFF = 128) and no extra memory is needed;
//   zero   2 KiB        source of the TMA bulk zero-fills;
//   per warp            input tile 32 x (CH+16) B and state tile D x 32 x (CH+16) B.
//
// Per tile
//   1. every lane zero-fills ITS rows of the sparse columns (substr ids, enable bitmaps, masked chars / ids) with TMA bulk
//      stores (cp.async.bulk shared->global) from the zero buffer — no LSU instructions, completion awaited lazily;
//   2. chunks of CH positions: the warp stages the 32 strings' bytes with coalesced 16-byte loads into the padded tile,
//      each lane walks its own string with conflict-free LDS.128 reads:  PRMT (address) -> LDS (entry) per byte on the
//      dependent chain, plus one ATOMS.POPC.INC (multiplicity bin), one PRMT (state byte into the output pack) and one
//      LOP3+branch (rare-row test) off the chain; states go back through the state tile and out with coalesced 16-byte stores;
//   3. rare rows are queued and replayed at the end of the string (rare.cuh), in lockstep across the warp.
#pragma once
#include "rare.cuh"

namespace b2r {

constexpr int DCH = 64;               // positions per staged chunk
constexpr int DPITCH = DCH + 16;      // tile row pitch: 5 x 16 B keeps per-lane LDS.128 / STS.128 conflict-free
constexpr int ZERO_BYTES = 2048;
constexpr uint32_t DROW = 260;                 // bytes per table row (byte value): 65 words, so bank = (c + s) mod 32
constexpr uint32_t DTAB_BYTES = 256 * DROW;    // 66,560 B per def
constexpr int DIRECT_MAX_STATES = 64;
constexpr int DIRECT_MAX_THREADS = 512;

// direct-table entry encoding (built by build_direct_table in kernels.cu)
constexpr uint32_t DE_RARE_MASK = 0xFFFF0000u;   // substr id + flags
constexpr uint32_t DE_SID_MASK = 0x00FF0000u;

// PTX prmt (default mode): selector nibble 0-7 picks a byte of {a (0-3), b (4-7)}; nibble bit 3 replicates that byte's
// sign bit instead (used to produce zero bytes from a byte whose msb is known to be 0).  __byte_perm masks bit 3 away.
__device__ __forceinline__ uint32_t prmt(uint32_t a, uint32_t b, uint32_t sel) {   // sel folds to an immediate after unrolling
    uint32_t d;
    asm("prmt.b32 %0, %1, %2, %3;" : "=r"(d) : "r"(a), "r"(b), "r"(sel));
    return d;
}

__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void bulk_store(void* gdst, const void* ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(smem_u32(ssrc)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

template <int D, bool HIST, bool HIST_IN_ROW>
__global__ void __launch_bounds__(DIRECT_MAX_THREADS, 1) walk_direct_kernel(const __grid_constant__ WalkParams p, const uint32_t* __restrict__ gtab) {
    constexpr uint32_t hist_off = HIST_IN_ROW ? 128u : D * DTAB_BYTES;
    constexpr bool want_hist = HIST;
    extern __shared__ __align__(1024) unsigned char dsmem[];
    unsigned char* const smem = dsmem;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;
    const int n_warps = blockDim.x >> 5;

    // ---- shared memory carve-up --------------------------------------------------------------------------------
    unsigned char* const tab = smem;                                   // D x 64 KiB (+ D x 64 KiB bins when hist_off = D*64 KiB)
    const uint32_t tab_bytes = D * DTAB_BYTES;
    const uint32_t bins_bytes = HIST_IN_ROW ? 0u : tab_bytes;
    unsigned char* const zero = smem + tab_bytes + bins_bytes;
    CtaCounters* const cc = reinterpret_cast<CtaCounters*>(zero + ZERO_BYTES);
    unsigned char* const ep_base = zero + ZERO_BYTES + sizeof(CtaCounters);
    uint32_t* ep_s[D];
    ep_smem_layout<D>(p, ep_base, ep_s);
    unsigned char* const tiles = ep_base + p.ep_smem_bytes;
    unsigned char* const in_tile = tiles + (size_t)warp * (32 * DPITCH * (1 + D));
    unsigned char* const st_tile = in_tile + 32 * DPITCH;
    {
        const uint4* g = reinterpret_cast<const uint4*>(gtab);
        uint4* s4 = reinterpret_cast<uint4*>(tab);
        for (uint32_t i = threadIdx.x; i < tab_bytes / 16; i += blockDim.x) s4[i] = g[i];   // bins inside table rows arrive zeroed
        uint4* z4 = reinterpret_cast<uint4*>(tab + tab_bytes);
        for (uint32_t i = threadIdx.x; i < (bins_bytes + ZERO_BYTES) / 16; i += blockDim.x) z4[i] = make_uint4(0, 0, 0, 0);
    }
    cta_counters_init<D>(p, ep_base, cc);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");        // zero buffer -> visible to the TMA (async proxy)
    __syncthreads();

    const uint32_t M = p.max_chars;
    const uint32_t Mpad = (M + 15u) & ~15u;                             // rows written (row_pitch >= Mpad by contract)
    const uint32_t n_chunks = (Mpad + DCH - 1) / DCH;
    const uint64_t rp = p.row_pitch;

    for (uint32_t tile = blockIdx.x * n_warps + warp; tile < p.n_tiles; tile += gridDim.x * n_warps) {
        const uint64_t tile_base = (uint64_t)tile * 32;
        const uint64_t idx = tile_base + lane;
        const bool valid = idx < p.n_strings;
        uint64_t off = 0, end = 0;
        if (valid) { off = p.offsets[idx]; end = p.offsets[idx + 1]; }
        Cold<D, DirectTables> k;
        bool dead = !valid;
        if (valid && (end < off || end - off > (uint64_t)(M - 1))) {    // len must be <= M-1
            dead = true; end = off;
            k.idx = idx;
            kill_string<D, DirectTables>(p, k);
        }
        const uint32_t L = (uint32_t)(end - off);
        k.init(idx, p.bytes + off, L);
        uint32_t cur[D], expect[D];                                     // cur = entry that led to the current state
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t f = p.def[d].first_state;
            cur[d] = (f << 2) | (f << 8); expect[d] = 0;
            k.tb[d].tab = tab + d * DTAB_BYTES; k.ep_s[d] = ep_s[d];
        }

        // (1) TMA zero-fill of this lane's rows of the sparse columns
        if (valid) {
            auto zero_row = [&](uint8_t* row, uint32_t bytes) {
                for (uint32_t o = 0; o < bytes; o += ZERO_BYTES) bulk_store(row + o, zero, min(bytes - o, (uint32_t)ZERO_BYTES));
            };
            if (p.masked_chars) zero_row(p.masked_chars + idx * rp, Mpad);
            if (p.masked_substr_ids) zero_row(p.masked_substr_ids + idx * rp, Mpad);
#pragma unroll
            for (int d = 0; d < D; d++) {
                if (p.def[d].substr_ids) zero_row(p.def[d].substr_ids + idx * rp, Mpad);
                if (p.def[d].start_enable) zero_row(p.def[d].start_enable + idx * p.bitmap_pitch, (uint32_t)p.bitmap_pitch);
                if (p.def[d].end_enable) zero_row(p.def[d].end_enable + idx * p.bitmap_pitch, (uint32_t)p.bitmap_pitch);
            }
            bulk_commit();
        }
        bool zero_done = false;

        // staging geometry: this lane moves vector kv of rows r0 + 8*i (i = 0..3)
        const uint32_t shift = (uint32_t)(off & 15);
        const bool any_shift = __any_sync(0xffffffffu, shift != 0);
        const int kv = lane & 3, r0 = lane >> 2;
        const uint8_t* in_ptr[4];
        uint32_t in_left[4];                                            // bytes of the row from vector kv of chunk 0 to the string end
        uint64_t st_off[4];
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = r0 + 8 * i;
            const uint64_t roff = __shfl_sync(0xffffffffu, off, row);
            const uint64_t rend = __shfl_sync(0xffffffffu, end, row);
            const uint64_t a = (roff & ~uint64_t(15)) + (uint32_t)kv * 16;
            in_ptr[i] = p.bytes + a;
            in_left[i] = rend > a ? (uint32_t)(rend - a) : 0u;
            st_off[i] = (tile_base + row) * rp + (uint32_t)kv * 16;
        }
        const uint32_t rows_here = (p.n_strings - tile_base < 32) ? (uint32_t)(p.n_strings - tile_base) : 32u;

#pragma unroll 1
        for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
            const uint32_t cbase = chunk * DCH;
            // (2a) stage the input chunk (coalesced 16-byte loads -> padded tile)
#pragma unroll
            for (int i = 0; i < 4; i++) {
                uint4 val = make_uint4(0, 0, 0, 0);
                if (cbase < in_left[i]) val = *reinterpret_cast<const uint4*>(in_ptr[i] + cbase);
                *reinterpret_cast<uint4*>(in_tile + (r0 + 8 * i) * DPITCH + kv * 16) = val;
            }
            if (any_shift) {   // unaligned strings need a fifth vector per row
                const uint64_t a = (off & ~uint64_t(15)) + cbase + 64;
                uint4 val = make_uint4(0, 0, 0, 0);
                if (a < end) val = *reinterpret_cast<const uint4*>(p.bytes + a);
                *reinterpret_cast<uint4*>(in_tile + lane * DPITCH + 64) = val;
            }
            __syncwarp();

            // (2b) walk this lane's string over [cbase, cbase + DCH)
            const unsigned char* my_in = in_tile + lane * DPITCH;
#pragma unroll 1
            for (int g = 0; g < DCH / 16; g++) {
                const uint32_t gbase = cbase + g * 16;
                if (gbase >= Mpad) break;
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
                uint32_t pk[D][4];
                if (gbase + 16 <= L && !dead) {
                    // ---- hot path: 16 real characters --------------------------------------------------------------
#pragma unroll
                    for (int b = 0; b < 16; b++) {
                        uint32_t nxt[D];
                        uint32_t rare = 0;
#pragma unroll
                        for (int d = 0; d < D; d++) {
                            // address = state<<2 (byte 0 of cur) | c<<8 (byte b of the input word); upper bytes = sign(flags byte) = 0
                            const uint32_t c = prmt(w[b >> 2], 0u, 0x4440u + (b & 3));              // off the chain
                            const uint32_t addr = prmt(cur[d], w[b >> 2], 0xBB40u + ((b & 3) << 4)) + (c << 2);   // c*260 + s*4
                            nxt[d] = *reinterpret_cast<const uint32_t*>(tab + d * DTAB_BYTES + addr);
                            if (want_hist) atomicAdd(reinterpret_cast<uint32_t*>(tab + d * DTAB_BYTES + hist_off + addr), 1u);
                            // state byte (byte 1 of cur) into byte (b&3) of the output pack
                            const uint32_t sel = (b & 3) == 0 ? 0x3215u : (b & 3) == 1 ? 0x3250u : (b & 3) == 2 ? 0x3510u : 0x5210u;
                            pk[d][b >> 2] = prmt(pk[d][b >> 2], cur[d], sel);
                            rare |= (nxt[d] ^ expect[d]) & DE_RARE_MASK;
                        }
                        if (rare) {
                            uint32_t inval = 0;
                            Event<D>& ev = k.q[k.nq];
                            ev.pos = gbase + b; ev.c = (w[b >> 2] >> ((b & 3) * 8)) & 0xFFu;
#pragma unroll
                            for (int d = 0; d < D; d++) {
                                ev.e[d] = nxt[d]; ev.s[d] = (cur[d] >> 8) & 0xFFu; ev.nx[d] = (nxt[d] >> 8) & 0xFFu;
                                expect[d] = nxt[d] & DE_SID_MASK;
                                inval |= nxt[d] & ENT_INVALID;
                            }
                            if (inval) { dead = true; kill_string<D, DirectTables>(p, k); break; }
                            if (++k.nq == QCAP) {
                                if (!zero_done) { bulk_wait_all(); zero_done = true; }
                                drain<D, DirectTables>(p, k);
                            }
                        }
#pragma unroll
                        for (int d = 0; d < D; d++) cur[d] = nxt[d];
                    }
                } else {
                    // ---- ragged end: characters, then the final-state row, then dummy rows -------------------------
#pragma unroll
                    for (int d = 0; d < D; d++) pk[d][0] = pk[d][1] = pk[d][2] = pk[d][3] = 0;
#pragma unroll 1
                    for (int b = 0; b < 16; b++) {
                        const uint32_t pos = gbase + b;
                        const uint32_t c = (w[b >> 2] >> ((b & 3) * 8)) & 0xFFu;
                        uint32_t stb[D];
                        if (pos < L && !dead) {
                            uint32_t nxt[D];
                            uint32_t rare = 0;
#pragma unroll
                            for (int d = 0; d < D; d++) {
                                const uint32_t addr = (cur[d] & 0xFFu) + c * DROW;
                                nxt[d] = *reinterpret_cast<const uint32_t*>(tab + d * DTAB_BYTES + addr);
                                if (want_hist) atomicAdd(reinterpret_cast<uint32_t*>(tab + d * DTAB_BYTES + hist_off + addr), 1u);
                                stb[d] = (cur[d] >> 8) & 0xFFu;
                                rare |= (nxt[d] ^ expect[d]) & DE_RARE_MASK;
                            }
                            if (rare) {
                                uint32_t inval = 0;
                                Event<D>& ev = k.q[k.nq];
                                ev.pos = pos; ev.c = c;
#pragma unroll
                                for (int d = 0; d < D; d++) {
                                    ev.e[d] = nxt[d]; ev.s[d] = stb[d]; ev.nx[d] = (nxt[d] >> 8) & 0xFFu;
                                    expect[d] = nxt[d] & DE_SID_MASK;
                                    inval |= nxt[d] & ENT_INVALID;
                                }
                                if (inval) { dead = true; kill_string<D, DirectTables>(p, k); }
                                else if (++k.nq == QCAP) {
                                    if (!zero_done) { bulk_wait_all(); zero_done = true; }
                                    drain<D, DirectTables>(p, k);
                                }
                            }
                            if (!dead) {
#pragma unroll
                                for (int d = 0; d < D; d++) cur[d] = nxt[d];
                            }
                        } else {
#pragma unroll
                            for (int d = 0; d < D; d++) stb[d] = (pos <= L) ? ((cur[d] >> 8) & 0xFFu) : p.def[d].num_states;   // final state, then dummy
                            if (pos == L && !dead) {
                                uint32_t fs[D];
#pragma unroll
                                for (int d = 0; d < D; d++) fs[d] = (cur[d] >> 8) & 0xFFu;
                                if (!zero_done) { bulk_wait_all(); zero_done = true; }
                                finish_string<D, DirectTables>(p, k, fs);
                            }
                        }
#pragma unroll
                        for (int d = 0; d < D; d++) {
                            const uint32_t v = stb[d] << ((b & 3) * 8);
                            if ((b >> 2) == 0) pk[d][0] |= v; else if ((b >> 2) == 1) pk[d][1] |= v; else if ((b >> 2) == 2) pk[d][2] |= v; else pk[d][3] |= v;
                        }
                    }
                }
#pragma unroll
                for (int d = 0; d < D; d++)
                    *reinterpret_cast<uint4*>(st_tile + (d * 32 + lane) * DPITCH + g * 16) = make_uint4(pk[d][0], pk[d][1], pk[d][2], pk[d][3]);
            }
            __syncwarp();

            // (2c) store the state tile (coalesced 16-byte vectors)
#pragma unroll
            for (int d = 0; d < D; d++) {
                if (!p.def[d].states) continue;
                uint8_t* base = reinterpret_cast<uint8_t*>(p.def[d].states) + cbase;
#pragma unroll
                for (int i = 0; i < 4; i++) {
                    const int row = r0 + 8 * i;
                    if (row < (int)rows_here && cbase + kv * 16 < Mpad)
                        *reinterpret_cast<uint4*>(base + st_off[i]) = *reinterpret_cast<const uint4*>(st_tile + (d * 32 + row) * DPITCH + kv * 16);
                }
            }
            __syncwarp();
        }
        if (!zero_done && valid) bulk_wait_all();

        // per-tile counters: rows with enable = 0 all look up table row 0 (src/lib.rs:218-232 with enable = 0)
        cta_counters_tile(cc, valid && !dead, M - L, (k.flags & B2R_ST_OVERLAP) != 0);
    }

    __syncthreads();
    cta_counters_flush<D>(p, ep_s, cc);

    // ---- flush the multiplicity bins: bin (c,s) of def d -> dense global histogram [c*S + s] -----------------------------
    if (want_hist) {
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t S = p.def[d].num_states;
            for (uint32_t i = threadIdx.x; i < 256u * 64u; i += blockDim.x) {
                const uint32_t c = i >> 6, s = i & 63u;
                if (s >= S) continue;
                const uint32_t v = *reinterpret_cast<const uint32_t*>(tab + d * DTAB_BYTES + hist_off + c * DROW + s * 4);
                if (v) atomicAdd(p.def[d].hist + c * S + s, (unsigned long long)v);
            }
        }
    }
}

}  // namespace b2r
