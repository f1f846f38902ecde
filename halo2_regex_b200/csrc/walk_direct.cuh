// walk_direct_kernel — the hot kernel for definitions with at most 64 states per def (all shipped DFAs: 13..29 states).
//
// One LANE per string, one WARP per tile of 32 strings, persistent CTAs (one per SM), tiles handed out by an atomic counter.
//
// Shared memory per CTA
//   table  D x 65 KiB   direct next-state table: entry (u32) of (byte c, state s) at c*260 + s*4.  The 65-word row stride
//                       makes the bank (c + s) mod 32, so lanes sitting in the same state with different bytes do not
//                       conflict.  entry = [ next<<2 | next<<8 | substr_id<<16 | flags<<24 ] (flags: is_start, is_end,
//                       invalid): ONE byte-permute of the previous entry (byte0 = s<<2) and the input word (byte1 = c)
//                       gives c*256 + s*4, and c*4 + table base is added off the dependent chain.  This is the "dense
//                       256 x S next-state table staged into shared memory" of the north star, padded to 65 slots per byte;
//   hist               multiplicity bins with the same (c,s) addressing: inside the unused upper half of each table row when
//                       S <= 32 (HIST_IN_ROW, +128 B), else a second D x 65 KiB region;
//   zero   8 KiB        source of the TMA bulk zero-fills;
//   per warp            2 input tiles (double buffer) of 32 x (DCH+16) B, a state tile of D x 32 x (DCH+16) B and the
//                       lanes' cold state (struct of arrays).
//
// Per tile
//   1. lane 0 zero-fills the tile's 32 adjacent rows of every sparse column (substr ids, enable bitmaps, masked chars / ids)
//      with TMA bulk stores (cp.async.bulk shared->global) from the zero buffer, one tile AHEAD of the walk;
//   2. chunks of DCH positions: cp.async (16 B, zero-filled past the string end) stages chunk k+1 of the 32 strings into
//      the spare input tile while chunk k is walked; each lane reads ITS string with conflict-free LDS.128 and walks it:
//      PRMT (address) -> IADD -> LDS (entry) per byte on the dependent chain, plus one shared-memory atomic (multiplicity
//      bin), PRMT (state byte into the output pack) and the rare-row test off the chain; states go back through the state
//      tile and out with coalesced 16-byte stores;
//   3. the hot path only flags words that contain a rare row; they are replayed exactly, in order, once per chunk by the
//      whole warp in lockstep (chunk_rare), with the lane's cold state in shared memory (rare.cuh).
#pragma once
#include "rare.cuh"

namespace b2r {

constexpr int DCH = 64;               // positions per staged chunk
constexpr int DPITCH = DCH + 16;      // tile row pitch: 5 x 16 B keeps per-lane LDS.128 / STS.128 conflict-free
constexpr int ZERO_BYTES = 8192;
constexpr uint32_t DROW = 260;                 // bytes per table row (byte value): 65 words, so bank = (c + s) mod 32
constexpr uint32_t DTAB_BYTES = 256 * DROW;    // 66,560 B per def
constexpr int DIRECT_MAX_STATES = 64;
constexpr int DIRECT_MAX_THREADS = 512;
__host__ __device__ constexpr int direct_tile_bytes_per_warp(int D) { return 32 * DPITCH * (2 + D) + cold_fields(D) * 128; }

// direct-table entry encoding (built by build_direct_table in defs.cpp)
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
__device__ __forceinline__ uint32_t lds32(uint32_t saddr) {
    uint32_t v;
    asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(saddr));
    return v;
}
__device__ __forceinline__ void bulk_store(void* gdst, uint32_t ssrc, uint32_t bytes) {
    asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(gdst), "r"(ssrc), "r"(bytes) : "memory");
}
__device__ __forceinline__ void bulk_commit() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
__device__ __forceinline__ void bulk_wait_all() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
// 16-byte async copy global -> shared; copies src_bytes (0 or 16) and zero-fills the rest
__device__ __forceinline__ void cp_async16(uint32_t sdst, const void* gsrc, uint32_t src_bytes) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst), "l"(gsrc), "r"(src_bytes) : "memory");
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;" ::: "memory"); }
template <int N>
__device__ __forceinline__ void cp_async_wait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }

// one rare row handled immediately (ragged path), out of line
template <int D>
__device__ __noinline__ void rare_row_now(const WalkParams& p, const Cold<D, 32>& k, const RowCtx<D, DirectTables>& x, uint32_t pos, const uint32_t* e,
                                          const uint32_t* s, uint32_t* expect) {
    uint32_t nx[D], run_sid[D];
#pragma unroll
    for (int d = 0; d < D; d++) { nx[d] = (e[d] >> 8) & 0xFFu; run_sid[d] = expect[d] >> 16; }
    (void)nx; (void)run_sid;
    push_row<D, 32, DirectTables>(p, k, x, pos, e, s);
#pragma unroll
    for (int d = 0; d < D; d++) expect[d] = e[d] & DE_SID_MASK;
}

// Exact, in-order replay of the rare rows of one chunk, run once per chunk by the whole warp in lockstep.
// The hot path only sets bit n of `bits` when word n (rows cbase+4n .. cbase+4n+3) contains an entry whose substr id /
// flags differ from `expect` AS IT WAS AT THE START OF THE CHUNK (stale).  Here every flagged word is re-examined row by row
// from shared memory (4 bytes from the input tile, 4 states from the state tile, 4 independent table lookups); while the
// true `expect` differs from the stale one the following words are examined too, flagged or not, so no id change is
// missed.  Rare rows are only QUEUED here (fire-and-forget stores into the lane's L2-resident queue slice); their heavy
// processing happens at the end of the string, in lockstep (rare.cuh).  All arguments are scalars in registers.
// n_words: words of the chunk the hot path has walked (later rows belong to the ragged path, which is exact by itself).
// in_s / st_s: shared addresses of this lane's input bytes (shift applied) and state bytes of def 0 for row cbase.
// Returns 0 = ok, 1 = invalid transition (the reference panics, src/lib.rs:817), 2 = the queue is full: the caller
// drains it (out of line) and calls again with the returned *bits / *next_word.
template <int D>
__device__ __noinline__ uint32_t chunk_rare(uint32_t* cold_base, uint32_t* qbase, uint32_t* bits_io, uint32_t* next_word_io, uint32_t n_words, uint32_t cbase,
                                            uint32_t in_s, uint32_t st_s, uint32_t tab_s, const uint32_t* stale, uint32_t* expect) {
    const Cold<D, 32> k{cold_base};
    uint32_t bits = *bits_io;
    uint32_t run[D];
#pragma unroll
    for (int d = 0; d < D; d++) run[d] = expect[d];
    uint32_t nq = k.f(CF_NQ);
    uint32_t n = *next_word_io;
    if (n == 0xFFFFFFFFu) n = bits ? (uint32_t)__ffs((int)bits) - 1u : 16u;
    uint32_t rc = 0;
    while (n < n_words) {
        if (nq + 4 > QCAP) { rc = 2; break; }                           // room for the 4 rows of a word
        uint32_t c[4], sv[D], e[4][D];
#pragma unroll
        for (int j = 0; j < 4; j++) c[j] = lds8(in_s + n * 4 + j);
#pragma unroll
        for (int d = 0; d < D; d++) sv[d] = lds32(st_s + d * (32 * DPITCH) + n * 4);   // 4 state bytes
#pragma unroll
        for (int j = 0; j < 4; j++)
#pragma unroll
            for (int d = 0; d < D; d++) e[j][d] = lds32(tab_s + d * DTAB_BYTES + c[j] * DROW + ((sv[d] >> (8 * j)) & 0xFFu) * 4);
#pragma unroll
        for (int j = 0; j < 4; j++) {
            uint32_t rare = 0, inval = 0;
#pragma unroll
            for (int d = 0; d < D; d++) { rare |= (e[j][d] ^ run[d]) & DE_RARE_MASK; inval |= e[j][d] & ENT_INVALID; }
            if (inval) { rc = 1; break; }
            if (rare) {
                uint32_t* q = qbase + nq * (1 + 2 * D) * 32;
                __stcg(q, cbase + n * 4 + j);
#pragma unroll
                for (int d = 0; d < D; d++) {
                    __stcg(q + (1 + 2 * d) * 32, e[j][d]);
                    __stcg(q + (2 + 2 * d) * 32, (sv[d] >> (8 * j)) & 0xFFu);
                    run[d] = e[j][d] & DE_SID_MASK;
                }
                nq++;
            }
        }
        if (rc) break;
        bits &= ~(1u << n);
        bool differs = false;
#pragma unroll
        for (int d = 0; d < D; d++) differs = differs || run[d] != stale[d];
        n = differs ? n + 1 : (bits ? (uint32_t)__ffs((int)bits) - 1u : 16u);
    }
    k.f(CF_NQ) = nq;
#pragma unroll
    for (int d = 0; d < D; d++) expect[d] = run[d];
    *bits_io = bits; *next_word_io = n;
    return rc;
}

// out-of-line drain of a full queue in the middle of a string (rare: more than QCAP-3 rare rows pending)
template <int D>
__device__ __noinline__ void drain_now(const WalkParams& p, const Cold<D, 32>& k, const RowCtx<D, DirectTables>& x) {
    drain<D, 32, DirectTables>(p, k, x);
}

template <int D, bool HIST, bool HIST_IN_ROW>
__global__ void __launch_bounds__(DIRECT_MAX_THREADS, 1) walk_direct_kernel(const __grid_constant__ WalkParams p, const uint32_t* __restrict__ gtab) {
    constexpr uint32_t hist_off = HIST_IN_ROW ? 128u : D * DTAB_BYTES;
    extern __shared__ __align__(1024) unsigned char dsmem[];
    unsigned char* const smem = dsmem;
    const int lane = threadIdx.x & 31;
    const int warp = threadIdx.x >> 5;

    // ---- shared memory carve-up --------------------------------------------------------------------------------
    unsigned char* const tab = smem;                                   // D tables (+ D bin regions unless HIST_IN_ROW)
    const uint32_t tab_bytes = D * DTAB_BYTES;
    const uint32_t bins_bytes = (HIST && !HIST_IN_ROW) ? tab_bytes : 0u;
    unsigned char* const zero = smem + tab_bytes + bins_bytes;
    CtaCounters* const cc = reinterpret_cast<CtaCounters*>(zero + ZERO_BYTES);
    unsigned char* const ep_base = zero + ZERO_BYTES + sizeof(CtaCounters);
    uint32_t* ep_s[D];
    ep_smem_layout<D>(p, ep_base, ep_s);
    unsigned char* const tiles = ep_base + p.ep_smem_bytes;
    unsigned char* const in_tile0 = tiles + (size_t)warp * direct_tile_bytes_per_warp(D);
    unsigned char* const st_tile = in_tile0 + 2 * 32 * DPITCH;
    const Cold<D, 32> k{reinterpret_cast<uint32_t*>(st_tile + D * 32 * DPITCH) + lane};   // cold state, struct of arrays
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
    const uint32_t tab_s = smem_u32(tab);
    const uint32_t zero_s = smem_u32(zero);
    const uint32_t in_s = smem_u32(in_tile0);
    const int kv = lane & 3, r0 = lane >> 2;                            // staging: this lane moves vector kv of rows r0 + 8*i

    // lane 0: TMA zero-fill of the 32 adjacent rows of tile t in every sparse column
    auto zero_fill_tile = [&](uint32_t t) {
        const uint64_t base = (uint64_t)t * 32;
        const uint64_t rows = (p.n_strings - base < 32) ? (p.n_strings - base) : 32;
        auto fill = [&](uint8_t* ptr, uint64_t bytes) {
            for (uint64_t o = 0; o < bytes; o += ZERO_BYTES) bulk_store(ptr + o, zero_s, (uint32_t)(bytes - o < ZERO_BYTES ? bytes - o : ZERO_BYTES));
        };
        if (p.masked_chars) fill(p.masked_chars + base * rp, rows * rp);
        if (p.masked_substr_ids) fill(p.masked_substr_ids + base * rp, rows * rp);
#pragma unroll
        for (int d = 0; d < D; d++) {
            if (p.def[d].substr_ids) fill(p.def[d].substr_ids + base * rp, rows * rp);
            if (p.def[d].start_enable) fill(p.def[d].start_enable + base * p.bitmap_pitch, rows * p.bitmap_pitch);
            if (p.def[d].end_enable) fill(p.def[d].end_enable + base * p.bitmap_pitch, rows * p.bitmap_pitch);
        }
        bulk_commit();
    };
    auto next_tile = [&]() -> uint32_t {
        unsigned long long t = 0;
        if (lane == 0) t = atomicAdd(&p.counters->tile_counter, 1ull);
        t = __shfl_sync(0xffffffffu, t, 0);
        return t < p.n_tiles ? (uint32_t)t : 0xFFFFFFFFu;
    };

    uint32_t tile = next_tile();
    if (tile != 0xFFFFFFFFu && lane == 0) zero_fill_tile(tile);

    while (tile != 0xFFFFFFFFu) {
        const uint32_t tile_after = next_tile();
        // the zero-fill of THIS tile was issued one tile ago: wait for it, then start the next one
        if (lane == 0) {
            bulk_wait_all();
            if (tile_after != 0xFFFFFFFFu) zero_fill_tile(tile_after);
        }
        __syncwarp();

        const uint64_t tile_base = (uint64_t)tile * 32;
        const uint64_t idx = tile_base + lane;
        const bool valid = idx < p.n_strings;
        uint64_t off = 0, end = 0;
        if (valid) { off = p.offsets[idx]; end = p.offsets[idx + 1]; }
        bool dead = !valid;
        if (valid && (end < off || end - off > (uint64_t)(M - 1))) {    // len must be <= M-1
            dead = true; end = off;
            kill_string(p, idx);
        }
        const uint32_t L = (uint32_t)(end - off);
        k.init();
        uint32_t cur[D], expect[D];                                     // cur = entry that led to the current state
#pragma unroll
        for (int d = 0; d < D; d++) {
            const uint32_t f = p.def[d].first_state;
            cur[d] = (f << 2) | (f << 8); expect[d] = 0;
        }
        auto make_ctx = [&](uint32_t tile_pos, uint32_t tile_s) {       // built from registers where the rare path is entered
            RowCtx<D, DirectTables> x;
            x.idx = idx; x.src = p.bytes + off; x.len = L; x.tile_pos = tile_pos; x.tile_s = tile_s;
#pragma unroll
            for (int d = 0; d < D; d++) { x.tb[d].tab = tab + d * DTAB_BYTES; x.ep_s[d] = ep_s[d]; }
            x.qbase = p.queue + ((size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * queue_words(D)) * 32 + lane;
            return x;
        };

        // staging geometry
        const uint32_t shift = (uint32_t)(off & 15);
        const bool any_shift = __any_sync(0xffffffffu, shift != 0);
        const uint8_t* in_ptr[4];                                       // vector kv of chunk 0 of row r0 + 8*i
        uint32_t in_left[4];                                            // bytes from there to the end of that string
#pragma unroll
        for (int i = 0; i < 4; i++) {
            const int row = r0 + 8 * i;
            const uint64_t roff = __shfl_sync(0xffffffffu, off, row);
            const uint64_t rend = __shfl_sync(0xffffffffu, end, row);
            const uint64_t a = (roff & ~uint64_t(15)) + (uint32_t)kv * 16;
            in_ptr[i] = p.bytes + a;
            in_left[i] = rend > a ? (uint32_t)(rend - a) : 0u;
        }
        const uint64_t tail_a = (off & ~uint64_t(15)) + 64;             // fifth vector of this lane's own row (unaligned strings)
        const uint8_t* const my_tail = p.bytes + tail_a;
        const uint32_t my_tail_left = end > tail_a ? (uint32_t)(end - tail_a) : 0u;
        const uint32_t rows_here = (p.n_strings - tile_base < 32) ? (uint32_t)(p.n_strings - tile_base) : 32u;

        auto stage = [&](uint32_t chunk) {   // async copy of chunk `chunk` into input tile (chunk & 1)
            const uint32_t cbase = chunk * DCH;
            const uint32_t dst = in_s + (chunk & 1) * (32 * DPITCH) + kv * 16;
#pragma unroll
            for (int i = 0; i < 4; i++) cp_async16(dst + (r0 + 8 * i) * DPITCH, in_ptr[i] + cbase, cbase < in_left[i] ? 16u : 0u);
            if (any_shift) cp_async16(in_s + (chunk & 1) * (32 * DPITCH) + lane * DPITCH + 64, my_tail + cbase, cbase < my_tail_left ? 16u : 0u);
            cp_async_commit();
        };
        stage(0);

#pragma unroll 1
        for (uint32_t chunk = 0; chunk < n_chunks; chunk++) {
            const uint32_t cbase = chunk * DCH;
            if (chunk + 1 < n_chunks) { stage(chunk + 1); cp_async_wait<1>(); } else cp_async_wait<0>();
            __syncwarp();

            // walk this lane's string over [cbase, cbase + DCH)
            const uint32_t my_in = in_s + (chunk & 1) * (32 * DPITCH) + lane * DPITCH;
            const uint32_t my_st = smem_u32(st_tile) + lane * DPITCH;
            uint32_t rare_bits = 0;                                     // words of this chunk with a rare row (hot path only)
            uint32_t hot_words = 0;                                     // words of this chunk walked by the hot path
            auto replay = [&]() {   // out-of-line exact replay; only copies escape, cur/expect stay in registers
                if (!dead) {
                    uint32_t tx[D], ts[D];
#pragma unroll
                    for (int d = 0; d < D; d++) { tx[d] = expect[d]; ts[d] = expect[d]; }
                    uint32_t tb = rare_bits, tn = 0xFFFFFFFFu;
                    uint32_t* const qb = p.queue + ((size_t)(blockIdx.x * (blockDim.x >> 5) + warp) * queue_words(D)) * 32 + lane;
                    for (;;) {
                        const uint32_t rc = chunk_rare<D>(k.base, qb, &tb, &tn, hot_words, cbase, my_in + shift, my_st, tab_s, ts, tx);
                        if (rc == 1) { dead = true; kill_string(p, idx); }
                        if (rc != 2) break;
                        const RowCtx<D, DirectTables> x = make_ctx(cbase, my_in + shift);
                        drain_now<D>(p, k, x);
                    }
#pragma unroll
                    for (int d = 0; d < D; d++) expect[d] = tx[d];
                }
                rare_bits = 0;
            };
#pragma unroll 1
            for (int g = 0; g < DCH / 16; g++) {
                const uint32_t gbase = cbase + g * 16;
                if (gbase >= Mpad) break;
                uint32_t w[4];
                if (!any_shift) {
                    asm volatile("ld.shared.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(w[0]), "=r"(w[1]), "=r"(w[2]), "=r"(w[3]) : "r"(my_in + g * 16));
                } else {
                    const uint32_t q = my_in + g * 16 + (shift & ~3u);
                    const uint32_t sh = (shift & 3u) * 8;
                    const uint32_t x0 = lds32(q), x1 = lds32(q + 4), x2 = lds32(q + 8), x3 = lds32(q + 12), x4 = lds32(q + 16);
                    w[0] = __funnelshift_r(x0, x1, sh); w[1] = __funnelshift_r(x1, x2, sh);
                    w[2] = __funnelshift_r(x2, x3, sh); w[3] = __funnelshift_r(x3, x4, sh);
                }
                uint32_t pk[D][4];
                if (gbase + 16 <= L && !dead) {
                    // ---- hot path: 16 real characters.  Per byte: PRMT+IMAD (byte*4 + table base, off the chain), PRMT -> IADD -> LDS
                    // (the chain), one shared atomic (multiplicity bin), PRMT (state byte into the pack), LOP3 (rare accumulate);
                    // the rare test runs once per 4 bytes and only records a bit.
#pragma unroll
                    for (int q = 0; q < 4; q++) {
                        uint32_t acc = 0;
#pragma unroll
                        for (int j = 0; j < 4; j++) {
#pragma unroll
                            for (int d = 0; d < D; d++) {
                                const uint32_t c = prmt(w[q], 0u, 0x4440u + j);
                                const uint32_t cb = tab_s + d * DTAB_BYTES + (c << 2);                   // off the chain
                                const uint32_t addr = prmt(cur[d], w[q], 0xBB40u + (j << 4)) + cb;       // base + c*260 + s*4
                                const uint32_t e = lds32(addr);
                                if (HIST) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr + hist_off) : "memory");
                                const uint32_t sel = j == 0 ? 0x3215u : j == 1 ? 0x3250u : j == 2 ? 0x3510u : 0x5210u;
                                pk[d][q] = prmt(pk[d][q], cur[d], sel);                                 // state byte into the output pack
                                acc |= e ^ expect[d];
                                cur[d] = e;
                            }
                        }
                        if (acc & DE_RARE_MASK) rare_bits |= 1u << (g * 4 + q);   // replayed once per chunk (chunk_rare)
                    }
                    hot_words = g * 4 + 4;
                } else {
                    // ---- ragged end: characters, then the final-state row, then dummy rows -------------------------
                    if (rare_bits) replay();   // keep the queued rows in position order
#pragma unroll
                    for (int d = 0; d < D; d++) pk[d][0] = pk[d][1] = pk[d][2] = pk[d][3] = 0;
                    uint32_t v0 = w[0], v1 = w[1], v2 = w[2], v3 = w[3];   // shifted along: no dynamic register indexing
#pragma unroll 1
                    for (int b = 0; b < 16; b++) {
                        const uint32_t pos = gbase + b;
                        const uint32_t c = v0 & 0xFFu;
                        v0 = __funnelshift_r(v0, v1, 8); v1 = __funnelshift_r(v1, v2, 8); v2 = __funnelshift_r(v2, v3, 8); v3 >>= 8;
                        uint32_t stb[D];
                        if (pos < L && !dead) {
                            uint32_t nxt[D];
                            uint32_t rare = 0, inval = 0;
#pragma unroll
                            for (int d = 0; d < D; d++) {
                                const uint32_t addr = tab_s + d * DTAB_BYTES + (cur[d] & 0xFFu) + c * DROW;
                                nxt[d] = lds32(addr);
                                if (HIST) asm volatile("red.shared.add.u32 [%0], 1;" ::"r"(addr + hist_off) : "memory");
                                stb[d] = (cur[d] >> 8) & 0xFFu;
                                rare |= (nxt[d] ^ expect[d]) & DE_RARE_MASK;
                                inval |= nxt[d] & ENT_INVALID;
                            }
                            if (inval) { dead = true; kill_string(p, idx); }
                            else if (rare) {
                                const RowCtx<D, DirectTables> x = make_ctx(cbase, my_in + shift);
                                uint32_t te[D], ts[D], tx[D];
#pragma unroll
                                for (int d = 0; d < D; d++) { te[d] = nxt[d]; ts[d] = stb[d]; tx[d] = expect[d]; }
                                rare_row_now<D>(p, k, x, pos, te, ts, tx);
#pragma unroll
                                for (int d = 0; d < D; d++) expect[d] = tx[d];
                            }
                            if (!dead) {
#pragma unroll
                                for (int d = 0; d < D; d++) cur[d] = nxt[d];
                            }
                        } else {
#pragma unroll
                            for (int d = 0; d < D; d++) stb[d] = (pos <= L) ? ((cur[d] >> 8) & 0xFFu) : p.def[d].num_states;   // final state, then dummy
                            if (pos == L && !dead) {
                                const RowCtx<D, DirectTables> x = make_ctx(cbase, my_in + shift);
                                uint32_t fs[D];
#pragma unroll
                                for (int d = 0; d < D; d++) fs[d] = (cur[d] >> 8) & 0xFFu;
                                finish_string<D, 32, DirectTables>(p, k, x, fs);
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
            if (__any_sync(0xffffffffu, rare_bits != 0)) replay();   // lockstep: costs the longest lane, not the sum over lanes

            // store the state tile (coalesced 16-byte vectors)
            if (cbase + kv * 16 < Mpad) {
#pragma unroll
                for (int d = 0; d < D; d++) {
                    if (!p.def[d].states) continue;
                    uint8_t* base = reinterpret_cast<uint8_t*>(p.def[d].states) + (tile_base + r0) * rp + cbase + kv * 16;
#pragma unroll
                    for (int i = 0; i < 4; i++) {
                        if (r0 + 8 * i < (int)rows_here)
                            *reinterpret_cast<uint4*>(base + (uint64_t)(8 * i) * rp) = *reinterpret_cast<const uint4*>(st_tile + (d * 32 + r0 + 8 * i) * DPITCH + kv * 16);
                    }
                }
            }
            __syncwarp();
        }

        // per-tile counters: rows with enable = 0 all look up table row 0 (src/lib.rs:218-232 with enable = 0)
        cta_counters_tile(cc, valid && !dead, M - L, (k.f(CF_FLAGS) & B2R_ST_OVERLAP) != 0);
        tile = tile_after;
    }

    __syncthreads();
    cta_counters_flush<D>(p, ep_s, cc);

    // ---- flush the multiplicity bins: bin (c,s) of def d -> dense global histogram [c*S + s] -----------------------------
    if (HIST) {
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
