// Small kernels (multiplicity finalisation, failure diagnosis) and the walk_kernel launch dispatch.
// The hot kernel itself lives in walk.cuh and is instantiated per number of defs in walk_inst.cu.
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>

#include "walk.cuh"
#include "walk_direct.cuh"

namespace b2r {

// ---- multiplicity finalisation: dense (byte,state) bins -> table rows (reference order, src/table.rs:101-122) ----
__global__ void finalize_kernel(const __grid_constant__ FinalizeParams p) {
    const uint32_t tid = blockIdx.x * blockDim.x + threadIdx.x;
    const uint32_t nth = gridDim.x * blockDim.x;
    for (uint32_t d = 0; d < p.n_defs; d++) {
        const auto& f = p.def[d];
        if (f.mult) {
            for (uint32_t r = tid; r < f.n_rows; r += nth) {
                const unsigned long long v = (r == 0) ? p.counters->pad_rows : f.hist[f.row_bin[r]];
                f.mult[r] = (p.accumulate ? f.mult[r] : 0ull) + v;
            }
        }
        if (f.endpoint_mult && tid == 0) {   // a handful of rows
            unsigned long long tot_s = 0, tot_e = 0;
            for (uint32_t r = 1; r < f.n_erows; r++) {
                const unsigned long long vs = f.erow_start_bin[r] != 0xFFFFFFFFu ? f.ep_start[f.erow_start_bin[r]] : 0ull;
                const unsigned long long ve = f.erow_end_bin[r] != 0xFFFFFFFFu ? f.ep_end[f.erow_end_bin[r]] : 0ull;
                tot_s += vs; tot_e += ve;
                f.endpoint_mult[r] = (p.accumulate ? f.endpoint_mult[r] : 0ull) + vs;
                f.endpoint_mult[f.n_erows + r] = (p.accumulate ? f.endpoint_mult[f.n_erows + r] : 0ull) + ve;
            }
            // every other row of the batch looks up (0, dummy, dummy) = endpoint row 0
            f.endpoint_mult[0] = (p.accumulate ? f.endpoint_mult[0] : 0ull) + (p.n_rows_total - tot_s);
            f.endpoint_mult[f.n_erows] = (p.accumulate ? f.endpoint_mult[f.n_erows] : 0ull) + (p.n_rows_total - tot_e);
        }
    }
}

// ---- diagnose: details of the failing string for the batch status (one thread) ------------------------------------
__global__ void diagnose_kernel(const __grid_constant__ WalkParams p, uint64_t j, b2r_batch_status* out) {
    b2r_batch_status r = diagnose_string(p, j);
    r.n_overlap_lo = (uint32_t)p.counters->n_overlap;
    *out = r;
}

// ---- launchers ---------------------------------------------------------------------------------------------------
template <typename ST>
static size_t per_warp_smem(int D) { return 32 * IN_PITCH + (size_t)D * 32 * StTile<ST>::PITCH; }

int walk_smem_bytes(const WalkParams& p, bool wide, int warps, bool smem_tables, bool smem_hist) {
    size_t n = 0;
    for (uint32_t d = 0; d < p.n_defs; d++) {
        if (smem_tables) n += (size_t)p.def[d].num_classes * p.def[d].num_states * 4 + 256;
        if (smem_hist) n += (size_t)256 * p.def[d].num_states * 4;
    }
    n = (n + 15) & ~size_t(15);
    n += sizeof(CtaCounters) + ((p.ep_smem_bytes + 15u) & ~15u);
    n += (size_t)warps * (wide ? per_warp_smem<uint16_t>(p.n_defs) : per_warp_smem<uint8_t>(p.n_defs));
    return (int)n;
}

template <> int launch_walk_d<1>(const WalkParams&, bool, bool, bool, size_t, int, cudaStream_t);
template <> int launch_walk_d<2>(const WalkParams&, bool, bool, bool, size_t, int, cudaStream_t);
template <> int launch_walk_d<3>(const WalkParams&, bool, bool, bool, size_t, int, cudaStream_t);
template <> int launch_walk_d<4>(const WalkParams&, bool, bool, bool, size_t, int, cudaStream_t);

int launch_walk(const WalkParams& p, bool wide, void* stream, WalkLaunch* chosen) {
    constexpr int WARPS = WALK_WARPS;
    int dev = 0, n_sm = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    // shared-memory budget: tables first, then the histogram; fall back to global (L2) when they do not fit
    bool ts = true, hs = true;
    const int budget = max_smem / 2 - 1024;   // keep two CTAs per SM when possible
    if (walk_smem_bytes(p, wide, WARPS, true, true) > budget) {
        if (walk_smem_bytes(p, wide, WARPS, true, true) <= max_smem - 1024) { /* one CTA per SM */ }
        else if (walk_smem_bytes(p, wide, WARPS, true, false) <= max_smem - 1024) hs = false;
        else { ts = false; hs = false; }
    }
    const size_t smem = (size_t)walk_smem_bytes(p, wide, WARPS, ts, hs);
    int per_sm = (int)((size_t)max_smem / (smem + 1024));
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 8) per_sm = 8;
    long long want = ((long long)p.n_tiles + WARPS - 1) / WARPS;
    int grid = n_sm * per_sm;
    if (grid > want) grid = (int)(want > 0 ? want : 1);
    if (chosen) { chosen->grid = grid; chosen->block = WARPS * 32; chosen->smem_bytes = smem; }
    cudaStream_t st = (cudaStream_t)stream;
    switch (p.n_defs) {
        case 1: return launch_walk_d<1>(p, wide, ts, hs, smem, grid, st);
        case 2: return launch_walk_d<2>(p, wide, ts, hs, smem, grid, st);
        case 3: return launch_walk_d<3>(p, wide, ts, hs, smem, grid, st);
        case 4: return launch_walk_d<4>(p, wide, ts, hs, smem, grid, st);
    }
    set_error("unsupported number of defs %u", p.n_defs);
    return B2R_ERR_UNSUPPORTED;
}

template <int D>
int launch_direct_d(const WalkParams& p, const uint32_t* tab, bool in_row, size_t smem, int grid, int block, cudaStream_t st);
template <> int launch_direct_d<1>(const WalkParams&, const uint32_t*, bool, size_t, int, int, cudaStream_t);
template <> int launch_direct_d<2>(const WalkParams&, const uint32_t*, bool, size_t, int, int, cudaStream_t);

int launch_walk_direct(const WalkParams& p, const uint32_t* d_direct_tab, uint32_t hist_off, void* stream, WalkLaunch* chosen) {
    int dev = 0, n_sm = 0, max_smem = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n_sm, cudaDevAttrMultiProcessorCount, dev);
    cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    const uint32_t D = p.n_defs;
    const bool in_row = hist_off == 128u;
    const bool bins = p.want_hist && !in_row;
    const size_t fixed = (size_t)D * DTAB_BYTES + (bins ? (size_t)D * DTAB_BYTES : 0) + ZERO_BYTES + sizeof(CtaCounters) + p.ep_smem_bytes;
    const size_t per_warp = (size_t)direct_tile_bytes_per_warp((int)D);
    int warps = (int)(((size_t)max_smem - fixed) / per_warp);
    if (warps > DIRECT_MAX_THREADS / 32) warps = DIRECT_MAX_THREADS / 32;
    if (warps < 1) { set_error("direct tables do not fit in shared memory"); return B2R_ERR_UNSUPPORTED; }
    // persistent: one CTA per SM, tiles strided over (CTA, warp); small batches use fewer CTAs
    long long ctas = ((long long)p.n_tiles + warps - 1) / warps;
    int grid = (int)(ctas < n_sm ? (ctas > 0 ? ctas : 1) : n_sm);
    const size_t smem = fixed + per_warp * warps;
    if (chosen) { chosen->grid = grid; chosen->block = warps * 32; chosen->smem_bytes = smem; }
    cudaStream_t st = (cudaStream_t)stream;
    if (D == 1) return launch_direct_d<1>(p, d_direct_tab, in_row, smem, grid, warps * 32, st);
    if (D == 2) return launch_direct_d<2>(p, d_direct_tab, in_row, smem, grid, warps * 32, st);
    set_error("direct kernel supports at most 2 defs");
    return B2R_ERR_UNSUPPORTED;
}

int launch_finalize(const FinalizeParams& p, void* stream) {
    finalize_kernel<<<32, 256, 0, (cudaStream_t)stream>>>(p);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("finalize_kernel launch: %s", cudaGetErrorString(e)); return B2R_ERR_CUDA; }
    return B2R_OK;
}

int launch_diagnose(const WalkParams& p, uint64_t string_idx, b2r_batch_status* d_out, void* stream) {
    diagnose_kernel<<<1, 1, 0, (cudaStream_t)stream>>>(p, string_idx, d_out);
    cudaError_t e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("diagnose_kernel launch: %s", cudaGetErrorString(e)); return B2R_ERR_CUDA; }
    return B2R_OK;
}

}  // namespace b2r
