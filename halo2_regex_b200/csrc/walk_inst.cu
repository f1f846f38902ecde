// Instantiates walk_kernel for one number of regex defs (compiled once per D with -DB2R_INST_D=<D>, in parallel).
#include "walk.cuh"
#include "walk_direct.cuh"

#ifndef B2R_INST_D
#error "compile with -DB2R_INST_D=<1..4>"
#endif

namespace b2r {

template <>
int launch_walk_d<B2R_INST_D>(const WalkParams& p, bool wide, bool ts, bool hs, size_t smem, int grid, cudaStream_t st) {
    constexpr int D = B2R_INST_D;
    if (wide) {
        if (ts && hs) return launch_one<D, uint16_t, true, true, WALK_WARPS>(p, smem, grid, st);
        if (ts) return launch_one<D, uint16_t, true, false, WALK_WARPS>(p, smem, grid, st);
        return launch_one<D, uint16_t, false, false, WALK_WARPS>(p, smem, grid, st);
    }
    if (ts && hs) return launch_one<D, uint8_t, true, true, WALK_WARPS>(p, smem, grid, st);
    if (ts) return launch_one<D, uint8_t, true, false, WALK_WARPS>(p, smem, grid, st);
    return launch_one<D, uint8_t, false, false, WALK_WARPS>(p, smem, grid, st);
}

template <int D>
int launch_direct_d(const WalkParams& p, const uint32_t* tab, bool in_row, size_t smem, int grid, int block, cudaStream_t st);

#if B2R_INST_D <= 2
template <int D, bool HIST, bool IN_ROW>
static int launch_direct_one(const WalkParams& p, const uint32_t* tab, size_t smem, int grid, int block, cudaStream_t st) {
    auto kern = walk_direct_kernel<D, HIST, IN_ROW>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) { set_error("cudaFuncSetAttribute: %s", cudaGetErrorString(e)); return B2R_ERR_CUDA; }
    kern<<<grid, block, smem, st>>>(p, tab);
    e = cudaGetLastError();
    if (e != cudaSuccess) { set_error("walk_direct_kernel launch: %s", cudaGetErrorString(e)); return B2R_ERR_CUDA; }
    return B2R_OK;
}
template <>
int launch_direct_d<B2R_INST_D>(const WalkParams& p, const uint32_t* tab, bool in_row, size_t smem, int grid, int block, cudaStream_t st) {
    constexpr int D = B2R_INST_D;
    if (!p.want_hist) return launch_direct_one<D, false, true>(p, tab, smem, grid, block, st);
    return in_row ? launch_direct_one<D, true, true>(p, tab, smem, grid, block, st) : launch_direct_one<D, true, false>(p, tab, smem, grid, block, st);
}
#endif

}  // namespace b2r
