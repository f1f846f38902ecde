// Instantiates walk_kernel for one number of regex defs (compiled once per D with -DB2R_INST_D=<D>, in parallel).
#include "walk.cuh"

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

}  // namespace b2r
