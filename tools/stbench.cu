// State-column store pattern microbenchmark (development aid).  A warp owns a tile of 32 rows (pitch 1056 B) and fills the
// rows left to right, W bytes per row and step, with st.global.v4 — W = 32 is what walk_kernel does per 32-byte chunk
// (one 32-byte sector per row and instruction pair); larger W means the warp collects more of each row before storing.
// Between steps every warp runs a dependent chain of NL shared-memory lookups (the walk).  Question: does the DRAM side
// care whether a row arrives as 32-byte sectors, 64-byte halves or full 128-byte lines?
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o stbench tools/stbench.cu && ./stbench
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

constexpr int PITCH = 1056, ROWB = 1024;

// FILL: every warp also zero-fills 3 x 33 KB + 2 x 4 KB per tile (the sparse columns of walk_kernel) with 2 KB TMA bulk stores
// spread over the steps, two ops per lane.
template <int W, bool FILL>
__global__ void __launch_bounds__(512, 1) fill_rows(unsigned char* out, unsigned char* zcols, int tiles_per_warp, int nl, uint32_t* sink) {
    __shared__ uint32_t tab[8192];
    __shared__ __align__(128) unsigned char zero[2048];
    for (int i = threadIdx.x; i < 8192; i += blockDim.x) tab[i] = (uint32_t)((i * 37 + 11) & 255) * 128;
    for (int i = threadIdx.x; i < 512; i += blockDim.x) reinterpret_cast<uint32_t*>(zero)[i] = 0;
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const uint32_t zero_s = (uint32_t)__cvta_generic_to_shared(zero);
    constexpr int ZT = 3 * 32 * PITCH + 2 * 4224;   // zero bytes per tile (107.8 KB = 52.7 ops of 2 KB)
    constexpr int NOPS = (ZT + 2047) / 2048;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t gw = (size_t)blockIdx.x * 16 + warp, warps = (size_t)gridDim.x * 16;
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab) + lane * 4;
    uint32_t a = lane * 128, acc = 0;
    constexpr int LPR = W / 16;            // lanes per row
    constexpr int RPI = 32 / LPR;          // rows per store instruction
    const uint4 z = make_uint4(gw, 1, 2, 3);
    for (int t = 0; t < tiles_per_warp; t++) {
        unsigned char* tile = out + ((size_t)t * warps + gw) * 32 * PITCH;
        unsigned char* ztile = zcols + ((size_t)t * warps + gw) * ZT;
        for (int step = 0; step < ROWB / W; step++) {
            if (FILL) {   // op j (j % 32 == lane) goes out at step j * nsteps / NOPS
#pragma unroll
                for (int h = 0; h < 2; h++) {
                    const int j = lane + 32 * h;
                    if (j < NOPS && step == j * (ROWB / W) / NOPS) {
                        const uint32_t nb = j == NOPS - 1 ? (uint32_t)(ZT - j * 2048) : 2048u;
                        asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(ztile + (size_t)j * 2048), "r"(zero_s), "r"(nb) : "memory");
                    }
                }
            }
            for (int k = 0; k < nl * (W / 32); k++) {
                uint32_t e;
                asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(tab_s + (a & 0x7F80u)));
                a = e + k; acc ^= e;
            }
#pragma unroll
            for (int i = 0; i < 32 / RPI; i++) {
                const int row = lane / LPR + RPI * i;
                *reinterpret_cast<uint4*>(tile + (size_t)row * PITCH + step * W + (lane % LPR) * 16) = z;
            }
        }
    }
    if (FILL) { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }
    if (acc == 0x12345u) *sink = acc;
}

template <typename F>
static float time_ms(F f) {
    cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e9f;
    for (int i = 0; i < 3; i++) { cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1); float ms; cudaEventElapsedTime(&ms, e0, e1); if (ms < best) best = ms; }
    return best;
}

int main() {
    const int tpw = 14;                                  // 2^20 strings / 32 / (148 * 16 warps) = 13.8
    const size_t bytes = (size_t)148 * 16 * tpw * 32 * PITCH;
    const size_t zbytes = (size_t)148 * 16 * tpw * (3 * 32 * PITCH + 2 * 4224);
    unsigned char *out, *zc; uint32_t* sink;
    cudaMalloc(&out, bytes); cudaMalloc(&zc, zbytes + 4096); cudaMalloc(&sink, 4);
    const double gb = 148.0 * 16 * tpw * 32 * ROWB / 1e9;
    printf("%.2f GB of rows (pitch %d), 148 x 16 warps, %d tiles per warp; fill %.2f GB\n", gb, PITCH, tpw, zbytes / 1e9);
    for (int nl : {0, 8, 16, 24, 32}) {
        const float t32 = time_ms([&] { fill_rows<32, false><<<148, 512>>>(out, zc, tpw, nl, sink); });
        const float t128 = time_ms([&] { fill_rows<128, false><<<148, 512>>>(out, zc, tpw, nl, sink); });
        const float f32 = time_ms([&] { fill_rows<32, true><<<148, 512>>>(out, zc, tpw, nl, sink); });
        const float f64 = time_ms([&] { fill_rows<64, true><<<148, 512>>>(out, zc, tpw, nl, sink); });
        const float f128 = time_ms([&] { fill_rows<128, true><<<148, 512>>>(out, zc, tpw, nl, sink); });
        const float f512 = time_ms([&] { fill_rows<512, true><<<148, 512>>>(out, zc, tpw, nl, sink); });
        printf("chain of %2d LDS per 32 B: rows only W=32 %.3f W=128 %.3f ms | rows + TMA fill: W=32 %.3f | W=64 %.3f | W=128 %.3f | W=512 %.3f ms\n", nl, t32, t128, f32, f64, f128, f512);
    }
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
