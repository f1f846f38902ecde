// Write-bandwidth microbenchmark (development aid): what does a pure write stream reach on this B200, and with which
// store flavour?  The emit stage is ~100 % writes, so this is its real ceiling.
//   nvcc -O3 -std=c++17 -gencode arch=compute_100a,code=sm_100a -o wbench tools/wbench.cu && ./wbench
#include <cuda_runtime.h>

#include <cstdint>
#include <cstdio>
#include <vector>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)

// each warp owns contiguous blocks of `blk` bytes (like a tile's rows of one column), blocks strided over all warps
template <int MODE>
__global__ void __launch_bounds__(256) wr_blocks(uint4* dst, size_t n_vec, size_t blk_vec) {
    const size_t warps = (size_t)gridDim.x * (blockDim.x >> 5);
    const size_t w = (size_t)blockIdx.x * (blockDim.x >> 5) + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t b = w * blk_vec; b < n_vec; b += warps * blk_vec) {
        for (size_t v = lane; v < blk_vec && b + v < n_vec; v += 32) {
            if (MODE == 0) dst[b + v] = z;
            else if (MODE == 1) __stcs(dst + b + v, z);
            else if (MODE == 2) __stcg(dst + b + v, z);
            else asm volatile("st.global.L1::no_allocate.v4.u32 [%0], {%1,%2,%3,%4};" ::"l"(dst + b + v), "r"(0), "r"(0), "r"(0), "r"(0) : "memory");
        }
    }
}

// plain grid-stride (every warp instruction writes 512 contiguous bytes, consecutive warps adjacent)
__global__ void __launch_bounds__(256) wr_stride(uint4* dst, size_t n_vec) {
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) dst[i] = z;
}

// TMA bulk stores from a zeroed shared buffer: one thread per CTA issues `chunk`-byte copies
__global__ void __launch_bounds__(128) wr_bulk(unsigned char* dst, size_t bytes, uint32_t chunk) {
    extern __shared__ __align__(128) unsigned char zsm[];
    for (uint32_t i = threadIdx.x * 16; i < chunk; i += blockDim.x * 16) *reinterpret_cast<uint4*>(zsm + i) = make_uint4(0, 0, 0, 0);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    if (threadIdx.x == 0) {
        const uint32_t s = (uint32_t)__cvta_generic_to_shared(zsm);
        int inflight = 0;
        for (size_t o = (size_t)blockIdx.x * chunk; o + chunk <= bytes; o += (size_t)gridDim.x * chunk) {
            asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], %2;" ::"l"(dst + o), "r"(s), "r"(chunk) : "memory");
            asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            if (++inflight >= 8) { asm volatile("cp.async.bulk.wait_group.read 4;" ::: "memory"); inflight = 4; }
        }
        asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
    }
}

__global__ void __launch_bounds__(256) copy_stride(uint4* dst, const uint4* src, size_t n_vec) {
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) dst[i] = src[i];
}
__global__ void __launch_bounds__(256) read_stride(const uint4* src, size_t n_vec, uint32_t* sink) {
    uint32_t acc = 0;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n_vec; i += (size_t)gridDim.x * blockDim.x) { const uint4 v = src[i]; acc ^= v.x ^ v.y ^ v.z ^ v.w; }
    if (acc == 0x12345678u) *sink = acc;
}

template <typename F>
static float time_ms(F f, int iters = 5) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0); cudaEventCreate(&e1);
    f(); cudaDeviceSynchronize();
    float best = 1e9f;
    for (int i = 0; i < iters; i++) {
        cudaEventRecord(e0); f(); cudaEventRecord(e1); cudaEventSynchronize(e1);
        float ms; cudaEventElapsedTime(&ms, e0, e1);
        if (ms < best) best = ms;
    }
    return best;
}

int main() {
    const size_t bytes = (size_t)3584 << 20;   // 3.5 GiB, the sparse columns of config 1
    unsigned char *a, *b;
    uint32_t* sink;
    CK(cudaMalloc(&a, bytes)); CK(cudaMalloc(&b, bytes)); CK(cudaMalloc(&sink, 4));
    CK(cudaMemset(a, 1, bytes)); CK(cudaMemset(b, 2, bytes));
    const size_t n_vec = bytes / 16;
    auto report = [&](const char* name, float ms, double moved) { printf("%-44s %8.3f ms  %8.1f GB/s\n", name, ms, moved / ms / 1e6); };
    report("cudaMemsetAsync", time_ms([&] { cudaMemsetAsync(a, 0, bytes); }), (double)bytes);
    for (int grid : {148 * 2, 148 * 4, 148 * 8, 148 * 16}) {
        char nm[96];
        snprintf(nm, sizeof nm, "st.v4 grid-stride, grid %d x 256", grid);
        report(nm, time_ms([&] { wr_stride<<<grid, 256>>>((uint4*)a, n_vec); }), (double)bytes);
    }
    for (size_t blk : {1056ull * 32, 4096ull, 65536ull}) {
        char nm[96];
        snprintf(nm, sizeof nm, "st.v4 warp-blocks of %zu B, grid 1184", blk);
        report(nm, time_ms([&] { wr_blocks<0><<<1184, 256>>>((uint4*)a, n_vec, blk / 16); }), (double)bytes);
        snprintf(nm, sizeof nm, "st.cs.v4 warp-blocks of %zu B", blk);
        report(nm, time_ms([&] { wr_blocks<1><<<1184, 256>>>((uint4*)a, n_vec, blk / 16); }), (double)bytes);
        snprintf(nm, sizeof nm, "st.cg.v4 warp-blocks of %zu B", blk);
        report(nm, time_ms([&] { wr_blocks<2><<<1184, 256>>>((uint4*)a, n_vec, blk / 16); }), (double)bytes);
        snprintf(nm, sizeof nm, "st.no_allocate.v4 warp-blocks of %zu B", blk);
        report(nm, time_ms([&] { wr_blocks<3><<<1184, 256>>>((uint4*)a, n_vec, blk / 16); }), (double)bytes);
    }
    for (uint32_t chunk : {4096u, 16384u, 32768u}) {
        for (int per_sm : {1, 4}) {
            char nm[96];
            snprintf(nm, sizeof nm, "TMA bulk store, %u B chunks, %d CTA/SM", chunk, per_sm);
            cudaFuncSetAttribute(wr_bulk, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)chunk);
            report(nm, time_ms([&] { wr_bulk<<<148 * per_sm, 128, chunk>>>(a, bytes, chunk); }), (double)bytes);
        }
    }
    report("copy (read + write), grid 1184", time_ms([&] { copy_stride<<<1184, 256>>>((uint4*)a, (const uint4*)b, n_vec); }), 2.0 * bytes);
    report("cudaMemcpyAsync D2D (read + write)", time_ms([&] { cudaMemcpyAsync(a, b, bytes, cudaMemcpyDeviceToDevice); }), 2.0 * bytes);
    report("read only, grid 1184", time_ms([&] { read_stride<<<1184, 256>>>((const uint4*)b, n_vec, sink); }), (double)bytes);
    // 1 part read : 4 parts write, the mix of the whole path
    report("read 0.875 GiB + write 3.5 GiB concurrently", time_ms([&] {
               cudaStream_t s2; cudaStreamCreate(&s2);
               read_stride<<<296, 256, 0, s2>>>((const uint4*)b, n_vec / 4, sink);
               wr_stride<<<888, 256>>>((uint4*)a, n_vec);
               cudaStreamSynchronize(s2); cudaStreamDestroy(s2);
           }), 1.25 * bytes);
    CK(cudaDeviceSynchronize());
    CK(cudaGetLastError());
    return 0;
}
