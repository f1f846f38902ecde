// Does store traffic share the SM's shared-memory pipe?  (development aid for the fused kernel's bottleneck analysis)
// Every warp runs ITER iterations of: NL conflict-free LDS.32 (a 32-lane wavefront each) + NS store instructions of a given
// flavour (0 none, 1 st.v4 contiguous 512 B/warp, 2 st.v4 scattered over 16 rows x 32 B like the state column,
// 3 one 4 KB TMA bulk store per 8 iterations from a shared zero buffer).  Timing of LDS-only, store-only and both tells
// whether the two add up (shared pipe) or overlap.
#include <cuda_runtime.h>
#include <cstdint>
#include <cstdio>

template <int MODE>
__global__ void __launch_bounds__(512, 1) mix(uint4* out, size_t out_vecs, int iters, int nl, int ns, uint32_t* sink) {
    extern __shared__ __align__(128) unsigned char sm[];
    uint32_t* tab = reinterpret_cast<uint32_t*>(sm);                   // 64 KB table, conflict-free column per lane
    for (int i = threadIdx.x; i < 16384; i += blockDim.x) tab[i] = (uint32_t)((i * 37 + 11) & 511) * 128;
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) reinterpret_cast<uint32_t*>(sm + 65536)[i] = 0;   // 4 KB zeros
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
    __syncthreads();
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    const size_t warps = (size_t)gridDim.x * 16, gw = (size_t)blockIdx.x * 16 + warp;
    const uint32_t tab_s = (uint32_t)__cvta_generic_to_shared(tab) + lane * 4, zero_s = (uint32_t)__cvta_generic_to_shared(sm + 65536);
    uint32_t a = lane * 128, acc = 0;
    size_t v = gw * 32 + lane;                                          // my vector index in the output stream
    const uint4 z = make_uint4(0, 0, 0, 0);
    for (int it = 0; it < iters; it++) {
        for (int k = 0; k < nl; k++) {                                  // dependent chain of conflict-free lookups (like the walk)
            uint32_t e;
            asm volatile("ld.shared.u32 %0, [%1];" : "=r"(e) : "r"(tab_s + (a & 0xFF80u)));
            a = e + k; acc ^= e;
        }
        if (MODE == 1) {
            for (int k = 0; k < ns; k++) { if (v < out_vecs) out[v] = z; v += warps * 32; }
        } else if (MODE == 2) {                                         // 16 rows of 1056 B apart, 2 lanes per row
            for (int k = 0; k < ns; k++) {
                const size_t base = (gw * iters * ns + (size_t)it * ns + k) % (out_vecs / 66 / 16) * 66 * 16;
                const size_t idx = base + (size_t)(lane >> 1) * 66 + (lane & 1) + (size_t)(it & 31) * 2;
                if (idx < out_vecs) out[idx] = z;
            }
        } else if (MODE == 3) {
            if ((it & 7) == 0 && lane < ns) {                           // ns lanes issue one 4 KB op each, every 8 iterations
                const size_t o = ((gw * (iters / 8) + it / 8) * ns + lane) * 256 % (out_vecs - 256);
                asm volatile("cp.async.bulk.global.shared::cta.bulk_group [%0], [%1], 4096;" ::"l"(out + o), "r"(zero_s) : "memory");
                asm volatile("cp.async.bulk.commit_group;" ::: "memory");
            }
        }
    }
    if (MODE == 3) asm volatile("cp.async.bulk.wait_group 0;" ::: "memory");
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
    const size_t bytes = (size_t)4 << 30;
    uint4* out; uint32_t* sink;
    cudaMalloc(&out, bytes); cudaMalloc(&sink, 4);
    const size_t vecs = bytes / 16;
    const int smem = 65536 + 4096, iters = 2048;
    cudaFuncSetAttribute(mix<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mix<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mix<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(mix<3>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int nl = 64;
    printf("148 CTAs x 16 warps, %d iterations, %d LDS wavefronts per warp-iteration\n", iters, nl);
    const float t_l = time_ms([&] { mix<0><<<148, 512, smem>>>(out, vecs, iters, nl, 0, sink); });
    printf("LDS only                                   %.3f ms\n", t_l);
    for (int ns : {2, 4, 8}) {
        const double gb = 148.0 * 16 * iters * ns * 512 / 1e9;
        const float s1 = time_ms([&] { mix<1><<<148, 512, smem>>>(out, vecs, iters, 0, ns, sink); });
        const float b1 = time_ms([&] { mix<1><<<148, 512, smem>>>(out, vecs, iters, nl, ns, sink); });
        const float s2 = time_ms([&] { mix<2><<<148, 512, smem>>>(out, vecs, iters, 0, ns, sink); });
        const float b2 = time_ms([&] { mix<2><<<148, 512, smem>>>(out, vecs, iters, nl, ns, sink); });
        printf("%d st.v4/iter (%.2f GB): contiguous alone %.3f, with LDS %.3f (sum %.3f) | scattered alone %.3f, with LDS %.3f (sum %.3f)\n",
               ns, gb, s1, b1, s1 + t_l, s2, b2, s2 + t_l);
    }
    for (int ns : {1, 2, 4}) {
        const double gb = 148.0 * 16 * (iters / 8) * ns * 4096 / 1e9;
        const float s3 = time_ms([&] { mix<3><<<148, 512, smem>>>(out, vecs, iters, 0, ns, sink); });
        const float b3 = time_ms([&] { mix<3><<<148, 512, smem>>>(out, vecs, iters, nl, ns, sink); });
        printf("%d x 4 KB TMA ops / 8 iters (%.2f GB): alone %.3f, with LDS %.3f (sum %.3f)\n", ns, gb, s3, b3, s3 + t_l);
    }
    cudaDeviceSynchronize();
    printf("%s\n", cudaGetErrorString(cudaGetLastError()));
    return 0;
}
