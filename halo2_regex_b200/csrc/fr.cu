// SURVEY 8(f) rank 1, the step right after the path: every witness value the path produces is placed in a halo2 cell as
// `Value::known(F::from(x as u64))` (reference src/lib.rs:342-347, 388-418), F = halo2curves::bn256::Fr in the reference's own
// circuits (src/lib.rs:896, 1079).  These kernels turn whole witness columns (u8 / u16 / bitmap / u64) into arrays of Fr in the
// in-memory layout of that type — four little-endian u64 limbs in Montgomery form, value * 2^256 mod r — so that the host shim
// hands `assign_advice` ready-made field elements instead of converting cell by cell.
//
// halo2curves (pinned only through halo2-base rev 9860acc, Cargo.toml:12-15; its source is not under /root/reference) defines
// `From<u64> for Fr` as `Fr([v, 0, 0, 0]) * R2`, i.e. one Montgomery multiplication by R^2 mod r.  fr_from_u64 below is that
// multiplication (CIOS, one non-zero limb) with the published constants of the BN254 scalar field; tests/test_fr_feed.py checks
// the constants and the outputs against a Python big-integer model (v * 2^256 mod r).
//
// HBM-bound: 1-8 bytes read, 32 bytes written per cell.  u8 and bitmap columns go through a 256-entry table of Fr in shared
// memory (8 KB, built per CTA by the same fr_from_u64); wider values are converted on the fly.
#include <cuda_runtime.h>

#include <cstdint>

#include "config.hpp"

namespace b2r {

// BN254 scalar field: r, R^2 mod r (R = 2^256), -r^{-1} mod 2^64
__device__ __constant__ uint64_t FR_MODULUS[4] = {0x43e1f593f0000001ull, 0x2833e84879b97091ull, 0xb85045b68181585dull, 0x30644e72e131a029ull};
__device__ __constant__ uint64_t FR_R2[4] = {0x1bb8e645ae216da7ull, 0x53fe3ab1e35c59e3ull, 0x8c49833d53bb8085ull, 0x0216d0b17f4e44a5ull};
constexpr uint64_t FR_INV = 0xc2e1f593efffffffull;

// v * R mod r: Montgomery product of [v,0,0,0] and R2
__device__ __forceinline__ void fr_from_u64(uint64_t v, uint64_t out[4]) {
    typedef unsigned __int128 u128;
    uint64_t t[5];
    u128 c = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) { c += (u128)v * FR_R2[i]; t[i] = (uint64_t)c; c >>= 64; }
    t[4] = (uint64_t)c;
#pragma unroll
    for (int round = 0; round < 4; round++) {                             // t = (t + m * r) / 2^64
        const uint64_t m = t[0] * FR_INV;
        c = (u128)m * FR_MODULUS[0] + t[0];
        c >>= 64;
#pragma unroll
        for (int i = 1; i < 4; i++) { c += (u128)m * FR_MODULUS[i] + t[i]; t[i - 1] = (uint64_t)c; c >>= 64; }
        c += t[4];
        t[3] = (uint64_t)c;
        t[4] = (uint64_t)(c >> 64);
    }
    // t < 2r: one conditional subtraction
    uint64_t d[4];
    u128 b = 0;
#pragma unroll
    for (int i = 0; i < 4; i++) {
        const u128 x = (u128)t[i] - FR_MODULUS[i] - (uint64_t)b;
        d[i] = (uint64_t)x;
        b = (x >> 64) & 1;
    }
    const bool ge = t[4] != 0 || b == 0;
#pragma unroll
    for (int i = 0; i < 4; i++) out[i] = ge ? d[i] : t[i];
}

__device__ __forceinline__ void store_fr(uint64_t* dst, const uint64_t v[4]) {
    ulonglong2* q = reinterpret_cast<ulonglong2*>(dst);
    q[0] = make_ulonglong2(v[0], v[1]);
    q[1] = make_ulonglong2(v[2], v[3]);
}

struct FrParams {
    const void* col;          // kind U8/U16/U64/BITMAP: the column; CHARS/ENABLE: the input bytes
    const uint64_t* offsets;  // CHARS/ENABLE
    uint64_t n_strings, rows, pitch;
    uint32_t kind;
    uint64_t* out;            // n_strings * rows cells of 4 limbs
};

constexpr int FR_THREADS = 256;

__global__ void __launch_bounds__(FR_THREADS) column_to_fr_kernel(const __grid_constant__ FrParams p) {
    __shared__ __align__(16) uint64_t table[256][4];
    const bool tabled = p.kind == B2R_COL_U8 || p.kind == B2R_COL_BITMAP || p.kind == B2R_COL_CHARS || p.kind == B2R_COL_ENABLE;
    if (tabled) {
        for (uint32_t v = threadIdx.x; v < 256; v += blockDim.x) fr_from_u64(v, table[v]);
        __syncthreads();
    }
    const uint64_t total = p.n_strings * p.rows;
    for (uint64_t cell = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; cell < total; cell += (uint64_t)gridDim.x * blockDim.x) {
        const uint64_t j = cell / p.rows, i = cell - j * p.rows;
        uint64_t v = 0;
        switch (p.kind) {
            case B2R_COL_U8: v = static_cast<const uint8_t*>(p.col)[j * p.pitch + i]; break;
            case B2R_COL_U16: v = static_cast<const uint16_t*>(p.col)[j * p.pitch + i]; break;
            case B2R_COL_U64: v = static_cast<const uint64_t*>(p.col)[j * p.pitch + i]; break;
            case B2R_COL_BITMAP: v = (static_cast<const uint8_t*>(p.col)[j * p.pitch + (i >> 3)] >> (i & 7)) & 1u; break;
            case B2R_COL_CHARS: {       // character_values: the byte for i < len, 0 after (src/lib.rs:341-348)
                const uint64_t off = p.offsets[j], len = p.offsets[j + 1] - off;
                v = i < len ? static_cast<const uint8_t*>(p.col)[off + i] : 0u;
                break;
            }
            case B2R_COL_ENABLE: {      // enable_values: 1 for i < len, 0 after
                const uint64_t len = p.offsets[j + 1] - p.offsets[j];
                v = i < len ? 1u : 0u;
                break;
            }
        }
        if (tabled) {
            const ulonglong2* t = reinterpret_cast<const ulonglong2*>(table[v]);
            ulonglong2* q = reinterpret_cast<ulonglong2*>(p.out + cell * 4);
            q[0] = t[0];
            q[1] = t[1];
        } else {
            uint64_t f[4];
            fr_from_u64(v, f);
            store_fr(p.out + cell * 4, f);
        }
    }
}

}  // namespace b2r

using namespace b2r;

extern "C" {

int b2r_column_to_fr(b2r_config* c, const void* d_col, uint32_t kind, const uint64_t* d_offsets, uint64_t n_strings, uint64_t rows, uint64_t pitch,
                     uint64_t* d_fr, void* cuda_stream) {
    if (!c || !d_fr || (!d_col && n_strings != 0 && rows != 0 && kind != B2R_COL_ENABLE)) { set_error("null argument"); return B2R_ERR_INVALID_ARG; }
    if (c->device < 0) { set_error("this handle was created without a device (device = -1): no CPU fallback exists"); return B2R_ERR_CUDA; }
    if (kind < B2R_COL_U8 || kind > B2R_COL_ENABLE) { set_error("unknown column kind %u", kind); return B2R_ERR_INVALID_ARG; }
    if ((kind == B2R_COL_CHARS || kind == B2R_COL_ENABLE) && !d_offsets && n_strings) { set_error("offsets are required for this column kind"); return B2R_ERR_INVALID_ARG; }
    if (kind != B2R_COL_CHARS && kind != B2R_COL_ENABLE && pitch < (kind == B2R_COL_BITMAP ? (rows + 7) / 8 : rows)) { set_error("pitch smaller than a row"); return B2R_ERR_INVALID_ARG; }
    if (reinterpret_cast<uintptr_t>(d_fr) & 15) { set_error("the Fr array must be 16-byte aligned"); return B2R_ERR_ALIGNMENT; }
    const uint64_t total = n_strings * rows;
    if (!total) return B2R_OK;
    DeviceGuard g(c->device);
    if (!g.ok) { set_error("cudaSetDevice(%d) failed", c->device); return B2R_ERR_CUDA; }
    int n_sm = 148, max_smem = 0;
    device_limits(&n_sm, &max_smem);
    FrParams p;
    p.col = d_col; p.offsets = d_offsets; p.n_strings = n_strings; p.rows = rows; p.pitch = pitch; p.kind = kind; p.out = d_fr;
    const uint64_t want = (total + FR_THREADS - 1) / FR_THREADS;
    const unsigned grid = (unsigned)(want < (uint64_t)n_sm * 8 ? want : (uint64_t)n_sm * 8);   // 8 resident CTAs of 256 threads per SM
    column_to_fr_kernel<<<grid, FR_THREADS, 0, (cudaStream_t)cuda_stream>>>(p);
    CUDA_TRY(cudaGetLastError());
    return B2R_OK;
}

}  // extern "C"
