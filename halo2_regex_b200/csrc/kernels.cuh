// Device-side parameter blocks shared by kernels.cu and api.cu.
#pragma once
#include <cstdint>

#include "../../include/b2r.h"

namespace b2r {

struct DefDev {
    const uint8_t* byte_class;       // [256]
    const uint32_t* trans;           // [num_classes][num_states] packed entries (defs.hpp)
    uint32_t num_states;             // S (= dummy state value)
    uint32_t num_classes;
    uint32_t first_state;
    uint32_t accepted_state;
    uint32_t sid_offset;
    uint32_t num_substrs;
    unsigned long long* hist;        // dense [256][S] multiplicity bins (global, u64)
    unsigned long long* ep_start;    // [num_substrs][S] start-endpoint counters
    unsigned long long* ep_end;      // [num_substrs][S]
    // outputs (may be null)
    void* states;
    uint8_t* substr_ids;
    uint8_t* start_enable;
    uint8_t* end_enable;
};

struct BatchCounters {               // zeroed (first_bad = ~0) before every batch
    unsigned long long first_bad;    // lowest string index with an invalid transition / too long
    unsigned long long n_overlap;
    unsigned long long pad_rows;     // sum over strings of (M - len): multiplicity of table row 0
    unsigned long long n_ok_strings; // strings that were walked to the end
    unsigned long long tile_counter; // next tile of 32 strings to hand out (dynamic scheduling of the persistent CTAs)
    unsigned long long reserved[3];
};

struct WalkParams {
    const uint8_t* bytes;
    const uint64_t* offsets;
    uint64_t n_strings;
    uint64_t total_bytes;
    uint64_t row_pitch, bitmap_pitch;
    uint32_t max_chars;              // M
    uint32_t n_defs;
    DefDev def[B2R_MAX_DEFS];
    uint8_t* masked_chars;
    uint8_t* masked_substr_ids;
    b2r_string_status* status;
    b2r_substr_record* records;
    uint8_t* compact_bytes;
    uint32_t max_records, compact_pitch;
    BatchCounters* counters;
    uint32_t n_tiles;                // ceil(n_strings / 32)
    uint32_t smem_tables;            // 1: class/transition tables staged in shared memory
    uint32_t smem_hist;              // 1: multiplicity bins accumulated in shared memory, flushed with global atomics
    uint32_t want_hist;              // 0: no multiplicity output was requested, skip the histogram
    uint32_t* queue;                 // global scratch for the per-lane rare-row queues: grid*block lanes x queue_words(D) words
    uint32_t ep_smem_bytes;          // bytes of shared memory for the endpoint counters (0: count with global atomics)
};

struct FinalizeParams {
    uint32_t n_defs;
    uint32_t accumulate;
    uint64_t n_rows_total;           // N*M, for the endpoint row-0 counts
    const BatchCounters* counters;
    struct {
        const unsigned long long* hist;
        const uint32_t* row_bin;     // [T]
        uint32_t n_rows;             // T
        const unsigned long long *ep_start, *ep_end;
        const uint32_t *erow_start_bin, *erow_end_bin;
        uint32_t n_erows;            // E
        unsigned long long* mult;    // out [T] or null
        unsigned long long* endpoint_mult;  // out [2E] or null
    } def[B2R_MAX_DEFS];
};

struct WalkLaunch {
    int grid, block;
    size_t smem_bytes;
};

// host-callable launchers (kernels.cu)
int launch_walk(const WalkParams& p, bool wide_states, void* stream, WalkLaunch* chosen);
int launch_finalize(const FinalizeParams& p, void* stream);
int launch_diagnose(const WalkParams& p, uint64_t string_idx, b2r_batch_status* d_out, void* stream);
int launch_walk_direct(const WalkParams& p, const uint32_t* d_direct_tab, uint32_t hist_off, void* stream, WalkLaunch* chosen);
int walk_smem_bytes(const WalkParams& p, bool wide_states, int warps, bool smem_tables, bool smem_hist);

}  // namespace b2r
