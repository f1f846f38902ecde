// The multi-device handle (b2r_config_new_multi, reference call site src/lib.rs:311-318 served from ONE process) through the
// C++ host mirror: a ragged batch is run on one GPU and on `n_dev` GPUs; every witness column must be identical and every
// multiplicity counter row must match (the per-device counters are summed by one NCCL all-reduce), dense and sparse D2H.
// Usage: test_multi_device <dir with the lookup texts> <n_dev>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "b2r.hpp"

using namespace halo2_regex;

static uint64_t rng_state = 0xB2000006ull;
static uint32_t rnd() { rng_state = rng_state * 6364136223846793005ull + 1442695040888963407ull; return (uint32_t)(rng_state >> 33); }

int main(int argc, char** argv) {
    if (argc < 3) { std::printf("usage: %s <defs dir> <n_dev>\n", argv[0]); return 2; }
    const std::string dir = std::string(argv[1]) + "/";
    const int n_dev = std::atoi(argv[2]);
    try {
        const std::vector<RegexDefs> defs = {
            {AllstrRegexDef::read_from_text(dir + "regex1_test_lookup.txt"), {SubstrRegexDef::read_from_text(dir + "substr1_test_lookup.txt")}},
            {AllstrRegexDef::read_from_text(dir + "regex2_test_lookup.txt"), {SubstrRegexDef::read_from_text(dir + "substr2_test_lookup.txt")}}};
        const size_t M = 200;
        std::vector<int> devices;
        for (int i = 0; i < n_dev; i++) devices.push_back(i);
        const RegexVerifyConfig one = RegexVerifyConfig::configure(M, defs, 0);
        const RegexVerifyConfig many = RegexVerifyConfig::configure_multi(M, defs, devices);
        if ((int)b2r_config_num_devices(many.handle()) != n_dev) { std::printf("FAIL device count\n"); return 1; }
        // 50 003 ragged strings over the regexes' alphabet with the two substrings planted in most of them
        const char alphabet[] = "abcdefghijklmnopqrstuvwxyz .@:<>\r\n";
        std::vector<std::vector<uint8_t>> strings(50003);
        for (auto& s : strings) {
            const size_t len = rnd() % M;
            for (size_t i = 0; i < len; i++) s.push_back((uint8_t)alphabet[rnd() % (sizeof alphabet - 1)]);
            const std::string plant = (rnd() & 1) ? "email was meant for @abc." : " Also for xyz.";
            if (len > plant.size() + 2 && (rnd() % 8)) std::memcpy(s.data() + rnd() % (len - plant.size()), plant.data(), plant.size());
        }
        int failures = 0;
        for (int sparse = 0; sparse < 2; sparse++) {
            const auto a = one.match_batch(strings, false), b = many.match_batch(strings, sparse != 0);
            auto same = [&](const char* what, bool ok) { if (!ok) { std::printf("FAIL %s (sparse=%d)\n", what, sparse); failures++; } };
            for (size_t d = 0; d < defs.size(); d++) {
                same("states", a.states[d] == b.states[d]); same("substr_ids", a.substr_ids[d] == b.substr_ids[d]);
                same("mult rows", a.mult[d] == b.mult[d]); same("endpoint_mult rows", a.endpoint_mult[d] == b.endpoint_mult[d]);
                // bitmaps: only the M defined bits are specified
                for (size_t j = 0; j < strings.size() && failures == 0; j++)
                    for (size_t i = 0; i < M; i++) {
                        const size_t o = j * a.bitmap_pitch + (i >> 3);
                        if (((a.start_enable[d][o] ^ b.start_enable[d][o]) | (a.end_enable[d][o] ^ b.end_enable[d][o])) >> (i & 7) & 1) { same("enable bitmaps", false); break; }
                    }
                uint64_t sum = 0;
                for (uint64_t v : b.mult[d]) sum += v;
                same("sum(mult) == N*M", sum == (uint64_t)strings.size() * M);
            }
            same("masked_chars", a.masked_chars == b.masked_chars); same("masked_substr_ids", a.masked_substr_ids == b.masked_substr_ids);
            bool st = true;
            for (size_t j = 0; j < strings.size(); j++) st = st && a.status[j].flags == b.status[j].flags && a.status[j].n_records == b.status[j].n_records;
            same("status", st);
        }
        if (failures) return 1;
        std::printf("multi-device handle on %d GPUs: every column and every multiplicity row equals the single-device result\n", n_dev);
        return 0;
    } catch (const std::exception& e) {
        std::printf("error: %s\n", e.what());
        return 1;
    }
}
