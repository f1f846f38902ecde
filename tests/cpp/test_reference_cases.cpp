// The reference's own tests, restated against the C++ host mirror (include/b2r.hpp) of its API:
//   TestCircuit1 (regex1+substr1, regex2+substr2)  test_substr_pass1 / pass2 / fail1     reference src/lib.rs:1068-1151
//   TestCircuit2 (regex3+substr3)                  test_substr_pass3 / pass4 / fail2-4   reference src/lib.rs:1317-1470
//   examples/regex.rs                              expected public instances             reference examples/regex.rs:185-206
// The reference checks `masked_characters` / `all_substr_ids` element-wise against expectations built from
// `correct_substrs: Vec<(usize, String)>` with id = list index + 1 (src/lib.rs:1043-1059), and expects `verify()` to fail for
// the fail cases (the accept constraint, src/lib.rs:442-457).  Usage: test_reference_cases <dir with the lookup texts>
#include <array>
#include <cstdio>
#include <cstring>
#include <string>
#include <vector>

#include "b2r.hpp"

using namespace halo2_regex;

static int failures = 0;
#define EXPECT(cond, name)                                                    \
    do {                                                                      \
        if (!(cond)) { std::printf("FAIL %s: %s\n", name, #cond); failures++; } \
    } while (0)

static std::vector<uint8_t> bytes(const std::string& s) { return std::vector<uint8_t>(s.begin(), s.end()); }

// src/lib.rs:1043-1059: expected_masked_chars / expected_substr_ids from correct_substrs
static void check_case(const RegexVerifyConfig& cfg, const char* name, const std::string& input, bool verify_ok,
                       const std::vector<std::pair<size_t, std::string>>& correct_substrs, bool compare) {
    const AssignedRegexResult r = cfg.match_substrs(bytes(input));
    bool all_accepted = true;
    for (bool a : r.accepted) all_accepted = all_accepted && a;
    EXPECT(all_accepted == verify_ok, name);
    if (compare) {
        std::vector<uint8_t> expected_masked_chars(cfg.max_chars_size, 0), expected_substr_ids(cfg.max_chars_size, 0);
        for (size_t substr_idx = 0; substr_idx < correct_substrs.size(); substr_idx++) {
            const auto& [start, chars] = correct_substrs[substr_idx];
            for (size_t idx = 0; idx < chars.size(); idx++) {
                expected_masked_chars[start + idx] = (uint8_t)chars[idx];
                expected_substr_ids[start + idx] = (uint8_t)(substr_idx + 1);
            }
        }
        EXPECT(r.masked_characters == expected_masked_chars, name);
        EXPECT(r.all_substr_ids == expected_substr_ids, name);
    }
    size_t enabled = 0;
    for (uint8_t e : r.all_enable_flags) enabled += e;
    EXPECT(enabled == input.size(), name);
    if (failures == 0) std::printf("ok   %s\n", name);
}

int main(int argc, char** argv) {
    if (argc < 2) { std::printf("usage: %s <defs dir>\n", argv[0]); return 2; }
    const std::string dir = std::string(argv[1]) + "/";
    try {
        const size_t MAX_STRING_LEN = 1024;   // src/lib.rs:930
        // TestCircuit1::configure, src/lib.rs:960-987
        const std::vector<RegexDefs> defs1 = {
            {AllstrRegexDef::read_from_text(dir + "regex1_test_lookup.txt"), {SubstrRegexDef::read_from_text(dir + "substr1_test_lookup.txt")}},
            {AllstrRegexDef::read_from_text(dir + "regex2_test_lookup.txt"), {SubstrRegexDef::read_from_text(dir + "substr2_test_lookup.txt")}}};
        const RegexVerifyConfig c1 = RegexVerifyConfig::configure(MAX_STRING_LEN, defs1);
        check_case(c1, "test_substr_pass1", "email was meant for @y. Also for x.", true, {{21, "y"}, {33, "x"}}, true);
        check_case(c1, "test_substr_pass2", "email was meant for @yajk. Also for swq.", true, {{21, "yajk"}, {36, "swq"}}, true);
        check_case(c1, "test_substr_fail1", "email was meant for @@", false, {}, true);
        // TestCircuit2::configure, src/lib.rs:1227-1242
        const std::vector<RegexDefs> defs2 = {
            {AllstrRegexDef::read_from_text(dir + "regex3_test_lookup.txt"), {SubstrRegexDef::read_from_text(dir + "substr3_test_lookup.txt")}}};
        const RegexVerifyConfig c2 = RegexVerifyConfig::configure(MAX_STRING_LEN, defs2);
        check_case(c2, "test_substr_pass3", "from:alice@gmail.com\r\n", true, {{5, "alice@gmail.com"}}, true);
        check_case(c2, "test_substr_pass4", "dummy\r\nfrom:alice<alice@gmail.com>\r\n", true, {{18, "alice@gmail.com"}}, true);
        check_case(c2, "test_substr_fail2", "from:alice<alicegmail.com>\r\n", false, {}, false);
        check_case(c2, "test_substr_fail3", "from:alice<alice@gmail.com>", false, {}, false);
        check_case(c2, "test_substr_fail4", "fromalice<alice@gmail.com>\r\n", false, {}, false);
        // examples/regex.rs: MAX_STRING_LEN = 128 (:21), expected instances (:185-206)
        const std::vector<RegexDefs> defs3 = {
            {AllstrRegexDef::read_from_text(dir + "ex_allstr.txt"), {SubstrRegexDef::read_from_text(dir + "ex_substr_id1.txt")}}};
        const RegexVerifyConfig c3 = RegexVerifyConfig::configure(128, defs3);
        check_case(c3, "examples/regex.rs", "email was meant for @vitalik.", true, {{21, "vitalik"}}, true);
        // derive_* shapes (src/lib.rs:804-888)
        const auto s = bytes("email was meant for @vitalik.");
        const auto st = c3.derive_states(s);
        EXPECT(st.size() == 1 && st[0].size() == s.size() + 1 && st[0][0] == 0 && st[0].back() == 2, "derive_states");
        const auto ids = c3.derive_substr_ids(s);
        EXPECT(ids[0].size() == s.size() && ids[0][20] == 0 && ids[0][21] == 1 && ids[0][27] == 1 && ids[0][28] == 0, "derive_substr_ids");
        const auto se = c3.derive_is_start_end(s);
        EXPECT(se.first[0].size() == s.size() + 1 && se.first[0][21] && !se.first[0].back() && !se.second[0][0], "derive_is_start_end");
        // the reference panics on a byte without a transition (src/lib.rs:817)
        bool threw = false;
        try { c3.match_substrs(bytes("email was Xeant")); } catch (const InvalidTransition& e) { threw = std::string(e.what()).find("by 88 is invalid!") != std::string::npos; }
        EXPECT(threw, "invalid transition panic text");
        // table row order (src/table.rs:101-122): row 0 = (0, dummy, dummy, 0)
        const auto rows = c1.table_rows(0);
        EXPECT(rows.size() == 2843 && rows[0][0] == 0 && rows[0][1] == 29 && rows[0][2] == 29 && rows[0][3] == 0, "table rows");
    } catch (const std::exception& e) {
        std::printf("FAIL exception: %s\n", e.what());
        return 1;
    }
    std::printf(failures ? "%d FAILURES\n" : "all reference cases pass\n", failures);
    return failures ? 1 : 0;
}
