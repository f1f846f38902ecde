// Host-side model of the reference's regex definitions and the dense device tables packed from them.
//   AllstrRegexDef / SubstrRegexDef : reference src/defs.rs:26-36, 115-132 (public fields keep their meaning)
//   PackedDef                       : replaces the HashMap/HashSet probes of src/lib.rs:804-888 by one table entry
#pragma once
#include <cstdint>
#include <map>
#include <set>
#include <string>
#include <utility>
#include <vector>

namespace b2r {

struct Transition {  // one surviving state_lookup entry: key (ch,cur) → value (line_idx,next)
    uint8_t ch;
    uint64_t cur, next, line_idx;
};

struct AllstrDef {
    uint64_t first_state_val = 0, accepted_state_val = 0, largest_state_val = 0;
    // key (cur, ch) → (line_idx, next); a later line with the same key replaces the value (HashMap::insert)
    std::map<std::pair<uint64_t, uint8_t>, std::pair<uint64_t, uint64_t>> state_lookup;
    std::vector<Transition> in_table_order() const;  // sorted by line index (src/table.rs:103-108)
};

struct SubstrDef {
    uint64_t max_length = 0, min_position = 0, max_position = 0;
    std::set<std::pair<uint64_t, uint64_t>> valid_state_transitions;
    std::vector<uint64_t> start_states, end_states;  // file order, duplicates kept (Vec)
};

// returns 0 or B2R_ERR_PARSE with *err_line set
int parse_allstr(const char* text, size_t len, AllstrDef& out, uint64_t* err_line);
int parse_substr(const char* text, size_t len, SubstrDef& out, uint64_t* err_line);
int read_file(const char* path, std::string& out);

// ---- packed tables ---------------------------------------------------------------------------------------
// Transition entry (u32), indexed [byte_class][state]:
//   bits  0..15  next state
//   bits 16..23  substr id of the transition (global id incl. the running offset; 0 = none)
//   bit  24      is_start : substr id != 0 and cur  in start_states of that substr   (src/lib.rs:857-868)
//   bit  25      is_end   : substr id != 0 and next in end_states   of that substr   (src/lib.rs:870-881)
//   bit  26      invalid  : no state_lookup entry for (byte, cur)                    (src/lib.rs:817 panics)
constexpr uint32_t ENT_NEXT_MASK = 0xFFFFu;
constexpr uint32_t ENT_SID_SHIFT = 16;
constexpr uint32_t ENT_SID_MASK = 0xFFu << ENT_SID_SHIFT;
constexpr uint32_t ENT_IS_START = 1u << 24;
constexpr uint32_t ENT_IS_END = 1u << 25;
constexpr uint32_t ENT_INVALID = 1u << 26;
constexpr uint32_t ENT_RARE_MASK = ENT_SID_MASK | ENT_IS_START | ENT_IS_END | ENT_INVALID;

struct TableRow { uint64_t ch, cur, next, sid; };       // src/table.rs:72-100 assign_row
struct EndpointRow { uint64_t sid, start, end; };       // src/table.rs:130-193

struct PackedDef {
    uint32_t num_states = 0;      // S = largest_state_val + 1 = dummy state value
    uint32_t first_state = 0;
    uint32_t accepted_state = 0;  // 0xFFFFFFFF when the accepted state id is not a real state
    uint32_t state_width = 1;     // bytes per state in the state column
    uint32_t substr_id_offset = 1, num_substrs = 0;
    uint32_t num_classes = 0;     // byte equivalence classes incl. the all-invalid class (if any byte has no edge)
    uint32_t hot_states[2] = {0, 0};   // bit s: some transition out of state s (< 64) carries a substr id
    std::vector<uint8_t> byte_class;   // [256]
    std::vector<uint32_t> trans;       // [num_classes][num_states]
    std::vector<TableRow> rows;        // RegexTableConfig::load order; rows[0] = (0,dummy,dummy,0)
    std::vector<uint32_t> row_bin;     // for r >= 1: dense histogram bin ch*S + cur of row r (row_bin[0] unused)
    std::vector<EndpointRow> erows;
    // endpoint bins: counter index k*S + state (k = substr index within the def) for each endpoint row that is the
    // FIRST row matching its tuple; 0xFFFFFFFF for rows shadowed by an identical earlier row (they receive 0)
    std::vector<uint32_t> erow_start_bin, erow_end_bin;
};

// Walk table of walk_kernel (walk.cuh): [num_classes][P] with P = the power of two >= S + 1, entry = next << 16 | rare.
// rare (bit 0) marks transitions with a non-zero substr id and invalid transitions; an invalid transition leads to the
// trap state S (= the dummy state value), which only leads to itself.  Slots s >= S are trap rows.
uint32_t padded_states(uint32_t num_states);
void build_walk_table(const PackedDef& def, std::vector<uint32_t>& out);

// returns 0, B2R_ERR_UNSUPPORTED or B2R_ERR_INVALID_ARG (message via set_error)
int pack_def(const AllstrDef& a, const std::vector<const SubstrDef*>& substrs, uint32_t substr_id_offset, PackedDef& out);

void set_error(const char* fmt, ...);
const char* get_error();

}  // namespace b2r
