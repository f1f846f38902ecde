"""CPU: the C-ABI library loads, exports every symbol include/b2r.h declares, and its host logic (definition loaders,
dense-table packer, table row order) agrees with the oracle.  No compute calls (no GPU here)."""
import os
import re
import subprocess

import numpy as np
import pytest

from conftest import DEF_SETS, DEFS, ROOT, oracle_config, product_config, read

import halo2_regex_b200 as H
from oracle import oracle as O
from oracle import pyref as P


def test_library_exports_every_declared_symbol():
    hdr = open(os.path.join(ROOT, "include", "b2r.h")).read()
    hdr = re.sub(r"/\*.*?\*/", "", hdr, flags=re.S)
    declared = set(re.findall(r"\b(b2r_[a-z0-9_]+)\s*\(", hdr))
    exported = set(re.findall(r" T (b2r_\w+)", subprocess.check_output(["nm", "-D", H.LIB_PATH]).decode()))
    assert declared, "no declarations parsed"
    assert declared <= exported, f"declared but not exported: {sorted(declared - exported)}"
    assert set(H.SYMBOLS) == declared
    assert b"sm_100a" in H.lib.b2r_version()


def test_library_is_sm100a_only():
    out = subprocess.check_output(["cuobjdump", "-lelf", H.LIB_PATH]).decode()
    archs = set(re.findall(r"sm_\d+a?", out))
    assert archs == {"sm_100a"}, archs


@pytest.mark.parametrize("fname", ["ex_allstr.txt", "regex1_test_lookup.txt", "regex2_test_lookup.txt", "regex3_test_lookup.txt"])
def test_allstr_loader_matches_reference_format(fname):
    a = H.AllstrRegexDef.read_from_text(os.path.join(DEFS, fname))
    p = P.PyAllstr(read(fname))
    assert (a.first_state_val, a.accepted_state_val, a.largest_state_val) == (p.first_state_val, p.accepted_state_val, p.largest_state_val)
    assert a.state_lookup == p.state_lookup
    assert a.get(ord("e"), 0) == p.state_lookup.get((ord("e"), 0))
    assert a.get(0xFF, 0) is None
    b = H.AllstrRegexDef.read_from_reader(read(fname))
    assert b.state_lookup == a.state_lookup


@pytest.mark.parametrize("fname", ["ex_substr_id1.txt", "substr1_test_lookup.txt", "substr2_test_lookup.txt", "substr3_test_lookup.txt"])
def test_substr_loader_matches_reference_format(fname):
    s = H.SubstrRegexDef.read_from_text(os.path.join(DEFS, fname))
    p = P.PySubstr(read(fname))
    assert (s.max_length, s.min_position, s.max_position) == (p.max_length, p.min_position, p.max_position)
    assert s.valid_state_transitions == p.valid_state_transitions
    assert s.start_states == p.start_states and s.end_states == p.end_states
    s2 = H.SubstrRegexDef.new(p.max_length, p.min_position, p.max_position, p.valid_state_transitions, p.start_states, p.end_states)
    assert s2.valid_state_transitions == s.valid_state_transitions and s2.start_states == s.start_states


# edge cases of src/defs.rs:75-110
ALLSTR_CASES = [
    (b"0\n1\n1\n0 1 97\n", True),
    (b"0\r\n1\r\n1\r\n0 1 97\r\n", True),                    # BufRead::lines strips \r\n
    (b"0\n1\n1\n0 1 97", True),                              # no trailing newline
    (b"0\n1\n1\n0 1 353\n", True),                           # 353 as u8 == 97
    (b"0\n1\n1\n0 1 97\n0 0 97\n", True),                    # duplicate key: later line wins, earlier row vanishes
    (b"0\n1\n1\n0 1 97 extra 9\n", False),                   # extra token must still parse as u64
    (b"0\n1\n1\n0 1 97 5 6\n", True),                        # extra numeric tokens are ignored
    (b"0\n1\n1\n\n0 1 97\n", False),                         # empty body line: index out of bounds → panic
    (b"0\n1\n1\n0 1\n", False),                              # short body line
    (b"0\n\n1\n", False),                                    # empty header line
    (b"0\n1\n1\n0 -1 97\n", False),                          # not a u64
    (b"0\n1\n1\n0 +1 97\n", True),                           # Rust's u64 parser accepts a leading '+'
    (b"0\n1\n1\n0 18446744073709551616 97\n", False),        # overflow
    (b"0\n1\n1\n0 18446744073709551615 97\n", True),
    (b"0\n1\n1\n0\t1\xc2\xa097\n", True),                    # Unicode whitespace (U+00A0) separates tokens
    (b"0\n1\n1\n0 1 97\xff\n", False),                       # invalid UTF-8
    (b"", True),
    (b"5\n", True),
    (b"0\n1\n1\n0 1 97\n\n", False),                         # a blank line before EOF is still a line
]


@pytest.mark.parametrize("text,ok", ALLSTR_CASES)
def test_allstr_parser_edge_cases(text, ok):
    if ok:
        a = H.AllstrRegexDef.read_from_reader(text)
        o = O.OracleAllstr(text)
        p = P.PyAllstr(text)
        assert a.state_lookup == p.state_lookup
        assert (a.first_state_val, a.accepted_state_val, a.largest_state_val) == (o.first_state_val, o.accepted_state_val, o.largest_state_val)
        assert len(a.state_lookup) == o.num_transitions
    else:
        with pytest.raises(H.RegexParseError):
            H.AllstrRegexDef.read_from_reader(text)
        with pytest.raises(O.OracleParseError):
            O.OracleAllstr(text)


SUBSTR_CASES = [
    (b"4\n0\n127\n21 \n22 23 \n21 22\n21 23\n", True),
    (b"4\n0\n127\n\n\n21 22\n", True),                       # empty start / end lines are allowed
    (b"4\n0\n127\n21\n22\n21 22\n21 22\n", True),            # duplicate pair collapses (HashSet)
    (b"4\n0\n127\n21\n22\n21\n", False),                     # short pair line
    (b"\n0\n127\n", False),
    (b"4\n0\n127\n21 21\n22\n", True),                       # duplicate start states are kept (Vec)
    (b"4\n0\n127\n21\n22\n21 22 99\n", True),
]


@pytest.mark.parametrize("text,ok", SUBSTR_CASES)
def test_substr_parser_edge_cases(text, ok):
    if ok:
        s = H.SubstrRegexDef.read_from_reader(text)
        p = P.PySubstr(text)
        o = O.OracleSubstr(text)
        assert s.valid_state_transitions == p.valid_state_transitions and len(p.valid_state_transitions) == o.num_transitions
        assert s.start_states == p.start_states and s.end_states == p.end_states
    else:
        with pytest.raises(H.RegexParseError):
            H.SubstrRegexDef.read_from_reader(text)
        with pytest.raises(O.OracleParseError):
            O.OracleSubstr(text)


def test_missing_file_is_an_error():
    with pytest.raises(FileNotFoundError):
        H.AllstrRegexDef.read_from_text("/nonexistent/allstr.txt")
    with pytest.raises(FileNotFoundError):
        H.SubstrRegexDef.read_from_text("/nonexistent/substr.txt")


@pytest.mark.parametrize("set_name", sorted(DEF_SETS))
def test_table_rows_follow_reference_order(set_name):
    """RegexTableConfig::load row order (src/table.rs:101-122, 129-193) incl. the running substr id offset."""
    cfg = product_config(set_name, 1024, device=-1)
    ocfg = oracle_config(set_name, 1024)
    off = 1
    for d, (a_name, s_names) in enumerate(DEF_SETS[set_name]):
        rows, erows = P.table_rows(P.PyAllstr(read(a_name)), [P.PySubstr(read(s)) for s in s_names], off)
        assert [tuple(int(x) for x in r) for r in cfg.table_rows(d)] == rows
        assert [tuple(int(x) for x in r) for r in cfg.endpoint_rows(d)] == erows
        assert np.array_equal(cfg.table_rows(d), ocfg.table_rows(d))
        assert np.array_equal(cfg.endpoint_rows(d), ocfg.endpoint_rows(d))
        assert cfg.substr_id_offsets[d] == off
        off += len(s_names)
    # SURVEY 8 fixture facts
    facts = {"example": ([75], [17]), "regex1": ([2843], [18]), "regex2": ([1275], [11]), "regex3": ([1961], [15])}
    if set_name in facts:
        assert (cfg.table_num_rows, cfg.num_byte_classes) == facts[set_name]


def test_duplicate_key_row_vanishes():
    a = H.AllstrRegexDef.read_from_reader(b"0\n1\n1\n0 1 97\n1 1 98\n0 0 97\n")
    cfg = H.RegexVerifyConfig.configure(8, [H.RegexDefs(a, [])], device=-1)
    assert [tuple(int(x) for x in r) for r in cfg.table_rows(0)] == [(0, 2, 2, 0), (98, 1, 1, 0), (97, 0, 0, 0)]


def test_unsupported_definitions_are_rejected():
    big = H.AllstrRegexDef.read_from_reader(b"0\n1\n70000\n0 1 97\n")
    with pytest.raises(RuntimeError, match="65535"):
        H.RegexVerifyConfig.configure(8, [H.RegexDefs(big, [])], device=-1)
    collide = H.AllstrRegexDef.read_from_reader(b"0\n1\n1\n0 2 97\n")      # state 2 > largest 1: dummy collides
    with pytest.raises(RuntimeError, match="dummy"):
        H.RegexVerifyConfig.configure(8, [H.RegexDefs(collide, [])], device=-1)
    a = H.AllstrRegexDef.read_from_reader(b"0\n1\n1\n0 1 97\n")
    with pytest.raises(RuntimeError):
        H.RegexVerifyConfig.configure(8, [H.RegexDefs(a, [])] * 5, device=-1)


def test_no_cpu_fallback():
    """A handle without a device can answer table queries but never computes."""
    cfg = product_config("regex1", 64, device=-1)
    with pytest.raises(RuntimeError, match="no CPU fallback"):
        cfg.match_strings([b"abc"])


def test_product_never_imports_oracle():
    import halo2_regex_b200
    pkg = os.path.dirname(halo2_regex_b200.__file__)
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".hpp", "Makefile")):
                src = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in src.lower() or f == "buffers.py" or f == "regex.py", f
    for f in ("buffers.py", "regex.py"):
        src = open(os.path.join(pkg, f)).read()
        assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src


def test_workload_generators_agree():
    """numpy and torch (CPU) generators produce byte-identical config-1 batches; planted matches are where planned."""
    import torch
    from halo2_regex_b200 import workloads as W
    a, plan = W.config1_numpy(64, 256)
    b = W.config1_torch(64, 256, device="cpu").numpy()
    assert np.array_equal(a, b)
    c, _ = W.config1_numpy(16, 256, first=48)
    assert np.array_equal(a[48:], c)
    assert set(np.unique(a)) <= set(W.ALPHABET)
    j = int(np.nonzero(plan["has_match"])[0][0])
    o = int(plan["offset"][j])
    assert bytes(a[j, o:o + 21]) == b"email was meant for @"
    assert a[j, o + 21 + int(plan["name_len"][j])] == ord(".")
    ocfg = oracle_config("regex1", 257)
    out, res = ocfg.match_batch(a.reshape(-1), np.arange(65, dtype=np.uint64) * 256)
    assert res.code == 0
    acc = (out.status["flags"] & 1).astype(bool)
    assert acc[plan["has_match"]].all()


def test_every_documented_option_is_accepted():
    """The knobs listed under b2r_config_set_option in include/b2r.h exist (a handle without a device takes options too);
    an unknown name is an error, not a silent no-op."""
    header = open(os.path.join(ROOT, "include", "b2r.h")).read()
    block = header[header.index("testing / tuning knobs of a handle"):header.index("int b2r_config_set_option")]
    names = ["table_mode", "hist_mode", "fuse", "stagger_ns", "slices", "host_threads", "small_path", "sparse_cap", "sparse_direct", "trace_host",
             "hist_cache_log2", "spread_fill", "long_fused", "debug", "host_debug"]
    for name in names:
        assert name in block, f"{name} is not documented in include/b2r.h"
    cfg = product_config("regex1", 64, device=-1)
    values = {"table_mode": "repl16", "hist_mode": "smem", "trace_host": "0", "debug": "0", "host_debug": "0"}
    for name in names:
        cfg.set_option(name, values.get(name, "1"))
    with pytest.raises(Exception):
        cfg.set_option("no_such_option", "1")
