import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
DEFS = os.path.join(ROOT, "tests", "golden", "defs")

# name -> [(allstr file, [substr files]), ...]  (one tuple per RegexDefs)
DEF_SETS = {
    "example": [("ex_allstr.txt", ["ex_substr_id1.txt"])],                                   # examples/regex.rs:58-77
    "regex1": [("regex1_test_lookup.txt", ["substr1_test_lookup.txt"])],                     # BASELINE config 1
    "regex2": [("regex2_test_lookup.txt", ["substr2_test_lookup.txt"])],
    "regex3": [("regex3_test_lookup.txt", ["substr3_test_lookup.txt"])],                     # TestCircuit2, src/lib.rs:1227-1242
    "test1": [("regex1_test_lookup.txt", ["substr1_test_lookup.txt"]),                       # TestCircuit1, src/lib.rs:960-987
              ("regex2_test_lookup.txt", ["substr2_test_lookup.txt"])],
    "regex3_k3": [("regex3_test_lookup.txt", ["substr1_test_lookup.txt", "substr2_test_lookup.txt", "substr3_test_lookup.txt"])],  # config 2 reading (i)
    "three": [("regex1_test_lookup.txt", ["substr1_test_lookup.txt"]), ("regex2_test_lookup.txt", ["substr2_test_lookup.txt"]),
              ("regex3_test_lookup.txt", ["substr3_test_lookup.txt"])],                     # config 2 reading (ii)
}


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run with -m gpu on the B200 box)")


def pytest_collection_modifyitems(config, items):
    """gpu-marked tests are skipped (not failed) on a box without a CUDA device; `-m "not gpu"` deselects them altogether."""
    try:
        import torch
        have_gpu = torch.cuda.is_available()
    except Exception:  # pragma: no cover
        have_gpu = False
    if have_gpu:
        return
    skip = pytest.mark.skip(reason="needs a CUDA device (B200); there is no CPU fallback to run instead")
    for item in items:
        if "gpu" in item.keywords:
            item.add_marker(skip)


def read(name):
    with open(os.path.join(DEFS, name), "rb") as f:
        return f.read()


def oracle_config(set_name_or_list, max_chars_size):
    from oracle import oracle as O
    spec = DEF_SETS[set_name_or_list] if isinstance(set_name_or_list, str) else set_name_or_list
    defs = [(O.OracleAllstr(read(a)), [O.OracleSubstr(read(s)) for s in ss]) for a, ss in spec]
    return O.OracleConfig(defs, max_chars_size)


def product_config(set_name_or_list, max_chars_size, device=None):
    import halo2_regex_b200 as H
    spec = DEF_SETS[set_name_or_list] if isinstance(set_name_or_list, str) else set_name_or_list
    defs = [H.RegexDefs(H.AllstrRegexDef.read_from_text(os.path.join(DEFS, a)),
                        [H.SubstrRegexDef.read_from_text(os.path.join(DEFS, s)) for s in ss]) for a, ss in spec]
    return H.RegexVerifyConfig.configure(max_chars_size, defs, device=device)


def pyref_defs(set_name_or_list):
    from oracle import pyref as P
    spec = DEF_SETS[set_name_or_list] if isinstance(set_name_or_list, str) else set_name_or_list
    return [(P.PyAllstr(read(a)), [P.PySubstr(read(s)) for s in ss]) for a, ss in spec]


@pytest.fixture(scope="session")
def golden_vectors():
    import json
    with open(os.path.join(ROOT, "tests", "golden", "golden_vectors.json")) as f:
        return json.load(f)
