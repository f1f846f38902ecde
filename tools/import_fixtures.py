#!/usr/bin/env python3
"""Copies the regex DEFINITION DATA files (DFA lookup texts / decomposed-regex JSON produced by the reference's VRM
tool) from the read-only reference checkout into tests/golden/defs/ and records their sha256.

These are data fixtures, not source: they are the inputs every BASELINE.json config names (ex_allstr.txt,
regex{1,2,3}_test_lookup.txt, substr{1,2,3}_test_lookup.txt) and /root/reference does not exist on the GPU box.
Run once in the build container:  python tools/import_fixtures.py
"""
import hashlib
import json
import os
import shutil

REF = "/root/reference"
DST = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "defs")
FILES = [
    "examples/ex_allstr.txt", "examples/ex_substr_id1.txt",
    "test_regexes/regex1_test_lookup.txt", "test_regexes/regex2_test_lookup.txt", "test_regexes/regex3_test_lookup.txt",
    "test_regexes/substr1_test_lookup.txt", "test_regexes/substr2_test_lookup.txt", "test_regexes/substr3_test_lookup.txt",
    "test_regexes/regex1_test.json", "test_regexes/regex2_test.json", "test_regexes/regex3_test.json",
]
os.makedirs(DST, exist_ok=True)
manifest = {}
for rel in FILES:
    src = os.path.join(REF, rel)
    dst = os.path.join(DST, os.path.basename(rel))
    shutil.copyfile(src, dst)
    manifest[os.path.basename(rel)] = {"from": rel, "sha256": hashlib.sha256(open(src, "rb").read()).hexdigest()}
json.dump(manifest, open(os.path.join(DST, "MANIFEST.json"), "w"), indent=1, sort_keys=True)
print("imported", len(FILES), "files")
