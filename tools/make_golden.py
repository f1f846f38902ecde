#!/usr/bin/env python3
"""Writes tests/golden/golden_vectors.json: the literal known-answer vectors held by the reference's own tests.

Every entry is transcribed from the cited reference test (input string, (start, substring) expectations, whether
MockProver::verify must succeed).  expected masked_chars / masked_substr_ids follow the reference's construction:
position start+k holds substring byte k and id = (index in correct_substrs)+1   (src/lib.rs:1043-1051, 1292-1299;
examples/regex.rs:185-192).  No reference code is executed (it cannot be built here).
"""
import json
import os

T1 = [["regex1_test_lookup.txt", ["substr1_test_lookup.txt"]], ["regex2_test_lookup.txt", ["substr2_test_lookup.txt"]]]
T2 = [["regex3_test_lookup.txt", ["substr3_test_lookup.txt"]]]
EX = [["ex_allstr.txt", ["ex_substr_id1.txt"]]]
V = [
    dict(name="G1", cite="src/lib.rs:1068-1093 test_substr_pass1", defs=T1, M=1024,
         input="email was meant for @y. Also for x.", substrs=[[21, "y"], [33, "x"]], verify_ok=True),
    dict(name="G2", cite="src/lib.rs:1095-1120 test_substr_pass2", defs=T1, M=1024,
         input="email was meant for @yajk. Also for swq.", substrs=[[21, "yajk"], [36, "swq"]], verify_ok=True),
    dict(name="G3", cite="src/lib.rs:1122-1151 test_substr_fail1", defs=T1, M=1024,
         input="email was meant for @@", substrs=[], verify_ok=False),
    dict(name="G4", cite="src/lib.rs:1317-1343 test_substr_pass3", defs=T2, M=1024,
         input="from:alice@gmail.com\r\n", substrs=[[5, "alice@gmail.com"]], verify_ok=True),
    dict(name="G5", cite="src/lib.rs:1345-1371 test_substr_pass4", defs=T2, M=1024,
         input="dummy\r\nfrom:alice<alice@gmail.com>\r\n", substrs=[[18, "alice@gmail.com"]], verify_ok=True),
    dict(name="G6", cite="src/lib.rs:1373-1404 test_substr_fail2", defs=T2, M=1024,
         input="from:alice<alicegmail.com>\r\n", substrs=None, verify_ok=False),
    dict(name="G7", cite="src/lib.rs:1406-1437 test_substr_fail3", defs=T2, M=1024,
         input="from:alice<alice@gmail.com>", substrs=None, verify_ok=False),
    dict(name="G8", cite="src/lib.rs:1439-1470 test_substr_fail4", defs=T2, M=1024,
         input="fromalice<alice@gmail.com>\r\n", substrs=None, verify_ok=False),
    dict(name="G9", cite="examples/regex.rs:150-206", defs=EX, M=128,
         input="email was meant for @vitalik.", substrs=[[21, "vitalik"]], verify_ok=True),
]
out = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden", "golden_vectors.json")
json.dump(V, open(out, "w"), indent=1)
print("wrote", out)
