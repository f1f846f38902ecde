#!/usr/bin/env python3
"""Known-answer vectors for halo2_regex_b200/vrm.py::regex_to_dfa beyond the three reference fixtures.

The expected DFA graphs were produced by the repo's FIRST implementation of the regex -> DFA step — the function-by-function
port of /root/reference/src/vrm/regex.js that round 1 shipped (commit 601e63a, loaded here with `git show`) — before that half
of vrm.py was rewritten (cursor parser, bit-set subset construction, Moore refinement).  The port reproduced the reference's
three `*_lookup.txt` fixtures byte for byte; these vectors pin the rewrite to it on ~400 more patterns.
    python tools/make_vrm_golden.py  ->  tests/golden/vrm_dfa_cases.json
"""
import hashlib, importlib.util, json, os, random, subprocess, sys, tempfile

ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..")
src = subprocess.run(["git", "show", "601e63a:halo2_regex_b200/vrm.py"], cwd=ROOT, capture_output=True, text=True, check=True).stdout
with tempfile.NamedTemporaryFile("w", suffix=".py", delete=False) as f:
    f.write(src)
spec = importlib.util.spec_from_file_location("vrm_port", f.name)
port = importlib.util.module_from_spec(spec)
sys.modules["vrm_port"] = port
spec.loader.exec_module(port)


def _key(s):   # the port's sort key mixed str and tuple keys (a crash on sets with non-ASCII labels); same order, one type
    b = s.encode("utf-16-be")
    return tuple(int.from_bytes(b[i:i + 2], "big") for i in range(0, len(b), 2))


port._utf16_key = _key

rng = random.Random(20261017)
ATOMS = ["a", "b", "c", "d", "0", "1", " ", "\\(", "\\|", "\\*", "\\\\", "@", ".", "\r", "\n", "\t", "ϵ", "é", "-", "_"]


def gen(depth):
    r = rng.random()
    if depth == 0 or r < 0.3:
        return rng.choice(ATOMS)
    if r < 0.5:
        return "(" + "|".join(gen(depth - 1) for _ in range(rng.randint(2, 4))) + ")"
    if r < 0.8:
        return gen(depth - 1) + gen(depth - 1)
    x = gen(depth - 1)
    if len(x) > 1 and not (len(x) == 2 and x[0] == "\\"):
        x = "(" + x + ")"
    return x + rng.choice("*+?")


hand = ["a", "ab|c", "(a|b)*abb", "a+b?c*", "(ab)+", "x(y|z)?w", "a(b(c|d)*e)+f", "ϵ|a", "(a|ϵ)b", "\\(a\\)\\*", "(a*)*", "a**", "((a))",
        "email was meant for @(a|b|c|d|e|f|g|h|i|j|k|l|m|n|o|p|q|r|s|t|u|v|w|x|y|z)+.", "(\r\n|^)from:(a|b| )+<(a|b|@|.)+>\r\n",
        port.catch_all_regex_str() + "*x", "a\\ϵb", "(0|1|2|3|4|5|6|7|8|9)+(.(0|1|2|3|4|5|6|7|8|9)+)?"]
cases = []
for rx in hand + [gen(rng.randint(1, 6)) for _ in range(400)]:
    text = port.dfa_json(port.regex_to_dfa(rx))
    cases.append({"regex": rx, "states": text.count('"type"'), "sha256": hashlib.sha256(text.encode("utf-8")).hexdigest(), "dfa_json": text if len(text) <= 400 else None})
bad = ["", "(ab", "*a", "a||b", "(|a)", "()", "a|", "|a", "+", "a(?b)", "\\"]
out = {"generated_by": "tools/make_vrm_golden.py (the round-1 port of regex.js at commit 601e63a)", "cases": cases, "syntax_errors": bad}
with open(os.path.join(ROOT, "tests", "golden", "vrm_dfa_cases.json"), "w") as f:
    json.dump(out, f, ensure_ascii=True, indent=0)
print(len(cases), "cases")
