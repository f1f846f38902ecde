"""Regex definition compiler: decomposed-regex JSON -> the allstr / substr lookup text files the matcher reads.

Host-side tooling; nothing here runs on the GPU or on the witness-generation path.  The
reference does this step with a JavaScript program evaluated in an embedded JS engine plus Rust glue around `petgraph`
and `fancy_regex`:

    regex text -> syntax tree -> epsilon-NFA -> subset construction -> Hopcroft -> renumbering -> JSON graph
        (/root/reference/src/vrm/regex.js:40-90, 236-762)
    JSON graph -> "<first>\\n<accepted>\\n<max>\\n<cur> <next> <byte>..." text
        (/root/reference/src/vrm/js_caller.rs:127-157)
    JSON graph + the decomposed parts -> substring transition sets and their start / end states
        (/root/reference/src/vrm/mod.rs:63-90, 265-600)

The files it produces are only reproducible if every tie is broken the way the reference breaks it, so what is kept of the
reference are the *orders* its output depends on and nothing of its shape: the pattern language and the epsilon-NFA states per
syntax node (they decide which subset-construction states exist), the naming / sorting rule of the final numbering, the
insertion-ordered property maps of the JS engine (integer-like keys first), the UTF-16 string order of `Array.prototype.sort`,
the swap-remove edge storage and newest-first adjacency lists of petgraph's `Graph`, and the leftmost-first match rule of the
regex engine.  The regex -> DFA step itself is this module's own design (cursor parser, array NFA with bit-set closures, Moore
partition refinement); round 1 shipped a function-by-function port of regex.js there, which now only survives as the generator
of the known-answer vectors (tools/make_vrm_golden.py -> tests/golden/vrm_dfa_cases.json).
The known-answer tests: byte-identical regeneration of `test_regexes/*_lookup.txt` from `test_regexes/*.json` and 418 recorded
DFA graphs (tests/test_vrm.py, fixtures under tests/golden/).  Deliberate difference on malformed patterns: an unbalanced ')'
is a syntax error here (the reference's parser reads it as a literal).
"""
from __future__ import annotations

import json
import re
import sys
from dataclasses import dataclass, field
from pathlib import Path
from typing import Dict, Iterable, List, Optional, Sequence, Set, Tuple

EPS = "ϵ"                      # the reference's epsilon label (regex.js:331, 406)
_ESCAPES = {"n": "\n", "r": "\r", "t": "\t", "v": "\v", "f": "\f"}      # regex.js:7

_ALNUM = "0123456789abcdefghijklmnopqrstuvwxyzABCDEFGHIJKLMNOPQRSTUVWXYZ"
_PUNCT_ALL = ["!", "\"", "#", "$", "%", "&", "'", "\\(", "\\)", "\\*", "\\+", ",", "-", ".", "/", ":", ";", "<", "=", ">",
              "\\?", "@", "[", "\\\\", "]", "^", "_", "`", "{", "\\|", "}", "~", " ", "\t"]


def catch_all_regex_str() -> str:
    """Alternation over every printable byte and white space (regex.js:10-12)."""
    return "(" + "|".join(list(_ALNUM) + _PUNCT_ALL + ["\n", "\r", "\x0b", "\x0c"]) + ")"


def catch_all_without_rn_regex_str() -> str:
    """The same without CR / LF (regex.js:14-16)."""
    return "(" + "|".join(list(_ALNUM) + _PUNCT_ALL + ["\x0b", "\x0c"]) + ")"


def text_context_prefix() -> str:
    """regex.js:18-20."""
    return 'Content-Type: text/plain; charset="UTF-8"\r\n\r\n'


# ---------------------------------------------------------------------------------------------------------------------
# The two engine behaviours every tie-break below rests on.
# ---------------------------------------------------------------------------------------------------------------------

def _utf16_key(s: str):
    """Sort key reproducing the JS default string order (by UTF-16 code unit)."""
    if s.isascii():
        return tuple(s.encode("ascii"))
    b = s.encode("utf-16-be")
    return tuple(int.from_bytes(b[i:i + 2], "big") for i in range(0, len(b), 2))


def _js_sorted(strings: Iterable[str]) -> List[str]:
    return sorted(strings, key=_utf16_key)


def _is_index_key(k: str) -> bool:
    """True for property names the JS engine treats as array indices (canonical decimal, < 2^32 - 1)."""
    if not k or not k.isascii() or not k.isdigit():
        return False
    if len(k) > 1 and k[0] == "0":
        return False
    return int(k) < 4294967295


def _js_keys(d: Dict[str, object]) -> List[str]:
    """Property enumeration order: index-like names ascending, then the others in insertion order."""
    idx = [k for k in d if _is_index_key(k)]
    if not idx:
        return list(d)
    idx.sort(key=int)
    return idx + [k for k in d if not _is_index_key(k)]


def js_json_string(s: str) -> str:
    """`JSON.stringify` of one string."""
    out = ['"']
    for ch in s:
        o = ord(ch)
        if ch == '"':
            out.append('\\"')
        elif ch == "\\":
            out.append("\\\\")
        elif ch == "\b":
            out.append("\\b")
        elif ch == "\f":
            out.append("\\f")
        elif ch == "\n":
            out.append("\\n")
        elif ch == "\r":
            out.append("\\r")
        elif ch == "\t":
            out.append("\\t")
        elif o < 0x20:
            out.append("\\u%04x" % o)
        elif 0xD800 <= o <= 0xDFFF:
            out.append("\\u%04x" % o)
        else:
            out.append(ch)
    out.append('"')
    return "".join(out)


def _symbol_set_key(chars: Iterable[str]) -> str:
    """`JSON.stringify(Object.keys(set).sort())`: the label of a merged edge (regex.js:741)."""
    return "[" + ",".join(js_json_string(c) for c in _js_sorted(chars)) + "]"


def format_regex_printable(s: str) -> str:
    """Part text -> the pattern handed to the regex engine (regex.js:22-38)."""
    e = js_json_string(s)[1:-1]
    for a, b in (("\\\\\\\\", "\\"), ("\\\\", "\\"), ("/", "\\/"), ("\x0b", "\\♥"), ("^", "\\^"), ("$", "\\$"),
                 ("|[|", "|\\[|"), ("|]|", "|\\]|"), ("|.|", "|\\.|"), ("|$|", "|\\$|"), ("|^|", "|\\^|")):
        e = e.replace(a, b)
    return e


format_regex_str = format_regex_printable            # js_caller.rs:36-41


# ---------------------------------------------------------------------------------------------------------------------
# regex text -> DFA graph.
#
# What the reference's output depends on (and what is therefore kept), as opposed to how its JavaScript computes it:
#   * the LANGUAGE of the pattern syntax: `\x` makes x a literal (n r t v f are the usual control characters), the unescaped
#     characters ( ) | * + ? are operators, the character "ϵ" is the empty word, `x+` means x x*, `x?` means (x|ϵ);
#   * the SHAPE of the epsilon-NFA: subset-construction states are sets of NFA states, epsilon-only states included, so two
#     subsets that differ only in such states are different DFA states before minimisation — and the names of those states
#     ("A", "B", ..., "AA": discovery order of a breadth-first subset construction over the symbols in UTF-16 order) are what
#     the final numbering is sorted by.  The NFA below therefore has the same states per syntax node as the reference's
#     (regex.js:390-452): none for a concatenation beyond one joint per gap, an entry/exit pair per alternative and per star;
#   * the final numbering rule (regex.js:700-762): equivalence classes ordered by the comma-joined names of their members
#     (UTF-16 order), the start state's class swapped to the front; merged edge labels = JSON list of the sorted symbols.
# Everything else is this module's own: a cursor-based recursive-descent parser, an array NFA with bit-set closures, and
# Moore-style partition refinement (the coarsest stable partition is unique, so it equals what the reference's worklist
# algorithm (regex.js:563-690) arrives at).
# ---------------------------------------------------------------------------------------------------------------------

class RegexSyntaxError(ValueError):
    pass


_OPERATORS = "()|*+?"

# syntax nodes: ("lit", ch) | ("eps",) | ("cat", [nodes]) | ("alt", [nodes]) | ("star", node)


class _Parser:
    """pattern := branch ('|' branch)*;  branch := piece+;  piece := atom ('*' | '+' | '?')*;  atom := literal | 'ϵ' | '(' pattern ')'."""

    def __init__(self, text: str):
        self.text = text
        self.pos = 0

    def _peek(self) -> Optional[str]:
        """The operator at the cursor, "" for a literal, None at the end."""
        if self.pos >= len(self.text):
            return None
        c = self.text[self.pos]
        return c if c in _OPERATORS else ""

    def _literal(self):
        """Consumes one literal; returns (character, was it escaped)."""
        c = self.text[self.pos]
        if c == "\\":
            if self.pos + 1 >= len(self.text):
                raise RegexSyntaxError("pattern ends inside an escape (offset %d)" % self.pos)
            c = self.text[self.pos + 1]
            self.pos += 2
            return _ESCAPES.get(c, c), True
        self.pos += 1
        return c, False

    def pattern(self, opened_at: int = -1):
        branches = [self._branch()]
        while self._peek() == "|":
            self.pos += 1
            branches.append(self._branch())
        if opened_at >= 0:
            if self._peek() != ")":
                raise RegexSyntaxError("no closing bracket for the group opened at offset %d" % opened_at)
            self.pos += 1
        elif self._peek() is not None:
            raise RegexSyntaxError("unbalanced ')' at offset %d" % self.pos)
        return branches[0] if len(branches) == 1 else ("alt", branches)

    def _branch(self):
        pieces = []
        while True:
            op = self._peek()
            if op is None or op == "|" or op == ")":
                break
            if op == "(":
                at = self.pos
                self.pos += 1
                pieces.append(self.pattern(opened_at=at))
            elif op != "":                      # * + ?
                if not pieces:
                    raise RegexSyntaxError("'%s' at offset %d has nothing to repeat" % (op, self.pos))
                self.pos += 1
                x = pieces[-1]
                pieces[-1] = ("star", x) if op == "*" else ("cat", [x, ("star", x)]) if op == "+" else ("alt", [x, ("eps",)])
            else:
                ch, _ = self._literal()
                pieces.append(("eps",) if ch == EPS else ("lit", ch))   # escaped or not: the reference's NFA labels its epsilon edges with this character
        if not pieces:
            raise RegexSyntaxError("empty alternative at offset %d" % self.pos)
        return pieces[0] if len(pieces) == 1 else ("cat", pieces)


def parse_regex(text: str):
    return _Parser(text).pattern()


class _Nfa:
    """States are integers; `eps[s]` lists the epsilon successors of s, `step[s]` maps a symbol to the bit set of its successors."""

    def __init__(self):
        self.eps: List[List[int]] = []
        self.step: List[Dict[str, int]] = []
        self._closure: Dict[int, int] = {}

    def new_state(self) -> int:
        self.eps.append([])
        self.step.append({})
        return len(self.eps) - 1

    def add(self, node, entry: int, exit_: int) -> None:
        """Wire the automaton of `node` between two existing states."""
        kind = node[0]
        if kind == "lit":
            self.step[entry][node[1]] = self.step[entry].get(node[1], 0) | (1 << exit_)
        elif kind == "eps":
            self.eps[entry].append(exit_)
        elif kind == "cat":
            at = entry
            for part in node[1][:-1]:
                joint = self.new_state()
                self.add(part, at, joint)
                at = joint
            self.add(node[1][-1], at, exit_)
        elif kind == "alt":
            for part in node[1]:
                a, b = self.new_state(), self.new_state()
                self.eps[entry].append(a)
                self.eps[b].append(exit_)
                self.add(part, a, b)
        else:                                   # star
            a, b = self.new_state(), self.new_state()
            self.eps[entry] += [a, exit_]
            self.eps[b] += [a, exit_]
            self.add(node[1], a, b)

    def closure(self, s: int) -> int:
        """Bit set of the states reachable from s over epsilon edges."""
        got = self._closure.get(s)
        if got is None:
            got, todo = 1 << s, [s]
            while todo:
                for t in self.eps[todo.pop()]:
                    if not got >> t & 1:
                        got |= 1 << t
                        todo.append(t)
            self._closure[s] = got
        return got

    def close(self, members: int) -> int:
        out = 0
        while members:
            low = members & -members
            out |= self.closure(low.bit_length() - 1)
            members ^= low
        return out


def _bits(x: int):
    while x:
        low = x & -x
        yield low.bit_length() - 1
        x ^= low


def _alpha_count(n: int) -> str:
    """0 -> A, 25 -> Z, 26 -> AA ...: the names the reference gives subset-construction states (regex.js:517-527)."""
    s = ""
    while n >= 0:
        s = chr(ord("A") + n % 26) + s
        n = n // 26 - 1
    return s


def _subset_dfa(text: str):
    """Breadth-first subset construction, symbols in UTF-16 order.  Returns (transitions, accepting): state i has the edges
    transitions[i] (symbol -> state), states are numbered in discovery order, 0 is the start."""
    sys.setrecursionlimit(max(sys.getrecursionlimit(), 20000))          # `add` recurses along the nesting depth of the pattern
    nfa = _Nfa()
    entry, final = nfa.new_state(), nfa.new_state()
    nfa.add(parse_regex(text), entry, final)
    index = {nfa.closure(entry): 0}
    order = [nfa.closure(entry)]
    transitions: List[Dict[str, int]] = []
    moved: Dict[int, int] = {}                                          # closure of a target set, by the set
    for members in order:                                              # grows while it is walked
        out: Dict[str, int] = {}
        for s in _bits(members):
            for sym, targets in nfa.step[s].items():
                out[sym] = out.get(sym, 0) | targets
        edges: Dict[str, int] = {}
        for sym in _js_sorted(out):
            closed = moved.get(out[sym])
            if closed is None:
                closed = moved[out[sym]] = nfa.close(out[sym])
            if closed not in index:
                index[closed] = len(order)
                order.append(closed)
            edges[sym] = index[closed]
        transitions.append(edges)
    return transitions, [bool(m >> final & 1) for m in order]


def _equivalence_classes(transitions, accepting) -> List[List[int]]:
    """Coarsest partition of the (partial) DFA's states that separates accepting from other states and is stable under every
    symbol; a missing edge is its own kind of target.  Refinement by signatures until nothing splits."""
    cls = [1 if a else 0 for a in accepting]
    count = len(set(cls))
    while True:
        seen: Dict[tuple, int] = {}
        nxt = []
        for i, edges in enumerate(transitions):
            sig = (cls[i], tuple(sorted((sym, cls[t]) for sym, t in edges.items())))
            nxt.append(seen.setdefault(sig, len(seen)))
        cls = nxt
        if len(seen) == count:
            break
        count = len(seen)
    groups: Dict[int, List[int]] = {}
    for i, c in enumerate(cls):
        groups.setdefault(c, []).append(i)
    return list(groups.values())


def regex_to_dfa(regex: str) -> List[dict]:
    """`regexToDfa` (regex.js:40-90) as the parsed value of the JSON it returns: one `{"type", "edges"}` object per
    state, edges keyed by the JSON text of the sorted list of characters that share a target."""
    transitions, accepting = _subset_dfa(regex)
    names = [_alpha_count(i) for i in range(len(transitions))]
    classes = _equivalence_classes(transitions, accepting)
    # the reference's numbering: classes ordered by the joined names of their members, then the start's class swapped to the front
    keyed = []
    for members in classes:
        members = sorted(members, key=lambda i: _utf16_key(names[i]))
        keyed.append((",".join(names[i] for i in members), members))
    keyed.sort(key=lambda km: _utf16_key(km[0]))
    at = next(k for k, (_, members) in enumerate(keyed) if 0 in members)
    keyed[0], keyed[at] = keyed[at], keyed[0]
    number = {}
    for k, (_, members) in enumerate(keyed):
        for i in members:
            number[i] = k
    # merged edges: every symbol that leads from class a to class b shares one label
    merged: List[Dict[int, Set[str]]] = [dict() for _ in keyed]
    for i, edges in enumerate(transitions):
        for sym, t in edges.items():
            merged[number[i]].setdefault(number[t], set()).add(sym)
    labelled = [{_symbol_set_key(syms): to for to, syms in by_to.items()} for by_to in merged]
    ordered = _js_sorted({lab for edges in labelled for lab in edges})
    return [{"type": "accept" if accepting[members[0]] else "", "edges": {lab: labelled[k][lab] for lab in ordered if lab in labelled[k]}}
            for k, (_, members) in enumerate(keyed)]


def get_dfa_json_value(regex: str) -> List[dict]:
    """js_caller.rs:43-48."""
    return regex_to_dfa(regex)


def dfa_json(graph: List[dict]) -> str:
    """The JSON text `regexToDfa` itself returns (insertion-ordered keys, no white space)."""
    parts = []
    for st in graph:
        if st is None:
            parts.append("null")
            continue
        edges = ",".join(js_json_string(k) + ":" + str(v) for k, v in st["edges"].items())
        parts.append('{"type":%s,"edges":{%s}}' % (js_json_string(st["type"]), edges))
    return "[" + ",".join(parts) + "]"


# ---------------------------------------------------------------------------------------------------------------------
# JSON graph -> allstr text (js_caller.rs:59-157)
# ---------------------------------------------------------------------------------------------------------------------

class VrmError(Exception):
    pass


def _edges_in_map_order(state: dict) -> List[Tuple[str, int]]:
    """serde_json's default object is a BTreeMap: iteration is by key bytes."""
    edges = state["edges"]
    if not isinstance(edges, dict):
        raise VrmError("Edges %r are not object" % (edges,))
    out = []
    for k in sorted(edges, key=lambda s: s.encode("utf-8")):
        v = edges[k]
        if not isinstance(v, int) or isinstance(v, bool) or v < 0:
            raise VrmError("node value %r is not u64" % (v,))
        out.append((k, v))
    return out


def get_accepted_state(graph: List[dict]) -> Optional[int]:
    for i, st in enumerate(graph):
        if st is not None and st.get("type") == "accept":
            return i
    return None


def get_max_state(graph: List[dict]) -> int:
    m = 0
    for st in graph:
        for _, nxt in _edges_in_map_order(st):
            m = max(m, nxt)
    return m


def dfa_to_regex_def_text(graph: List[dict]) -> str:
    accepted = get_accepted_state(graph)
    if accepted is None:
        raise VrmError("No accepted state")
    lines = ["0", str(accepted), str(get_max_state(graph))]
    for i, st in enumerate(graph):
        for key, nxt in _edges_in_map_order(st):
            for ch in json.loads(key):
                lines.append("%d %d %d" % (i, nxt, ord(ch[0]) & 0xFF))
    return "\n".join(lines) + "\n"


# ---------------------------------------------------------------------------------------------------------------------
# Substring definitions (mod.rs:309-600).  The reference enumerates simple paths accepted-state -> state 0 over a
# petgraph `Graph` whose edges point backwards, removing self loops while it walks adjacency lists; which paths it
# finds depends on that container's edge numbering, so the container is modelled, not abstracted.
# ---------------------------------------------------------------------------------------------------------------------

_END = -1


class _BackGraph:
    """Directed multigraph with per-node singly linked out / in lists (newest edge first) and swap-remove edge storage."""

    def __init__(self, n_nodes: int):
        self.node_next = [[_END, _END] for _ in range(n_nodes)]
        self.e_node: List[List[int]] = []
        self.e_next: List[List[int]] = []
        self.e_weight: List[str] = []

    def add_edge(self, a: int, b: int, w: str) -> int:
        e = len(self.e_node)
        if a == b:
            nxt = list(self.node_next[a])
            self.node_next[a] = [e, e]
        else:
            nxt = [self.node_next[a][0], self.node_next[b][1]]
            self.node_next[a][0] = e
            self.node_next[b][1] = e
        self.e_node.append([a, b])
        self.e_next.append(nxt)
        self.e_weight.append(w)
        return e

    def _edge(self, e: int) -> bool:
        return 0 <= e < len(self.e_node)

    def find_edge(self, a: int, b: int) -> Optional[int]:
        e = self.node_next[a][0]
        while self._edge(e):
            if self.e_node[e][1] == b:
                return e
            e = self.e_next[e][0]
        return None

    def _relink(self, nodes: List[int], e: int, nxt: List[int]) -> None:
        for k in (0, 1):
            head = self.node_next[nodes[k]]
            if head[k] == e:
                head[k] = nxt[k]
            else:
                cur = head[k]
                while self._edge(cur):
                    if self.e_next[cur][k] == e:
                        self.e_next[cur][k] = nxt[k]
                        break
                    cur = self.e_next[cur][k]

    def remove_edge(self, e: int) -> None:
        if not self._edge(e):
            return
        self._relink(self.e_node[e], e, list(self.e_next[e]))
        last = len(self.e_node) - 1
        if e != last:
            self.e_node[e], self.e_next[e], self.e_weight[e] = self.e_node[last], self.e_next[last], self.e_weight[last]
        self.e_node.pop()
        self.e_next.pop()
        self.e_weight.pop()
        if e != last:
            self._relink(self.e_node[e], last, [e, e])


@dataclass
class RegexPartConfig:
    """mod.rs:39-50."""
    is_public: bool
    regex_def: str
    max_size: int
    solidity: Optional[dict] = None


@dataclass
class DecomposedRegexConfig:
    """mod.rs:30-37."""
    max_byte_size: int
    parts: List[RegexPartConfig]

    @classmethod
    def from_json(cls, text: str) -> "DecomposedRegexConfig":
        v = json.loads(text)
        return cls(int(v["max_byte_size"]),
                   [RegexPartConfig(bool(p["is_public"]), str(p["regex_def"]), int(p["max_size"]), p.get("solidity"))
                    for p in v["parts"]])

    @classmethod
    def from_file(cls, path) -> "DecomposedRegexConfig":
        return cls.from_json(Path(path).read_text(encoding="utf-8"))

    # -- mod.rs:309-527 -------------------------------------------------------------------------------------------
    def extract_substr_ids(self, graph: List[dict]):
        max_state = get_max_state(graph)
        g = _BackGraph(max_state + 1)
        for i, st in enumerate(graph):
            for key, nxt in _edges_in_map_order(st):
                chars = json.loads(key)
                assert all(len(c.encode("utf-8")) == 1 for c in chars)
                g.add_edge(nxt, i, "".join(chars))
        accepted = get_accepted_state(graph)
        if accepted is None:
            raise VrmError("No accepted state")

        self_char: Dict[int, int] = {}
        for s in range(max_state + 1):
            e = g.find_edge(s, s)
            if e is not None:
                self_char[s] = ord(g.e_weight[e][0])

        self_nodes: Set[int] = set()
        paths: List[List[int]] = []
        stack: List[Tuple[int, List[int]]] = [(accepted, [accepted])]
        while stack:
            node, path = stack.pop()
            walk = g.node_next[node][0]
            while g._edge(walk):
                e = walk
                walk = g.e_next[e][0]
                parent = g.e_node[e][1]
                if parent == node:
                    self_nodes.add(node)
                    g.remove_edge(e)
                    continue
                if parent not in path:
                    if parent == 0:
                        paths.append(list(path))
                        continue
                    stack.append((parent, path + [parent]))

        public_idx = [i for i, p in enumerate(self.parts) if p.is_public]
        patterns: List[str] = []
        for i, p in enumerate(self.parts):
            patterns.append((patterns[i - 1] if i else "") + format_regex_printable(p.regex_def))
        compiled = [re.compile(p) for p in patterns]

        defs: List[Set[Tuple[int, int]]] = [set() for _ in public_idx]
        ends: List[Tuple[Set[int], Set[int]]] = [(set(), set()) for _ in public_idx]
        for path in paths:
            full = path + [0]
            labels = []
            for a, b in zip(full[:-1], full[1:]):
                e = g.find_edge(a, b)
                if e is None:
                    raise VrmError("No edge from %d to %d in the graph" % (a, b))
                labels.append(g.e_weight[e])
            states = full[::-1]
            labels = labels[::-1]
            for k, (sub_states, sub) in enumerate(self._substr_defs_from_path(states, labels, compiled, public_idx)):
                d = defs[k]
                ends[k][0].add(sub_states[0])
                ends[k][1].add(sub_states[-1])
                for j in range(len(sub_states) - 1):
                    d.add((sub_states[j], sub_states[j + 1]))
                    if sub_states[j] in self_nodes:
                        d.add((sub_states[j], sub_states[j]))
                    for pre in range(j + 1):
                        if g.find_edge(sub_states[pre], sub_states[j + 1]) is not None:
                            d.add((sub_states[j + 1], sub_states[pre]))
                tail = sub_states[-1]
                if tail in self_nodes:
                    if compiled[public_idx[k]].search(sub + chr(self_char[tail])) is not None:
                        d.add((tail, tail))
        return defs, ends, public_idx

    # -- mod.rs:529-600 -------------------------------------------------------------------------------------------
    @staticmethod
    def _substr_defs_from_path(states: List[int], labels: List[str], compiled, public_idx):
        assert len(states) == len(labels) + 1
        concat = "".join(lab[0] for lab in labels)
        index_ends = []
        for rx in compiled:
            m = rx.search(concat)
            if m is None:
                raise VrmError("part pattern %r does not match the path string %r" % (rx.pattern, concat))
            index_ends.append(m.end() + 1 if m.start() == m.end() else m.end())
        out = []
        for idx in public_idx:
            start = 0 if idx == 0 else index_ends[idx - 1]
            end = index_ends[idx]
            out.append((states[start:end + 1], concat[:end]))
        return out

    # -- mod.rs:63-90, 265-307 ------------------------------------------------------------------------------------
    def gen_regex_texts(self) -> Tuple[str, List[str]]:
        """The contents `gen_regex_files` writes: (allstr text, [substr text per public part])."""
        graph = get_dfa_json_value("".join(p.regex_def for p in self.parts))
        allstr = dfa_to_regex_def_text(graph)
        defs, ends, public_idx = self.extract_substr_ids(graph)
        substrs = []
        for k, d in enumerate(defs):
            lines = [str(self.parts[public_idx[k]].max_size), "0", str(self.max_byte_size - 1),
                     "".join("%d " % s for s in sorted(ends[k][0])),
                     "".join("%d " % s for s in sorted(ends[k][1]))]
            lines += ["%d %d" % t for t in sorted(d)]
            substrs.append("\n".join(lines) + "\n")
        return allstr, substrs

    def gen_regex_files(self, allstr_file_path, substr_file_pathes: Sequence) -> None:
        allstr, substrs = self.gen_regex_texts()
        Path(allstr_file_path).write_text(allstr, encoding="utf-8")
        for k, text in enumerate(substrs):
            Path(substr_file_pathes[k]).write_text(text, encoding="utf-8")


def main(argv: Optional[Sequence[str]] = None) -> int:
    """`vrm gen-halo2-texts` (/root/reference/src/bin/vrm.rs:22-70)."""
    import argparse
    ap = argparse.ArgumentParser(prog="python -m halo2_regex_b200.vrm")
    sub = ap.add_subparsers(dest="command", required=True)
    g = sub.add_parser("gen-halo2-texts")
    g.add_argument("-d", "--decomposed-regex-path", required=True)
    g.add_argument("-a", "--allstr-file-path", required=True)
    g.add_argument("-s", "--substrs-dir-path", required=True)
    args = ap.parse_args(argv)
    cfg = DecomposedRegexConfig.from_file(args.decomposed_regex_path)
    n_public = sum(1 for p in cfg.parts if p.is_public)
    cfg.gen_regex_files(args.allstr_file_path,
                        [Path(args.substrs_dir_path) / ("substr%d.txt" % i) for i in range(n_public)])
    return 0


if __name__ == "__main__":
    raise SystemExit(main())
