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

The files it produces are only reproducible if every tie in those algorithms is broken the way the reference breaks it,
so the restatement below keeps the reference's *orders* and nothing else of its shape: the insertion-ordered property
maps of the JS engine (integer-like keys first), the UTF-16 string order of `Array.prototype.sort`, the swap-remove edge
storage and newest-first adjacency lists of petgraph's `Graph`, and the leftmost-first match rule of the regex engine.
The known-answer test is byte-identical regeneration of `test_regexes/*_lookup.txt` from `test_regexes/*.json`
(tests/test_vrm.py, fixtures under tests/golden/regexes/).
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
        return s
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
# regex text -> syntax tree (regex.js:236-382)
# ---------------------------------------------------------------------------------------------------------------------

class RegexSyntaxError(ValueError):
    pass


@dataclass
class _Tok:
    ch: str
    escaped: bool


@dataclass
class _Ast:
    kind: str                           # "or" | "cat" | "star" | "empty" | "text"
    parts: List["_Ast"] = field(default_factory=list)
    sub: Optional["_Ast"] = None
    text: str = ""


def _tokenise(text: str) -> List[_Tok]:
    toks: List[_Tok] = []
    i = 0
    while i < len(text):
        if text[i] == "\\":
            if i + 1 >= len(text):
                raise RegexSyntaxError("Error: dangling escape at %d." % i)
            c = text[i + 1]
            toks.append(_Tok(_ESCAPES.get(c, c), True))
            i += 2
        else:
            toks.append(_Tok(text[i], False))
            i += 1
    return toks


def _is(tok: _Tok, ch: str) -> bool:
    return (not tok.escaped) and tok.ch == ch


def _parse_alternation(toks: Sequence[_Tok], begin: int) -> _Ast:
    if len(toks) == 0:
        raise RegexSyntaxError("Error: empty input at %d." % begin)
    parts: List[_Ast] = []
    depth = 0
    last = 0
    for i in range(len(toks) + 1):
        if i == len(toks) or (_is(toks[i], "|") and depth == 0):
            if last == 0 and i == len(toks):
                return _parse_sequence(toks, begin)
            parts.append(_parse_alternation(toks[last:i], begin + last))
            last = i + 1
        elif _is(toks[i], "("):
            depth += 1
        elif _is(toks[i], ")"):
            depth -= 1
    if len(parts) == 1:
        return parts[0]
    return _Ast("or", parts=parts)


def _parse_sequence(toks: Sequence[_Tok], begin: int) -> _Ast:
    if len(toks) == 0:
        raise RegexSyntaxError("Error: empty input at %d." % begin)
    parts: List[_Ast] = []
    i = 0
    n = len(toks)
    while i < n:
        t = toks[i]
        if _is(t, "("):
            last = i + 1
            i += 1
            depth = 1
            while i < n and depth != 0:
                if _is(toks[i], "("):
                    depth += 1
                elif _is(toks[i], ")"):
                    depth -= 1
                i += 1
            if depth != 0:
                raise RegexSyntaxError("Error: missing right bracket for %d." % (begin + last))
            i -= 1
            parts.append(_parse_alternation(toks[last:i], begin + last))
        elif _is(t, "*"):
            if not parts:
                raise RegexSyntaxError("Error: unexpected * at %d." % (begin + i))
            parts[-1] = _Ast("star", sub=parts[-1])
        elif _is(t, "+"):
            if not parts:
                raise RegexSyntaxError("Error: unexpected + at %d." % (begin + i))
            parts[-1] = _Ast("cat", parts=[parts[-1], _Ast("star", sub=parts[-1])])
        elif _is(t, "?"):
            if not parts:
                raise RegexSyntaxError("Error: unexpected + at %d." % (begin + i))
            parts[-1] = _Ast("or", parts=[parts[-1], _Ast("empty")])
        elif _is(t, EPS):
            parts.append(_Ast("empty"))
        else:
            parts.append(_Ast("text", text=t.ch))
        i += 1
    if len(parts) == 1:
        return parts[0]
    return _Ast("cat", parts=parts)


def parse_regex(text: str) -> _Ast:
    return _parse_alternation(_tokenise(text), 0)


# ---------------------------------------------------------------------------------------------------------------------
# syntax tree -> epsilon-NFA (regex.js:390-452).  State numbers are handed out in the order the reference's recursive
# construction first enters / leaves a state; the subset keys, and through them every later order, depend on them.
# ---------------------------------------------------------------------------------------------------------------------

class _NfaState:
    __slots__ = ("id", "accept", "edges")

    def __init__(self, accept: bool = False):
        self.id = -1
        self.accept = accept
        self.edges: List[Tuple[str, "_NfaState"]] = []


def _thompson(node: _Ast, start: _NfaState, end: _NfaState, count: int) -> int:
    if start.id < 0:
        start.id = count
        count += 1
    k = node.kind
    if k == "empty":
        start.edges.append((EPS, end))
    elif k == "text":
        start.edges.append((node.text, end))
    elif k == "cat":
        last = start
        for part in node.parts[:-1]:
            mid = _NfaState()
            count = _thompson(part, last, mid, count)
            last = mid
        count = _thompson(node.parts[-1], last, end, count)
    elif k == "or":
        for part in node.parts:
            s, e = _NfaState(), _NfaState()
            e.edges.append((EPS, end))
            start.edges.append((EPS, s))
            count = _thompson(part, s, e, count)
    elif k == "star":
        s, e = _NfaState(), _NfaState()
        e.edges.append((EPS, s))
        e.edges.append((EPS, end))
        start.edges.append((EPS, s))
        start.edges.append((EPS, end))
        count = _thompson(node.sub, s, e, count)
    else:  # pragma: no cover
        raise AssertionError(k)
    if end.id < 0:
        end.id = count
        count += 1
    return count


def regex_to_nfa(text: str) -> _NfaState:
    ast = parse_regex(text)
    start, accept = _NfaState(), _NfaState(accept=True)
    limit = sys.getrecursionlimit()
    sys.setrecursionlimit(max(limit, 100000))
    try:
        _thompson(ast, start, accept, 0)
    finally:
        sys.setrecursionlimit(limit)
    return start


# ---------------------------------------------------------------------------------------------------------------------
# subset construction (regex.js:460-552)
# ---------------------------------------------------------------------------------------------------------------------

class _DfaState:
    __slots__ = ("key", "items", "symbols", "accept", "trans", "id")

    def __init__(self, key, items, symbols, accept):
        self.key = key
        self.items: List[_NfaState] = items
        self.symbols: List[str] = symbols
        self.accept: bool = accept
        self.trans: Dict[str, "_DfaState"] = {}
        self.id = ""


def _alpha_count(n: int) -> str:
    """0 -> A, 25 -> Z, 26 -> AA ... (regex.js:517-527)."""
    s = ""
    while n >= 0:
        s = chr(ord("A") + n % 26) + s
        n = n // 26 - 1
    return s


def _closure(seed: Sequence[_NfaState]) -> _DfaState:
    seen: Dict[int, _NfaState] = {}
    stack: List[_NfaState] = []
    symbols: Set[str] = set()
    accept = False
    for s in seed:
        stack.append(s)
        seen[s.id] = s
        accept |= s.accept
    while stack:
        top = stack.pop()
        for sym, nxt in top.edges:
            if sym == EPS:
                if nxt.id not in seen:
                    seen[nxt.id] = nxt
                    stack.append(nxt)
                    accept |= nxt.accept
            else:
                symbols.add(sym)
    items = [seen[i] for i in sorted(seen)]
    return _DfaState(",".join(str(s.id) for s in items), items, _js_sorted(symbols), accept)


def _closed_move(state: _DfaState, symbol: str, memo: Dict[Tuple[int, ...], _DfaState]) -> _DfaState:
    nexts: Dict[int, _NfaState] = {}
    for item in state.items:
        for sym, nxt in item.edges:
            if sym == symbol:
                nexts[nxt.id] = nxt
    seed = tuple(sorted(nexts))             # a closure is a function of its seed set alone
    hit = memo.get(seed)
    if hit is None:
        hit = memo[seed] = _closure(list(nexts.values()))
    return hit


def nfa_to_dfa(nfa: _NfaState) -> _DfaState:
    first = _closure([nfa])
    first.id = _alpha_count(0)
    states = {first.key: first}
    memo: Dict[Tuple[int, ...], _DfaState] = {}
    queue = [first]
    front = 0
    while front < len(queue):
        top = queue[front]
        front += 1
        for sym in top.symbols:
            c = _closed_move(top, sym, memo)
            known = states.get(c.key)
            if known is None:
                c.id = _alpha_count(len(states))
                states[c.key] = c
                queue.append(c)
                known = c
            top.trans[sym] = known
    return first


# ---------------------------------------------------------------------------------------------------------------------
# Hopcroft minimisation and the merged-edge automaton (regex.js:560-762)
# ---------------------------------------------------------------------------------------------------------------------

class _MinState:
    __slots__ = ("number", "accept", "trans")

    def __init__(self, number: int, accept: bool):
        self.number = number                # 1-based, the reference's "nature"
        self.accept = accept
        self.trans: Dict[str, "_MinState"] = {}     # merged label -> target


def _reverse_edges(start: _DfaState):
    symbols: Dict[str, bool] = {}
    id_map: Dict[str, _DfaState] = {}
    rev: Dict[str, Dict[str, List[str]]] = {}
    visited = {start.id}
    queue = [start]
    front = 0
    while front < len(queue):
        top = queue[front]
        front += 1
        id_map[top.id] = top
        for sym in top.symbols:
            symbols.setdefault(sym, True)
            nxt = top.trans[sym]
            rev.setdefault(nxt.id, {}).setdefault(sym, []).append(top.id)
            if nxt.id not in visited:
                visited.add(nxt.id)
                queue.append(nxt)
    return _js_keys(symbols), id_map, rev


def _hopcroft(symbols: List[str], id_map: Dict[str, _DfaState], rev) -> List[List[str]]:
    ids = _js_sorted(id_map.keys())
    partitions: Dict[str, List[str]] = {}
    queue: List[Optional[str]] = []
    visited: Dict[str, int] = {}
    g1 = [i for i in ids if id_map[i].accept]
    g2 = [i for i in ids if not id_map[i].accept]
    key = ",".join(g1)
    partitions[key] = g1
    queue.append(key)
    visited[key] = 0
    if g2:
        key = ",".join(g2)
        partitions[key] = g2
        queue.append(key)
    front = 0
    while front < len(queue):
        top = queue[front]
        front += 1
        if not top:
            continue
        members = top.split(",")
        for sym in symbols:
            rev_group: Set[str] = set()
            for m in members:
                srcs = rev.get(m)
                if srcs is not None and sym in srcs:
                    rev_group.update(srcs[sym])
            for key in _js_keys(partitions):
                block = partitions[key]
                g1 = [x for x in block if x in rev_group]
                if not g1 or len(g1) == len(block):
                    continue
                g2 = [x for x in block if x not in rev_group]
                del partitions[key]
                k1, k2 = ",".join(g1), ",".join(g2)
                partitions[k1] = g1
                partitions[k2] = g2
                if k1 in visited:
                    queue[visited[k1]] = None
                    visited[k1] = len(queue)
                    queue.append(k1)
                    visited[k2] = len(queue)
                    queue.append(k2)
                elif len(g1) <= len(g2):
                    visited[k1] = len(queue)
                    queue.append(k1)
                else:
                    visited[k2] = len(queue)
                    queue.append(k2)
    return [partitions[k] for k in _js_keys(partitions)]


def _merge(start: _DfaState, partitions: List[List[str]], id_map, rev) -> List[_MinState]:
    partitions.sort(key=lambda p: _utf16_key(",".join(p)))
    for i, p in enumerate(partitions):
        if start.id in p:
            if i > 0:
                partitions[i], partitions[0] = partitions[0], partitions[i]
            break
    group: Dict[str, int] = {}
    nodes: List[_MinState] = []
    for i, p in enumerate(partitions):
        nodes.append(_MinState(i + 1, id_map[p[0]].accept))
        for x in p:
            group[x] = i
    edges: List[Dict[int, Set[str]]] = [dict() for _ in partitions]
    for to, by_sym in rev.items():
        for sym, sources in by_sym.items():
            for src in sources:
                edges[group[src]].setdefault(group[to], set()).add(sym)
    for frm, by_to in enumerate(edges):
        for to in sorted(by_to):
            nodes[frm].trans[_symbol_set_key(by_to[to])] = nodes[to]
    return nodes


def min_dfa(dfa: _DfaState) -> List[_MinState]:
    symbols, id_map, rev = _reverse_edges(dfa)
    return _merge(dfa, _hopcroft(symbols, id_map, rev), id_map, rev)


def regex_to_dfa(regex: str) -> List[dict]:
    """`regexToDfa` (regex.js:40-90) as the parsed value of the JSON it returns: one `{"type", "edges"}` object per
    state, edges keyed by the JSON text of the sorted list of characters that share a target."""
    nodes = min_dfa(nfa_to_dfa(regex_to_nfa(regex)))
    reach: Dict[int, _MinState] = {}
    stack = [nodes[0]]
    labels: Set[str] = set()
    while stack:
        top = stack.pop()
        if top.number in reach:
            continue
        reach[top.number] = top
        for label, nxt in top.trans.items():
            labels.add(label)
            stack.append(nxt)
    ordered = _js_sorted(labels)
    size = max(reach)
    graph: List[Optional[dict]] = [None] * size
    for number in sorted(reach):
        st = reach[number]
        graph[number - 1] = {"type": "accept" if st.accept else "",
                             "edges": {lab: st.trans[lab].number - 1 for lab in ordered if lab in st.trans}}
    return graph


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
