"""Second, independently written restatement of the reference path in pure Python (dict / set / list shapes).
TEST INFRASTRUCTURE ONLY — used to cross-check oracle.c on small inputs and to re-check the reference's constraint
system (first-state gate, transition lookup, endpoint lookups, accept rule) on the generated witness.

Follows reference src/defs.rs:75-110, 209-265; src/table.rs:101-122, 129-193; src/lib.rs:311-888.
"""
import re

_WS = re.compile(r"\s+", re.UNICODE)


def _u64(tok):
    if not re.fullmatch(r"\+?[0-9]+", tok, re.ASCII):
        raise ValueError(tok)
    v = int(tok)
    if v >= 1 << 64:
        raise ValueError(tok)
    return v


def _lines(text):
    if isinstance(text, bytes):
        text = text.decode("utf-8")
    parts = text.split("\n")
    if parts and parts[-1] == "":
        parts.pop()
    return [p[:-1] if p.endswith("\r") else p for p in parts]


class PyAllstr:
    """AllstrRegexDef, src/defs.rs:26-36, read_from_reader :75-110"""

    def __init__(self, text):
        self.state_lookup = {}
        self.first_state_val = self.accepted_state_val = self.largest_state_val = 0
        for idx, line in enumerate(_lines(text)):
            el = [_u64(t) for t in line.split()]
            if idx == 0:
                self.first_state_val = el[0]
            elif idx == 1:
                self.accepted_state_val = el[0]
            elif idx == 2:
                self.largest_state_val = el[0]
            else:
                self.state_lookup[(el[2] & 0xFF, el[0])] = (idx, el[1])


class PySubstr:
    """SubstrRegexDef, src/defs.rs:115-132, read_from_reader :209-265"""

    def __init__(self, text):
        self.valid_state_transitions = set()
        self.max_length = self.min_position = self.max_position = 0
        self.start_states, self.end_states = [], []
        for idx, line in enumerate(_lines(text)):
            el = [_u64(t) for t in line.split()]
            if idx == 0:
                self.max_length = el[0]
            elif idx == 1:
                self.min_position = el[0]
            elif idx == 2:
                self.max_position = el[0]
            elif idx == 3:
                self.start_states = el
            elif idx == 4:
                self.end_states = el
            else:
                self.valid_state_transitions.add((el[0], el[1]))


def table_rows(allstr, substrs, substr_id_offset):
    """RegexTableConfig::load, src/table.rs:61-198 → (transition rows, endpoint rows)"""
    dummy = allstr.largest_state_val + 1
    rows = [(0, dummy, dummy, 0)]
    for (ch, cur), (idx, nxt) in sorted(allstr.state_lookup.items(), key=lambda kv: kv[1][0]):
        sid = 0
        for j, sd in enumerate(substrs):
            if (cur, nxt) in sd.valid_state_transitions:
                sid = substr_id_offset + j
                break
        rows.append((ch, cur, nxt, sid))
    erows = [(0, dummy, dummy)]
    for j, sd in enumerate(substrs):
        for s in sd.start_states:
            erows.append((substr_id_offset + j, s, dummy))
        for e in sd.end_states:
            erows.append((substr_id_offset + j, dummy, e))
    return rows, erows


class InvalidTransition(Exception):
    pass


def match_substrs(regex_defs, max_chars_size, characters):
    """regex_defs: list of (PyAllstr, [PySubstr]).  Returns a dict of per-row lists (M rows)."""
    M, L, D = max_chars_size, len(characters), len(regex_defs)
    assert L <= M - 1
    # derive_states :804-823
    states = []
    for d, (a, _) in enumerate(regex_defs):
        st = [a.first_state_val]
        for c in characters:
            nxt = a.state_lookup.get((c, st[-1]))
            if nxt is None:
                raise InvalidTransition(f"The transition from {st[-1]} by {c} is invalid!", d, len(st) - 1, st[-1], c)
            st.append(nxt[1])
        states.append(st)
    # derive_substr_ids :825-845
    substr_ids, off = [], 1
    for d, (_, subs) in enumerate(regex_defs):
        ids = [0] * L
        for i in range(L):
            for k, sd in enumerate(subs):
                if (states[d][i], states[d][i + 1]) in sd.valid_state_transitions:
                    ids[i] = off + k
                    break
        substr_ids.append(ids)
        off += len(subs)
    # derive_is_start_end :847-888
    is_starts, is_ends, off = [], [], 1
    for d, (_, subs) in enumerate(regex_defs):
        s = [sid != 0 and states[d][i] in subs[sid - off].start_states for i, sid in enumerate(substr_ids[d])] + [False]
        e = [False] + [sid != 0 and states[d][i + 1] in subs[sid - off].end_states for i, sid in enumerate(substr_ids[d])]
        is_starts.append(s)
        is_ends.append(e)
        off += len(subs)
    enable = [1] * L + [0] * (M - L)
    chars = list(characters) + [0] * (M - L)
    sid_sum = [0] * M
    is_start_sum = [0] * (M + 1)
    is_end_sum = [0] * (M + 1)
    out = {"enable": enable, "chars": chars, "states": [], "substr_ids": [], "start_enable": [], "end_enable": [], "accepted": []}
    for d, (a, _) in enumerate(regex_defs):
        dummy = a.largest_state_val + 1
        state_values = states[d][:L]
        sid_values = list(substr_ids[d])
        is_start_values = is_starts[d][:L]
        is_end_values = is_ends[d][:L]
        for idx in range(L, M):                                    # :404-418
            sid_values.append(0)
            if idx == L:
                state_values.append(states[d][idx]); is_start_values.append(is_starts[d][idx]); is_end_values.append(is_ends[d][idx])
            else:
                state_values.append(dummy); is_start_values.append(False); is_end_values.append(False)
        out["states"].append(state_values)
        out["substr_ids"].append(sid_values)
        for i in range(M):
            sid_sum[i] += sid_values[i]
        out["start_enable"].append([enable[i] * int(is_start_values[i]) for i in range(M)])
        for i in range(M):
            is_start_sum[i] += int(is_start_values[i])
        ee = [0] * M
        for i in range(M - 1):                                     # :501-519
            ee[i] = enable[i] * int(is_end_values[i + 1])
            is_end_sum[i + 1] += int(is_end_values[i + 1])
        out["end_enable"].append(ee)
        out["accepted"].append(states[d][L] == a.accepted_state_val)
    sel = lambda a, b, s: s * (a - b) + b  # noqa: E731  halo2-base select(a,b,sel)
    start_mask, last = [], 0
    for idx in range(M):                                           # :598-645
        pre = 0 if idx == 0 else sid_sum[idx - 1]
        chg = 1 - int(pre == sid_sum[idx])
        is_set = is_start_sum[idx] * chg
        is_reset = (1 - is_start_sum[idx]) * is_end_sum[idx] * chg
        m = sel(0, sel(1, last, is_set), is_reset)
        start_mask.append(m)
        last = m
    end_mask, last = [], 0
    for idx in range(M):                                           # :663-714
        pre = 0 if idx == 0 else sid_sum[M - idx]
        chg = 1 - int(pre == sid_sum[M - 1 - idx])
        is_set = is_end_sum[M - idx] * chg
        is_reset = (1 - is_end_sum[M - idx]) * is_start_sum[M - idx] * chg
        m = sel(0, sel(1, last, is_set), is_reset)
        end_mask.append(m)
        last = m
    end_mask.reverse()
    mask = [start_mask[i] * end_mask[i] for i in range(M)]
    out["mask"] = mask
    out["masked_chars"] = [mask[i] * chars[i] for i in range(M)]    # :740-764
    out["masked_substr_ids"] = [mask[i] * sid_sum[i] for i in range(M)]
    out["overlap"] = any(v > 1 for v in is_start_sum) or any(v > 1 for v in is_end_sum)
    return out


def check_constraints(regex_defs, max_chars_size, w):
    """Re-check what MockProver checks on the witness `w` of match_substrs (src/lib.rs:173-284, 442-457).
    Returns (satisfied_except_accept, accepted)."""
    M = max_chars_size
    off = 1
    ok = True
    for d, (a, subs) in enumerate(regex_defs):
        rows, erows = table_rows(a, subs, off)
        rows, erows = set(rows), set(erows)
        dummy = a.largest_state_val + 1
        st, sid, en, ch = w["states"][d], w["substr_ids"][d], w["enable"], w["chars"]
        if en[0] and st[0] != a.first_state_val:                   # q_first gate :173-191
            ok = False
        for i in range(M):
            nxt = st[i + 1] if i + 1 < M else 0
            tup = (en[i] * ch[i], en[i] * st[i] + (1 - en[i]) * dummy, en[i] * nxt + (1 - en[i]) * dummy, en[i] * sid[i])
            if tup not in rows:                                    # :207-233
                ok = False
            se, ee = w["start_enable"][d][i], w["end_enable"][d][i]
            if (se * sid[i], se * st[i] + (1 - se) * dummy, dummy) not in erows:   # :235-258
                ok = False
            if (ee * sid[i], dummy, ee * nxt + (1 - ee) * dummy) not in erows:     # :260-284
                ok = False
        off += len(subs)
    return ok, all(w["accepted"])
