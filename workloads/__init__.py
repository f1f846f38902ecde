"""Seeded synthetic string batches for the BASELINE.json configs.

Every generator is a pure function of (seed, index) built from splitmix64's finaliser, written once over an array
namespace so that numpy (CPU, tests / CPU baseline) and torch (GPU, bench) produce byte-identical batches.
"""
import numpy as np

GOLDEN = 0x9E3779B97F4A7C15
# the 98-byte alphabet the test_regexes DFAs are total over: \t \n \r and 0x20..0x7e (fixture facts)
ALPHABET = bytes([9, 10, 13] + list(range(0x20, 0x7F)))
assert len(ALPHABET) == 98

SEED_CONFIG1 = 0xB2000001
SEED_CONFIG2 = 0xB2000002
SEED_CONFIG3 = 0xB2000003
SEED_CONFIG4 = 0xB2000004


class _NP:
    """uint64 arithmetic in numpy"""
    name = "numpy"

    @staticmethod
    def arange(n, start=0):
        return np.arange(start, start + n, dtype=np.uint64)

    @staticmethod
    def u64(x):
        return np.uint64(x & 0xFFFFFFFFFFFFFFFF)

    @staticmethod
    def shr(x, k):
        return x >> np.uint64(k)

    @staticmethod
    def mul(x, c):
        with np.errstate(over="ignore"):
            return x * np.uint64(c & 0xFFFFFFFFFFFFFFFF)

    @staticmethod
    def add(x, c):
        with np.errstate(over="ignore"):
            return x + (np.uint64(c & 0xFFFFFFFFFFFFFFFF) if isinstance(c, int) else c)

    @staticmethod
    def mod(x, m):
        return x % np.uint64(m)

    @staticmethod
    def band(x, m):
        return x & np.uint64(m)


class _TorchI64:
    """the same arithmetic on torch int64 (two's complement wrap-around == uint64 wrap-around)"""
    name = "torch"

    def __init__(self, device):
        import torch
        self.t, self.device = torch, device

    def arange(self, n, start=0):
        return self.t.arange(start, start + n, dtype=self.t.int64, device=self.device)

    @staticmethod
    def _s(c):
        c &= 0xFFFFFFFFFFFFFFFF
        return c - (1 << 64) if c >= (1 << 63) else c

    def u64(self, x):
        return self._s(x)

    def shr(self, x, k):  # logical shift right
        return (x >> k) & ((1 << (64 - k)) - 1)

    def mul(self, x, c):
        return x * self._s(c)

    def add(self, x, c):
        return x + (self._s(c) if isinstance(c, int) else c)

    def mod(self, x, m):  # x is non-negative here (callers shift first)
        return x % m

    def band(self, x, m):
        return x & self._s(m)


def _mix(ns, z):
    """splitmix64 finaliser"""
    z = ns.mul(z ^ ns.shr(z, 30), 0xBF58476D1CE4E5B9)
    z = ns.mul(z ^ ns.shr(z, 27), 0x94D049BB133111EB)
    return z ^ ns.shr(z, 31)


def _rand(ns, seed, idx):
    """counter-based random u64: element i of the splitmix64 stream started at `seed`"""
    return _mix(ns, ns.add(ns.mul(ns.add(idx, 1), GOLDEN), seed))


def _filler(ns, seed, n_bytes, start=0):
    """n_bytes (multiple of 8) bytes uniform-ish over ALPHABET as an index array into ALPHABET (values 0..97)"""
    assert n_bytes % 8 == 0 and start % 8 == 0
    r = _rand(ns, seed, ns.arange(n_bytes // 8, start // 8))
    cols = [ns.mod(ns.band(ns.shr(r, 8 * k), 0xFF) if k else ns.band(r, 0xFF), 98) for k in range(8)]
    return cols  # 8 arrays, byte k of word i is position 8*i+k


def config1_numpy(n_strings, length=1024, seed=SEED_CONFIG1, first=0):
    """BASELINE config 1 (regex1+substr1): filler || 'email was meant for @' || [a-z]{1..4} || '.' || filler.
    Match offset uniform in [0, L-27]; 1 string in 16 has no match.  Returns (uint8 array (n, L), plan dict).
    `first` = index of the first string (so shards of one batch are slices of the same global batch)."""
    ns = _NP
    L = length
    assert L % 8 == 0 and L >= 32
    cols = _filler(ns, seed, n_strings * L, first * L)
    alpha = np.frombuffer(ALPHABET, dtype=np.uint8)
    data = np.empty((n_strings * L // 8, 8), dtype=np.uint8)
    for k in range(8):
        data[:, k] = alpha[cols[k]]
    data = data.reshape(n_strings, L)
    j = ns.arange(n_strings, first)
    r1 = _rand(ns, seed ^ 0x5DEECE66D, j)
    r2 = _rand(ns, seed ^ 0x1234567, j)
    has = ns.band(r1, 15) != 0
    name_len = (ns.band(ns.shr(r1, 4), 3) + np.uint64(1)).astype(np.int64)
    off = ns.mod(ns.shr(r1, 8), L - 27 + 1).astype(np.int64)
    prefix = np.frombuffer(b"email was meant for @", dtype=np.uint8)
    rows = np.nonzero(has)[0]
    for k in range(21):
        data[rows, off[rows] + k] = prefix[k]
    for k in range(4):
        sel = rows[name_len[rows] > k]
        ch = (ns.mod(ns.band(ns.shr(r2[sel], 8 * k), 0xFF), 26) + np.uint64(97)).astype(np.uint8)
        data[sel, off[sel] + 21 + k] = ch
    data[rows, off[rows] + 21 + name_len[rows]] = ord(".")
    return data, {"has_match": has, "offset": off, "name_len": name_len}


def config1_torch(n_strings, length=1024, seed=SEED_CONFIG1, first=0, device="cuda"):
    """Same batch as config1_numpy, generated on `device` (uint8 tensor (n, L))."""
    import torch
    ns = _TorchI64(device)
    L = length
    assert L % 8 == 0 and L >= 32
    alpha = torch.tensor(list(ALPHABET), dtype=torch.uint8, device=device)
    data = torch.empty((n_strings * L // 8, 8), dtype=torch.uint8, device=device)
    step = 1 << 24  # words per slab: bounds the int64 temporaries
    for w0 in range(0, n_strings * L // 8, step):
        w1 = min(n_strings * L // 8, w0 + step)
        r = _rand(ns, seed, ns.arange(w1 - w0, first * L // 8 + w0))
        for k in range(8):
            b = ns.band(ns.shr(r, 8 * k), 0xFF) if k else ns.band(r, 0xFF)
            data[w0:w1, k] = alpha[ns.mod(b, 98)]
    data = data.reshape(n_strings, L)
    j = ns.arange(n_strings, first)
    r1 = _rand(ns, seed ^ 0x5DEECE66D, j)
    r2 = _rand(ns, seed ^ 0x1234567, j)
    has = ns.band(r1, 15) != 0
    name_len = ns.band(ns.shr(r1, 4), 3) + 1
    off = ns.mod(ns.shr(r1, 8), L - 27 + 1)
    prefix = torch.tensor(list(b"email was meant for @"), dtype=torch.uint8, device=device)
    rows = torch.nonzero(has).squeeze(1)
    for k in range(21):
        data[rows, off[rows] + k] = prefix[k]
    for k in range(4):
        sel = rows[name_len[rows] > k]
        ch = (ns.mod(ns.band(ns.shr(r2[sel], 8 * k), 0xFF), 26) + 97).to(torch.uint8)
        data[sel, off[sel] + 21 + k] = ch
    data[rows, off[rows] + 21 + name_len[rows]] = ord(".")
    return data


def ragged_from_fixed(data, seed, min_len=0):
    """Parity-only variant of a fixed-length batch: string j keeps its first len_j bytes, len_j uniform in
    [min_len, L] (second set).  Returns (flat uint8 array, uint64 offsets)."""
    n, L = data.shape
    r = _rand(_NP, seed ^ 0x7A66ED, _NP.arange(n))
    lens = (np.uint64(min_len) + _NP.mod(_NP.shr(r, 11), L - min_len + 1)).astype(np.int64)
    offs = np.zeros(n + 1, dtype=np.uint64)
    offs[1:] = np.cumsum(lens)
    mask = np.arange(L)[None, :] < lens[:, None]
    return data[mask], offs


# ---- BASELINE config 2: regex3 (+ substr defs), "from:" header lines ------------------------------------------------------
def config2_numpy(n_strings, length=1024, seed=SEED_CONFIG2, first=0):
    """optional `filler\\r\\n` || 'from:' || optional `name<` || local@domain || optional '>' || '\\r\\n', padded IN FRONT with
    filler to `length` bytes.  regex3_test accepts exactly the strings whose last header line is a
    well-formed from: line; the filler before it must end in \\r\\n unless the from: line starts the string, so the padding is a
    filler line terminated by \\r\\n.  Returns (uint8 array (n, L), plan dict)."""
    ns = _NP
    L = length
    assert L % 8 == 0 and L >= 96
    lower = np.frombuffer(b"abcdefghijklmnopqrstuvwxyz", dtype=np.uint8)
    # filler without \r and \n (they would start header lines of their own): map them to spaces
    cols = _filler(ns, seed, n_strings * L, first * L)
    alpha = np.frombuffer(ALPHABET, dtype=np.uint8).copy()
    alpha[alpha == 10] = 32
    alpha[alpha == 13] = 32
    data = np.empty((n_strings * L // 8, 8), dtype=np.uint8)
    for k in range(8):
        data[:, k] = alpha[cols[k]]
    data = data.reshape(n_strings, L)
    j = ns.arange(n_strings, first)
    r1 = _rand(ns, seed ^ 0x5DEECE66D, j)
    r2 = _rand(ns, seed ^ 0x1234567, j)
    r3 = _rand(ns, seed ^ 0x7654321, j)
    with_name = (ns.band(r1, 1) != 0)
    name_len = (ns.band(ns.shr(r1, 1), 7) + np.uint64(1)).astype(np.int64)       # 1..8
    local_len = (ns.band(ns.shr(r1, 4), 7) + np.uint64(1)).astype(np.int64)      # 1..8
    dom_len = (ns.band(ns.shr(r1, 7), 7) + np.uint64(2)).astype(np.int64)        # 2..9
    addr = np.zeros((n_strings, 24), dtype=np.uint8)
    starts = np.zeros(n_strings, dtype=np.int64)
    for i in range(n_strings):   # small per-string assembly (tests / CPU baseline sizes)
        nm = bytes(lower[(int(r2[i]) >> (5 * k)) % 26] for k in range(int(name_len[i])))
        lo = bytes(lower[(int(r3[i]) >> (5 * k)) % 26] for k in range(int(local_len[i])))
        do = bytes(lower[(int(r3[i]) >> (5 * (k + 8))) % 26] for k in range(int(dom_len[i]) - 4 if dom_len[i] > 5 else 1)) + b".com"
        tail = b"from:" + ((nm + b"<" + lo + b"@" + do + b">") if with_name[i] else (lo + b"@" + do)) + b"\r\n"
        pos = L - len(tail)
        data[i, pos - 2:pos] = (13, 10)                                         # the filler line ends in \r\n
        data[i, pos:] = np.frombuffer(tail, dtype=np.uint8)
        starts[i] = pos + 5 + (len(nm) + 1 if with_name[i] else 0)
        a = lo + b"@" + do
        addr[i, :len(a)] = np.frombuffer(a, dtype=np.uint8)
    return data, {"addr_start": starts, "addr": addr, "with_name": with_name}


# ---- BASELINE config 4: a synthetic DFA with >= 512 states in the reference's lookup-text format ---------------------------
HEADER_NAMES = """accept accept-language alternate-recipient archived-at authentication-results auto-submitted autoforwarded
autosubmitted bcc cc comments content-description content-disposition content-id content-identifier content-language
content-location content-md5 content-return content-transfer-encoding content-type conversion conversion-with-loss date
deferred-delivery delivered-to delivery-date discarded-x400-ipms-extensions discarded-x400-mts-extensions disclose-recipients
disposition-notification-options disposition-notification-to dkim-signature dl-expansion-history encoding encrypted expires
expiry-date from generate-delivery-report importance in-reply-to incomplete-copy keywords language latest-delivery-time
list-archive list-help list-id list-owner list-post list-subscribe list-unsubscribe message-context message-id message-type
mime-version mmhs-primary-precedence obsoletes organization original-encoded-information-types original-from
original-message-id original-recipient originator-return-address pics-label prevent-nondelivery-report priority received
received-spf references reply-by reply-to require-recipient-valid-since resent-bcc resent-cc resent-date resent-from
resent-message-id resent-sender resent-to return-path sender sensitivity subject supersedes to x400-content-identifier
x400-content-return x400-content-type x400-mts-identifier x400-originator x400-received x400-recipients x400-trace""".split()


def large_dfa_texts():
    """(allstr text, substr text) of a DFA with >= 512 states: the Aho-Corasick automaton of `\\r\\n<header-name>:` for the
    header names above, made total over ALPHABET, plus an address sub-automaton behind `\\r\\nfrom:`; accepting state = the
    state after a complete from: address line.  One substring def: the address after `\\r\\nfrom:`."""
    words = [b"\r\n" + h.encode() + b":" for h in HEADER_NAMES]
    goto, fail, term = [{}], [0], [None]
    for w in words:
        s = 0
        for ch in w:
            if ch not in goto[s]:
                goto.append({}); fail.append(0); term.append(None)
                goto[s][ch] = len(goto) - 1
            s = goto[s][ch]
        term[s] = w
    from collections import deque
    q = deque(goto[0].values())
    while q:
        s = q.popleft()
        for ch, t in goto[s].items():
            f = fail[s]
            while f and ch not in goto[f]:
                f = fail[f]
            fail[t] = goto[f].get(ch, 0) if goto[f].get(ch, 0) != t else 0
            q.append(t)
    n_trie = len(goto)

    def delta(s, ch):
        while s and ch not in goto[s]:
            s = fail[s]
        return goto[s].get(ch, 0)
    from_state = [i for i, w in enumerate(term) if w == b"\r\nfrom:"][0]
    A = n_trie          # inside the address
    DONE_R = n_trie + 1  # address followed by \r
    ACC = n_trie + 2     # ... and \n: accepted; behaves like the trie state of "\r\n" afterwards
    addr_chars = set(b"abcdefghijklmnopqrstuvwxyz0123456789._@-")
    crlf_state = delta(delta(0, 13), 10)
    lines = []
    for s in range(n_trie + 3):
        for ch in ALPHABET:
            if s < n_trie:
                t = A if (s == from_state and ch in addr_chars) else delta(s, ch)
            elif s == A:
                t = A if ch in addr_chars else (DONE_R if ch == 13 else delta(0, ch))
            elif s == DONE_R:
                t = ACC if ch == 10 else delta(delta(0, 13), ch)
            else:
                t = delta(crlf_state, ch)
            lines.append(f"{s} {t} {ch}\n")
    S = n_trie + 3
    allstr = f"0\n{ACC}\n{S - 1}\n" + "".join(lines)
    substr = f"64\n0\n4096\n{from_state}\n{A}\n{from_state} {A}\n{A} {A}\n"
    return allstr.encode(), substr.encode(), {"states": S, "from_state": from_state, "addr_state": A, "accept": ACC}


def config4_numpy(n_strings, length=4096, seed=SEED_CONFIG4, first=0, want_mask=False, _background=None):
    """Header blocks for the large DFA: lines `<header-name>: <filler>\\r\\n`, one of them (1 string in 8: none) a from: line with
    an address; the last line of 7 strings in 8 is the from: line (accepted).  Returns (uint8 array (n, L), plan).
    want_mask: plan["structural"] marks the bytes that are not free filler (header names, line ends, the from: line)."""
    ns = _NP
    L = length
    if _background is not None:
        data = _background.copy()
    else:
        cols = _filler(ns, seed, n_strings * L, first * L)
        alpha = np.frombuffer(ALPHABET, dtype=np.uint8).copy()
        alpha[alpha == 10] = 32
        alpha[alpha == 13] = 32
        data = np.empty((n_strings * L // 8, 8), dtype=np.uint8)
        for k in range(8):
            data[:, k] = alpha[cols[k]]
        data = data.reshape(n_strings, L)
    j = ns.arange(n_strings, first)
    r1 = _rand(ns, seed ^ 0x5DEECE66D, j)
    names = [h.encode() for h in HEADER_NAMES]
    has_from = np.zeros(n_strings, dtype=bool)
    filler_only = data.copy() if want_mask else None
    for i in range(n_strings):
        r = int(r1[i])
        pos = 0
        k = 0
        while pos + 96 < L - 64:                       # header lines of 40..100 bytes
            nm = names[(r >> (k % 40)) % len(names)]
            if nm == b"from":
                nm = b"sender"
            ln = 40 + ((r >> (7 * (k % 8))) & 63)
            data[i, pos:pos + len(nm) + 1] = np.frombuffer(nm + b":", dtype=np.uint8)
            data[i, pos + ln - 2:pos + ln] = (13, 10)
            pos += ln
            k += 1
        if (r & 7) != 0:
            has_from[i] = True
            tail = b"from:" + bytes(97 + ((r >> (5 * q)) % 26) for q in range(4 + (r >> 40) % 8)) + b"@example.org\r\n"
            data[i, pos - 2:pos] = (13, 10)
            rest = L - pos - len(tail)
            data[i, pos:pos + rest] = 32
            data[i, pos:pos + 2] = (120, 58)            # "x:" -> an unknown header line absorbs the slack
            data[i, L - len(tail) - 2:L - len(tail)] = (13, 10)
            data[i, L - len(tail):] = np.frombuffer(tail, dtype=np.uint8)
    plan = {"has_from": has_from}
    if want_mask:
        # a second pass over the same skeleton on a complemented background: a byte is structural iff both passes agree on it
        other = config4_numpy(n_strings, length, seed, first, want_mask=False, _background=filler_only ^ 0xFF)[0]
        plan["structural"] = data == other
    return data, plan


# ---- device-side generators of the full-size bench batches ---------------------------------------------------------------------
def _filler_torch(n_bytes, start, seed, alpha, device):
    """n_bytes of filler (the same stream as _filler) mapped through `alpha` (98-entry uint8 tensor), as a flat uint8 tensor."""
    import torch
    ns = _TorchI64(device)
    assert n_bytes % 8 == 0 and start % 8 == 0
    data = torch.empty((n_bytes // 8, 8), dtype=torch.uint8, device=device)
    step = 1 << 24  # words per slab: bounds the int64 temporaries
    for w0 in range(0, n_bytes // 8, step):
        w1 = min(n_bytes // 8, w0 + step)
        r = _rand(ns, seed, ns.arange(w1 - w0, start // 8 + w0))
        for k in range(8):
            b = ns.band(ns.shr(r, 8 * k), 0xFF) if k else ns.band(r, 0xFF)
            data[w0:w1, k] = alpha[ns.mod(b, 98)]
    return data.reshape(-1)


def _alpha_no_crlf(device):
    import torch
    a = np.frombuffer(ALPHABET, dtype=np.uint8).copy()
    a[a == 10] = 32
    a[a == 13] = 32
    return torch.from_numpy(a).to(device)


def config2_torch(n_strings, length=1024, seed=SEED_CONFIG2, first=0, device="cuda", base=1 << 14):
    """BASELINE config 2 at bench size on `device`: every string has its own filler (the stream of config2_numpy); the from: line
    at its end is the one of string (index mod `base`) of config2_numpy (the skeleton repeats every `base` strings, the bytes in
    front of it never do, so nothing is served from L2 twice)."""
    import torch
    L = length
    data = _filler_torch(n_strings * L, first * L, seed, _alpha_no_crlf(device), device).reshape(n_strings, L)
    tails = torch.from_numpy(config2_numpy(base, L, seed)[0][:, L - 48:].copy()).to(device)   # "\r\n" + from: line <= 37 bytes
    idx = (torch.arange(n_strings, device=device, dtype=torch.int64) + first) % base
    data[:, L - 48:] = tails[idx]
    return data


def config4_torch(n_strings, length=4096, seed=SEED_CONFIG4, first=0, device="cuda", base=1 << 12):
    """BASELINE config 4 at bench size on `device`: the header skeleton (names, line ends, the from: line) of string
    (index mod `base`) of config4_numpy over filler of its own."""
    import torch
    L = length
    data = _filler_torch(n_strings * L, first * L, seed, _alpha_no_crlf(device), device).reshape(n_strings, L)
    b, plan = config4_numpy(base, L, seed, want_mask=True)
    skel, mask = torch.from_numpy(b).to(device), torch.from_numpy(plan["structural"]).to(device)
    step = max(base, (1 << 28) // L // base * base)                       # strings per slab, a multiple of the base
    for lo in range(0, n_strings, step):
        hi = min(n_strings, lo + step)
        idx = (torch.arange(lo, hi, device=device, dtype=torch.int64) + first) % base
        data[lo:hi] = torch.where(mask[idx], skel[idx], data[lo:hi])
    return data


def config3_torch(length=1 << 26, seed=SEED_CONFIG3, device="cuda"):
    """BASELINE config 3: ONE string of `length` bytes over ALPHABET with ` Also for xyz.` (regex2 / substr2) planted at a seeded
    offset.  Returns (uint8 tensor (length,), offset of the planted text)."""
    import torch
    alpha = torch.from_numpy(np.frombuffer(ALPHABET, dtype=np.uint8).copy()).to(device)
    data = _filler_torch((length + 7) // 8 * 8, 0, seed, alpha, device)[:length].contiguous()
    plant = b" Also for xyz."
    at = (seed * 2654435761) % max(1, length - 64)
    if length >= 64:
        data[at:at + len(plant)] = torch.tensor(list(plant), dtype=torch.uint8, device=device)
    return data, at


def config3_numpy(length, seed=SEED_CONFIG3):
    """The same string as config3_torch, on the host."""
    cols = _filler(_NP, seed, (length + 7) // 8 * 8)
    alpha = np.frombuffer(ALPHABET, dtype=np.uint8)
    data = np.empty(((length + 7) // 8, 8), dtype=np.uint8)
    for k in range(8):
        data[:, k] = alpha[cols[k]]
    data = data.reshape(-1)[:length].copy()
    plant = b" Also for xyz."
    at = (seed * 2654435761) % max(1, length - 64)
    if length >= 64:
        data[at:at + len(plant)] = np.frombuffer(plant, dtype=np.uint8)
    return data, at
