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
