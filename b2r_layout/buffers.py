"""Caller-owned host buffers for one batch of witness columns (layout of include/b2r.h `b2r_outputs`).  Pure numpy / ctypes:
shared by the product package, the oracle wrapper and the benchmark; loads no native library."""
import ctypes as C

import numpy as np

from . import abi as _abi


def round_up(x, m):
    return (x + m - 1) // m * m


def aligned_empty(nbytes, align=64):
    """uint8 numpy array whose data pointer is `align`-byte aligned."""
    raw = np.empty(nbytes + align, dtype=np.uint8)
    off = (-raw.ctypes.data) % align
    return raw[off:off + nbytes]


class HostOutputs:
    """All witness columns of a batch as numpy arrays plus the matching `b2r_outputs` struct.

    Column semantics follow the reference's assignment: `states[d][j, i]`, `substr_ids[d][j, i]`,
    bitmaps `start_enable[d]`, `end_enable[d]` (LSB-first), `masked_chars`, `masked_substr_ids`; `mult[d]` has one
    u64 per row of RegexTableConfig::load's transition table (reference src/table.rs:101-122), `endpoint_mult[d]`
    two halves over the endpoint table rows (src/table.rs:129-193).
    """

    def __init__(self, n_strings, max_chars_size, state_widths, table_rows, endpoint_rows, row_pitch=None,
                 bitmap_pitch=None, max_records=8, compact_pitch=64, want=None, fill=None, allocator=None):
        n, m = int(n_strings), int(max_chars_size)
        self.n, self.m = n, m
        self.n_defs = len(state_widths)
        self.row_pitch = int(row_pitch) if row_pitch else round_up(m, 32)
        self.bitmap_pitch = int(bitmap_pitch) if bitmap_pitch else round_up((m + 7) // 8, 32)
        assert self.row_pitch >= m and self.row_pitch % 16 == 0
        assert self.bitmap_pitch >= (m + 7) // 8 and self.bitmap_pitch % 4 == 0
        want = want or {"states", "substr_ids", "start_enable", "end_enable", "masked_chars", "masked_substr_ids",
                        "status", "records", "compact_bytes", "mult", "endpoint_mult"}
        self.want = set(want)
        self.max_records, self.compact_pitch = int(max_records), int(compact_pitch)
        rp, bp = self.row_pitch, self.bitmap_pitch

        def col(nbytes):
            a = allocator(nbytes) if allocator else aligned_empty(nbytes)   # e.g. page-locked memory for fast D2H
            assert a.ctypes.data % 16 == 0
            if fill is not None:
                a[:] = fill
            return a

        def zeros(shape, dtype):
            # every buffer comes from the allocator: one pageable destination makes its D2H copy synchronous
            if not allocator:
                return np.zeros(shape, dtype=dtype)
            count = int(np.prod(shape))
            a = allocator(count * np.dtype(dtype).itemsize)
            a[:] = 0
            return a.view(dtype).reshape(shape)

        self.states, self.substr_ids, self.start_enable, self.end_enable, self.mult, self.endpoint_mult = [], [], [], [], [], []
        for d in range(self.n_defs):
            w = state_widths[d]
            self.states.append(col(n * rp * w).view(np.uint8 if w == 1 else np.uint16).reshape(n, rp) if "states" in self.want else None)
            self.substr_ids.append(col(n * rp).reshape(n, rp) if "substr_ids" in self.want else None)
            self.start_enable.append(col(n * bp).reshape(n, bp) if "start_enable" in self.want else None)
            self.end_enable.append(col(n * bp).reshape(n, bp) if "end_enable" in self.want else None)
            self.mult.append(zeros(table_rows[d], np.uint64) if "mult" in self.want else None)
            self.endpoint_mult.append(zeros(2 * endpoint_rows[d], np.uint64) if "endpoint_mult" in self.want else None)
        self.masked_chars = col(n * rp).reshape(n, rp) if "masked_chars" in self.want else None
        self.masked_substr_ids = col(n * rp).reshape(n, rp) if "masked_substr_ids" in self.want else None
        self.status = zeros(n, _abi.STATUS_DTYPE) if "status" in self.want else None
        self.records = zeros((n, self.max_records), _abi.RECORD_DTYPE) if "records" in self.want else None
        self.compact_bytes = zeros((n, self.compact_pitch), np.uint8) if "compact_bytes" in self.want else None

    @staticmethod
    def _p(a):
        return None if a is None else a.ctypes.data

    def all_arrays(self):
        cols = self.states + self.substr_ids + self.start_enable + self.end_enable + self.mult + self.endpoint_mult
        cols += [self.masked_chars, self.masked_substr_ids, self.status, self.records, self.compact_bytes]
        return [c for c in cols if c is not None]

    def struct(self, flags=0):
        o = _abi.Outputs()
        o.row_pitch, o.bitmap_pitch = self.row_pitch, self.bitmap_pitch
        for d in range(self.n_defs):
            o.states[d] = self._p(self.states[d])
            o.substr_ids[d] = self._p(self.substr_ids[d])
            o.start_enable[d] = self._p(self.start_enable[d])
            o.end_enable[d] = self._p(self.end_enable[d])
            o.mult[d] = self._p(self.mult[d])
            o.endpoint_mult[d] = self._p(self.endpoint_mult[d])
        o.masked_chars = self._p(self.masked_chars)
        o.masked_substr_ids = self._p(self.masked_substr_ids)
        o.status = self._p(self.status)
        o.records = self._p(self.records)
        o.max_records = self.max_records if self.records is not None else 0
        o.compact_pitch = self.compact_pitch if self.compact_bytes is not None else 0
        o.compact_bytes = self._p(self.compact_bytes)
        o.flags = flags
        return o

    # ---- views restricted to the M defined rows / bits -------------------------------------------------------
    def bits(self, bitmap):
        """(n, M) bool array from an LSB-first bitmap column."""
        return np.unpackbits(bitmap, axis=1, bitorder="little")[:, :self.m].astype(bool)

    def defined(self):
        """dict of every defined output restricted to rows [0, M), for exact comparison between implementations."""
        out = {}
        for d in range(self.n_defs):
            if self.states[d] is not None:
                out[f"states{d}"] = self.states[d][:, :self.m]
            if self.substr_ids[d] is not None:
                out[f"substr_ids{d}"] = self.substr_ids[d][:, :self.m]
            if self.start_enable[d] is not None:
                out[f"start_enable{d}"] = self.bits(self.start_enable[d])
            if self.end_enable[d] is not None:
                out[f"end_enable{d}"] = self.bits(self.end_enable[d])
            if self.mult[d] is not None:
                out[f"mult{d}"] = self.mult[d]
            if self.endpoint_mult[d] is not None:
                out[f"endpoint_mult{d}"] = self.endpoint_mult[d]
        if self.masked_chars is not None:
            out["masked_chars"] = self.masked_chars[:, :self.m]
        if self.masked_substr_ids is not None:
            out["masked_substr_ids"] = self.masked_substr_ids[:, :self.m]
        return out


def compare_outputs(a, b, skip_rows_mask=None):
    """Bit-exact comparison of two HostOutputs (rows [0,M) only).  Returns a list of mismatch descriptions.

    Strings flagged invalid-transition / too-long have unspecified rows and strings flagged overlap have unspecified
    masked outputs / records (include/b2r.h); they are excluded accordingly (the flags themselves must agree)."""
    errs = []
    sa, sb = a.status, b.status
    dead = (sa["flags"] & (_abi.B2R_ST_INVALID_TRANSITION | _abi.B2R_ST_TOO_LONG)) != 0
    overlap = (sa["flags"] & _abi.B2R_ST_OVERLAP) != 0
    for f in ("flags", "err_pos", "err_state", "err_byte", "err_def", "n_records", "n_compact"):
        keep = ~(dead | overlap) if f in ("n_records", "n_compact") else np.ones(len(sa), bool)
        if not np.array_equal(sa[f][keep], sb[f][keep]):
            bad = np.nonzero((sa[f] != sb[f]) & keep)[0]
            errs.append(f"status.{f}: {len(bad)} strings differ, first {bad[0]}: {sa[f][bad[0]]} vs {sb[f][bad[0]]}")
    da, db = a.defined(), b.defined()
    for k in da:
        if k.startswith("mult") or k.startswith("endpoint_mult"):
            if dead.any():  # the batch failed (the reference panics): multiplicities are unspecified
                continue
            if not np.array_equal(da[k], db[k]):
                bad = np.nonzero(da[k] != db[k])[0]
                errs.append(f"{k}: {len(bad)} bins differ, first {bad[0]}: {da[k][bad[0]]} vs {db[k][bad[0]]}")
            continue
        skip = dead | overlap if k.startswith("masked") else dead
        x, y = da[k][~skip], db[k][~skip]
        if not np.array_equal(x, y):
            bad = np.argwhere(x != y)
            errs.append(f"{k}: {len(bad)} cells differ, first (row-in-kept {bad[0][0]}, col {bad[0][1]}): {x[tuple(bad[0])]} vs {y[tuple(bad[0])]}")
    ok = ~(dead | overlap)
    if a.records is not None and b.records is not None:
        for j in np.nonzero(ok)[0]:
            k = min(int(sa["n_records"][j]), a.max_records)
            if not np.array_equal(a.records[j, :k], b.records[j, :k]):
                errs.append(f"records of string {j} differ: {a.records[j, :k]} vs {b.records[j, :k]}")
                break
    if a.compact_bytes is not None and b.compact_bytes is not None:
        for j in np.nonzero(ok)[0]:
            k = min(int(sa["n_compact"][j]), a.compact_pitch)
            if not np.array_equal(a.compact_bytes[j, :k], b.compact_bytes[j, :k]):
                errs.append(f"compact bytes of string {j} differ")
                break
    return errs
