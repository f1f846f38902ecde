"""GPU (-m gpu): an independent reader of the C-ABI output layout.

Every other parity test reads the witness columns through b2r_layout / halo2_regex_b200.buffers — the same helper the oracle
wrapper uses, so a pitch or offset mistake common to both would cancel.  Here nothing is shared: the `b2r_outputs` struct is
declared from the text of include/b2r.h alone, the buffers are raw ctypes arrays with pitches no helper would choose, every
element is addressed by hand as `j * pitch + i`, and the expected values come from oracle/pyref.py (pure-Python lists written
independently of oracle.c)."""
import ctypes as C
import os
import random

import pytest

from conftest import DEFS, pyref_defs

pytestmark = pytest.mark.gpu

MAX_DEFS = 4


class Outputs(C.Structure):      # include/b2r.h: typedef struct b2r_outputs
    _fields_ = [("row_pitch", C.c_uint64), ("bitmap_pitch", C.c_uint64),
                ("states", C.c_void_p * MAX_DEFS), ("substr_ids", C.c_void_p * MAX_DEFS),
                ("start_enable", C.c_void_p * MAX_DEFS), ("end_enable", C.c_void_p * MAX_DEFS),
                ("masked_chars", C.c_void_p), ("masked_substr_ids", C.c_void_p),
                ("status", C.c_void_p), ("records", C.c_void_p),
                ("max_records", C.c_uint32), ("compact_pitch", C.c_uint32), ("compact_bytes", C.c_void_p),
                ("mult", C.c_void_p * MAX_DEFS), ("endpoint_mult", C.c_void_p * MAX_DEFS),
                ("flags", C.c_uint32), ("reserved", C.c_uint32)]


def _buf(n):
    raw = (C.c_uint8 * (n + 64))()
    off = (-C.addressof(raw)) % 64                                        # 64-byte aligned start inside the allocation
    C.memset(C.addressof(raw) + off, 0xA5, n)                             # poison: every byte read below must have been written
    return raw, C.addressof(raw) + off


@pytest.mark.parametrize("set_name,files", [("regex1", [("regex1_test_lookup.txt", ["substr1_test_lookup.txt"])]),
                                            ("test1", [("regex1_test_lookup.txt", ["substr1_test_lookup.txt"]), ("regex2_test_lookup.txt", ["substr2_test_lookup.txt"])])])
def test_columns_addressed_by_hand_match_the_pure_python_model(set_name, files):
    from oracle import pyref as P
    lib = C.CDLL(os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "halo2_regex_b200", "libb2r.so"))
    lib.b2r_last_error.restype = C.c_char_p
    M, RP, BP, n_defs = 77, 96, 12, len(files)                           # row_pitch 96 (>= M, multiple of 16), bitmap_pitch 12 (>= ceil(77/8) = 10, multiple of 4)
    allstr, substrs, nsub = [], [], []
    for a, ss in files:
        h = C.c_void_p()
        assert lib.b2r_allstr_read_from_text(os.path.join(DEFS, a).encode(), C.byref(h), None) == 0
        allstr.append(h)
        arr = (C.c_void_p * len(ss))()
        for k, sfile in enumerate(ss):
            sh = C.c_void_p()
            assert lib.b2r_substr_read_from_text(os.path.join(DEFS, sfile).encode(), C.byref(sh), None) == 0
            arr[k] = sh
        substrs.append(arr)
        nsub.append(len(ss))
    cfg = C.c_void_p()
    a_arr = (C.c_void_p * n_defs)(*[h.value for h in allstr])
    s_arr = (C.POINTER(C.c_void_p) * n_defs)(*[C.cast(x, C.POINTER(C.c_void_p)) for x in substrs])
    n_arr = (C.c_uint32 * n_defs)(*nsub)
    rc = lib.b2r_config_new(a_arr, s_arr, n_arr, C.c_uint32(n_defs), C.c_uint64(M), C.c_int(0), C.byref(cfg))
    assert rc == 0, lib.b2r_last_error()

    rng = random.Random(99)
    alphabet = b"abcdefghijklmnopqrstuvwxyz @.ABC"
    strings = []
    for _ in range(70):                                                   # 70 strings: two full tiles and a partial one
        s = bytes(rng.choice(alphabet) for _ in range(rng.randint(0, M - 1)))
        if rng.random() < 0.6 and len(s) > 40:
            at = rng.randint(0, len(s) - 36)
            s = s[:at] + b"email was meant for @" + bytes(rng.choice(b"xyzw") for _ in range(rng.randint(1, 6))) + b". Also for ab." + s[at + 36:]
            s = s[:M - 1]
        strings.append(s)
    n = len(strings)
    data = b"".join(strings)
    offs = (C.c_uint64 * (n + 1))()
    for j, s in enumerate(strings):
        offs[j + 1] = offs[j] + len(s)
    keep, out = [], Outputs()
    out.row_pitch, out.bitmap_pitch = RP, BP

    def col(nbytes):
        raw, addr = _buf(nbytes)
        keep.append(raw)
        return addr
    for d in range(n_defs):
        out.states[d] = col(n * RP); out.substr_ids[d] = col(n * RP)
        out.start_enable[d] = col(n * BP); out.end_enable[d] = col(n * BP)
    out.masked_chars = col(n * RP); out.masked_substr_ids = col(n * RP)
    out.status = col(n * 32)
    rc = lib.b2r_match_batch_host(cfg, data, offs, C.c_uint64(n), C.byref(out), None)
    any_invalid = False

    def byte_at(addr, j, i, pitch):
        return C.c_uint8.from_address(addr + j * pitch + i).value

    def bit_at(addr, j, i):
        return (C.c_uint8.from_address(addr + j * BP + (i >> 3)).value >> (i & 7)) & 1     # LSB-first bitmaps
    pdefs = pyref_defs(files)
    for j, s in enumerate(strings):
        flags = C.c_uint32.from_address(out.status + j * 32).value                       # b2r_string_status.flags is the first word
        try:
            w = P.match_substrs(pdefs, M, s)
        except P.InvalidTransition:
            assert flags & (1 << 9)                                                       # B2R_ST_INVALID_TRANSITION
            any_invalid = True
            continue
        assert not flags & (1 << 9)
        for d in range(n_defs):
            assert [byte_at(out.states[d], j, i, RP) for i in range(M)] == w["states"][d], (j, d)
            assert [byte_at(out.substr_ids[d], j, i, RP) for i in range(M)] == w["substr_ids"][d], (j, d)
            assert [bit_at(out.start_enable[d], j, i) for i in range(M)] == w["start_enable"][d], (j, d)
            assert [bit_at(out.end_enable[d], j, i) for i in range(M)] == w["end_enable"][d], (j, d)
            assert bool(flags & (1 << d)) == w["accepted"][d], (j, d)                     # B2R_ST_ACCEPTED(d) = 1u << d
        if not w["overlap"]:
            assert [byte_at(out.masked_chars, j, i, RP) for i in range(M)] == w["masked_chars"], j
            assert [byte_at(out.masked_substr_ids, j, i, RP) for i in range(M)] == w["masked_substr_ids"], j
    assert (rc != 0) == any_invalid, lib.b2r_last_error()
    lib.b2r_config_free(cfg)
