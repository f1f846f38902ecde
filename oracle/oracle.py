"""ctypes wrapper around oracle/liboracle.so — the CPU restatement of the reference path.  TEST INFRASTRUCTURE ONLY.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module.
See oracle.c for the reference file:line each step follows.
"""
import ctypes as C
import os
import subprocess

import numpy as np

import sys

_HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(_HERE))
from b2r_layout import abi as _abi  # noqa: E402  (struct layouts only; neutral package, loads no native library)
from b2r_layout.buffers import HostOutputs  # noqa: E402

_LIB = None


def build():
    subprocess.check_call(["make", "-s", "-C", _HERE])


def lib():
    global _LIB
    if _LIB is None:
        path = os.path.join(_HERE, "liboracle.so")
        if not os.path.exists(path):
            build()
        L = C.CDLL(path)
        L.orc_allstr_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.orc_substr_parse.argtypes = [C.c_char_p, C.c_size_t, C.POINTER(C.c_void_p), C.POINTER(C.c_uint64)]
        L.orc_allstr_free.argtypes = [C.c_void_p]
        L.orc_substr_free.argtypes = [C.c_void_p]
        L.orc_config_new.argtypes = [C.POINTER(C.c_void_p), C.POINTER(C.POINTER(C.c_void_p)), C.POINTER(C.c_uint32),
                                     C.c_uint32, C.c_uint64, C.POINTER(C.c_void_p)]
        L.orc_config_free.argtypes = [C.c_void_p]
        for f in (L.orc_table_num_rows, L.orc_endpoint_num_rows):
            f.argtypes = [C.c_void_p, C.c_uint32]
            f.restype = C.c_uint64
        L.orc_table_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_endpoint_rows.argtypes = [C.c_void_p, C.c_uint32, C.c_void_p]
        L.orc_allstr_num_transitions.argtypes = [C.c_void_p]
        L.orc_allstr_num_transitions.restype = C.c_uint64
        L.orc_allstr_header.argtypes = [C.c_void_p, C.c_int]
        L.orc_allstr_header.restype = C.c_uint64
        L.orc_substr_num_transitions.argtypes = [C.c_void_p]
        L.orc_substr_num_transitions.restype = C.c_uint64
        L.orc_match_batch.argtypes = [C.c_void_p, C.c_void_p, C.c_void_p, C.c_uint64, C.POINTER(_abi.Outputs), C.c_int,
                                      C.POINTER(_abi.BatchStatus)]
        _LIB = L
    return _LIB


class OracleParseError(Exception):
    def __init__(self, line):
        super().__init__(f"parse error at line {line}")
        self.line = line


class OracleAllstr:
    def __init__(self, text):
        if isinstance(text, str):
            text = text.encode()
        h, line = C.c_void_p(), C.c_uint64()
        rc = lib().orc_allstr_parse(text, len(text), C.byref(h), C.byref(line))
        if rc != 0:
            raise OracleParseError(line.value)
        self.h = h
        self.first_state_val = lib().orc_allstr_header(h, 0)
        self.accepted_state_val = lib().orc_allstr_header(h, 1)
        self.largest_state_val = lib().orc_allstr_header(h, 2)
        self.num_transitions = lib().orc_allstr_num_transitions(h)

    @classmethod
    def read_from_text(cls, path):
        with open(path, "rb") as f:
            return cls(f.read())


class OracleSubstr:
    def __init__(self, text):
        if isinstance(text, str):
            text = text.encode()
        h, line = C.c_void_p(), C.c_uint64()
        rc = lib().orc_substr_parse(text, len(text), C.byref(h), C.byref(line))
        if rc != 0:
            raise OracleParseError(line.value)
        self.h = h
        self.num_transitions = lib().orc_substr_num_transitions(h)

    @classmethod
    def read_from_text(cls, path):
        with open(path, "rb") as f:
            return cls(f.read())


class OracleConfig:
    """regex_defs: list of (OracleAllstr, [OracleSubstr, ...])"""

    def __init__(self, regex_defs, max_chars_size):
        self.regex_defs = regex_defs
        self.m = int(max_chars_size)
        D = len(regex_defs)
        allstr = (C.c_void_p * D)(*[a.h for a, _ in regex_defs])
        sub_arrays = [(C.c_void_p * max(1, len(s)))(*[x.h for x in s]) for _, s in regex_defs]
        subs = (C.POINTER(C.c_void_p) * D)(*[C.cast(a, C.POINTER(C.c_void_p)) for a in sub_arrays])
        ns = (C.c_uint32 * D)(*[len(s) for _, s in regex_defs])
        h = C.c_void_p()
        rc = lib().orc_config_new(allstr, subs, ns, D, self.m, C.byref(h))
        if rc != 0:
            raise ValueError(f"orc_config_new failed: {rc}")
        self.h, self.n_defs = h, D
        self._keep = (allstr, sub_arrays, subs, ns)
        wide = any(a.largest_state_val + 1 > 255 for a, _ in regex_defs)
        self.state_widths = [2 if wide else 1] * D
        self.table_num_rows = [lib().orc_table_num_rows(h, d) for d in range(D)]
        self.endpoint_num_rows = [lib().orc_endpoint_num_rows(h, d) for d in range(D)]

    def table_rows(self, d):
        out = np.zeros((self.table_num_rows[d], 4), dtype=np.uint64)
        lib().orc_table_rows(self.h, d, out.ctypes.data)
        return out

    def endpoint_rows(self, d):
        out = np.zeros((self.endpoint_num_rows[d], 3), dtype=np.uint64)
        lib().orc_endpoint_rows(self.h, d, out.ctypes.data)
        return out

    def new_outputs(self, n, **kw):
        return HostOutputs(n, self.m, self.state_widths, self.table_num_rows, self.endpoint_num_rows, **kw)

    def match_batch(self, data, offsets, out=None, nthreads=1, flags=0, **kw):
        """data: uint8 array of concatenated strings; offsets: uint64 array (N+1).  Returns (HostOutputs, BatchStatus)."""
        data = np.ascontiguousarray(data, dtype=np.uint8)
        offsets = np.ascontiguousarray(offsets, dtype=np.uint64)
        n = len(offsets) - 1
        if out is None:
            out = self.new_outputs(n, fill=0xEE, **kw)
        st = out.struct(flags)
        res = _abi.BatchStatus()
        rc = lib().orc_match_batch(self.h, data.ctypes.data, offsets.ctypes.data, n, C.byref(st), nthreads, C.byref(res))
        assert rc == 0
        return out, res

    def match_strings(self, strings, **kw):
        data = np.frombuffer(b"".join(strings), dtype=np.uint8) if strings else np.zeros(0, np.uint8)
        offs = np.zeros(len(strings) + 1, dtype=np.uint64)
        offs[1:] = np.cumsum([len(s) for s in strings])
        return self.match_batch(data, offs, **kw)
