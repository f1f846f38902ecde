"""Regex definition model — host mirror of the reference's `halo2_regex::defs` (src/defs.rs).

`AllstrRegexDef`, `SubstrRegexDef` and `RegexDefs` keep the reference's names, constructors, public fields and text
formats; parsing is done by the C++ loader behind the C ABI (include/b2r.h), which reproduces the reference's edge
cases (`as u8` truncation, later duplicate key wins, empty/short body line is an error, Unicode whitespace, `\\r\\n`).
"""
import ctypes as C

import numpy as np

from . import _abi
from ._ffi import last_error, lib


class RegexParseError(ValueError):
    """The reference panics (`expect` / index out of bounds, src/defs.rs:85-100, 219-238); we raise instead."""

    def __init__(self, msg, line=None):
        super().__init__(msg)
        self.line = line


def _check(rc, line=None):
    if rc == _abi.B2R_ERR_PARSE:
        raise RegexParseError(last_error(), line.value if line is not None else None)
    if rc == _abi.B2R_ERR_IO:
        raise FileNotFoundError(last_error())
    if rc != 0:
        raise RuntimeError(f"b2r error {rc}: {last_error()}")


class AllstrRegexDef:
    """Regex that the whole input string must satisfy (reference src/defs.rs:26-36).

    Fields: `first_state_val`, `accepted_state_val`, `largest_state_val`, `state_lookup` — a dict
    `(char: u8, cur_state) -> (line_idx, next_state)` exactly like the reference's HashMap.
    """

    def __init__(self, handle):
        self._h = handle
        self.first_state_val = lib.b2r_allstr_first_state_val(handle)
        self.accepted_state_val = lib.b2r_allstr_accepted_state_val(handle)
        self.largest_state_val = lib.b2r_allstr_largest_state_val(handle)
        self._lookup = None

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b2r_allstr_free(h)

    @classmethod
    def read_from_text(cls, file_path):
        """reference src/defs.rs:54-58"""
        h, line = C.c_void_p(), C.c_uint64()
        _check(lib.b2r_allstr_read_from_text(str(file_path).encode(), C.byref(h), C.byref(line)), line)
        return cls(h)

    @classmethod
    def read_from_reader(cls, reader):
        """reference src/defs.rs:75-110; `reader` is bytes, str or a file-like object"""
        data = reader.read() if hasattr(reader, "read") else reader
        if isinstance(data, str):
            data = data.encode("utf-8")
        h, line = C.c_void_p(), C.c_uint64()
        _check(lib.b2r_allstr_parse(data, len(data), C.byref(h), C.byref(line)), line)
        return cls(h)

    def entries(self):
        """(n, 4) uint64 array of (char, cur, next, line_idx) in table order (sorted by line index)."""
        n = lib.b2r_allstr_num_transitions(self._h)
        out = np.zeros((n, 4), dtype=np.uint64)
        _check(lib.b2r_allstr_entries(self._h, out.ctypes.data, n))
        return out

    @property
    def state_lookup(self):
        if self._lookup is None:
            self._lookup = {(int(c), int(cur)): (int(li), int(nx)) for c, cur, nx, li in self.entries()}
        return self._lookup

    def get(self, char, state):
        """state_lookup.get(&(char, state)) → (line_idx, next) or None"""
        li, nx = C.c_uint64(), C.c_uint64()
        if lib.b2r_allstr_lookup(self._h, char, state, C.byref(li), C.byref(nx)):
            return li.value, nx.value
        return None


class SubstrRegexDef:
    """Regex that each substring must satisfy (reference src/defs.rs:115-132)."""

    def __init__(self, max_length=None, min_position=0, max_position=0, valid_state_transitions=(), start_states=(),
                 end_states=(), _handle=None):
        """SubstrRegexDef::new (reference src/defs.rs:147-163)"""
        if _handle is None:
            pairs = np.array(sorted(valid_state_transitions), dtype=np.uint64).reshape(-1, 2)
            st = np.array(list(start_states), dtype=np.uint64)
            en = np.array(list(end_states), dtype=np.uint64)
            h = C.c_void_p()
            _check(lib.b2r_substr_new(int(max_length), int(min_position), int(max_position), pairs.ctypes.data, len(pairs),
                                      st.ctypes.data, len(st), en.ctypes.data, len(en), C.byref(h)))
            _handle = h
        self._h = _handle
        self.max_length = lib.b2r_substr_max_length(self._h)
        self.min_position = lib.b2r_substr_min_position(self._h)
        self.max_position = lib.b2r_substr_max_position(self._h)
        n = lib.b2r_substr_num_transitions(self._h)
        tr = np.zeros((n, 2), dtype=np.uint64)
        _check(lib.b2r_substr_transitions(self._h, tr.ctypes.data, n))
        self.valid_state_transitions = {(int(a), int(b)) for a, b in tr}
        ns, ne = lib.b2r_substr_num_start_states(self._h), lib.b2r_substr_num_end_states(self._h)
        s, e = np.zeros(ns, dtype=np.uint64), np.zeros(ne, dtype=np.uint64)
        _check(lib.b2r_substr_start_states(self._h, s.ctypes.data, ns))
        _check(lib.b2r_substr_end_states(self._h, e.ctypes.data, ne))
        self.start_states, self.end_states = [int(x) for x in s], [int(x) for x in e]

    new = classmethod(lambda cls, *a, **k: cls(*a, **k))

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            lib.b2r_substr_free(h)

    @classmethod
    def read_from_text(cls, file_path):
        """reference src/defs.rs:184-188"""
        h, line = C.c_void_p(), C.c_uint64()
        _check(lib.b2r_substr_read_from_text(str(file_path).encode(), C.byref(h), C.byref(line)), line)
        return cls(_handle=h)

    @classmethod
    def read_from_reader(cls, reader):
        """reference src/defs.rs:209-265"""
        data = reader.read() if hasattr(reader, "read") else reader
        if isinstance(data, str):
            data = data.encode("utf-8")
        h, line = C.c_void_p(), C.c_uint64()
        _check(lib.b2r_substr_parse(data, len(data), C.byref(h), C.byref(line)), line)
        return cls(_handle=h)


class RegexDefs:
    """reference src/defs.rs:17-22"""

    def __init__(self, allstr, substrs):
        self.allstr = allstr
        self.substrs = list(substrs)
