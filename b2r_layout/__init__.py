"""Struct layouts of include/b2r.h (ctypes) and numpy host buffers in that layout.  Neutral: imports neither the CUDA library
nor the oracle, so both sides (and bench.py's CPU reference arm) can share it."""
from . import abi  # noqa: F401
from .buffers import HostOutputs, aligned_empty, compare_outputs, round_up  # noqa: F401
