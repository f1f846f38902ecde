"""ctypes mirror of include/b2r.h: lives in the neutral b2r_layout package (shared with the test infrastructure)."""
from b2r_layout.abi import *  # noqa: F401,F403
from b2r_layout.abi import B2R_ST_ACCEPTED  # noqa: F401
