"""Synthetic workload generators live in the neutral top-level `workloads` module (pure numpy / torch, no dependency on this
package or its CUDA library, so the CPU reference arm of bench.py can use them without loading libb2r.so); re-exported here."""
from workloads import *  # noqa: F401,F403
from workloads import ALPHABET, SEED_CONFIG1, SEED_CONFIG2, SEED_CONFIG4, _NP, _TorchI64, _filler, _mix, _rand  # noqa: F401
