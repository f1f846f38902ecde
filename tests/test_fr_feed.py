"""witness columns -> BN254 Fr cells in Montgomery form (b2r_column_to_fr), against a Python big-integer
model of halo2curves' `Fr::from(u64)` (= x * 2^256 mod r, four little-endian u64 limbs).  The reference does this per cell:
src/lib.rs:342-347, 388-418 (`Value::known(F::from(..))`)."""
import random

import numpy as np
import pytest

from conftest import product_config
from test_gpu_parity import _pack
from test_oracle_golden import SNIPPETS, _random_strings

# BN254 scalar field (halo2curves::bn256::Fr): modulus and the published Montgomery constants
R_MOD = 21888242871839275222246405745257275088548364400416034343698204186575808495617
MASK = (1 << 64) - 1


def fr_limbs(x):
    m = (int(x) << 256) % R_MOD
    return [(m >> (64 * i)) & MASK for i in range(4)]


def test_constants_of_the_kernel_match_the_field():
    """The constants compiled into csrc/fr.cu are the field's: r, R^2 mod r, -r^{-1} mod 2^64; and the model reproduces the
    library's published R = Fr::one() (0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb)."""
    import os
    import re
    src = open(os.path.join(os.path.dirname(__file__), "..", "halo2_regex_b200", "csrc", "fr.cu")).read()

    def limbs(name):
        body = re.search(name + r"\[4\] = \{([^}]*)\}", src).group(1)
        return [int(x.strip().rstrip("ul"), 16) for x in body.split(",")]

    assert sum(v << (64 * i) for i, v in enumerate(limbs("FR_MODULUS"))) == R_MOD
    assert sum(v << (64 * i) for i, v in enumerate(limbs("FR_R2"))) == pow(2, 512, R_MOD)
    inv = int(re.search(r"FR_INV = (0x[0-9a-f]+)ull", src).group(1), 16)
    assert (inv * R_MOD + 1) % (1 << 64) == 0
    assert sum(v << (64 * i) for i, v in enumerate(fr_limbs(1))) == 0x0e0a77c19a07df2f666ea36f7879462e36fc76959f60cd29ac96341c4ffffffb
    assert R_MOD == 0x30644e72e131a029b85045b68181585d2833e84879b9709143e1f593f0000001


def _as_u64(t):
    return t.cpu().numpy().view(np.uint64)


@pytest.mark.gpu
def test_every_u8_value_and_wide_values():
    import torch
    cfg = product_config("regex1", 64)
    col = torch.arange(256, dtype=torch.uint8, device="cuda").reshape(4, 64)
    got = _as_u64(cfg.column_to_fr(col, rows=64))
    for v in range(256):
        assert list(got[v // 64, v % 64]) == fr_limbs(v), v
    rng = random.Random(1)
    vals = [0, 1, 2, 255, 256, 65535, (1 << 32) - 1, 1 << 32, (1 << 63) - 1, 1 << 63, (1 << 64) - 1, R_MOD & MASK] + [rng.getrandbits(64) for _ in range(500)]
    t64 = torch.tensor([v - (1 << 64) if v >= (1 << 63) else v for v in vals], dtype=torch.int64, device="cuda").reshape(1, -1)
    got = _as_u64(cfg.column_to_fr(t64, rows=len(vals)))
    for k, v in enumerate(vals):
        assert list(got[0, k]) == fr_limbs(v), hex(v)
    v16 = [rng.getrandbits(16) for _ in range(300)] + [0, 65535, 1023]
    t16 = torch.tensor(np.array(v16, dtype=np.uint16).view(np.int16), device="cuda").reshape(3, 101)
    got = _as_u64(cfg.column_to_fr(t16, rows=100))          # pitch 101, 100 rows
    for j in range(3):
        for i in range(100):
            assert list(got[j, i]) == fr_limbs(v16[j * 101 + i])


@pytest.mark.gpu
@pytest.mark.parametrize("set_name", ["regex1", "test1"])
def test_witness_columns_of_a_batch_as_fr(set_name):
    """Every column of a batch -> Fr, cell by cell equal to F::from of the oracle's value (the reference's assignment values)."""
    import torch
    import halo2_regex_b200 as H
    from conftest import oracle_config
    M = 70
    rng = random.Random(4)
    strings = _random_strings(rng, 300, M - 1, SNIPPETS)
    data, offs = _pack(strings)
    cfg = product_config(set_name, M)
    o, _ = oracle_config(set_name, M).match_batch(data, offs)
    d = torch.from_numpy(np.concatenate([data, np.zeros(16, np.uint8)])).cuda()
    d_offs = torch.from_numpy(offs.astype(np.int64)).cuda()
    out = H.DeviceOutputs(cfg, len(strings))
    cfg.match_batch_device(d[:len(data)], d_offs, out)
    cfg.batch_result(check=False)
    table = np.array([fr_limbs(v) for v in range(256)], dtype=np.uint64)
    ok = (o.status["flags"] & (H._abi.B2R_ST_INVALID_TRANSITION | H._abi.B2R_ST_TOO_LONG)) == 0

    def check(fr, expect):          # expect: (n, M) small integers
        got = _as_u64(fr)
        assert np.array_equal(got[ok], table[expect[ok].astype(np.int64)])

    for dd in range(cfg.n_defs):
        check(cfg.column_to_fr(out.states[dd]), o.states[dd][:, :M])
        check(cfg.column_to_fr(out.substr_ids[dd]), o.substr_ids[dd][:, :M])
        check(cfg.column_to_fr(out.start_enable[dd], kind="bitmap"), o.bits(o.start_enable[dd]).astype(np.uint8))
        check(cfg.column_to_fr(out.end_enable[dd], kind="bitmap"), o.bits(o.end_enable[dd]).astype(np.uint8))
    check(cfg.column_to_fr(out.masked_chars), o.masked_chars[:, :M])
    check(cfg.column_to_fr(out.masked_substr_ids), o.masked_substr_ids[:, :M])
    lens = np.diff(offs.astype(np.int64))
    chars = np.zeros((len(strings), M), dtype=np.uint8)
    for j, s in enumerate(strings):
        chars[j, :len(s)] = np.frombuffer(s, dtype=np.uint8)
    ok[:] = True
    check(cfg.column_to_fr(d, kind="chars", offsets=d_offs), chars)
    check(cfg.column_to_fr(d, kind="enable", offsets=d_offs), (np.arange(M)[None, :] < lens[:, None]).astype(np.uint8))
