"""GPU (-m gpu): the host-pointer entry points of the C ABI beyond the plain sliced pipeline — sparse device-to-host mode,
the small-batch path, page-locked buffers from the library, out-of-range offsets, the multi-device handle — every one
compared with the CPU oracle on identical inputs."""
import random
import zlib

import numpy as np
import pytest

from conftest import DEF_SETS, oracle_config, product_config
from test_gpu_parity import _pack
from test_oracle_golden import SNIPPETS, _random_strings

pytestmark = pytest.mark.gpu


def _oracle(set_name, M, data, offs, **kw):
    return oracle_config(set_name, M).match_batch(data, offs, **kw)


@pytest.mark.parametrize("set_name", sorted(DEF_SETS))
def test_sparse_d2h_equals_dense_equals_oracle(set_name):
    """B2R_OUT_SPARSE_D2H: compacted on the device, expanded by host threads — the same bits as the dense copies and the oracle."""
    import halo2_regex_b200 as H
    M = 97
    rng = random.Random(zlib.crc32(set_name.encode()))
    strings = _random_strings(rng, 20011, M - 1, SNIPPETS)              # >= 16384: eight slices; not a multiple of 32
    data, offs = _pack(strings, lead=3)
    cfg = product_config(set_name, M)
    kw = dict(max_records=4, compact_pitch=32)
    o, ores = _oracle(set_name, M, data, offs, **kw)
    dense, dres = cfg.match_batch_host(data, offs, check=False, fill=0xCD, **kw)
    sparse, sres = cfg.match_batch_host(data, offs, check=False, fill=0xAB, sparse=True, **kw)
    assert dres.code == ores.code == sres.code
    assert H.compare_outputs(dense, o) == []
    assert H.compare_outputs(sparse, o) == []
    h2d_d, d2h_s = cfg.last_host_bytes()
    assert d2h_s < 0.6 * sum(a.nbytes for a in sparse.all_arrays())       # fewer bytes crossed PCIe than the columns hold
    # pinned buffers from the library's own allocator take the same path
    alloc = H.PinnedAllocator()
    try:
        out = cfg.new_host_outputs(len(strings), allocator=alloc, **kw)
        cfg.match_batch_host(data, offs, out=out, check=False, sparse=True)
        assert H.compare_outputs(out, o) == []
    finally:
        alloc.free()


@pytest.mark.parametrize("set_name", ["regex1", "three"])
def test_sparse_reuse_clears_only_what_the_last_call_wrote(set_name):
    """B2R_OUT_SPARSE_REUSE: the same host buffers batch after batch — the library clears the sectors of the previous batch instead
    of zeroing the columns; different strings every time, a dense-fallback slice in between, a foreign buffer (full zeroing)."""
    import halo2_regex_b200 as H
    M = 97
    rng = random.Random(zlib.crc32(set_name.encode()) + 9)
    cfg = product_config(set_name, M)
    kw = dict(max_records=4, compact_pitch=32)
    n = 18011
    out = cfg.new_host_outputs(n, fill=0xAB, **kw)
    other = cfg.new_host_outputs(n, fill=0x77, **kw)
    for step in range(5):
        strings = _random_strings(rng, n, M - 1, SNIPPETS)
        data, offs = _pack(strings, lead=step)
        o, ores = _oracle(set_name, M, data, offs, **kw)
        if step == 2:
            cfg.set_option("sparse_cap", 64)                               # this batch: most column slices cross densely
        if step == 3:
            cfg.set_option("sparse_cap", 0)
        target = other if step == 4 else out                               # step 4: buffers the handle has never seen, still full of 0x77
        g, gres = cfg.match_batch_host(data, offs, out=target, check=False, sparse=True, reuse=step > 0)
        assert gres.code == ores.code
        assert H.compare_outputs(g, o) == [], step


def test_sparse_d2h_dense_fallback_and_threads():
    """A column slice with more non-zero sectors than the compaction arena holds crosses densely; any thread count works."""
    import halo2_regex_b200 as H
    M = 130
    rng = random.Random(77)
    strings = _random_strings(rng, 17000, M - 1, SNIPPETS)
    data, offs = _pack(strings)
    o, _ = _oracle("test1", M, data, offs)
    cfg = product_config("test1", M)
    for cap, threads in ((1, 1), (300, 3), (0, 16)):
        cfg.set_option("sparse_cap", cap)
        cfg.set_option("host_threads", threads)
        g, _ = cfg.match_batch_host(data, offs, check=False, fill=0x5A, sparse=True)
        assert H.compare_outputs(g, o) == [], (cap, threads)


@pytest.mark.parametrize("set_name", ["regex1", "three", "example"])
@pytest.mark.parametrize("n", [1, 2, 31, 32, 33])
def test_small_batches_one_copy_each_way(set_name, n):
    """n <= 32 strings take the one-tile path (one H2D, one D2H): same bits as the sliced pipeline and the oracle."""
    import halo2_regex_b200 as H
    M = 257
    rng = random.Random(n * 131 + len(set_name))
    strings = _random_strings(rng, n, M - 1, SNIPPETS)
    strings[0] = b""
    data, offs = _pack(strings, lead=5)
    o, ores = _oracle(set_name, M, data, offs)
    cfg = product_config(set_name, M)
    g, gres = cfg.match_batch_host(data, offs, check=False, fill=0xCD)
    assert gres.code == ores.code and H.compare_outputs(g, o) == []
    cfg.set_option("small_path", 0)
    g2, _ = cfg.match_batch_host(data, offs, check=False, fill=0xCD)
    assert H.compare_outputs(g2, o) == []
    # accumulate on the small path
    cfg.set_option("small_path", 1)
    g3, _ = cfg.match_batch_host(data, offs, out=g, check=False, flags=H._abi.B2R_OUT_ACCUMULATE_MULT)
    if ores.code == 0:
        for d in range(cfg.n_defs):
            assert np.array_equal(g3.mult[d], 2 * o.mult[d]) and np.array_equal(g3.endpoint_mult[d], 2 * o.endpoint_mult[d])


def test_small_batch_failure_is_reported():
    import halo2_regex_b200 as H
    cfg = product_config("regex1", 64)
    with pytest.raises(H.InvalidTransitionError) as e:
        cfg.match_strings([b"ok then", b"bad \x01 byte"])
    assert (e.value.string_idx, e.value.pos, e.value.char) == (1, 4, 1)
    with pytest.raises(H.StringTooLongError):
        cfg.match_strings([b"x" * 64])


def test_unaligned_bitmap_pitch_across_slices():
    """bitmap_pitch only has to be a multiple of 4 (include/b2r.h): slices are cut at multiples of 32 strings so that every
    column slice stays 16-byte aligned (n = 16391 used to fail with B2R_ERR_ALIGNMENT in the middle of the pipeline)."""
    import halo2_regex_b200 as H
    M = 1025
    rng = random.Random(3)
    strings = _random_strings(rng, 16391, 120, SNIPPETS)
    data, offs = _pack(strings)
    cfg = product_config("regex1", M)
    o, _ = oracle_config("regex1", M).match_batch(data, offs, bitmap_pitch=132, row_pitch=1040)
    for sparse in (False, True):
        g, _ = cfg.match_batch_host(data, offs, bitmap_pitch=132, row_pitch=1040, fill=0xEE, sparse=sparse)
        assert H.compare_outputs(g, o) == [], sparse


def test_offsets_outside_the_buffer_are_flagged_not_read():
    """Device entry point: an offset pair that leaves [0, total_bytes] marks the string (B2R_ST_TOO_LONG) instead of being read;
    host entry point: decreasing offsets are rejected before anything is enqueued."""
    import torch
    import halo2_regex_b200 as H
    M = 64
    cfg = product_config("regex1", M)
    strings = [b"email was meant for @ab.", b"second", b"third one"]
    data, offs = _pack(strings)
    bad = offs.copy().astype(np.int64)
    bad_t = torch.tensor([0, 24, 1 << 40, (1 << 40) + 5], dtype=torch.int64, device="cuda")    # string 1 ends far outside, string 2 lies outside
    d = torch.from_numpy(np.concatenate([data, np.zeros(16, np.uint8)])).cuda()
    out = H.DeviceOutputs(cfg, 3)
    cfg.match_batch_device(d[:len(data)], bad_t, out)
    res = cfg.batch_result(check=False)
    assert res.code == H._abi.B2R_ERR_TOO_LONG and res.string_idx == 1
    st = out.to_host().status
    assert not st["flags"][0] & H._abi.B2R_ST_TOO_LONG and st["flags"][1] & H._abi.B2R_ST_TOO_LONG and st["flags"][2] & H._abi.B2R_ST_TOO_LONG
    dec = offs.copy()
    dec[2] = dec[1] - 1
    with pytest.raises(RuntimeError, match="non-decreasing"):
        cfg.match_batch_host(data, dec)
    for n in (3, 20000):                                                    # the sliced pipeline checks every offset, not only the cuts
        o2 = np.arange(n + 1, dtype=np.uint64) * 2
        o2[n // 2] = 10 ** 9
        with pytest.raises(RuntimeError, match="non-decreasing"):
            cfg.match_batch_host(np.zeros(2 * n, np.uint8), o2)


def test_options_follow_the_environment_once(monkeypatch):
    """The B2R_* hooks are read when the handle is created, never on the batch path; set_option changes them afterwards."""
    monkeypatch.setenv("B2R_TABLE_MODE", "plain16")
    cfg = product_config("regex1", 64)
    monkeypatch.setenv("B2R_TABLE_MODE", "global")                          # too late for this handle
    cfg.match_strings([b"email was meant for @ab."])
    assert cfg.last_plan()[0] == "plain16"
    cfg.set_option("table_mode", "repl")
    cfg.match_strings([b"email was meant for @ab."])
    assert cfg.last_plan()[0] == "repl"
    with pytest.raises(RuntimeError):
        cfg.set_option("no_such_option", 1)


def _multi_or_skip(n):
    import torch
    if torch.cuda.device_count() < n:
        pytest.skip(f"needs {n} GPUs")


@pytest.mark.parametrize("set_name,n_dev", [("regex1", 2), ("three", 2), ("test1", 4)])
def test_multi_device_handle_matches_oracle(set_name, n_dev):
    """b2r_config_new_multi: one process, several GPUs, strings sharded by bytes, ONE NCCL all-reduce of the multiplicity block —
    every column and every counter row equals the single-thread oracle on the whole batch."""
    _multi_or_skip(n_dev)
    import halo2_regex_b200 as H
    M = 130
    rng = random.Random(n_dev * 7 + len(set_name))
    strings = _random_strings(rng, 40037, M - 1, SNIPPETS)
    data, offs = _pack(strings, lead=1)
    o, ores = _oracle(set_name, M, data, offs)
    spec = DEF_SETS[set_name]
    import os
    from conftest import DEFS
    defs = [H.RegexDefs(H.AllstrRegexDef.read_from_text(os.path.join(DEFS, a)), [H.SubstrRegexDef.read_from_text(os.path.join(DEFS, s)) for s in ss]) for a, ss in spec]
    cfg = H.RegexVerifyConfig.configure(M, defs, devices=list(range(n_dev)))
    for sparse in (False, True):
        g, gres = cfg.match_batch_host(data, offs, check=False, fill=0xCD, sparse=sparse)
        assert gres.code == ores.code
        assert H.compare_outputs(g, o) == [], sparse
    # accumulate: the caller's counters are uploaded by the first device only
    g2, _ = cfg.match_batch_host(data, offs, out=g, check=False, flags=H._abi.B2R_OUT_ACCUMULATE_MULT)
    if ores.code == 0:
        for d in range(cfg.n_defs):
            assert np.array_equal(g2.mult[d], 2 * o.mult[d])
    # fewer strings than devices x 32: some devices get nothing and still join the all-reduce
    few, foffs = _pack(strings[:40])
    o3, _ = _oracle(set_name, M, few, foffs)
    g3, _ = cfg.match_batch_host(few, foffs, check=False)
    assert H.compare_outputs(g3, o3) == []
    r = cfg.match_substrs(strings[5])
    assert r is not None
