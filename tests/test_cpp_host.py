"""The compiled-language host mirror of the reference API (include/b2r.hpp, C++ on top of the C ABI): it builds here on CPU,
and on the GPU box the reference's own test cases (tests/cpp/test_reference_cases.cpp) run through it."""
import os
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SRC = os.path.join(ROOT, "tests", "cpp", "test_reference_cases.cpp")
LIBDIR = os.path.join(ROOT, "halo2_regex_b200")


def _build(tmp_path, src=SRC):
    exe = os.path.join(str(tmp_path), os.path.splitext(os.path.basename(src))[0])
    cmd = ["g++", "-std=c++17", "-O1", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"), "-o", exe, src, "-L", LIBDIR, "-lb2r", f"-Wl,-rpath,{LIBDIR}"]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_cpp_host_mirror_builds_and_refuses_to_run_without_a_gpu(tmp_path):
    import torch
    exe = _build(tmp_path)
    if not torch.cuda.is_available():
        r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "defs")], capture_output=True, text=True)
        assert r.returncode == 1 and "no CPU fallback" in r.stdout      # fails loudly, does not fall back


@pytest.mark.gpu
def test_reference_cases_through_the_cpp_host(tmp_path):
    exe = _build(tmp_path)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "defs")], capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and "all reference cases pass" in r.stdout, r.stdout + r.stderr


MULTI = os.path.join(ROOT, "tests", "cpp", "test_multi_device.cpp")


def test_multi_device_cpp_test_builds(tmp_path):
    _build(tmp_path, MULTI)


@pytest.mark.gpu
@pytest.mark.parametrize("n_dev", [2, 8])
def test_multi_device_handle_through_the_cpp_host(tmp_path, n_dev):
    import torch
    if torch.cuda.device_count() < n_dev:
        pytest.skip(f"needs {n_dev} GPUs")
    exe = _build(tmp_path, MULTI)
    r = subprocess.run([exe, os.path.join(ROOT, "tests", "golden", "defs"), str(n_dev)], capture_output=True, text=True, timeout=300)
    assert r.returncode == 0 and "equals the single-device result" in r.stdout, r.stdout + r.stderr
