"""CPU: the reference arm of bench.py (`--impl reference`) prints one JSON line with the contract's keys and never loads libb2r.so."""
import json
import os
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_reference_arm_prints_the_contract_line():
    env = dict(os.environ, PYTHONDONTWRITEBYTECODE="1")
    out = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "1"], cwd=ROOT, env=env,
                         capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
    line = json.loads(out.stdout.strip().splitlines()[-1])
    assert line["impl"] == "reference" and line["metric"] == "input bytes/sec (DFA witness gen)" and line["unit"] == "GB/s"
    assert line["higher_is_better"] is True and line["steps"] == 1 and line["warmup"] == 1 and line["value"] > 0
    assert line["config"]["strings_per_gpu"] == 1 << 20 and line["config"]["max_chars_size"] == 1025          # the GPU arm's config object
    cb = line["cpu_baseline"]
    assert cb["kind"] == "port" and cb["cores"] >= 1 and cb["value"] == line["value"] and "sample" in cb
    assert line["e2e"] == {"value": line["value"], "unit": "GB/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}


def test_reference_arm_does_not_load_the_cuda_library():
    code = ("import sys, runpy; sys.argv = ['bench.py', '--impl', 'reference', '--steps', '1', '--warmup', '0']\n"
            "runpy.run_path('bench.py', run_name='__main__')\n"
            "import os\n"
            "maps = open('/proc/self/maps').read()\n"
            "assert 'libb2r.so' not in maps, 'the reference arm mapped libb2r.so'\n"
            "assert 'halo2_regex_b200' not in sys.modules\n")
    out = subprocess.run([sys.executable, "-c", code], cwd=ROOT, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stderr[-2000:]
