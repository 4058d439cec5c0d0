"""bench.py contract checks that need no GPU: the reference arm prints ONE JSON line with the
keys the driver reads; the product arm fails loudly (no CPU fallback) when there is no CUDA
device; the product package never imports the oracle."""
import json
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def run_bench(*args, timeout=600):
    env = dict(os.environ, CUDA_VISIBLE_DEVICES="" if not torch.cuda.is_available() else os.environ.get("CUDA_VISIBLE_DEVICES", ""))
    return subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), *args], capture_output=True, text=True,
                          timeout=timeout, cwd=ROOT, env=env)


def test_reference_arm_json_line():
    out = run_bench("--impl", "reference", "--steps", "2", "--warmup", "1", "--log2-keys-per-gpu", "22")
    assert out.returncode == 0, out.stderr[-2000:]
    lines = [l for l in out.stdout.splitlines() if l.strip()]
    assert len(lines) == 1, lines
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "Gkeys/s" and d["higher_is_better"] is True
    assert d["value"] > 0 and d["steps"] == 2
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["sample"]
    assert d["e2e"] == {"value": d["value"], "unit": "Gkeys/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["config"]["workload"].startswith("uint32 inclusive scan")
    assert d["warmup"] == 1, "the reference arm honours --warmup exactly"
    # the same `config` object as the product arm prints for the same flags
    sys.path.insert(0, ROOT)
    import bench

    import argparse

    ns = argparse.Namespace(algo="csdl", log2_keys_per_gpu=22, log2_total=0, carry="p2p")
    assert d["config"] == bench.workload_config(ns, 1)
    assert set(d["config"]) == {"workload", "keys_total", "keys_per_gpu", "in_place", "l2"}


def test_reference_arm_never_loads_the_product_library():
    """--impl reference must not map libgps_b200.so (nor import the package) at any point."""
    code = (
        "import sys, types; sys.argv=['bench.py','--impl','reference','--steps','1','--warmup','0','--log2-keys-per-gpu','20'];"
        f"sys.path.insert(0, {ROOT!r}); import bench; bench.main();"
        "maps=open('/proc/self/maps').read();"
        "assert 'libgps_b200' not in maps, 'product library mapped in the reference arm';"
        "assert not any(m.startswith('gpuprefixsums_b200') for m in sys.modules), 'product package imported';"
        "assert 'liboracle_scan' in maps"
    )
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=600, cwd=ROOT)
    assert out.returncode == 0, out.stderr[-2000:]


@pytest.mark.skipif(torch.cuda.is_available(), reason="needs a box without a GPU")
def test_product_arm_fails_loudly_without_gpu():
    out = run_bench("--steps", "1", "--warmup", "1", "--log2-keys-per-gpu", "20", timeout=300)
    assert out.returncode != 0
    assert not [l for l in out.stdout.splitlines() if l.startswith("{")], "no bench line may be printed without a GPU"


def test_product_package_never_touches_the_oracle():
    pkg = os.path.join(ROOT, "gpuprefixsums_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h", ".hpp", ".cpp")):
                text = open(os.path.join(dirpath, f), errors="ignore").read()
                # (comments may NAME the oracle; nothing may import, include, load or execute it)
                pat = r'(?:^|\n)\s*(?:import|from)\s+\S*oracle|#include\s+"[^"]*oracle|CDLL\([^)]*oracle|liboracle|oracle_lib|libref_shim'
                assert not re.search(pat, text), os.path.join(dirpath, f)
