"""CPU tests of the drop-in boundary: the C-ABI library loads and exports every symbol
that include/gps_b200.h declares; compute entry points fail loudly without a GPU."""
import ctypes as C
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "gps_b200.h")


def declared_symbols():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(gps_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_expected_surface():
    syms = declared_symbols()
    for must in ["gps_create", "gps_destroy", "gps_scan", "gps_scan_carry_slots", "gps_reduce",
                 "gps_reduce_publish", "gps_scan_emulated_deadlock", "gps_scan_host", "gps_init_ones",
                 "gps_validate_ones", "gps_sharded_scan", "gps_ipc_export", "gps_last_error"]:
        assert must in syms


def test_library_exports_every_declared_symbol():
    from gpuprefixsums_b200 import _capi as capi

    lib = C.CDLL(capi.LIB_PATH)
    missing = [s for s in declared_symbols() if not hasattr(lib, s)]
    assert not missing, f"declared in gps_b200.h but not exported: {missing}"
    # and the Python binding covers the same set
    assert sorted(capi.SIGNATURES) == declared_symbols()
    assert lib.gps_abi_version() == 1


def test_library_contains_sm100a_code_and_tma():
    from gpuprefixsums_b200 import _capi as capi

    try:
        out = subprocess.run(["cuobjdump", "-lelf", capi.LIB_PATH], capture_output=True, text=True, timeout=60)
    except FileNotFoundError:
        pytest.skip("cuobjdump not on PATH")
    assert "sm_100a" in out.stdout
    sass = subprocess.run(["cuobjdump", "-sass", capi.LIB_PATH], capture_output=True, text=True, timeout=300).stdout
    assert "csdl_tma_kernel" in sass and "csdl_reg_kernel" in sass
    assert "UBLKCP" in sass          # cp.async.bulk -> TMA bulk copy engine
    assert re.search(r"LDG\.E(\.\w+)*\.128", sass) and re.search(r"STG\.E(\.\w+)*\.128", sass)  # 128-bit path
    assert "HMMA" not in sass and "UTCHMMA" not in sass  # scan is not a contraction: no tensor cores


def test_no_cpu_fallback_without_gpu():
    import torch

    if torch.cuda.is_available():
        pytest.skip("GPU present: covered by the gpu tests")
    from gpuprefixsums_b200 import _capi as capi

    ctx = C.c_void_p()
    rc = capi.lib.gps_create(C.byref(ctx), 0, 1 << 20)
    assert rc == capi.GPS_ERR_CUDA and not ctx.value
    assert b"failed" in capi.lib.gps_last_error()
    from gpuprefixsums_b200 import PrefixSum

    with pytest.raises(RuntimeError):
        PrefixSum(1 << 20)


def test_cub_baseline_library_is_separate():
    from gpuprefixsums_b200 import _capi as capi

    assert hasattr(C.CDLL(capi.CUB_LIB_PATH), "gps_cub_scan")
    assert not hasattr(C.CDLL(capi.LIB_PATH), "gps_cub_scan")


def test_host_mirror_keeps_the_reference_class_surface():
    """The C++ dispatcher headers keep the reference's class and method names
    (SURVEY.md 8b)."""
    host = os.path.join(ROOT, "gpuprefixsums_b200", "host")
    want = {
        "ChainedScanDecoupledLookbackDispatcher.hpp": ["class ChainedScanDecoupledLookbackDispatcher",
                                                       "void TestAllExclusive()", "void TestAllInclusive()",
                                                       "void BatchTimingInclusive(uint32_t size, uint32_t batchCount)"],
        "ReduceThenScanDispatcher.hpp": ["class ReduceThenScanDispatcher", "void TestAllExclusive()",
                                         "void TestAllInclusive()",
                                         "void BatchTimingInclusive(uint32_t size, uint32_t batchCount)"],
        "EmulatedDeadlockingDispatcher.hpp": ["class EmulatedDeadlockingDispatcher", "void TestAllInclusive()",
                                              "double BatchTimingInclusive(uint32_t size, uint32_t batchCount, uint32_t partMask",
                                              "void RecordTimingData(uint32_t spinCount)"],
        "CubDispatcher.hpp": ["class CubDispatcher", "BatchTimingCubChainedScan(uint32_t size, uint32_t batchCount)"],
    }
    for fname, needles in want.items():
        text = " ".join(open(os.path.join(host, fname)).read().split())
        for n in needles:
            assert " ".join(n.split()) in text, (fname, n)
    assert os.path.exists(os.path.join(ROOT, "gpuprefixsums_b200", "bin", "gps_main"))


def test_reference_main_compiles_against_the_host_mirror():
    """The literal drop-in, guarded: the reference's main() with only its #include names swapped
    (.cuh -> .hpp) compiles against gpuprefixsums_b200/host/ (oracle/Makefile ref_main_dropin).
    Needs the reference sources, i.e. the build container."""
    ref = "/root/reference/GPUPrefixSumsCUDA/GPUPrefixSumsCUDA.cu"
    if not os.path.exists(ref):
        pytest.skip("reference sources not present")
    out = subprocess.run(["make", "-C", os.path.join(ROOT, "oracle"), "-B", os.path.join(ROOT, "oracle", "_ref", "ref_main_dropin")],
                         capture_output=True, text=True, timeout=300)
    assert out.returncode == 0, out.stdout[-1500:] + out.stderr[-1500:]
    syms = subprocess.run(["nm", "-D", "--undefined-only", os.path.join(ROOT, "oracle", "_ref", "ref_main_dropin")],
                          capture_output=True, text=True).stdout
    assert "gps_scan" in syms and "gps_scan_emulated_deadlock" in syms  # it really binds the C ABI
