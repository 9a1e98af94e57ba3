"""The C-ABI library loads and exports every symbol include/gmr_b200.h declares (no compute)."""
import ctypes
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_functions():
    src = open(os.path.join(ROOT, "include", "gmr_b200.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(gmr_[a-z0-9_]+)\s*\(", src)))


def test_header_lists_expected_entry_points():
    names = declared_functions()
    for n in ("gmr_model_create", "gmr_model_destroy", "gmr_retarget_batch", "gmr_retarget_batch_f64",
              "gmr_retarget_batch_host", "gmr_last_error", "gmr_launch_count", "gmr_kernel_info",
              "gmr_oracle_retarget_batch"):
        assert n in names


def test_cuda_library_exports_every_declared_symbol(built):
    from general_motion_retargeting_b200 import _native
    lib = ctypes.CDLL(str(_native.LIB_PATH))
    for n in declared_functions():
        if n.startswith("gmr_oracle_"):
            continue
        assert hasattr(lib, n), n
    assert sorted(_native.EXPORTED_SYMBOLS) == [n for n in declared_functions() if not n.startswith("gmr_oracle_")]
    assert lib.gmr_launch_count() == 0


def test_oracle_library_exports_its_symbol(built):
    lib = ctypes.CDLL(os.path.join(ROOT, "oracle", "liboracle.so"))
    assert hasattr(lib, "gmr_oracle_retarget_batch")


def test_cuda_library_has_blackwell_sass(built):
    """sm_100a cubin with the TMA bulk copy (UBLKCP) and cp.async (LDGSTS) the design relies on."""
    import shutil
    import subprocess
    from general_motion_retargeting_b200 import _native
    if not shutil.which("cuobjdump"):
        pytest.skip("cuobjdump not available")
    out = subprocess.run(["cuobjdump", "-sass", str(_native.LIB_PATH)], capture_output=True, text=True).stdout
    assert "sm_100a" in out
    assert "UBLKCP" in out and "LDGSTS" in out
    assert "HMMA" not in out          # no tensor-core path: the contractions are 6x6


def test_missing_library_fails_loudly(monkeypatch):
    from general_motion_retargeting_b200 import _native
    monkeypatch.setattr(_native, "_LIB", None)
    monkeypatch.setenv("GMR_B200_LIB", "/nonexistent/libgmr_b200.so")
    with pytest.raises(_native.NativeLibraryMissing):
        _native.load_library()
    from general_motion_retargeting_b200 import GeneralMotionRetargeting
    with pytest.raises(_native.NativeLibraryMissing):
        GeneralMotionRetargeting("smplx", "unitree_g1")


def test_product_never_imports_the_oracle():
    pkg = os.path.join(ROOT, "general_motion_retargeting_b200")
    for dp, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dp, f)).read()
                assert "import oracle" not in txt and "from oracle" not in txt and "liboracle" not in txt, f
                assert "libgmr_emu" not in txt and "tests/emu" not in txt.replace("tests/emu debugs", ""), f


def test_bench_reference_arm_prints_the_contract_line(built):
    """`bench.py --impl reference` runs the CPU restatement on a bounded sample and prints ONE JSON line with the
    keys the driver reads (no GPU needed)."""
    import json
    import subprocess
    import sys
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--steps", "1", "--warmup", "0",
                        "--frames", "12"], capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["unit"] == "frames/s" and d["higher_is_better"] is True and d["value"] > 0
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "frames/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert "workload" in d["config"]
