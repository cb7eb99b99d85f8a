"""The C-ABI library builds, loads without a GPU and exports exactly what include/liftreg_b200.h declares."""
import ctypes
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "liftreg_b200.h")


def declared_symbols():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"LR_API\s+[\w\s\*]+?\b(lr_\w+)\s*\(", src)))


@pytest.fixture(scope="module")
def lib_path():
    from liftreg_b200 import build
    return build.build()


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for must in ["lr_drr_forward", "lr_drr_backward", "lr_backproject_forward", "lr_backproject_backward",
                 "lr_warp_forward", "lr_warp_backward", "lr_drr_forward_host", "lr_backproject_forward_host",
                 "lr_warp_forward_host", "lr_last_error", "lr_project_grid", "lr_backproj_grid", "lr_identity_map"]:
        assert must in syms
    assert len(syms) >= 20


def test_library_exports_every_declared_symbol(lib_path):
    out = subprocess.run(["nm", "-D", "--defined-only", lib_path], capture_output=True, text=True, check=True).stdout
    exported = set(re.findall(r"\sT\s+(lr_\w+)", out))
    missing = [s for s in declared_symbols() if s not in exported]
    assert not missing, "declared in the header but not exported: %s" % missing
    extra = sorted(exported - set(declared_symbols()))
    assert not extra, "exported but not declared in the header: %s" % extra
    # nothing but the C-ABI is visible (no C++ symbols leak)
    leaked = [l for l in out.splitlines() if " T " in l and "lr_" not in l]
    assert not leaked, leaked


def test_ctypes_table_matches_header(lib_path):
    from liftreg_b200 import _native
    assert sorted(_native.SIGNATURES) == declared_symbols()
    lib = _native.lib()          # binds every signature; raises on a missing symbol
    assert lib.lr_abi_version() == 1


def test_library_contains_sm_100a_code_only(lib_path):
    out = subprocess.run(["cuobjdump", "-lelf", lib_path], capture_output=True, text=True).stdout
    archs = set(re.findall(r"sm_(\d+a?)", out))
    assert archs == {"100a"}, archs


def test_no_gpu_calls_fail_cleanly_without_a_device(lib_path):
    """On the CPU-only build box the library must report, not crash; on a GPU box this just counts devices."""
    from liftreg_b200 import _native
    lib = _native.lib()
    n = lib.lr_device_count()
    assert n >= 1 or (n == -4 and lib.lr_last_error())
    # argument validation happens before any CUDA call
    assert lib.lr_warp_forward(None, None, 1, 1, 2, 2, 2, 0, 0, 1, 0, None, None) == -1
    assert b"null" in lib.lr_last_error()
    assert lib.lr_drr_forward_host_workspace_bytes(1, 160, 160, 160, 4, 240, 240) == 4 * (160 ** 3 + 4 * 240 * 240)
    assert lib.lr_warp_forward_host_workspace_bytes(0, 1, 2, 2, 2) == 0


def test_missing_library_raises_loudly(monkeypatch, tmp_path):
    from liftreg_b200 import _native
    monkeypatch.setattr(_native, "_lib", None)
    monkeypatch.setattr(_native, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(_native.NativeLibraryError, match="no CPU/PyTorch fallback"):
        _native.lib()


def test_oracle_library_builds_and_exports():
    from oracle import c_oracle
    so = c_oracle.build()
    out = subprocess.run(["nm", "-D", "--defined-only", so], capture_output=True, text=True, check=True).stdout
    for s in ["lro_drr_forward", "lro_backproject_forward", "lro_warp_forward", "lro_project_grid"]:
        assert s in out
    assert ctypes.CDLL(so).lro_version() == 1
