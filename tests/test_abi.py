"""CPU tests of the drop-in boundary: the C-ABI library loads, exports every symbol include/*.h declares,
and refuses to compute without a GPU (no CPU fallback)."""
import ctypes as C
import subprocess

import pytest


def test_library_exports_every_declared_symbol(ks):
    names = ks.declared_symbols()
    assert "ks_select" in names and "ks_check_cell" in names and "ks_snapshot_create" in names
    missing = [n for n in names if not hasattr(ks.lib, n)]
    assert not missing, f"declared in include/*.h but not exported by libksched.so: {missing}"


def test_product_does_not_link_or_import_oracle(ks):
    out = subprocess.run(["ldd", ks.LIB_PATH], capture_output=True, text=True).stdout
    assert "oracle" not in out
    sym = subprocess.run(["nm", "-D", ks.LIB_PATH], capture_output=True, text=True).stdout
    assert "orc_" not in sym
    import os
    pkg = os.path.dirname(ks.LIB_PATH)
    for base, _, files in os.walk(pkg):
        if os.path.basename(base) == "build":
            continue
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".cpp", ".h")):
                text = open(os.path.join(base, f), errors="replace").read()
                assert "liboracle" not in text and "orc_run" not in text and "import orc" not in text, f


def test_mask_row_bytes(ks):
    assert ks.mask_row_bytes(1) == 32
    assert ks.mask_row_bytes(256) == 32
    assert ks.mask_row_bytes(257) == 64
    assert ks.mask_row_bytes(10_000) == 1280
    assert ks.mask_row_bytes(50_000) == 6272


def test_no_gpu_means_loud_failure_not_fallback(ks):
    if ks.device_count() > 0:
        pytest.skip("GPU present")
    h = C.c_void_p()
    rc = ks.lib.ks_snapshot_create(0, C.byref(h))
    assert rc == -8  # KS_ERR_NO_DEVICE
    assert b"no CPU fallback" in ks.lib.ks_last_error()
    with pytest.raises(ks.KsError):
        ks.Snapshot(0)
