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


def _build_c_example(ks, tmp_path):
    import os
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    exe = str(tmp_path / "reconcile_loop")
    libdir = os.path.dirname(ks.LIB_PATH)
    cmd = ["gcc", "-std=c99", "-Wall", "-Wextra", "-pedantic", "-Werror", "-I" + os.path.join(root, "include"),
           os.path.join(root, "examples", "reconcile_loop.c"), "-L" + libdir, "-lksched", "-Wl,-rpath," + libdir, "-o", exe]
    r = subprocess.run(cmd, capture_output=True, text=True)
    assert r.returncode == 0, r.stderr
    return exe


def test_headers_are_plain_c_and_the_c_example_links(ks, tmp_path):
    """include/*.h must be consumable by a C (not C++) host: the example controller loop compiles as strict C99,
    links against libksched.so alone and, without a GPU, fails loudly instead of computing on the CPU."""
    exe = _build_c_example(ks, tmp_path)
    r = subprocess.run([exe], capture_output=True, text=True, timeout=120)
    if ks.device_count() > 0:
        assert r.returncode == 0, r.stderr
        assert "p9: already bound, skipped" in r.stdout
    else:
        assert r.returncode == 3 and "no CPU fallback" in r.stderr


@pytest.mark.gpu
@pytest.mark.parametrize("mode", [[], ["--batch"]])
def test_c_example_drives_the_abi_on_a_gpu(ks, tmp_path, mode):
    """examples/reconcile_loop.c (strict C99, links libksched.so only) run on the B200: the reference's reconcile loop
    through the C ABI, pod by pod and as a drained batch."""
    exe = _build_c_example(ks, tmp_path)
    r = subprocess.run([exe] + mode, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-2000:]
    assert ("round(s)" if mode else "already bound, skipped") in r.stdout


def test_rust_shim_binds_only_exported_symbols(ks):
    """examples/rust_shim/ksched_sys.rs cannot be compiled here (no Rust toolchain); at least every function its
    `extern "C"` block names must be exported by libksched.so, and every struct field list must match the ctypes
    mirror of the same header (field names in order)."""
    import os
    import re
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = open(os.path.join(root, "examples", "rust_shim", "ksched_sys.rs")).read()
    fns = re.findall(r"pub fn (ksh?_[a-z0-9_]+)\(", src)
    assert len(fns) > 40
    for name in fns:
        assert hasattr(ks.lib, name), f"{name} is declared in ksched_sys.rs but not exported"
    assert set(fns) <= set(ks.declared_symbols())

    def rust_fields(struct):
        body = re.search(r"pub struct %s \{(.*?)\n\}" % struct, src, re.S).group(1)
        return re.findall(r"pub ([a-z_0-9]+):", body)

    assert rust_fields("ks_pod_obj") == [f[0] for f in ks.objects.ks_pod_obj._fields_]
    assert rust_fields("ks_node_obj") == [f[0] for f in ks.objects.ks_node_obj._fields_]
    assert rust_fields("ks_bindings") == [f[0] for f in ks.capi.ks_bindings._fields_]
    assert rust_fields("ks_exchange") == [f[0] for f in ks.capi.ks_exchange._fields_]
    assert rust_fields("ks_pods") == [f[0] for f in ks.capi.ks_pods._fields_]
