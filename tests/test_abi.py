"""CPU-side checks of the drop-in boundary: the library loads, exports every symbol that
include/mindrec_b200.h declares, and validates its arguments without touching a GPU."""
import ctypes
import os
import re

import pytest

from mindrec_b200 import _lib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "mindrec_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(mrec_\w+)\s*\(", text)))


def test_header_declares_symbols():
    syms = _declared_symbols()
    assert "mrec_gather" in syms and "mrec_unique" in syms and "mrec_sparse_lazy_adam" in syms
    assert len(syms) >= 15


def test_library_exports_every_declared_symbol(built_lib):
    lib = ctypes.CDLL(built_lib)
    missing = [s for s in _declared_symbols() if not hasattr(lib, s)]
    assert not missing, "declared in include/mindrec_b200.h but not exported: %s" % missing


def test_version_and_counters(built_lib):
    assert "sm_100a" in _lib.version()
    assert _lib.launch_count() >= 0


def test_bad_nparam_is_rejected_without_gpu(built_lib):
    # validation happens before any CUDA call, so this is safe on a CPU-only box
    rc = _lib.aot_call_raw("mrec_gather", [16, 32], [(4, 4), (2,)], ["float32", "int32"])
    assert rc == 1  # MREC_ERR_NPARAM
    assert "expected" in _lib.last_error()


def test_bad_dtype_is_rejected_without_gpu(built_lib):
    rc = _lib.aot_call_raw("mrec_gather", [16, 32, 64], [(4, 4), (2,), (2, 4)],
                           ["float16", "int32", "float32"])
    assert rc == 2  # MREC_ERR_DTYPE
    rc = _lib.aot_call_raw("mrec_gather", [16, 32, 64], [(4, 4), (2,), (2, 4)],
                           ["float32", "float32", "float32"])
    assert rc == 2


def test_bad_shape_and_alignment_are_rejected_without_gpu(built_lib):
    rc = _lib.aot_call_raw("mrec_gather", [16, 32, 64], [(4, 4), (2,), (3, 4)],
                           ["float32", "int32", "float32"])
    assert rc == 3  # MREC_ERR_SHAPE
    rc = _lib.aot_call_raw("mrec_gather", [20, 32, 64], [(4, 4), (2,), (2, 4)],
                           ["float32", "int32", "float32"])
    assert rc == 4  # MREC_ERR_ALIGN
    rc = _lib.aot_call_raw("mrec_gather", [0, 32, 64], [(4, 4), (2,), (2, 4)],
                           ["float32", "int32", "float32"])
    assert rc == 8  # MREC_ERR_NULL


def test_workspace_size_helpers(built_lib):
    lib = ctypes.CDLL(built_lib)
    for name in ("mrec_unique_workspace_bytes", "mrec_unique_first_workspace_bytes",
                 "mrec_sparse_opt_workspace_bytes"):
        f = getattr(lib, name)
        f.restype = ctypes.c_size_t
        f.argtypes = [ctypes.c_int64, ctypes.c_int]
    assert lib.mrec_unique_workspace_bytes(624000, 4) >= 624000 * 12
    assert lib.mrec_unique_workspace_bytes(0, 4) > 0
    assert lib.mrec_sparse_opt_workspace_bytes(624000, 80) >= 624000 // 32 * 2 * 80 * 4


def test_missing_library_fails_loudly(monkeypatch, tmp_path):
    monkeypatch.setattr(_lib, "_lib", None)
    monkeypatch.setattr(_lib, "LIB_PATH", str(tmp_path / "nope.so"))
    with pytest.raises(RuntimeError, match="no CPU / PyTorch fallback"):
        _lib.lib()


def test_every_aot_entry_rejects_an_empty_argument_pack(built_lib):
    """Every aot symbol of the header validates nparam before touching CUDA (error 1), never crashes."""
    helpers = ("_workspace_bytes", "mrec_version", "mrec_last_error", "mrec_launch_count", "mrec_peer_alloc",
               "mrec_peer_free", "mrec_ipc_", "mrec_tfrecord_", "mrec_crc32c", "mrec_varint_pack", "mrec_rt_")
    for sym in _declared_symbols():
        if any(h in sym for h in helpers):
            continue
        rc = _lib.aot_call_raw(sym, [], [], [])
        assert rc == 1, "%s returned %d for nparam = 0" % (sym, rc)
        assert _lib.last_error()


def test_dense_glue_validation_without_gpu(built_lib):
    lib = ctypes.CDLL(built_lib)
    lib.mrec_relu_bwd_bias_workspace_bytes.restype = ctypes.c_size_t
    lib.mrec_relu_bwd_bias_workspace_bytes.argtypes = [ctypes.c_int64]
    ws = lib.mrec_relu_bwd_bias_workspace_bytes(1024)
    assert ws >= 1024 * 4
    # g int32: dtype error
    rc = _lib.aot_call_raw("mrec_relu_bwd_bias", [16, 32, 48, 64, 80], [(4, 8), (4, 8), (4, 8), (8,), (ws,)],
                           ["int32", "int32", "int32", "float32", "uint8"])
    assert rc == 2
    # y of another shape
    rc = _lib.aot_call_raw("mrec_relu_bwd_bias", [16, 32, 48, 64, 80], [(4, 8), (4, 4), (4, 8), (8,), (ws,)],
                           ["float16", "float16", "float16", "float32", "uint8"])
    assert rc == 3
    # workspace too small
    rc = _lib.aot_call_raw("mrec_relu_bwd_bias", [16, 32, 48, 64, 80], [(4, 8), (4, 8), (4, 8), (8,), (16,)],
                           ["float16", "float16", "float16", "float32", "uint8"])
    assert rc == 7
    # head: w of the wrong length
    rc = _lib.aot_call_raw("mrec_dense_head_fwd", [16, 32, 48, 64], [(4, 8), (7,), (1,), (4,)],
                           ["float16", "float16", "float16", "float32"])
    assert rc == 3
    rc = _lib.aot_call_raw("mrec_dense_head_fwd", [16, 32, 48, 64], [(4, 8), (8,), (1,), (4,)],
                           ["float16", "float32", "float16", "float32"])
    assert rc == 2


def test_peer_exchange_validation_without_gpu(built_lib):
    # shard_offsets: every param int32
    rc = _lib.aot_call_raw("mrec_shard_offsets", [16] * 6, [(6,), (2,), (2,), (3,), (2,), (1,)],
                           ["int32", "int32", "float32", "int32", "int32", "int32"])
    assert rc == 2
    # push_rows: the modulo transform is for int32 keys only
    rc = _lib.aot_call_raw("mrec_push_rows_to_peers", [16] * 7, [(8, 4), (3,), (2,), (2,), (8, 0), (5, 0), (1,)],
                           ["float32", "int32", "int32", "int64", "float32", "float32", "int32"])
    assert rc == 2
    # more ranks than the kernels support
    rc = _lib.aot_call_raw("mrec_peer_wait", [16, 32, 48], [(4096,), (1,), (1,)], ["int32", "int32", "int32"])
    assert rc == 3


def test_every_exported_mrec_symbol_is_declared_in_the_header(built_lib):
    """The reverse direction: nothing is exported that include/mindrec_b200.h does not document."""
    import shutil
    import subprocess
    if shutil.which("nm") is None:
        pytest.skip("binutils nm not available")
    out = subprocess.run(["nm", "-D", "--defined-only", built_lib], capture_output=True, text=True, check=True).stdout
    exported = {l.split()[-1] for l in out.splitlines() if " T mrec_" in l}
    assert exported, "no mrec_* symbols exported?"
    assert not (exported - set(_declared_symbols())), sorted(exported - set(_declared_symbols()))


def test_c_harness_compiles_against_the_header_and_runs(built_lib, tmp_path):
    """The boundary is usable from plain C: tests/c/abi_harness.c includes include/mindrec_b200.h (C99), dlopens the
    library and checks the documented error codes — no Python, no torch in that process."""
    import shutil
    import subprocess
    if shutil.which("gcc") is None:
        pytest.skip("gcc not available")
    exe = str(tmp_path / "abi_harness")
    subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-I", os.path.join(ROOT, "include"),
                    os.path.join(ROOT, "tests", "c", "abi_harness.c"), "-ldl", "-o", exe], check=True)
    r = subprocess.run([exe, built_lib], capture_output=True, text=True)
    assert r.returncode == 0 and "ABI HARNESS OK" in r.stdout, r.stdout + r.stderr
