"""The C-ABI library loads and exports every symbol include/crgpu.h declares (no compute calls:
this runs without a GPU)."""
import ctypes as C
import os
import re

import pytest

from cellranger_b200 import _lib, build

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def _declared_symbols():
    src = open(os.path.join(ROOT, "include", "crgpu.h")).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(crgpu_[a-z0-9_]+)\s*\(", src)))


def test_header_declares_entry_points():
    syms = _declared_symbols()
    for must in ("crgpu_ctx_create", "crgpu_whitelist_add", "crgpu_pass1", "crgpu_pass2", "crgpu_count",
                 "crgpu_correct_barcodes", "crgpu_matrix_get", "crgpu_keys_partition"):
        assert must in syms


def test_library_builds_and_exports_every_symbol():
    so = build.build_library()
    assert os.path.exists(so)
    L = C.CDLL(so)
    missing = [s for s in _declared_symbols() if not hasattr(L, s)]
    assert not missing, f"symbols declared in crgpu.h but not exported: {missing}"


def test_struct_sizes_match_header(tmp_path):
    """The ctypes mirrors of the ABI structs against what a C compiler makes of include/crgpu.h (size and the
    offset of every pointer field)."""
    import subprocess

    src = tmp_path / "layout.c"
    src.write_text('#include <stddef.h>\n#include <stdio.h>\n#include "crgpu.h"\n'
                   'int main(void) { printf("%zu %zu %zu %zu %zu %zu %zu %zu\\n", sizeof(crgpu_library_def), '
                   'sizeof(crgpu_read_batch), offsetof(crgpu_read_batch, r1_seq), offsetof(crgpu_read_batch, feature), '
                   'offsetof(crgpu_read_batch, r2_seq), offsetof(crgpu_read_batch, on_device), '
                   'offsetof(crgpu_read_batch, select_key), sizeof(crgpu_synth_params)); return 0; }\n')
    exe = tmp_path / "layout"
    res = subprocess.run(["gcc", "-std=c99", "-I" + os.path.join(ROOT, "include"), str(src), "-o", str(exe)],
                         capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
    got = [int(x) for x in subprocess.run([str(exe)], capture_output=True, text=True).stdout.split()]
    rb = _lib.ReadBatch
    assert got == [C.sizeof(_lib.LibraryDef), C.sizeof(rb), rb.r1_seq.offset, rb.feature.offset, rb.r2_seq.offset,
                   rb.on_device.offset, rb.select_key.offset, C.sizeof(_lib.SynthParams)]
    assert C.sizeof(_lib.LibraryDef) == 40


def test_no_cpu_fallback_without_device():
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    L = _lib.load()
    ctx = C.c_void_p()
    rc = L.crgpu_ctx_create(0, C.byref(ctx))
    assert rc == -2  # CRGPU_E_CUDA
    assert b"no CPU fallback" in L.crgpu_last_error()
    from cellranger_b200 import GemWell, CrgpuError

    with pytest.raises(CrgpuError):
        GemWell()


def test_product_never_imports_oracle():
    pkg = os.path.join(ROOT, "cellranger_b200")
    for dirpath, _, files in os.walk(pkg):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                txt = open(os.path.join(dirpath, f), errors="replace").read()
                assert "oracle" not in txt.lower().replace("# oracle", ""), f"{f} mentions the oracle"


def test_header_is_plain_c(tmp_path):
    """include/crgpu.h must be consumable by a C compiler (the boundary is a C ABI, not a C++ one)."""
    import os
    import subprocess

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    src = tmp_path / "use_header.c"
    src.write_text('#include "crgpu.h"\n'
                   "int main(void) { crgpu_library_def d; crgpu_read_batch b; (void)d; (void)b;\n"
                   "  return sizeof(d) == 40 && CRGPU_STAT_COUNT == 17 ? 0 : 1; }\n")
    res = subprocess.run(["gcc", "-std=c99", "-Wall", "-Werror", "-pedantic", "-I" + os.path.join(root, "include"),
                          "-fsyntax-only", str(src)], capture_output=True, text=True)
    assert res.returncode == 0, res.stderr
