"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports exactly what
include/las_b200.h declares, and the ctypes mirror of las_dec_args matches the C layout."""
import ctypes
import os
import re
import subprocess

import pytest

from tests.util import pkg

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "las_b200.h")


def declared_functions():
    src = open(HEADER).read()
    src = re.sub(r"/\*.*?\*/", "", src, flags=re.S)
    return sorted(set(re.findall(r"\b(las_[A-Za-z0-9_]+)\s*\(", src)))


def test_library_exports_every_declared_symbol():
    lib = pkg("_lib")
    assert os.path.exists(lib.LIB_PATH), "run __graft_entry__.build() first"
    h = lib.lib()
    names = declared_functions()
    assert len(names) >= 20
    for n in names:
        assert hasattr(h, n), f"{n} declared in include/las_b200.h but not exported"
    assert sorted(lib.SIGNATURES) == names, "ctypes signature table out of sync with the header"
    assert h.las_version() == 2


def test_dec_args_layout_matches_c(tmp_path):
    lib = pkg("_lib")
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "las_b200.h"\n'
                   'int main(){printf("%zu %zu %zu %zu\\n", sizeof(las_dec_args), offsetof(las_dec_args, enc_h),'
                   ' offsetof(las_dec_args, ws), offsetof(las_dec_args, denc));return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.check_call(["gcc", "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    size, o_enc, o_ws, o_denc = map(int, subprocess.check_output([str(exe)]).split())
    D = lib.DecArgs
    assert ctypes.sizeof(D) == size
    assert D.enc_h.offset == o_enc and D.ws.offset == o_ws and D.denc.offset == o_denc


def test_compute_entry_points_refuse_to_run_without_cuda():
    import torch
    if torch.cuda.is_available():
        pytest.skip("CUDA present")
    lib = pkg("_lib")
    with pytest.raises(lib.LasError):
        lib.call("las_add2", None, None, None, 0)
