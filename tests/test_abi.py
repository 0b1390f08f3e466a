"""The C-ABI library loads and exports every symbol include/ipf_b200.h declares
(no compute calls: runs without a GPU)."""
import ctypes
import os
import re

from ipfitting_jl_b200 import _capi

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "ipf_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ipf_[a-z0-9_]+)\s*\(", text)))


def test_header_declares_the_path():
    names = declared_symbols()
    for must in ("ipf_init", "ipf_assemble", "ipf_neighbours", "ipf_solve_begin", "ipf_solve_gram",
                 "ipf_solve_factor", "ipf_solve_finish", "ipf_solve", "ipf_residual_stats", "ipf_predict"):
        assert must in names


def test_library_exports_every_declared_symbol():
    path = _capi.library_path()
    assert os.path.exists(path), "run __graft_entry__.build() first"
    lib = ctypes.CDLL(path)
    missing = [n for n in declared_symbols() if not hasattr(lib, n)]
    assert not missing, f"declared but not exported: {missing}"


def test_binding_table_matches_header():
    names = set(declared_symbols())
    bound = set(_capi.SIGNATURES)
    assert bound <= names, f"bound but not declared: {sorted(bound - names)}"
    assert names <= bound, f"declared but not bound: {sorted(names - bound)}"


def test_no_cpu_fallback_without_gpu():
    """Without a CUDA device the product path must fail loudly."""
    import torch
    if torch.cuda.is_available():
        return
    import pytest
    with pytest.raises(_capi.IPFError):
        _capi.Context(0)


def test_block_spec_layout_matches_the_c_compiler(tmp_path):
    """`ipf_block_spec` is passed by pointer from ctypes (and from the Julia shim's `BlockSpec`): size and field offsets of
    the ctypes mirror must be what gcc lays out for the header."""
    import shutil
    import subprocess
    cc = shutil.which("gcc") or shutil.which("cc")
    if cc is None:
        import pytest
        pytest.skip("no C compiler")
    fields = [f[0] for f in _capi.BlockSpec._fields_]
    src = tmp_path / "layout.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "ipf_b200.h"\nint main(void) {\n'
                   '  printf("%zu", sizeof(ipf_block_spec));\n'
                   + "".join(f'  printf(" %zu", offsetof(ipf_block_spec, {f}));\n' for f in fields)
                   + "  return 0;\n}\n")
    exe = tmp_path / "layout"
    subprocess.check_call([cc, "-I", os.path.join(ROOT, "include"), str(src), "-o", str(exe)])
    out = [int(v) for v in subprocess.check_output([str(exe)]).split()]
    assert out[0] == ctypes.sizeof(_capi.BlockSpec)
    assert out[1:] == [getattr(_capi.BlockSpec, f).offset for f in fields]
