"""The C-ABI library loads without a GPU/driver and exports every symbol the header declares."""
import ctypes
import re

from vattention_b200 import _lib


def _declared(repo_root):
    text = (repo_root / "include" / "vattn_b200.h").read_text()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(vattn_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_exported(repo_root):
    names = _declared(repo_root)
    assert len(names) >= 30
    lib = ctypes.CDLL(str(_lib.LIB_PATH))
    missing = [n for n in names if not hasattr(lib, n)]
    assert not missing, missing


def test_binding_table_matches_header(repo_root):
    assert sorted(_lib.EXPORTED_SYMBOLS) == _declared(repo_root)


def test_version_and_error_strings():
    assert b"sm_100a" in _lib.lib.vattn_version()
    assert _lib.lib.vattn_destroy(None) == 0
    # null handle -> error code, message retrievable
    assert _lib.lib.vattn_cleanup(None) == _lib.ERR_INVALID
    assert "null allocator" in _lib.last_error()


def test_no_cpu_fallback_for_cuda_backend():
    """Without a driver the CUDA backend must fail loudly, not degrade to something else."""
    import torch
    if torch.cuda.is_available():
        return
    h = ctypes.c_void_p()
    rc = _lib.lib.vattn_create(ctypes.byref(h), _lib.BACKEND_CUDA)
    assert rc < 0
    assert "libcuda" in _lib.last_error() or "driver" in _lib.last_error().lower()


def test_header_is_plain_c_and_struct_layouts_match_the_binding(repo_root, tmp_path):
    """include/vattn_b200.h compiles as C11 (no C++, no torch types) and every struct the ctypes
    binding mirrors has the same size and field offsets as the compiler gives the header's."""
    import shutil
    import subprocess
    gcc = shutil.which("gcc")
    if gcc is None:
        import pytest
        pytest.skip("no C compiler")
    structs = {"vattn_fwd_params_t": _lib.FwdParams, "vattn_config_t": _lib.VattnConfig,
               "vattn_step_stats_t": _lib.StepStats}
    lines = ['#include <stddef.h>', '#include <stdio.h>', '#include "vattn_b200.h"', "int main(void) {"]
    for cname, cls in structs.items():
        lines.append(f'  printf("{cname} size %zu\\n", sizeof({cname}));')
        for fname, _ in cls._fields_:
            lines.append(f'  printf("{cname} {fname} %zu\\n", offsetof({cname}, {fname}));')
    lines += ["  return 0;", "}"]
    src = tmp_path / "abi.c"
    src.write_text("\n".join(lines))
    exe = tmp_path / "abi"
    subprocess.run([gcc, "-std=c11", "-Wall", "-Werror", "-pedantic", f"-I{repo_root / 'include'}", str(src),
                    "-o", str(exe)], check=True, capture_output=True)
    out = subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout
    got = {}
    for ln in out.splitlines():
        s, f, v = ln.split()
        got[(s, f)] = int(v)
    for cname, cls in structs.items():
        assert got[(cname, "size")] == ctypes.sizeof(cls), cname
        for fname, _ in cls._fields_:
            assert got[(cname, fname)] == getattr(cls, fname).offset, (cname, fname)
