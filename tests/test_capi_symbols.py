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
