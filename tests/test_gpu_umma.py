"""GPU: the tcgen05 descriptor encodings the tensor-core kernels rely on (device self test)."""
import ctypes

import pytest

from vattention_b200 import _lib

pytestmark = pytest.mark.gpu


def test_umma_descriptor_selftest():
    import torch
    torch.zeros(1, device="cuda")
    buf = ctypes.create_string_buffer(2048)
    rc = _lib.lib.vattn_selftest_umma(buf, len(buf), None)
    report = buf.value.decode()
    print(report)
    assert rc == 0, report
