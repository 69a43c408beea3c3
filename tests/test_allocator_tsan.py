"""The allocator's API thread and mapper thread under ThreadSanitizer (host mock driver, no GPU).

step_async may queue a step behind the mapper pass in flight: the fast path reads a snapshot under the
mutex while the mapper works on the live bookkeeping without it.  tests/native/allocator_tsan.cpp
drives ~1500 decode steps (most of them queued) mixed with the calls that wait for the mapper; any
data race or lock-order problem fails the run."""
import os
import shutil
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
CSRC = os.path.join(ROOT, "vattention_b200", "csrc")


@pytest.mark.timeout(300)
def test_allocator_threads_are_race_free(tmp_path):
    gxx = shutil.which("g++")
    if gxx is None:
        pytest.skip("no g++")
    exe = str(tmp_path / "allocator_tsan")
    cmd = [gxx, "-fsanitize=thread", "-O1", "-g", "-std=c++17", "-I/usr/local/cuda/include",
           os.path.join(ROOT, "tests", "native", "allocator_tsan.cpp"),
           *(os.path.join(CSRC, f) for f in ("kv_allocator.cpp", "vmm_driver.cpp", "capi_alloc.cpp")),
           "-ldl", "-lpthread", "-o", exe]
    b = subprocess.run(cmd, capture_output=True, text=True)
    if b.returncode != 0 and "tsan" in (b.stderr or "").lower():
        pytest.skip("this toolchain has no ThreadSanitizer runtime")
    assert b.returncode == 0, b.stderr[-2000:]
    r = subprocess.run([exe], capture_output=True, text=True, timeout=240,
                       env=dict(os.environ, TSAN_OPTIONS="halt_on_error=1 exitcode=66"))
    if "unexpected memory mapping" in r.stderr or "ThreadSanitizer: CHECK failed" in r.stderr:
        pytest.skip("ThreadSanitizer cannot start under this kernel's address-space layout")
    assert "WARNING: ThreadSanitizer" not in r.stderr, r.stderr[-4000:]
    assert r.returncode == 0, (r.returncode, r.stdout[-500:], r.stderr[-2000:])   # 2 = nothing was queued
    assert "queued" in r.stdout
