"""Generate tests/golden/alloc_trace_*.json from the REFERENCE allocator itself.

Runs on the GPU box:  python oracle/gen_alloc_golden.py [outdir]   (default gpurun_out/golden)
For every named trace (oracle/alloc_traces.py) the concrete API call list is replayed on
oracle/_ref/vattention_ref*.so by oracle/ref_driver.py in a fresh process (the reference keeps
its state in process globals) and the per-call snapshots are stored next to the calls.
TEST INFRASTRUCTURE ONLY.
"""
import json
import os
import subprocess
import sys
import tempfile

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
from oracle import alloc_traces as T  # noqa: E402


def run_reference(concrete):
    with tempfile.TemporaryDirectory() as td:
        tp, op = os.path.join(td, "trace.json"), os.path.join(td, "out.json")
        json.dump(concrete, open(tp, "w"))
        r = subprocess.run([sys.executable, os.path.join(HERE, "ref_driver.py"), tp, op],
                           capture_output=True, text=True, timeout=600)
        if r.returncode != 0:
            raise RuntimeError(f"reference run failed ({r.returncode}):\n{r.stdout[-2000:]}\n{r.stderr[-2000:]}")
        return json.load(open(op))


def main(outdir):
    os.makedirs(outdir, exist_ok=True)
    for name in T.TRACE_NAMES:
        mb = T.ModelBackend()
        concrete = T.expand(T.build_trace(name), mb)
        snaps = run_reference(concrete)
        for s in snaps:
            s.pop("op", None)
        json.dump({"name": name, "source": "reference vattention.cu @ef3fff25 on B200 (oracle/_ref)",
                   "trace": concrete, "snaps": snaps},
                  open(os.path.join(outdir, f"alloc_trace_{name}.json"), "w"), separators=(",", ":"))
        print(name, len(concrete), "ops ->", outdir)


if __name__ == "__main__":
    main(sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(HERE), "gpurun_out", "golden"))
