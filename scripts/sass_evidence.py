#!/usr/bin/env python
"""Per-kernel SASS mnemonic counts of the built library (no GPU needed): which kernels issue tcgen05
MMAs (UTC*MMA), TMEM loads/stores (LDTM/STTM), TMA (UTMALDG/UTMASTG/UBLKCP), the legacy tensor path
(HMMA) and system-scope peer stores/loads.  `python scripts/sass_evidence.py > profiles/r2_sass_evidence.txt`"""
import collections
import re
import subprocess
import sys

LIB = sys.argv[1] if len(sys.argv) > 1 else "vattention_b200/libvattn_b200.so"
PAT = collections.OrderedDict([
    ("UTC*MMA (tcgen05.mma)", re.compile(r"\bUTC[A-Z]*MMA")),
    ("LDTM (tcgen05.ld)", re.compile(r"\bLDTM")),
    ("STTM (tcgen05.st)", re.compile(r"\bSTTM")),
    ("UTMALDG (TMA load)", re.compile(r"\bUTMALDG")),
    ("UTMASTG/UBLKCP", re.compile(r"\b(UTMASTG|UBLKCP)")),
    ("SYNCS (mbarrier)", re.compile(r"\bSYNCS")),
    ("HMMA (legacy mma.sync)", re.compile(r"\bHMMA")),
    ("MUFU.EX2", re.compile(r"\bMUFU\.EX2")),
    ("ST.E.*SYS (peer store)", re.compile(r"\bST\.E[.A-Z0-9]*\.SYS|\bST\.E\.STRONG\.SYS|\bSTG\.E[.A-Z0-9]*SYS")),
    ("LD.E.*SYS (peer load)", re.compile(r"\bLD\.E[.A-Z0-9]*\.SYS|\bLDG\.E[.A-Z0-9]*SYS")),
])


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    fn, counts, total = None, collections.OrderedDict(), {}
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            fn = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = fn.replace("(anonymous namespace)::", "").replace("vattn::", "").replace("void ", "")
            fn = re.sub(r"\(.*", "", fn)
            counts[fn] = collections.Counter()
            total[fn] = 0
            continue
        if fn is None or "/*" not in line:
            continue
        total[fn] += 1
        for name, pat in PAT.items():
            if pat.search(line):
                counts[fn][name] += 1
    names = list(PAT)
    print(f"# {LIB}: SASS mnemonic counts per kernel (cuobjdump -sass, sm_100a)")
    print("kernel | instructions | " + " | ".join(names))
    for fn in sorted(counts):
        print(f"{fn} | {total[fn]} | " + " | ".join(str(counts[fn][n]) for n in names))


if __name__ == "__main__":
    main()
