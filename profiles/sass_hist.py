#!/usr/bin/env python
"""Instruction evidence for the Blackwell paths: per kernel of libdet3d_b200.so, counts of the SASS mnemonics that prove
tcgen05 (UTC*MMA), TMEM access (LDTM / STTM), TMA (UTMALDG = tensor-map loads, UBLKCP = bulk copies), mbarrier
(SYNCS), async copies (LDGSTS) and reductions (RED / ATOM).  Run here (no GPU needed):
    python profiles/sass_hist.py > profiles/r2_sass_hist.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "det3d_b200", "lib", "libdet3d_b200.so")
PAT = re.compile(r"\b(UTC[A-Z]*MMA|UTCBAR|UTCCP|LDTM|STTM|UTMALDG|UTMASTG|UTMAPF|UBLKCP|SYNCS|LDGSTS|RED|ATOMG|ATOMS|HMMA|FFMA|DFMA)\b")


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    name = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            per[name] = collections.Counter()
            continue
        if name:
            for tok in PAT.findall(line):
                per[name][tok] += 1
    total = collections.Counter()
    print("# SASS mnemonic histogram of %s (cuobjdump -sass, sm_100a)" % os.path.relpath(LIB, ROOT))
    print("# UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UTMALDG = cp.async.bulk.tensor (tensor map), UBLKCP = cp.async.bulk,")
    print("# SYNCS = mbarrier, LDGSTS = cp.async, UTCBAR = tcgen05.commit\n")
    for k, c in per.items():
        total.update(c)
        interesting = {t: n for t, n in c.items() if t not in ("FFMA", "DFMA")}
        if interesting:
            print("%-70s %s" % (k[:70], "  ".join("%s=%d" % (t, n) for t, n in sorted(c.items()))))
    print("\nTOTAL  " + "  ".join("%s=%d" % (t, n) for t, n in sorted(total.items())))


if __name__ == "__main__":
    main()
