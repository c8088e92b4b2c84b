#!/usr/bin/env python
"""SASS census of libdmvs_b200.so: which Blackwell instructions each kernel family carries (profiles/r2_sass_census.txt).

    python tools/sass_census.py > profiles/r2_sass_census.txt

UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG = cp.async.bulk.tensor (TMA), UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit,
SYNCS = mbarrier, REDG = red.global (vector reductions of the W1 backward), HMMA = legacy mma.sync (must be absent).
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "dmvsnet_b200", "libdmvs_b200.so")
WATCH = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "HGMMA", "REDG", "LDGSTS", "LDS", "LDG", "MUFU"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            cur = re.sub(r"\(.*", "", name).replace("void ", "").replace("dmvs::", "")
            per.setdefault(cur, collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and cur:
            op = m.group(1)
            per[cur]["_total"] += 1
            for w in WATCH:
                if op.startswith(w):
                    per[cur][w] += 1
    fam = collections.OrderedDict()
    for k, c in per.items():
        f = re.sub(r"<.*", "", k)
        fam.setdefault(f, [0, collections.Counter()])
        fam[f][0] += 1
        fam[f][1].update(c)
    cols = [w for w in WATCH if any(c[w] for _, c in fam.values())]
    print("%-34s %5s %8s " % ("kernel family", "inst.", "SASS") + " ".join("%8s" % c for c in cols))
    tot = collections.Counter()
    for f, (n, c) in fam.items():
        print("%-34s %5d %8d " % (f[:34], n, c["_total"]) + " ".join("%8d" % c[w] for w in cols))
        tot.update(c)
    print("%-34s %5s %8d " % ("TOTAL", "", tot["_total"]) + " ".join("%8d" % tot[w] for w in cols))
    assert tot["HMMA"] == 0 and tot["HGMMA"] == 0, "legacy tensor instructions present"


if __name__ == "__main__":
    sys.exit(main())
