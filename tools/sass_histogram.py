#!/usr/bin/env python
"""SASS evidence: per-kernel counts of the Blackwell-native opcodes in libditree.so (cuobjdump -sass).

    python tools/sass_histogram.py [--out profiles/r02_sass_opcodes.md]

The PTX names never appear in SASS: tcgen05.mma -> UTC*MMA, tcgen05.ld -> LDTM, cp.async.bulk.tensor -> UTMALDG,
cp.async.bulk -> UBLKCP, cp.async -> LDGSTS; HMMA would be the legacy mma.sync path (there is none).
"""
import argparse
import collections
import os
import re
import subprocess
import sys

REPO = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(REPO, "ditreeonlineplanner_b200", "libditree.so")
WATCH = ["UTCHMMA", "UTCHMMA.2CTA", "UTCBAR", "LDTM", "UTMALDG", "UBLKCP", "LDGSTS", "SYNCS", "PREEXIT", "ACQBULK", "MUFU", "HMMA", "STL",
         "LDL"]


def demangle(names):
    out = subprocess.run(["c++filt"], input="\n".join(names), capture_output=True, text=True).stdout.splitlines()
    return dict(zip(names, out))


def histogram(lib=LIB):
    """-> {mangled kernel name: Counter(opcode family -> count, '_total' -> instructions)}"""
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_.]*)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            base = op.split(".")[0]
            if base in WATCH:
                cur[base] += 1
            if op.startswith("UTCHMMA.2CTA"):
                cur["UTCHMMA.2CTA"] += 1
    return per


def short(name):
    name = re.sub(r"\(.*$", "", name)
    return name.replace("void ", "")


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=None)
    a = ap.parse_args()
    per = histogram()
    dm = demangle(list(per))
    cols = [c for c in WATCH if any(per[k][c] for k in per)]
    lines = ["# SASS opcode histogram of libditree.so (sm_100a), per kernel", "",
             "`cuobjdump -sass ditreeonlineplanner_b200/libditree.so`, counted by `tools/sass_histogram.py`. "
             "`UTCHMMA` = tcgen05.mma.kind::f16 (`.2CTA` = cta_group::2), `LDTM` = tcgen05.ld, `UTMALDG` = "
             "cp.async.bulk.tensor (TMA tile loads), `UBLKCP` = cp.async.bulk (1-D TMA: the occupancy grid), `UTCBAR` = "
             "tcgen05.commit, `SYNCS` = mbarrier ops, `LDGSTS` = cp.async, `PREEXIT` / `ACQBULK` = griddepcontrol.launch_dependents / "
             ".wait (programmatic dependent launch), `MUFU` = SFU (sin/cos/ex2/rcp), `STL`/`LDL` = "
             "local-memory spills. No `HMMA` (mma.sync) anywhere.", "",
             "| kernel | instr | " + " | ".join(cols) + " |", "|---|---|" + "---|" * len(cols)]
    for k, c in per.items():
        lines.append(f"| `{short(dm[k])}` | {c['_total']} | " + " | ".join(str(c[x]) if c[x] else "" for x in cols) + " |")
    tot = collections.Counter()
    for c in per.values():
        tot.update(c)
    lines.append("| **total** | %d | " % tot["_total"] + " | ".join(str(tot[x]) for x in cols) + " |")
    text = "\n".join(lines) + "\n"
    if a.out:
        with open(a.out, "w") as f:
            f.write(text)
    sys.stdout.write(text)


if __name__ == "__main__":
    main()
