#!/usr/bin/env python
"""Opcode evidence for the tensor-core / TMA claims: disassembles libfi_b200.so (cuobjdump -sass, runs without a GPU)
and writes, per kernel, how many tcgen05 MMA (UTCHMMA, .2CTA = cta_group::2), TMA load/store/prefetch (UTMALDG /
UTMASTG / UTMAPF), TMEM load (LDTM), TMEM alloc (UTCATOMSWS), commit/barrier (UTCBAR, SYNCS), programmatic-dependent-launch
(PREEXIT = griddepcontrol.launch_dependents, ACQBULK = griddepcontrol.wait) and legacy-MMA (HMMA: must be 0)
instructions it contains.

    python tools/sass_summary.py > profiles/sass_summary.txt
"""
import collections
import re
import subprocess
import sys
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "ai-based-frame-interpolation_b200" / "libfi_b200.so"
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "UTMALDG", "UTMASTG", "UTMAPF", "LDTM", "UTCATOMSWS", "UTCBAR", "SYNCS", "PREEXIT",
        "ACQBULK", "HMMA", "FFMA", "IDP", "ATOM", "RED"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    demangle = {}
    names = re.findall(r"Function : (\S+)", sass)
    if names:
        out = subprocess.run(["cu++filt"] + names, capture_output=True, text=True).stdout.splitlines()
        demangle = dict(zip(names, out)) if len(out) == len(names) else {}
    per = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = per.setdefault(m.group(1), collections.Counter())
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            cur["_total"] += 1
            base = op.split(".")[0]
            if base in KEYS:
                cur[base] += 1
            if base == "UTCHMMA" and ".2CTA" in op:
                cur["UTCHMMA.2CTA"] += 1
    total = collections.Counter()
    print(f"# SASS opcode summary of {LIB.name} (sm_100a), {len(per)} kernels; produced by tools/sass_summary.py")
    print("# columns: " + " ".join(KEYS) + " | total instructions | kernel")
    for fn, c in per.items():
        total.update(c)
        name = demangle.get(fn, fn)
        # drop the parameter list, keep template arguments such as <(int)256, (int)2, (bool)0>
        name = name[:name.rfind(">(") + 1] if ">(" in name else re.sub(r"\(.*\)$", "", name)
        name = name.replace("(int)", "").replace("(bool)", "")
        print(" ".join(f"{c.get(k, 0):6d}" for k in KEYS) + f" | {c['_total']:7d} | {name}")
    print("# library totals")
    print(" ".join(f"{total.get(k, 0):6d}" for k in KEYS) + f" | {total['_total']:7d} | ALL")
    if total.get("HMMA", 0):
        print("# WARNING: legacy mma.sync instructions present", file=sys.stderr)


if __name__ == "__main__":
    main()
