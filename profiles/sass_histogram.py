#!/usr/bin/env python
"""SASS instruction histogram per kernel of libamb200.so (cuobjdump -sass): the mnemonics that prove
the Blackwell-native paths (UTC*MMA = tcgen05.mma, LDTM = tcgen05.ld, UBLKCP = cp.async.bulk,
UTCBAR = tcgen05.commit, LDGSTS = cp.async, DFMA = FP64 pipe).
usage: python profiles/sass_histogram.py [lib] > profiles/r02_sass_histogram.txt"""
import collections
import re
import subprocess
import sys
from pathlib import Path

lib = sys.argv[1] if len(sys.argv) > 1 else str(Path(__file__).resolve().parents[1] / "audio_metrics_b200" / "libamb200.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True, check=True).stdout
KEY = ("UTCHMMA", "UTCIMMA", "UTCQMMA", "UTCBAR", "UTCATOM", "LDTM", "STTM", "UBLKCP", "UTMALDG", "LDGSTS", "SYNCS",
       "DFMA", "DMUL", "DADD", "HMMA", "IMMA", "DMMA", "FFMA", "MUFU", "ATOM", "RED", "LDS", "STS", "LDG", "STG",
       "BAR", "SHFL", "VOTE", "ELECT", "UCGABAR", "ACQBULK")
per = collections.OrderedDict()
name = None
for ln in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        name = re.sub(r"\(.*", "", name).replace("void ", "")
        per[name] = collections.Counter()
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)((?:\.[A-Z0-9_]+)*)", ln)
    if m and name:
        op, mods = m.group(1), m.group(2)
        per[name]["_total"] += 1
        if op in KEY:
            per[name][op + (".2CTA" if ".2CTA" in mods else "")] += 1
print(f"SASS instruction histogram of {Path(lib).name} (sm_100a); only the mnemonics that identify a pipe are listed\n")
for name, c in per.items():
    total = c.pop("_total", 0)
    if not total:
        continue
    items = "  ".join(f"{k} {v}" for k, v in sorted(c.items(), key=lambda kv: -kv[1]))
    print(f"{name[:96]}\n    {total} instructions:  {items}\n")
