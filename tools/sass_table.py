#!/usr/bin/env python3
"""Opcode table of the built library (profiles/r02_sass_opcodes.txt): per kernel, the SASS mnemonics that prove which
hardware path it uses -- UTCHMMA / LDTM / UTCBAR (tcgen05 + TMEM), UTMALDG (TMA tensor loads), UBLKCP (TMA bulk copies),
UTMAPF (TMA L2 prefetch), SYNCS (mbarrier), LDGSTS (cp.async), FFMA / HMMA counts.  Run after `make`."""
import collections
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else "fish_speech_rs_b200/libfsb.so"
txt = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
OPS = ["UTCHMMA", "LDTM", "UTCBAR", "UTMALDG", "UTMAPF", "UBLKCP", "SYNCS", "LDGSTS", "HMMA", "FFMA", "DFMA", "MUFU", "BAR"]
cur, rows = None, collections.OrderedDict()
for ln in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", ln)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "")
        cur = re.sub(r"\(.*", "", cur).replace("void ", "").replace("fsb::", "")
        rows[cur] = collections.Counter()
        continue
    if cur:
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", ln)
        if m:
            op = m.group(1).split(".")[0]
            rows[cur]["_total"] += 1
            if op in OPS:
                rows[cur][op] += 1
print(f"{'kernel':72s} {'instr':>7s} " + " ".join(f"{o:>7s}" for o in OPS))
for k, c in rows.items():
    if c["_total"] < 40:
        continue
    print(f"{k[:72]:72s} {c['_total']:7d} " + " ".join(f"{c[o]:7d}" for o in OPS))
