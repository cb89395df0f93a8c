#!/usr/bin/env python3
"""Per-kernel SASS instruction census of gauss-process_b200/lib/libb200id.so (cuobjdump -sass), so that the
Blackwell-native claims -- and the documented FP64 exception to tcgen05 -- can be checked without rebuilding.

    python profiles/sass_summary.py > profiles/rNN_sass_summary.txt

Columns: DFMA/DADD/DMUL (FP64 CUDA-core pipe), DMMA (FP64 tensor pipe, mma.sync.m8n8k4.f64), FFMA2 (packed FP32),
LDGSTS (cp.async), UBLKCP (cp.async.bulk = TMA 1-D bulk copy), UTMALDG/UTMASTG (TMA tensor copies), SYNCS (mbarrier),
UTC*MMA / LDTM / STTM (tcgen05 + TMEM; FP64 has no tcgen05 kind)."""
import re
import subprocess
import sys
from collections import Counter, OrderedDict
from pathlib import Path

ROOT = Path(__file__).resolve().parent.parent
LIB = ROOT / "gauss-process_b200" / "lib" / "libb200id.so"
COLS = ["DFMA", "DADD", "DMUL", "DMMA", "FFMA2", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "UTCMMA", "LDTM", "STTM", "total"]


def main():
    txt = subprocess.run(["cuobjdump", "-sass", str(LIB)], capture_output=True, text=True, check=True).stdout
    per = OrderedDict()
    cur = None
    for line in txt.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            name = re.sub(r"\(.*", "", name).replace("b200id::", "").replace("void ", "")
            cur = per.setdefault(name, Counter())
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and cur is not None:
            op = m.group(1)
            base = op.split(".")[0]
            cur["total"] += 1
            if base in ("DFMA", "DADD", "DMUL", "DMMA", "FFMA2", "LDGSTS", "UBLKCP", "UTMALDG", "UTMASTG", "SYNCS", "LDTM", "STTM"):
                cur[base] += 1
            if base.startswith("UTC") and "MMA" in base:
                cur["UTCMMA"] += 1
    w = max(len(k) for k in per) + 2
    print(f"# cuobjdump -sass {LIB.relative_to(ROOT)}  (sm_100a); one row per kernel, instruction counts in the SASS text")
    print("kernel".ljust(w) + "".join(c.rjust(9) for c in COLS))
    tot = Counter()
    for k, c in per.items():
        print(k.ljust(w) + "".join(str(c.get(col, 0)).rjust(9) for col in COLS))
        tot.update(c)
    print("TOTAL".ljust(w) + "".join(str(tot.get(col, 0)).rjust(9) for col in COLS))


if __name__ == "__main__":
    sys.exit(main())
