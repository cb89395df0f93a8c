#!/usr/bin/env python3
"""Hot spots of an ncu source-page export: python profiles/ncu_source_hot.py prof.ncu-rep [top]
Lists the SASS instructions with the most stall samples, their dominant stall reasons and shared-memory
wavefront excess (bank conflicts), plus totals per opcode."""
import csv, io, subprocess, sys, collections
rep = sys.argv[1]; top = int(sys.argv[2]) if len(sys.argv) > 2 else 40
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.reader(io.StringIO("\n".join(lines[start:]))))
hdr = rows[0]; data = [r for r in rows[1:] if len(r) == len(hdr)]
ix = {h: i for i, h in enumerate(hdr)}
stall_cols = [h for h in hdr if h.startswith("stall_") and "Not Issued" not in h]
def f(r, h):
    try: return float(r[ix[h]])
    except Exception: return 0.0
tot = sum(f(r, "# Samples") for r in data)
print(f"total samples {tot:.0f}, instructions {len(data)}")
agg = collections.Counter(); aggs = collections.Counter(); exc = collections.Counter()
for r in data:
    op = r[ix["Source"]].split()[0] if r[ix["Source"]].split() else "?"
    if op.startswith("@"): op = r[ix["Source"]].split()[1]
    op = op.split(".")[0]
    agg[op] += f(r, "Instructions Executed"); aggs[op] += f(r, "# Samples"); exc[op] += f(r, "L1 Wavefronts Shared Excessive")
print("opcode: executed warp-instr, samples, excess smem wavefronts")
for op, c in agg.most_common(22):
    print(f"  {op:10s} {c:14.0f} {aggs[op]:9.0f} {exc[op]:12.0f}")
sc = collections.Counter()
for r in data:
    for h in stall_cols: sc[h] += f(r, h)
print("stall totals:", ", ".join(f"{h[6:]}={v:.0f}" for h, v in sc.most_common(8)))
print(f"top {top} instructions by samples:")
for r in sorted(data, key=lambda r: -f(r, "# Samples"))[:top]:
    st = sorted(((f(r, h), h[6:]) for h in stall_cols), reverse=True)[:2]
    print(f"  {r[ix['Address']][-5:]} {f(r,'# Samples'):7.0f} exec={f(r,'Instructions Executed'):10.0f} xs={f(r,'L1 Wavefronts Shared Excessive'):9.0f} {st[0][1]}:{st[0][0]:.0f} {st[1][1]}:{st[1][0]:.0f} | {r[ix['Source']][:70]}")
