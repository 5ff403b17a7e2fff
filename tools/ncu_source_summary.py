"""Per-opcode instruction and stall-sample totals from `ncu --page source --csv` output.
   python tools/ncu_source_summary.py <report.ncu-rep> [top]"""
import csv
import io
import subprocess
import sys
from collections import defaultdict

rep = sys.argv[1]
top = int(sys.argv[2]) if len(sys.argv) > 2 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
lines = out.splitlines()
start = next(i for i, l in enumerate(lines) if l.startswith('"Address"'))
rows = list(csv.DictReader(io.StringIO("\n".join(lines[start:]))))
inst = defaultdict(float)
samp = defaultdict(float)
tot_i = tot_s = 0.0
stall_cols = [c for c in rows[0].keys() if c.startswith("stall_") and "Not Issued" not in c]
stalls = defaultdict(float)
for r in rows:
    src = r["Source"].strip()
    op = src.split()[0] if src else "?"
    if op.startswith("@"):
        op = src.split()[1]
    op = op.rstrip(";")
    base = ".".join(op.split(".")[:2]) if op.startswith(("SYNCS", "MUFU", "LDS", "STS", "UTC")) else op.split(".")[0]
    try:
        n = float(r["Instructions Executed"] or 0)
        s = float(r["# Samples"] or 0)
    except ValueError:
        continue
    inst[base] += n
    samp[base] += s
    tot_i += n
    tot_s += s
    for c in stall_cols:
        try:
            stalls[c] += float(r[c] or 0)
        except ValueError:
            pass
print("total warp instructions %.1f M, samples %d" % (tot_i / 1e6, tot_s))
print("%-28s %12s %7s %10s %7s" % ("opcode", "inst (M)", "%", "samples", "%"))
for k in sorted(inst, key=lambda k: -inst[k])[:top]:
    print("%-28s %12.2f %6.1f%% %10d %6.1f%%" % (k, inst[k] / 1e6, 100 * inst[k] / tot_i, samp[k], 100 * samp[k] / max(tot_s, 1)))
print("stall samples:", ", ".join("%s %.1f%%" % (c[6:], 100 * v / max(tot_s, 1)) for c, v in sorted(stalls.items(), key=lambda x: -x[1])[:10]))
