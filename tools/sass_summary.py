"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native kernel, from cuobjdump of the built
libraries:  python tools/sass_summary.py [out.txt]
UTCHMMA = tcgen05.mma, LDTM / STTM = tcgen05.ld / st, UTMALDG / UTMASTG / UTMAREDG = TMA load / store / reduce,
UBLKCP = cp.async.bulk, HMMA = legacy mma.sync (allowed only in the debug library's A/B kernel)."""
import os
import re
import subprocess
import sys
from collections import OrderedDict

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
OPS = ["UTCHMMA", "UTCBAR", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UTMAREDG", "UBLKCP", "SYNCS", "HMMA", "MUFU.EX2", "FFMA2", "FADD2", "STG.E.ENL2.256"]


def demangle(name):
    out = subprocess.run(["c++filt", name], capture_output=True, text=True).stdout.strip()
    out = re.sub(r"\(CUtensorMap_st.*", "(...)", out)
    return re.sub(r"p5::\(anonymous namespace\)::", "", out)


def summarize(lib):
    sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
    kernels = OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            cur = demangle(m.group(1))
            kernels[cur] = {op: 0 for op in OPS}
            kernels[cur]["instructions"] = 0
            continue
        if cur and re.match(r"\s+/\*[0-9a-f]{4,}\*/", line):
            kernels[cur]["instructions"] += 1
            for op in OPS:
                if re.search(r"\b" + re.escape(op), line):
                    kernels[cur][op] += 1
    return kernels


def main():
    head = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    lines = [f"SASS summary (cuobjdump -sass, sm_100a) at commit {head}", ""]
    for lib in ("libprostt5_b200.so", "libprostt5_b200_debug.so"):
        path = os.path.join(ROOT, "unicore_b200", "lib", lib)
        ks = summarize(path)
        lines.append(f"== {lib}: {len(ks)} kernels")
        lines.append("%-78s %6s " % ("kernel", "instr") + " ".join("%8s" % o[:8] for o in OPS))
        for name, c in ks.items():
            lines.append("%-78s %6d " % (name[:78], c["instructions"]) + " ".join("%8d" % c[o] for o in OPS))
        tot = {o: sum(c[o] for c in ks.values()) for o in OPS}
        lines.append("%-78s %6s " % ("TOTAL", "") + " ".join("%8d" % tot[o] for o in OPS))
        lines.append("")
    text = "\n".join(lines)
    print(text)
    if len(sys.argv) > 1:
        open(sys.argv[1], "w").write(text + "\n")


if __name__ == "__main__":
    main()
