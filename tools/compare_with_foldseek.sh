#!/bin/bash
# True-reference comparison, for a machine that has BOTH a real `foldseek` (>= 10) and the real ProstT5
# weights (neither exists in the build container):
#   tools/compare_with_foldseek.sh <fasta> <weights_dir> [workdir]
# Runs the reference's exact CPU invocation [REF src/modules/createdb.rs:158-162] and this repo's shim
# on the same FASTA, then compares the 3Di strings entry by entry.
set -euo pipefail
FASTA=$1; W=$2; OUT=${3:-/tmp/p5_compare}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
command -v foldseek >/dev/null || { echo "foldseek not on PATH" >&2; exit 2; }
[ -f "$W/prostt5-f16.gguf" ] || { echo "$W/prostt5-f16.gguf missing (foldseek databases ProstT5 $W tmp)" >&2; exit 2; }
mkdir -p "$OUT/ref" "$OUT/b200"
foldseek createdb "$FASTA" "$OUT/ref/db" --prostt5-model "$W" --threads "$(nproc)"
"$ROOT/unicore_b200/bin/foldseek-b200" createdb "$FASTA" "$OUT/b200/db" --prostt5-model "$W" --threads "$(nproc)" --gpu 1
python - "$OUT" <<'PY'
import sys
out = sys.argv[1]
def read(db):
    names = [l.lstrip("\0") for l in open(db + "_h").read().split("\n") if l.strip("\0")]
    ss = [l.lstrip("\0") for l in open(db + "_ss").read().split("\n") if l.strip("\0")]
    return dict(zip((n.split()[0] for n in names), ss))
a, b = read(out + "/ref/db"), read(out + "/b200/db")
assert a.keys() == b.keys(), "entry sets differ"
res = sum(len(v) for v in a.values())
mism = sum(sum(x != y for x, y in zip(a[k], b[k])) + abs(len(a[k]) - len(b[k])) for k in a)
print(f"{len(a)} entries, {res} residues, {mism} 3Di mismatches ({100.0 * mism / max(res, 1):.4f} %)")
sys.exit(0 if mism == 0 else 1)
PY
