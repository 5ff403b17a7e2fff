#!/bin/bash
# True-reference comparison, for a machine that has BOTH a real `foldseek` (>= 10) and the real ProstT5
# weights (neither exists in the build container):
#   tools/compare_with_foldseek.sh <fasta> <weights_dir> [workdir]
# Runs the reference's exact CPU invocation [REF src/modules/createdb.rs:158-162] once and this repo's shim on the same
# FASTA under every combination of the three policies that cannot be settled without a real Foldseek (SURVEY.md Q2, Q3
# and the rare-residue mapping): split length 1024 / 0, </s> row in the CNN head's input 1 / 0, U/Z/O/B -> X or their
# own tokens; then compares the 3Di strings entry by entry.  The combination with zero mismatches is the one to make the
# default of the host tools (today: 1024, 1, x).
set -euo pipefail
FASTA=$1; W=$2; OUT=${3:-/tmp/p5_compare}
ROOT=$(cd "$(dirname "$0")/.." && pwd)
command -v foldseek >/dev/null || { echo "foldseek not on PATH" >&2; exit 2; }
[ -f "$W/prostt5-f16.gguf" ] || { echo "$W/prostt5-f16.gguf missing (foldseek databases ProstT5 $W tmp)" >&2; exit 2; }
mkdir -p "$OUT/ref"
foldseek createdb "$FASTA" "$OUT/ref/db" --prostt5-model "$W" --threads "$(nproc)"
best=1
for split in 1024 0; do for eos in 1 0; do for rare in x own; do
  d="$OUT/b200_s${split}_e${eos}_${rare}"; mkdir -p "$d"
  "$ROOT/unicore_b200/bin/foldseek-b200" createdb "$FASTA" "$d/db" --prostt5-model "$W" --threads "$(nproc)" --gpu 1 \
      --prostt5-split-length "$split" --prostt5-head-eos "$eos" --prostt5-rare-residues "$rare"
  if python - "$OUT/ref/db" "$d/db" "split_len=$split head_eos=$eos rare=$rare" <<'PY'
import sys
ref, mine, tag = sys.argv[1:4]
def read(db):
    names = [l.lstrip("\0") for l in open(db + "_h").read().split("\n") if l.strip("\0")]
    ss = [l.lstrip("\0") for l in open(db + "_ss").read().split("\n") if l.strip("\0")]
    return dict(zip((n.split()[0] for n in names), ss))
a, b = read(ref), read(mine)
assert a.keys() == b.keys(), "entry sets differ"
res = sum(len(v) for v in a.values())
mism = sum(sum(x != y for x, y in zip(a[k], b[k])) + abs(len(a[k]) - len(b[k])) for k in a)
long_mism = sum(sum(x != y for x, y in zip(a[k], b[k])) for k in a if len(a[k]) > 1024)
print(f"{tag}: {len(a)} entries, {res} residues, {mism} 3Di mismatches ({100.0 * mism / max(res, 1):.4f} %), {long_mism} of them in sequences > 1024 aa")
sys.exit(0 if mism == 0 else 1)
PY
  then best=0; fi
done; done; done
exit $best
