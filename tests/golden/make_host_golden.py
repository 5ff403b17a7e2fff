#!/usr/bin/env python
"""Generates tests/golden/example_data_host.json from /root/reference/example/data with the host oracle
(oracle/host_oracle.py): per-file record counts, totals, and an md5 over the sorted .map lines.  The
example FASTA files themselves are reference content and are NOT copied; the CPU test that consumes
this fixture re-reads them only when /root/reference exists (it does not on the GPU box)."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import host_oracle as H  # noqa: E402

SRC = "/root/reference/example/data"
data, lines = H.collect(SRC)
out = {
    "files": len(H.list_inputs(SRC)),
    "map_lines": len(lines),
    "unique_records": len(data),
    "residues": sum(len(v) for v in data.values()),
    "map_sorted_md5": hashlib.md5("".join("\t".join(l) + "\n" for l in sorted(lines)).encode()).hexdigest(),
    "first_sorted_lines": ["\t".join(l) for l in sorted(lines)[:3]],
}
json.dump(out, open(os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_data_host.json"), "w"), indent=1)
print(out)
