#!/usr/bin/env python
"""Generates tests/golden/example_data_host.json from the reference's example/data (copied as a fixture to
tests/golden/example_data) with the host oracle (oracle/host_oracle.py): record counts, totals, and an md5
over the sorted .map lines.  NOTE: this is "two restatements agree" (the Python restatement of the reference's
Rust here, the C++ host in the tests), not output of the compiled Rust: there is no Rust toolchain in this image."""
import hashlib
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import host_oracle as H  # noqa: E402

SRC = os.path.join(os.path.dirname(os.path.abspath(__file__)), "example_data")
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
