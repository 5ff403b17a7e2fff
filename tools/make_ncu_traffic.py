"""profiles/ncu_traffic.json from one `ncu --set full` capture of the four projection GEMMs of a layer (config 2):

    ncu --set full --clock-control none -k regex:gemm_tcgen05 -s 9 -c 4 -f -o gpurun_out/prof_gemm python tools/profile_target.py
    python tools/make_ncu_traffic.py gpurun_out/prof_gemm.ncu-rep [commit]

bench.py prints `gemm_dram_bytes_per_launch` (mean over the four shapes, read + write) as roofline.traffic together with
the commit and date recorded here, so the figure is tied to the kernels it was measured on."""
import csv
import datetime
import json
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
M, D, INNER, FF = 90112, 1024, 4096, 16384
SHAPES = {  # name: (N, K, output bytes per element, algorithmic bytes = A read + B read + C written (+ fp32 C read-modify at L2))
    "qkv": (3 * INNER, D, 2), "o": (D, INNER, 4), "ffn_in": (FF, D, 2), "ffn_out": (D, FF, 4)}


def main():
    rep = sys.argv[1]
    commit = sys.argv[2] if len(sys.argv) > 2 else None
    if not commit:
        try:
            commit = open(os.path.join(ROOT, "unicore_b200", "lib", "BUILD_COMMIT")).read().strip()
        except OSError:
            commit = subprocess.run(["git", "-C", ROOT, "rev-parse", "--short", "HEAD"], capture_output=True, text=True).stdout.strip()
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr = rows[0]
    col = {h: i for i, h in enumerate(hdr)}
    launches = []
    for r in rows[2:]:
        name = r[col["Kernel Name"]]
        def val(k):
            return float(r[col[k]].replace(",", ""))
        unit = {h: rows[1][i] for i, h in enumerate(hdr)}
        def to_bytes(k):
            u = unit[k].lower()
            return val(k) * {"byte": 1, "kbyte": 1e3, "mbyte": 1e6, "gbyte": 1e9}[u]
        dur = val("gpu__time_duration.sum") * {"ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}[unit["gpu__time_duration.sum"].lower()]
        launches.append({"name": name, "read": to_bytes("dram__bytes_read.sum"), "write": to_bytes("dram__bytes_write.sum"), "ms": dur,
                         "tensor_pipe_pct": val("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active")})
    assert len(launches) == 4, [l["name"] for l in launches]
    # identify the shapes: epilogue template argument, and of the two residual GEMMs the shorter one is O
    by = {}
    def epi(l):  # epilogue template argument: "...(p5::Epi)2>" or "<2, 256, 6, 2>" depending on the ncu version
        m = re.search(r"\(p5::Epi\)(\d+)", l["name"]) or re.search(r"<\s*\d+,\s*\d+,\s*\d+,\s*(\d+)\s*>", l["name"])
        return int(m.group(1)) if m else -1
    adds = sorted([l for l in launches if epi(l) in (2, 5)], key=lambda l: l["ms"])
    by["o"], by["ffn_out"] = adds[0], adds[1]
    by["qkv"] = next(l for l in launches if epi(l) == 0)
    by["ffn_in"] = next(l for l in launches if epi(l) == 1)
    per = {}
    for k, (N, K, osz) in SHAPES.items():
        alg = M * K * 2 + N * K * 2 + M * N * osz * (2 if osz == 4 else 1)  # the fp32 residual is read and written
        l = by[k]
        per[f"{k} [{M}x{K}]x[{N}x{K}]^T"] = {"read": l["read"], "write": l["write"], "ms": round(l["ms"], 4), "tensor_pipe_pct": round(l["tensor_pipe_pct"], 1),
                                               "algorithmic_bytes": alg, "traffic_over_algorithmic": round((l["read"] + l["write"]) / alg, 3),
                                               "tflops": round(2.0 * M * N * K / (l["ms"] * 1e-3) / 1e12, 1)}
    mean = sum(l["read"] + l["write"] for l in by.values()) / 4
    doc = {"source": f"{os.path.basename(rep)}: ncu --set full --clock-control none, one launch of each projection shape of a layer, config 2 (M = {M} tokens)",
           "commit": commit, "when": datetime.datetime.now(datetime.timezone.utc).strftime("%Y-%m-%dT%H:%M:%SZ"),
           "gemm_dram_bytes_per_launch": mean,
           "gemm_algorithmic_bytes_per_launch": sum(v["algorithmic_bytes"] for v in per.values()) / 4,
           "per_shape": per}
    json.dump(doc, open(os.path.join(ROOT, "profiles", "ncu_traffic.json"), "w"), indent=1)
    print(json.dumps(doc, indent=1))


if __name__ == "__main__":
    main()
