"""Dev tool: raw-page CSV of an `ncu --set full` capture (ncu -i x.ncu-rep --page raw --csv) -> per-kernel summary:
launches, mean duration, DRAM bytes read + written per launch, DRAM throughput, tensor-pipe activity.
Writes a readable table to stdout and, with --json KEY, merges {kernel class: dram bytes per STEP} into
profiles/traffic.json under KEY (e.g. "c3/bf16") -- the `traffic` field of bench.py's roofline reads that file.

    python tools/ncu_traffic.py profiles/r02_c3_bf16_raw.csv --json c3/bf16 --launches-per-step tc_gemm_fwd=4,...
"""
import argparse
import csv
import json
import os
import re
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
UNIT = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "Tbyte": 1e12,
        "ns": 1e-6, "us": 1e-3, "ms": 1.0, "s": 1e3}
# kernel name pattern -> the class name bench.py uses for its per-kernel work table
CLASSES = [("k_adam_rows", "adam_rows"), ("TcEpiDpStore", "gemm_bwd_dp"), ("TcEpiStore", "gemm_fwd"),
           ("TcEpiRowDot", "gemm_rowdot"), ("TcEpiAdam", "gemm_bwd_adam"), ("k_softmax_rows", "softmax_rows"),
           ("k_loss_reduce", "loss_reduce"), ("k_dy_assemble", "dy_assemble"), ("k_scale_rows_bf16", "scale_rows"),
           ("k_row_norm", "row_norm"), ("k_rowdot_finalize", "rowdot_finalize"), ("k_col_finalize", "col_finalize"),
           ("k_loss_scalars", "loss_scalars"), ("k_spatial_colstats", "spatial_colstats")]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("csv")
    ap.add_argument("--json", default=None)
    ap.add_argument("--per-step", default="", help="class=launches per iteration, comma separated (default 1)")
    a = ap.parse_args()
    rows = list(csv.reader(open(a.csv)))
    hdr, units, data = rows[0], rows[1], rows[2:]
    col = {h: i for i, h in enumerate(hdr)}

    def val(r, name):
        i = col.get(name)
        if i is None or r[i] in ("", "n/a"):
            return float("nan")
        return float(r[i].replace(",", "")) * UNIT.get(units[i], 1.0)

    per_step = dict(x.split("=") for x in a.per_step.split(",") if x)
    agg = {}
    for r in data:
        name = r[col["Kernel Name"]]
        cls = next((c for pat, c in CLASSES if pat in name), re.sub(r"\(.*", "", name)[:40])
        d = agg.setdefault(cls, dict(n=0, ms=0.0, rd=0.0, wr=0.0, tensor=0.0, name=name))
        d["n"] += 1
        d["ms"] += val(r, "gpu__time_duration.sum")
        d["rd"] += val(r, "dram__bytes_read.sum")
        d["wr"] += val(r, "dram__bytes_write.sum")
        t = val(r, "sm__mem_tensor_cycles_active.avg.pct_of_peak_sustained_active")       # tcgen05: tensor-memory pipe activity
        d["tensor"] += 0.0 if t != t else t
    out = {}
    print(f"{'kernel':18s} {'launches':>8s} {'ms/launch':>10s} {'DRAM rd GB':>11s} {'DRAM wr GB':>11s} {'TB/s':>7s} {'tensor%':>8s}")
    for cls, d in sorted(agg.items(), key=lambda kv: -kv[1]["ms"]):
        n = d["n"]
        ms, rd, wr = d["ms"] / n, d["rd"] / n, d["wr"] / n
        print(f"{cls:18s} {n:8d} {ms:10.4f} {rd / 1e9:11.4f} {wr / 1e9:11.4f} {(rd + wr) / (ms * 1e-3) / 1e12:7.3f} {d['tensor'] / n:8.1f}")
        out[cls] = (rd + wr) * int(per_step.get(cls, 1))
    if a.json:
        p = os.path.join(ROOT, "profiles", "traffic.json")
        j = json.load(open(p)) if os.path.exists(p) else {}
        j[a.json] = dict(out, _source=os.path.relpath(os.path.abspath(a.csv), ROOT),
                         _what="dram__bytes_read.sum + dram__bytes_write.sum per iteration (per launch x launches per iteration)")
        json.dump(j, open(p, "w"), indent=1, sort_keys=True)
        print("merged into", p)


if __name__ == "__main__":
    sys.exit(main())
