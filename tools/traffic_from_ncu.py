"""Summarise an ncu CSV (--metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum
--csv) of `tools/c4_steps.py <workload> N` into per-stage DRAM bytes per scan, and merge it into
profiles/r2_traffic.json (read by bench.py as roofline.traffic).

    python tools/traffic_from_ncu.py <workload> <ncu.csv> [n_scans_in_csv]"""
import csv
import json
import re
import sys
from collections import defaultdict
from pathlib import Path

STAGE_OF = [
    ("preprocess_bin_kernel", "preprocess_bin"),
    ("commit_move_clear_kernel", "commit_move_clear"),
    ("scatter_records_kernel", "sort_by_cell"),
    ("tile_estimate", "segreduce_estimate"),
    ("voxel_", "voxel_raycast"), ("DeviceRadixSort", "voxel_raycast"), ("ray_bin", "voxel_raycast"),
    ("raycast_", "voxel_raycast"),
]


def main():
    wl, path = sys.argv[1], sys.argv[2]
    rows = [r for r in csv.reader(open(path)) if len(r) > 10 and r[0].isdigit()]
    per = defaultdict(lambda: defaultdict(float))   # stage -> metric -> sum
    scans = 0
    for r in rows:
        name, metric, val = r[4], r[-3], float(r[-1].replace(",", ""))
        unit = r[-2]
        if "preprocess_bin_kernel" in name and metric == "gpu__time_duration.sum":
            scans += 1
        stage = next((s for pat, s in STAGE_OF if pat in name), None)
        if stage is None:
            continue
        if metric.startswith("dram__bytes"):
            mult = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(unit, 1.0)
            per[stage]["dram_bytes"] += val * mult
        elif metric == "gpu__time_duration.sum":
            mult = {"ns": 1e-3, "us": 1.0, "ms": 1e3}.get(unit, 1e-3)
            per[stage]["us"] += val * mult
    scans = int(sys.argv[3]) if len(sys.argv) > 3 else max(scans, 1)
    out = {s: round(v["dram_bytes"] / scans) for s, v in per.items()}
    times = {s: round(v["us"] / scans, 2) for s, v in per.items()}
    p = Path(__file__).resolve().parent.parent / "profiles" / "r2_traffic.json"
    d = json.loads(p.read_text()) if p.exists() else {}
    d[wl] = out
    d.setdefault("_kernel_us_per_scan_under_ncu", {})[wl] = times
    d["_how"] = ("dram__bytes_read.sum + dram__bytes_write.sum per scan, summed over the kernels of each pipeline "
                 "stage, from `ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum "
                 "--clock-control none python tools/c4_steps.py <workload> N` (steady-state scans)")
    p.write_text(json.dumps(d, indent=1) + "\n")
    print(wl, scans, "scans:", out, times)


if __name__ == "__main__":
    main()
