#!/usr/bin/env python
"""Summaries of gpurun_out artefacts: bench JSON lines and ncu launch lists."""
import collections, csv, json, sys

def bench(path):
    d = json.load(open(path))
    r = d["roofline"]
    print(f"{d['config']['workload']:18s} value {d['value']:8.0f} scans/s ({d['ms_per_step']*1e3:6.1f} us) "
          f"{d['mpoints_per_s']:7.0f} Mpts/s | e2e {d['e2e']['value']:7.0f} (sync {d['e2e'].get('sync_value',0):.0f}, h2d {d['e2e'].get('pinned_h2d_gbs',0):.1f} GB/s) | cpu {d['cpu_baseline']['value']:7.1f} "
          f"({d['cpu_baseline']['ms_per_scan']:.2f} ms) | stages us "
          f"{ {k: round(v*1e3,1) for k,v in r['stage_ms'].items()} } | dom {r['kernel']} frac {r['frac']:.4f} "
          f"pipe {r['pipeline']['frac']:.4f} | launches {d['gpu_launches']}+{d['library_launches']} | {d['last_scan']}")

def launches(path):
    lines = [l for l in open(path) if not l.startswith("==")]
    agg = collections.OrderedDict()
    for row in csv.DictReader(lines):
        try:
            v = float(row["Metric Value"].replace(",", ""))
        except Exception:
            continue
        u = row["Metric Unit"]
        v = v / 1000 if u == "ns" else v * 1000 if u == "ms" else v
        a = agg.setdefault(row["Kernel Name"][:64], [0, 0.0, 1e9])
        a[0] += 1; a[1] += v; a[2] = min(a[2], v)
    for k, (c, t, mn) in agg.items():
        print(f"{c:5d} avg {t/c:8.2f} us  min {mn:8.2f}  {k}")

if __name__ == "__main__":
    for p in sys.argv[1:]:
        (launches if p.endswith(".csv") else bench)(p)
