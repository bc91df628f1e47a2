"""tensor_metrics.csv (tools/gpu_profile.sh) -> markdown table + JSON with per-launch averages per kernel.
usage: python tools/summarize_metrics.py tensor_metrics.csv out.json >> profiles/x.md"""
import collections
import csv
import json
import sys


def main(path, out_json):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    byid = collections.OrderedDict()
    for r in rows:
        d = byid.setdefault(r["ID"], {"name": r["Kernel Name"], "grid": r["Grid Size"]})
        d[r["Metric Name"]] = (float(r["Metric Value"].replace(",", "")), r["Metric Unit"])
    tosec = lambda v, u: v * {"ns": 1e-9, "us": 1e-6, "ms": 1e-3, "s": 1}.get(u, 1e-9)
    tob = lambda v, u: v * {"byte": 1, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}.get(u, 1)
    agg = collections.OrderedDict()
    for d in byid.values():
        n = d["name"].split("(")[0].replace("void ", "").replace("tnb::", "")
        t = tosec(*d["gpu__time_duration.sum"])
        rd, wr = tob(*d["dram__bytes_read.sum"]), tob(*d["dram__bytes_write.sum"])
        l2 = tob(*d["lts__t_bytes.sum"]) if "lts__t_bytes.sum" in d else 0.0
        tp = d["sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active"][0]
        a = agg.setdefault(n, [0, 0.0, 0.0, 0.0, 0.0, 0.0])
        a[0] += 1; a[1] += t; a[2] += rd; a[3] += wr; a[4] += tp * t; a[5] += l2
    print("| kernel | launches / step | ms / step | DRAM read GB | DRAM write GB | DRAM TB/s | L2 GB | tensor-pipe active (time-weighted) |")
    print("|---|---:|---:|---:|---:|---:|---:|---:|")
    js = {}
    for n, a in agg.items():
        print(f"| `{n}` | {a[0]} | {a[1] * 1e3:.3f} | {a[2] / 1e9:.2f} | {a[3] / 1e9:.2f} | {(a[2] + a[3]) / a[1] / 1e12:.2f} | "
              f"{a[5] / 1e9:.1f} | {a[4] / a[1]:.1f}% |")
        js[n] = {"launches": a[0], "ms": a[1] * 1e3, "dram_bytes_per_launch": (a[2] + a[3]) / a[0],
                 "tensor_pipe_pct": a[4] / a[1]}
    json.dump(js, open(out_json, "w"), indent=1)


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
