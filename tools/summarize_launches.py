"""Turn an ncu launch list (--metrics gpu__time_duration.sum --csv) of tools/profile_step.py into the
markdown summary committed under profiles/. usage: python tools/summarize_launches.py launches.csv > profiles/x.md"""
import collections
import csv
import re
import sys


def main(path):
    lines = [l for l in open(path) if l.startswith('"')]
    rows = list(csv.DictReader(lines))
    names = [r["Kernel Name"] for r in rows]
    starts = [i for i, n in enumerate(names) if "pack_input" in n]
    step = rows[starts[-1]:]  # the last (steady-state) train step
    agg = collections.OrderedDict()
    total = 0.0
    for r in step:
        n = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("tnb::", "")[:48]
        d = float(r["Metric Value"]) / 1e6
        e = agg.setdefault(n, [0, 0.0])
        e[0] += 1
        e[1] += d
        total += d
    print(f"ncu launch list of one steady-state train step (bs 10, 288x512, fp32x3): {len(step)} launches, "
          f"{total:.3f} ms summed device time (serialised, cold-cache replays: compare SHARES, not absolutes)\n")
    print("| kernel | launches | ms | share |")
    print("|---|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.3f} | {100 * v[1] / total:.1f}% |")
    print("\nPer-launch list of the tensor-core and BN-backward kernels (step order):\n")
    print("| ms | grid | kernel |")
    print("|---:|---|---|")
    for r in step:
        n = r["Kernel Name"]
        if any(t in n for t in ("conv3x3", "wgrad", "bn_bwd_kernel")):
            print(f"| {float(r['Metric Value']) / 1e6:.3f} | {r['Grid Size']} | {re.sub(r'[(].*', '', n).replace('void tnb::', '')} |")


if __name__ == "__main__":
    main(sys.argv[1])
