"""Summarise an `ncu --set full` capture (one kernel) into the markdown committed under profiles/.
usage: python tools/summarize_ncu.py file.ncu-rep [algorithmic_flops] [algorithmic_bytes] >> profiles/x.md"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "duration"),
    ("launch__grid_size", "grid"),
    ("launch__block_size", "block"),
    ("launch__registers_per_thread", "registers / thread"),
    ("launch__shared_mem_per_block_dynamic", "dynamic smem / block"),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor pipe active (% of peak)"),
    ("sm__throughput.avg.pct_of_peak_sustained_elapsed", "SM throughput (% of peak)"),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "DRAM throughput (% of peak)"),
    ("dram__bytes_read.sum", "DRAM bytes read"),
    ("dram__bytes_write.sum", "DRAM bytes written"),
    ("lts__t_bytes.sum", "L2 bytes"),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.sum", "shared-memory wavefronts (LSU)"),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "achieved occupancy (% of peak warps)"),
    ("smsp__inst_executed.sum", "warp instructions executed"),
    ("sm__cycles_elapsed.max", "SM cycles elapsed"),
]


def main():
    path = sys.argv[1]
    flops = float(sys.argv[2]) if len(sys.argv) > 2 else None
    nbytes = float(sys.argv[3]) if len(sys.argv) > 3 else None
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units, vals = rows[0], rows[1], rows[2]
    d = {h: (v, u) for h, u, v in zip(hdr, units, vals)}
    print(f"### `{d.get('Kernel Name', ('?', ''))[0]}` — {path}\n")
    print("| metric | value |")
    print("|---|---|")
    for key, label in WANT:
        if key in d:
            print(f"| {label} (`{key}`) | {d[key][0]} {d[key][1]} |")
    try:
        dur = float(d["gpu__time_duration.sum"][0].replace(",", ""))
        unit = d["gpu__time_duration.sum"][1]
        sec = dur * {"ns": 1e-9, "us": 1e-6, "usecond": 1e-6, "ms": 1e-3, "msecond": 1e-3, "nsecond": 1e-9, "second": 1}.get(unit, 1e-9)
        if flops:
            print(f"| algorithmic TFLOP/s (under ncu, cold-cache replay) | {flops / sec / 1e12:.1f} |")
        if nbytes:
            print(f"| algorithmic GB/s | {nbytes / sec / 1e9:.0f} |")
    except Exception:
        pass
    print()


if __name__ == "__main__":
    main()
