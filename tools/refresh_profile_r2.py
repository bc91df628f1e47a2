"""profiles/r2_final.md sections 1-3 (launch list, per-kernel tensor / DRAM metrics, full captures) and
profiles/kernel_metrics_r2.json from a tools/gpu_profile_r2.sh output directory. The JSON is stamped with the hash of the
conv kernel sources and the git HEAD it was taken on: bench.py refuses it (roofline.traffic = null) once they change.
usage: python tools/refresh_profile_r2.py gpurun_out/<dir>"""
import json
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench  # noqa: E402

P = sys.argv[1]
run = lambda *a: subprocess.run([sys.executable] + list(a), capture_output=True, text=True, cwd=ROOT).stdout
shutil.copy(os.path.join(P, "launches.csv"), os.path.join(ROOT, "profiles", "launches_r2_final.csv"))
shutil.copy(os.path.join(P, "tensor_metrics.csv"), os.path.join(ROOT, "profiles", "tensor_metrics_r2_final.csv"))
head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
body = f"# profiles/r2_final - ncu passes of one steady-state train step (pass `{os.path.basename(P)}`, HEAD {head})\n\n"
body += ("Workload: BASELINE configs[1] (TrackNet seq_len 8, bg concat, bs 10, 288x512, fwd + WBCE + bwd, fp32x3), driver\n"
         "`tools/profile_step.py`, commands in `tools/gpu_profile_r2.sh` (`--clock-control none`, TNB_GRAPHS=0, TNB_PDL=0).\n\n")
body += "## 1. Launch list (`--metrics gpu__time_duration.sum`)\n\n"
body += run("tools/summarize_launches.py", os.path.join(P, "launches.csv"))
body += "\n## 2. Tensor-pipe / DRAM metrics of every tensor-core, BatchNorm-backward and view kernel launch of the step\n\n"
js = os.path.join(ROOT, "profiles", "kernel_metrics_r2.json")
body += run("tools/summarize_metrics.py", os.path.join(P, "tensor_metrics.csv"), js)
km = json.load(open(js))
km["_meta"] = {"sources_sha": bench.sources_sha(), "head": head, "pass": os.path.basename(P),
               "sources": bench.TRAFFIC_SOURCES}
json.dump(km, open(js, "w"), indent=1)
body += "\n## 3. Full captures (`ncu --set full --import-source on`; the .ncu-rep files stay in gpurun_out/)\n\n"
caps = (("prof_fwd_64_64", "1.0872e11"), ("prof_fwd_768_256", "3.2615e11"), ("prof_fwd_192_64", "3.2615e11"),
        ("prof_dgrad_256_256", "1.0872e11"), ("prof_wgrad_pair", "1.0872e11"), ("prof_wgrad_stacked", "1.0872e11"),
        ("prof_bn_bwd_apply", None))
for f, fl in caps:
    rep = os.path.join(P, f + ".ncu-rep")
    if os.path.exists(rep):
        body += run("tools/summarize_ncu.py", rep, *([fl] if fl else []))
open(os.path.join(ROOT, "profiles", "r2_final.md"), "w").write(body)
print("wrote profiles/r2_final.md, profiles/kernel_metrics_r2.json")
