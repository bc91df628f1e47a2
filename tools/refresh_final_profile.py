"""Regenerate sections 1-3 of profiles/r1_final.md (launch list, per-kernel tensor/DRAM metrics, full captures) and
profiles/kernel_metrics_r1.json from a tools/gpu_profile.sh output directory. usage: python tools/refresh_final_profile.py gpurun_out/prof3"""
import os
import shutil
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
P = sys.argv[1]
run = lambda *a: subprocess.run([sys.executable] + list(a), capture_output=True, text=True, cwd=ROOT).stdout
md = os.path.join(ROOT, "profiles", "r1_final.md")
s = open(md).read()
i, j = s.index("## 1. Launch list"), s.index("## 4. BASELINE configs[3]")
shutil.copy(os.path.join(P, "launches.csv"), os.path.join(ROOT, "profiles", "launches_r1_final.csv"))
shutil.copy(os.path.join(P, "tensor_metrics.csv"), os.path.join(ROOT, "profiles", "tensor_metrics_r1_final.csv"))
body = "## 1. Launch list (`--metrics gpu__time_duration.sum`; pass `%s`)\n\n" % os.path.basename(P)
body += run("tools/summarize_launches.py", os.path.join(P, "launches.csv"))
body += "\n## 2. Tensor-pipe / DRAM metrics of every tensor-core, BN-backward and view kernel launch of the step\n\n"
body += run("tools/summarize_metrics.py", os.path.join(P, "tensor_metrics.csv"), "profiles/kernel_metrics_r1.json")
body += "\n## 3. Full captures (`ncu --set full --import-source on`; the .ncu-rep files stay in gpurun_out/)\n\n"
body += ("conv forward: decoder `up1.conv1` (768 -> 256 channels at 72x128, bs 10), 0.326 TFLOP algorithmic per launch; dgrad and\n"
         "wgrad: 256 -> 256 at 72x128, 0.109 TFLOP; stacked wgrad: `up3.conv2` 64 -> 64 at 288x512, 0.109 TFLOP (all x3 executed);\n"
         "BatchNorm-backward apply pass of `up3.conv2` (HBM-bound: achieved DRAM bytes / duration).\n\n")
for f, fl in (("prof_conv_u1c1", "3.2615e11"), ("prof_dgrad_256", "1.0872e11"), ("prof_wgrad_256", "1.0872e11"),
              ("prof_wgrad_stacked", "1.0872e11"), ("prof_bn_bwd_apply", None)):
    rep = os.path.join(P, f + ".ncu-rep")
    if os.path.exists(rep):
        body += run("tools/summarize_ncu.py", rep, *([fl] if fl else []))
open(md, "w").write(s[:i] + body + s[j:])
print("refreshed", md)
