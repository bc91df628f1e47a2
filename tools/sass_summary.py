"""Per-kernel counts of the SASS mnemonics that prove which hardware paths the shipped library uses (tcgen05 MMA = UTCHMMA,
TMEM loads = LDTM, bulk TMA = UBLKCP, tensor TMA = UTMALDG, cp.async = LDGSTS, programmatic dependent launch =
ACQBULK (griddepcontrol.wait) / PREEXIT (launch_dependents)). Runs in the build container (no GPU).
usage: python tools/sass_summary.py > profiles/sass_summary_r2.txt"""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "tracknetv3_b200", "libtracknet_b200.so")
OPS = ["UTCHMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "UTMASTG", "LDGSTS", "SYNCS", "ACQBULK", "PREEXIT", "ATOMG", "RED", "HMMA"]


def main():
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
    demangle = lambda n: subprocess.run(["cu++filt", n], capture_output=True, text=True).stdout.strip() or n
    counts = collections.OrderedDict()
    cur = None
    for line in sass.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            cur = counts.setdefault(m.group(1), collections.Counter())
            continue
        if cur is None:
            continue
        m = re.match(r"\s*/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            op = m.group(1)
            for o in OPS:
                if op.startswith(o):
                    cur[o] += 1
            cur["_all"] += 1
    print(f"SASS summary of {os.path.relpath(LIB, ROOT)} ({os.path.getsize(LIB)} bytes, {len(counts)} kernels)\n")
    print(f"{'kernel':88s} {'instr':>7s} " + " ".join(f"{o:>7s}" for o in OPS))
    tot = collections.Counter()
    for name, c in counts.items():
        d = demangle(name).rsplit("(", 1)[0].replace("void ", "").replace("(tnb::SrcMode)", "").replace("tnb::", "").replace("(anonymous namespace)::", "")
        print(f"{d[:88]:88s} {c['_all']:7d} " + " ".join(f"{c[o]:7d}" for o in OPS))
        tot.update(c)
    print(f"{'TOTAL':88s} {tot['_all']:7d} " + " ".join(f"{tot[o]:7d}" for o in OPS))


if __name__ == "__main__":
    main()
