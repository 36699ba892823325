"""Per-kernel SASS instruction counts of libpnmol_b200.so (cuobjdump -sass): profiles/rNN_sass_summary.txt.

    python tools/sass_summary.py profiles/r02_sass_summary.txt
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "pnmol-experiments_b200", "pnmol_b200", "libpnmol_b200.so")
KEYS = ("DMMA", "DFMA", "DADD", "DMUL", "UTMALDG", "UTMASTG", "UBLKCP", "LDGSTS", "SYNCS", "LDG", "STG", "LDS", "STS", "LDL", "STL",
        "SHFL", "BAR", "MUFU")


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_summary.txt")
    sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True, check=True).stdout
    head = subprocess.run(["git", "rev-parse", "--short", "HEAD"], capture_output=True, text=True, cwd=ROOT).stdout.strip()
    counts, total, name = collections.OrderedDict(), {}, None
    for line in sass.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1).split(".")[0]
            total[name] += 1
            if op in KEYS:
                counts[name][op] += 1
    with open(out, "w") as f:
        f.write(f"# cuobjdump -sass {os.path.relpath(LIB, ROOT)} (sm_100a), git head {head}: instruction counts per kernel\n")
        f.write("# UTMALDG/UTMASTG = cp.async.bulk.tensor (TMA tile), UBLKCP = cp.async.bulk (TMA 1-D), LDGSTS = cp.async, SYNCS = mbarrier\n")
        f.write("kernel,total," + ",".join(KEYS) + "\n")
        for k, c in counts.items():
            f.write(f"{k},{total[k]}," + ",".join(str(c[x]) for x in KEYS) + "\n")
    print(open(out).read())


if __name__ == "__main__":
    main()
