"""Per-kernel SASS mnemonic counts of repmode_b200/librepmode_b200.so (cuobjdump -sass): the evidence that the hot kernels are
tcgen05 / TMA code (UTCHMMA = tcgen05.mma, .2CTA = cta_group::2, LDTM / STTM = tcgen05.ld / st, UTMALDG = cp.async.bulk.tensor,
UBLKCP = cp.async.bulk, UTCBAR = tcgen05.commit, SYNCS = mbarrier).   python tools/sass_summary.py > profiles/r2_sass_summary.csv"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
SO = os.path.join(ROOT, "repmode_b200", "librepmode_b200.so")
KEYS = ["UTCHMMA", "UTCHMMA.2CTA", "LDTM", "STTM", "UTMALDG", "UBLKCP", "UTCBAR", "UTCBAR.2CTA", "SYNCS", "HMMA", "FFMA", "LDG",
        "STG", "LD", "ST", "LDS", "STS", "BAR"]


def main():
    out = subprocess.run(["cuobjdump", "-sass", SO], capture_output=True, text=True, check=True).stdout
    name, counts, total = None, collections.OrderedDict(), {}
    for line in out.splitlines():
        m = re.match(r"\s*Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
            name = name.replace("void ", "").replace("mode::", "")
            counts[name] = collections.Counter()
            total[name] = 0
            continue
        m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
        if m and name:
            op = m.group(1)
            total[name] += 1
            base = op.split(".")[0]
            counts[name][base] += 1
            if ".2CTA" in op:
                counts[name][base + ".2CTA"] += 1
    print("# cuobjdump -sass repmode_b200/librepmode_b200.so (sm_100a), instruction counts per kernel; tools/sass_summary.py")
    print("kernel,instructions," + ",".join(KEYS))
    for k, c in counts.items():
        print(f"{k},{total[k]}," + ",".join(str(c.get(x, 0)) for x in KEYS))


if __name__ == "__main__":
    sys.exit(main())
