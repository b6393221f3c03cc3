"""Per-kernel SASS instruction counts of the tensor / TMA / async-copy paths in libdic.so (cuobjdump -sass, sm_100a).
    python scripts/sass_summary.py > profiles/r02_sass_summary.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "depth_image_captioning_pub_b200", "libdic.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
names = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True,
                       text=True).stdout.splitlines()
COLS = ["UTCHMMA", "UTMALDG", "UTMASTG", "LDTM", "UTCBAR", "HMMA", "LDSM", "LDGSTS", "SYNCS", "MUFU"]
rows, cur, k = [], None, -1
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        k += 1
        cur = [names[k] if k < len(names) else m.group(1), 0, collections.Counter()]
        rows.append(cur)
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\w+\s+)?([A-Z0-9_]+)", line)
    if m and cur is not None:
        cur[1] += 1
        cur[2][m.group(1)] += 1
print("cuobjdump -sass depth_image_captioning_pub_b200/libdic.so  (sm_100a), instruction counts per kernel that uses the "
      "tensor / TMA / async-copy paths")
print("(UTCHMMA = tcgen05.mma, UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR = tcgen05.commit, "
      "HMMA = mma.sync, LDSM = ldmatrix, LDGSTS = cp.async, SYNCS = mbarrier, MUFU = ex2 / rcp / ...)\n")
print(f"{'kernel':84s} {'instr':>6s} " + " ".join(f"{c:>7s}" for c in COLS))
tot = collections.Counter()
for name, n, cnt in sorted(rows):
    name = re.sub(r"^(void )?dic::", "", name)
    name = re.sub(r"\(.*", "", name)
    if not any(cnt[c] for c in COLS[:-1]):
        continue
    print(f"{name[:84]:84s} {n:6d} " + " ".join(f"{cnt[c]:7d}" for c in COLS))
    for c in COLS:
        tot[c] += cnt[c]
print(f"\n{'total over ' + str(len(rows)) + ' kernels':84s} {sum(r[1] for r in rows):6d} " + " ".join(f"{tot[c]:7d}" for c in COLS))
