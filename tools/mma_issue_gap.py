"""Static check of the tcgen05.mma issue streams: for every kernel of libplk.so, the number of SASS
instructions between consecutive UTCHMMA (cuobjdump -sass).  A gap of 1-3 means the descriptors stay on the
uniform datapath; ~12 means the elect / R2UR.BROADCAST / branch sequence of a divergent issuer (see
tc_common.cuh::elect_one).   usage: python tools/mma_issue_gap.py [path/to/lib.so]"""
import collections
import os
import re
import statistics
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "multimodal_plankton_recognition_b200", "libplk.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
kern, pos, idx = None, collections.OrderedDict(), 0
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        pos[kern], idx = [], 0
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m and kern:
        idx += 1
        if m.group(1).startswith("UTCHMMA"):
            pos[kern].append(idx)
print(f"{'kernel':64s} {'MMAs':>5s} {'median gap':>10s} {'gaps <= 4':>10s}")
for k, p in pos.items():
    if len(p) < 2:
        continue
    gaps = [b - a for a, b in zip(p, p[1:])]
    print(f"{k[:64]:64s} {len(p):5d} {statistics.median(gaps):10.0f} {sum(g <= 4 for g in gaps):6d}/{len(gaps)}")
