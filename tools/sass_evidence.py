"""Count the Blackwell-specific SASS mnemonics per kernel of libplk.so (cuobjdump -sass): UTC*MMA =
tcgen05.mma (UTCHMMA.2CTA = cta_group::2), LDTM / STTM = tcgen05.ld / .st, UTMALDG / UTMASTG = TMA loads / stores, HMMA = legacy
mma.sync (must be absent).  usage: python tools/sass_evidence.py > profiles/r2_sass_evidence.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "multimodal_plankton_recognition_b200", "libplk.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
want = ["UTCHMMA", "UTCHMMA.2CTA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "MUFU.EX2", "HMMA",
        "ACQBULK", "UCGABAR", "RED", "ATOM"]
kern, counts = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        counts[kern] = collections.Counter()
        continue
    if kern is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        counts[kern]["_total"] += 1
        for w in want:
            if op.startswith(w):
                counts[kern][w] += 1
print(f"# cuobjdump -sass {os.path.relpath(so, ROOT)} -- static instruction counts per kernel (sm_100a)")
print(f"{'kernel':70s} " + " ".join(f"{w.replace('UTCHMMA.2CTA', 'MMA.2CTA'):>8s}" for w in want) + "    total")
for k, c in counts.items():
    if not any(c[w] for w in ("UTCHMMA", "LDTM", "UTMALDG", "UTMASTG")) and "simt" not in k and "vec" not in k:
        continue
    print(f"{k[:70]:70s} " + " ".join(f"{c[w]:8d}" for w in want) + f" {c['_total']:8d}")
tc = [k for k, c in counts.items() if c["UTCHMMA"]]
print(f"\n{len(tc)} kernels issue tcgen05.mma (UTCHMMA), {sum(1 for c in counts.values() if c['UTCHMMA.2CTA'])} of them as "
      f"cta_group::2 pairs (UTCHMMA.2CTA); kernels with legacy HMMA: "
      f"{sum(1 for c in counts.values() if c['HMMA'])}")
