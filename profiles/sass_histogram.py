"""Per-kernel histogram of the SASS opcodes that prove a Blackwell-native path (B200_PROFILING.md):
UTC*MMA = tcgen05.mma, LDTM/STTM = tcgen05.ld/st, UTMALDG/UTMASTG = TMA tensor load/store, UBLKCP = bulk copy,
UTCBAR = tcgen05.commit, SYNCS = mbarrier, HMMA = legacy mma.sync (must be absent).
usage: python profiles/sass_histogram.py [libddpm.so]   (needs cuobjdump; runs without a GPU)"""
import collections
import os
import re
import subprocess
import sys

so = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))),
                                                        "imagegenerationdiffusionmodels.jl_b200", "libddpm.so")
sass = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True, check=True).stdout
filt = subprocess.run(["c++filt"], input="\n".join(re.findall(r"Function : (\S+)", sass)), capture_output=True, text=True).stdout.split("\n")
OPS = ["UTCHMMA", "UTCQMMA", "UTCIMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "UTCBAR", "SYNCS", "HMMA", "LDGSTS", "REDUX", "SHFL"]
per = collections.OrderedDict()
cur = None
names = iter(filt)
arch = set(re.findall(r"arch = (sm_\w+)", sass))
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = re.sub(r"ddpm::|\(.*$", "", next(names))[:96]
        per.setdefault(cur, collections.Counter())
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1).split(".")[0]
        per[cur]["_total"] += 1
        for o in OPS:
            if op.startswith(o):
                per[cur][o] += 1
print(f"{so}: arch {sorted(arch)}, {len(per)} kernels")
tot = collections.Counter()
print(f"{'total':>7} " + " ".join(f"{o:>7}" for o in OPS) + "  kernel")
for k, c in per.items():
    tot.update(c)
    if any(c[o] for o in OPS[:9]):
        print(f"{c['_total']:7d} " + " ".join(f"{c[o]:7d}" for o in OPS) + f"  {k}")
print(f"{tot['_total']:7d} " + " ".join(f"{tot[o]:7d}" for o in OPS) + "  ALL KERNELS (incl. those without tensor/TMA opcodes)")
