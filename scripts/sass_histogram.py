"""SASS evidence for profiles/: per kernel of liblowdin_itgpu.so the count of the instructions that carry the design --
DMMA (FP64 tensor), UTMALDG (TMA loads), SYNCS (mbarrier), USETMAXREG (register rebalancing), REDG/ATOMG (reductions), LDS/STS.128,
and spills (STL/LDL).  Usage: python scripts/sass_histogram.py > profiles/<round>_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
so = os.path.join(ROOT, "openlowdin_b200", "liblowdin_itgpu.so")
out = subprocess.run(["cuobjdump", "-sass", so], capture_output=True, text=True).stdout
arch = re.search(r"arch = (sm_\w+)", out)
keys = ["DMMA", "UTMALDG", "SYNCS", "USETMAXREG", "REDG", "ATOMG", "LDS.128", "STS.128", "LDG", "STG", "STL", "LDL"]
per = collections.OrderedDict()
cur = None
for line in out.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    m = re.search(r"^\s+/\*[0-9a-f]+\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_.]+)", line)
    if m:
        op = m.group(1)
        for k in keys:
            if op == k or op.startswith(k + ".") or (k in ("LDS.128", "STS.128") and op.startswith(k)):
                per[cur][k] += 1
        per[cur]["total"] += 1
demangle = subprocess.run(["c++filt"] + list(per), capture_output=True, text=True).stdout.splitlines()
print(f"# {os.path.basename(so)}: {len(per)} kernels, arch {arch.group(1) if arch else '?'}")
print("# " + " ".join(f"{k:>10}" for k in ["total"] + keys) + "  kernel")
tot = collections.Counter()
for (name, c), dn in zip(per.items(), demangle):
    tot.update(c)
    short = re.sub(r"\(.*", "", dn).replace("void lowdin::", "").replace("lowdin::", "")
    print("  " + " ".join(f"{c[k]:>10}" for k in ["total"] + keys) + "  " + short[:150])
print("# " + " ".join(f"{tot[k]:>10}" for k in ["total"] + keys) + "  ALL")
