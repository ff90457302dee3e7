"""Per-kernel SASS opcode histogram of the shipped library (cuobjdump -sass): which hardware paths each kernel
really uses (TMA: UTMALDG / UTMASTG / UTMAPF / UBLKCP, mbarrier: SYNCS, cluster barriers: UCGABAR_*, FP64 math:
DFMA / DMUL / DADD, FP64 tensor cores: DMMA).  Usage: python tools/sass_histogram.py [lib.so] > profiles/...md"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "dqc_simulator_b200", "lib", "libdqcengine.so")
txt = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
arch = sorted(set(re.findall(r"arch = (sm_\w+)", txt)))
kernels, cur = collections.OrderedDict(), None
for line in txt.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = kernels.setdefault(m.group(1), collections.Counter())
        continue
    m = re.match(r"\s*/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+(?:\.[A-Z0-9_.]+)?)", line)
    if m and cur is not None:
        cur[m.group(1).split(".")[0]] += 1
demangle = subprocess.run(["c++filt"], input="\n".join(kernels), capture_output=True, text=True).stdout.splitlines()
KEY = ["UTMALDG", "UTMASTG", "UTMAPF", "UBLKCP", "SYNCS", "UCGABAR_ARV", "UCGABAR_WAIT", "DMMA", "DFMA", "DMUL", "DADD",
       "LDS", "STS", "LDL", "STL", "BAR", "SHFL", "MUFU", "ATOM", "ATOMG", "RED"]
print(f"# SASS opcode histogram of `{os.path.relpath(lib, ROOT)}` (architectures in the fatbin: {', '.join(arch)})\n")
print("Static instruction counts per kernel (cuobjdump -sass); tensor-core / TMA / cluster opcodes first.\n")
print("| kernel | total | " + " | ".join(KEY) + " |")
print("|---|---|" + "---|" * len(KEY))
for name, ctr in zip(demangle, kernels.values()):
    short = re.sub(r"\(.*", "", name.replace("dqe::", "").replace("void ", ""))
    print(f"| `{short}` | {sum(ctr.values())} | " + " | ".join(str(ctr.get(k, 0)) for k in KEY) + " |")
print("\nTop 12 opcodes per kernel:\n")
for name, ctr in zip(demangle, kernels.values()):
    short = re.sub(r"\(.*", "", name.replace("dqe::", "").replace("void ", ""))
    print(f"* `{short}`: " + ", ".join(f"{k} {v}" for k, v in ctr.most_common(12)))
