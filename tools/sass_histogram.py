#!/usr/bin/env python
"""Per-kernel SASS opcode histogram of the device objects of libicc_b200.so (cuobjdump -sass): makes the claims of DESIGN.md checkable
from the repository -- FP64 tensor-core MMAs (DMMA) in the evaluation / solver kernels, tensor-memory parking (LDTM / STTM) and the TMEM
allocator (UTCATOMSWS) in eval_tmem_kernel, no library kernels, no HMMA / UTC*MMA (tcgen05.mma has no FP64 kind).
usage: python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["DMMA", "DFMA", "DMUL", "DADD", "LDTM", "STTM", "UTCATOMSWS", "REDG", "RED", "ATOMG", "LDS", "STS", "LDG", "STG", "LDL", "STL", "SHFL", "BAR", "MUFU", "HMMA", "UTCHMMA", "UTCQMMA", "UTMALDG", "UBLKCP"]
for obj in sorted(glob.glob(os.path.join(ROOT, "openimucameracalibrator_b200", "csrc", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    if "Function :" not in out:
        continue
    print(f"== {os.path.basename(obj)}")
    fn, hist = None, {}
    for line in out.splitlines():
        m = re.search(r"Function : (\S+)", line)
        if m:
            name = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
            fn = re.sub(r"\(icc::.*|\(double.*|\(int.*|\(\)$", "", name.replace("(anonymous namespace)::", "").replace("void ", "").replace("icc::", "", 1))
            hist[fn] = collections.Counter()
            continue
        m = re.match(r"\s+/\*[0-9a-f]{4,5}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m and fn:
            hist[fn][m.group(1)] += 1
    for fn, h in hist.items():
        total = sum(h.values())
        keys = " ".join(f"{k}={h[k]}" for k in KEY if h.get(k))
        print(f"  {fn[:70]:70s} {total:6d} instr  {keys}")
