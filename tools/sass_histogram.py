"""Opcode histogram of every object of libmatcouply_b200.so (`cuobjdump -sass`), the evidence behind "TMA + mbarrier +
fp64 tensor pipe are real": UTMALDG / UBLKCP = TMA tensor / bulk copies, SYNCS = mbarrier operations, DMMA = fp64
tensor-core MMA, LDGSTS = cp.async.  tcgen05 / TMEM (UTC*MMA, LDTM) are absent on purpose: tcgen05.mma has no fp64 kind.
    python tools/sass_histogram.py > profiles/r2_sass_opcodes.txt"""
import collections, glob, os, re, subprocess, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
KEY = ["UTMALDG", "UBLKCP", "SYNCS", "DMMA", "DFMA", "DADD", "DMUL", "FFMA", "LDGSTS", "LDS", "STS", "LDG", "STG", "ATOMS",
       "SHFL", "BAR", "MUFU", "HMMA", "UTCHMMA", "UTCQMMA", "LDTM", "STTM"]
print("object".ljust(24) + "kernels".rjust(8) + "instr".rjust(9) + "".join(k.rjust(9) for k in KEY))
for obj in sorted(glob.glob(os.path.join(ROOT, "matcouply_b200", "csrc", "*.o"))):
    out = subprocess.run(["cuobjdump", "-sass", obj], capture_output=True, text=True).stdout
    c, n_k, n_i = collections.Counter(), 0, 0
    for line in out.splitlines():
        if "Function :" in line:
            n_k += 1
        m = re.search(r"/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
        if m:
            n_i += 1
            c[m.group(1)] += 1
    if n_k:
        print(os.path.basename(obj).ljust(24) + str(n_k).rjust(8) + str(n_i).rjust(9) + "".join(str(c.get(k, 0)).rjust(9) for k in KEY))
