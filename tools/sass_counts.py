"""Per-kernel counts of the SASS mnemonics that prove a Blackwell-native path (B200_PROFILING.md): tcgen05.mma -> UTC*MMA,
tcgen05.ld/st -> LDTM/STTM, TMA bulk copies -> UBLKCP, tensor-map TMA -> UTMALDG/UTMASTG, stmatrix -> STSM, packed fp32x2.
Usage: python tools/sass_counts.py [lib.so] > profiles/r02_sass_counts.txt   (runs here: cuobjdump needs no GPU)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "index-tts-ipex_b200", "lib", "libbigvgan_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, text=True, check=True).stdout
pats = collections.OrderedDict([("UTC*MMA (tcgen05.mma)", r"\bUTC\w*MMA"), ("LDTM (tcgen05.ld)", r"\bLDTM"), ("STTM (tcgen05.st)", r"\bSTTM"),
                                ("UBLKCP (cp.async.bulk)", r"\bUBLKCP"), ("UTMALDG/STG (tensor TMA)", r"\bUTMA(LDG|STG)"),
                                ("UBLKPF (bulk L2 prefetch)", r"\bUBLKPF"), ("STSM (stmatrix)", r"\bSTSM"), ("SYNCS (mbarrier)", r"\bSYNCS"),
                                ("FFMA2/FMUL2/FADD2", r"\bF(FMA|MUL|ADD)2\b"), ("MUFU", r"\bMUFU"), ("HMMA (legacy mma.sync)", r"\bHMMA"),
                                ("USETMAXREG", r"\bUSETMAXREG"),
                                # a tcgen05.mma whose operands ptxas could not prove warp-uniform is issued through an
                                # ELECT / R2UR.BROADCAST loop (~150 cycles per MMA instead of ~10): must stay at a handful
                                ("R2UR.BROADCAST", r"\bR2UR\.BROADCAST")])
cur, counts, sizes = None, collections.OrderedDict(), {}
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip()
        cur = cur.replace("(anonymous namespace)::", "").replace("<unnamed>::", "").replace("bvg::", "")
        cur = re.sub(r"^void ", "", cur)
        cur = re.sub(r"\((?!anonymous).*$", "", cur)
        while cur in counts:
            cur += "'"
        counts[cur] = collections.Counter()
        sizes[cur] = 0
        continue
    if cur and re.match(r"\s*/\*[0-9a-f]{4,}\*/", line):
        sizes[cur] += 1
        for name, pat in pats.items():
            if re.search(pat, line):
                counts[cur][name] += 1
print(f"# SASS mnemonic counts per kernel of {os.path.relpath(lib, ROOT)} (cuobjdump -sass; instructions, not executions)")
hdr = list(pats)
print("kernel | instrs | " + " | ".join(hdr))
for k, c in counts.items():
    if sizes[k] == 0:
        continue
    print(f"{k} | {sizes[k]} | " + " | ".join(str(c[h]) for h in hdr))
