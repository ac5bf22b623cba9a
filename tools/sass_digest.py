"""SASS digest of the shipped library (run anywhere cuobjdump exists, no GPU): per kernel, the instruction count and the counts of the
mnemonics that prove which hardware paths the code uses -- tcgen05 tensor cores (UTCHMMA / UTCQMMA, TMEM loads LDTM, UTCBAR commits),
TMA bulk copies (UBLKCP / UTMALDG), mbarriers (SYNCS), packed FP32 (FFMA2), shared-memory atomics (ATOMS), local-memory spills
(LDL / STL).

    python tools/sass_digest.py [path/to/libsiftb200.so] > profiles/r02_sass_digest.md
"""
import collections
import os
import re
import subprocess
import sys

lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "sift-gpu_b200", "libsiftb200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
elf = subprocess.run(["cuobjdump", "-lelf", lib], capture_output=True, text=True).stdout
KEYS = ["UTCHMMA", "UTCQMMA", "LDTM", "UTCBAR", "UBLKCP", "UTMALDG", "SYNCS", "FFMA2", "FFMA", "MUFU", "LDS", "STS", "ATOMS", "LDL", "STL", "LDG", "STG", "SHFL"]
per = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        per[cur] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\w+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and cur:
        per[cur]["_n"] += 1
        per[cur][m.group(1)] += 1


def demangle(n):
    r = subprocess.run(["c++filt", n], capture_output=True, text=True).stdout.strip()
    r = r.replace("(anonymous namespace)::", "")
    r = re.sub(r"\(.*", "", r)
    r = re.sub(r"^void\s+", "", r)
    return r.split("::")[-1] if r else n


print("# SASS digest of `sift-gpu_b200/libsiftb200.so` (tools/sass_digest.py)\n")
print("ELF images: " + ", ".join(sorted(set(re.findall(r"sm_\d+a?", elf)))) + "\n")
print("| kernel | instructions | " + " | ".join(KEYS) + " |")
print("|---|---|" + "---|" * len(KEYS))
tot = collections.Counter()
for fn, c in per.items():
    print(f"| {demangle(fn)} | {c['_n']} | " + " | ".join(str(c[k]) if c[k] else "" for k in KEYS) + " |")
    tot.update(c)
print(f"| **all kernels** | {tot['_n']} | " + " | ".join(str(tot[k]) if tot[k] else "" for k in KEYS) + " |")
print("\nReadings: `UTCHMMA` = tcgen05.mma (kind::f16) -- only the L2 matcher uses tensor cores; `LDTM` = tcgen05.ld from tensor memory; `UTCBAR` = "
      "tcgen05.commit; `UBLKCP` = cp.async.bulk (1-D TMA bulk copies of the matcher's operand tiles); `SYNCS` = mbarrier operations; `FFMA2` = packed "
      "fma.rn.f32x2 of the pyramid kernels; `ATOMS` would be a shared-memory atomic (none on the descriptor path by design); `LDL`/`STL` = local-memory traffic.")
