"""Opcode histogram per kernel of the built library (no GPU needed).
    python profiles/sass_hist.py > profiles/r02_sass_opcodes.txt
Shows which hardware paths the kernels use: UBLKCP (1-D TMA bulk copy), SYNCS (mbarrier), REDUX (warp reductions in
the integer unit), ACQBULK / UBLKPF, BAR (named barriers), MUFU (ex2/lg2), DADD/DMUL/DFMA (fp64 beam scores and the
exact loss path), ATOMS (shared-memory atomics), LDL/STL (local memory: spills or run-time indexed arrays)."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = os.path.join(ROOT, "silent-speech-decoder_b200", "libssd_ctc.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
func, hist = None, collections.OrderedDict()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        func = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip()
        func = re.sub(r"\(.*", "", func)
        hist[func] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z0-9_]+)", line)
    if m and func:
        hist[func][m.group(1)] += 1
archs = sorted(set(re.findall(r"arch = (sm_\w+)", sass)))
print(f"# {os.path.relpath(lib, ROOT)}: cubin architectures {archs}")
KEY = ["UBLKCP", "UBLKPF", "SYNCS", "REDUX", "BAR", "MUFU", "ATOMS", "ATOMG", "ATOM", "SHFL", "VOTE", "DADD", "DMUL", "DFMA", "FFMA", "FADD", "FMUL",
       "LDS", "STS", "LDG", "STG", "LDL", "STL", "NANOSLEEP", "ACQBULK", "UTCHMMA", "UTMALDG", "LDTM", "HMMA"]
for f, h in hist.items():
    total = sum(h.values())
    print(f"\n## {f}: {total} SASS instructions")
    print("   " + "  ".join(f"{k} {h[k]}" for k in KEY if h.get(k)))
    print("   top: " + ", ".join(f"{k} {v}" for k, v in h.most_common(14)))
