"""Per-kernel histogram of the Blackwell-specific SASS opcodes in the built library (no GPU needed):
UTC*MMA (tcgen05.mma: H = kind::f16, Q = kind::mxf8f6f4 block-scaled), UTMA* (TMA), LDTM/STTM (tcgen05.ld/st), UTCBAR (tcgen05.commit),
SYNCS (mbarrier), UTCATOMSWS (TMEM alloc).  usage: python tools/sass_histogram.py [lib.so] > profiles/rNN_sass_histogram.txt"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "rusty_sr_b200", "lib", "librsr_b200.so")
sass = subprocess.run(["cuobjdump", "-sass", lib], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True).stdout
pat = re.compile(r"\b(UTC[A-Z]*MMA[.A-Z0-9_]*|UTMA[A-Z.0-9_]*|LDTM[.xA-Z0-9_]*|STTM[.xA-Z0-9_]*|UTCBAR[.A-Z0-9_]*|UTCATOMSWS[.A-Z0-9_]*|SYNCS[.A-Z0-9_]*|UCGABAR[_A-Z]*|MEMBAR[.A-Z]*)")
kern, hist, total = None, collections.OrderedDict(), collections.Counter()
for line in sass.splitlines():
    m = re.search(r"Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], stdout=subprocess.PIPE, text=True).stdout.strip().replace("rsr::", "")
        hist[kern] = collections.Counter()
        continue
    if kern:
        for op in pat.findall(line):
            hist[kern][op] += 1
            total[op] += 1
print(f"# {os.path.basename(lib)}: {len(hist)} kernels, sm_100a SASS (cuobjdump -sass)")
print("# whole library:", ", ".join(f"{k} {v}" for k, v in sorted(total.items())))
for k, h in hist.items():
    if h:
        print(f"\n{k}\n   " + ", ".join(f"{op} {n}" for op, n in sorted(h.items())))
