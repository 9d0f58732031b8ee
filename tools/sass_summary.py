#!/usr/bin/env python
"""Static SASS opcode histogram of every kernel in libips.so -> profiles/sass_summary.txt.

    python tools/sass_summary.py [path/to/libips.so] > profiles/sass_summary.txt

Per kernel: instruction count, the opcodes that prove the Blackwell paths (UTCHMMA = tcgen05.mma,
UTMALDG = TMA tensor load, UBLKCP = cp.async.bulk, LDTM = tcgen05.ld, FFMA2 / FADD2 / FMUL2 = packed f32x2,
SYNCS = mbarrier), and the ten most frequent opcodes."""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
lib = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "image_processing_suite_b200", "libips.so")
out = subprocess.run(["cuobjdump", "-sass", lib], capture_output=True, text=True).stdout
kern, hist = None, collections.OrderedDict()
for line in out.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        kern = subprocess.run(["c++filt", m.group(1)], capture_output=True, text=True).stdout.strip().split("(")[0]
        hist[kern] = collections.Counter()
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P\d+\s+)?([A-Z][A-Z0-9_]*)", line)
    if m and kern:
        hist[kern][m.group(1)] += 1
MARK = ["UTCHMMA", "UTMALDG", "UBLKCP", "LDTM", "SYNCS", "FFMA2", "FADD2", "FMUL2", "REDG", "ATOMG", "LDL", "STL"]
print("# static SASS summary of %s (cuobjdump -sass), sm_100a" % os.path.basename(lib))
print("# kernel | instructions | marker opcodes | top opcodes")
for k, h in hist.items():
    n = sum(h.values())
    marks = " ".join("%s=%d" % (m, h[m]) for m in MARK if h[m])
    top = " ".join("%s:%d" % kv for kv in h.most_common(10))
    print("%s | %d | %s | %s" % (k, n, marks or "-", top))
