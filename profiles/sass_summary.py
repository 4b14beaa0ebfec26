"""Per-kernel counts of the SASS mnemonics that prove (or disprove) a Blackwell-native path, from the shipped library:

    python profiles/sass_summary.py > profiles/r2_sass_summary.txt

UTCHMMA = tcgen05.mma (kind::tf32 / kind::f16), LDTM / STTM = tcgen05.ld / st, UTMALDG = TMA tensor loads,
HMMA = legacy warp-level mma.sync."""
import collections
import os
import re
import subprocess

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "gaussian-processes-after-pre-processing-with-normalising-flows-2_b200", "libflowk.so")
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
demangle = lambda s: subprocess.run(["c++filt", s], capture_output=True, text=True).stdout.strip()   # noqa: E731
counts = collections.OrderedDict()
cur = None
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        cur = m.group(1)
        counts[cur] = collections.Counter()
        continue
    if cur is None:
        continue
    for key in ("UTCHMMA", "UTCQMMA", "LDTM", "STTM", "UTMALDG", "UTMASTG", "UBLKCP", "HMMA", "MUFU.EX2", "SYNCS"):
        if re.search(r"\b" + re.escape(key), line):
            counts[cur][key] += 1
print("%-100s %8s %6s %6s %8s %6s" % ("kernel", "UTCHMMA", "LDTM", "STTM", "UTMALDG", "HMMA"))
for name, c in counts.items():
    if not (c["UTCHMMA"] or c["LDTM"] or c["UTMALDG"] or c["HMMA"]):
        continue
    d = re.sub(r"\(.*", "", demangle(name))[:100]
    print("%-100s %8d %6d %6d %8d %6d" % (d, c["UTCHMMA"], c["LDTM"], c["STTM"], c["UTMALDG"], c["HMMA"]))
print("\n%d kernels in libflowk.so; kernels without tensor-core / TMA instructions are not listed" % len(counts))
