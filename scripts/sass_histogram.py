"""Opcode histogram of the shipped library's SASS (no GPU needed): evidence that the hot path is tcgen05 / TMEM / TMA code.

UTCHMMA = tcgen05.mma (.2CTA = cta_group::2), UTMALDG / UTMASTG = TMA tensor load / store, LDTM = tcgen05.ld, UTCBAR =
tcgen05.commit -> mbarrier, HMMA = mma.sync (linear / full attention), SYNCS = mbarrier ops.
  python scripts/sass_histogram.py [out.txt]
"""
import collections
import os
import re
import subprocess
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
LIB = os.path.join(ROOT, "image-restoration-sde_b200", "libirsde_b200.so")
out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, "profiles", "r02_sass_opcode_histogram.txt")
info = dict(l.strip().split("=", 1) for l in open(os.path.join(ROOT, "image-restoration-sde_b200", "BUILD_INFO")) if "=" in l)
sass = subprocess.run(["cuobjdump", "-sass", LIB], capture_output=True, text=True).stdout
tot, per = collections.Counter(), collections.defaultdict(collections.Counter)
fn = "?"
KEY = re.compile(r"^(UTCHMMA|UTCQMMA|UTMALDG|UTMASTG|LDTM|STTM|UTCBAR|UTCCP|HMMA|UTMAPF|UTMACMDFLUSH)")
for line in sass.splitlines():
    m = re.match(r"\s*Function : (\S+)", line)
    if m:
        fn = m.group(1)
        d = subprocess.run(["c++filt", fn], capture_output=True, text=True).stdout.strip() or fn
        fn = re.sub(r"\(.*", "", d.replace("(anonymous namespace)::", "")).replace("irsde::", "").replace("void ", "")
        continue
    m = re.match(r"\s+/\*[0-9a-f]{4,}\*/\s+(?:@!?U?P[0-9T]+\s+)?([A-Z0-9_.a-z]+)", line)
    if not m:
        continue
    op = m.group(1).rstrip(";")
    base = op.split(".")[0]
    tot[base] += 1
    if KEY.match(op):
        per[fn][op] += 1
with open(out, "w") as f:
    f.write("# cuobjdump -sass libirsde_b200.so: opcode histogram; build commit=%s csrc_sha256=%s\n" % (info.get("commit"), info.get("csrc_sha256")))
    f.write("# whole library, top 60 opcodes\n")
    for op, n in tot.most_common(60):
        f.write("%8d %s\n" % (n, op))
    f.write("# tensor-core / TMA / TMEM opcodes per kernel (full modifiers)\n")
    for k in sorted(per):
        f.write("%s\n" % k)
        for op, n in per[k].most_common():
            f.write("%8d   %s\n" % (n, op))
print(out)
