"""Aggregate an IRSDE_PROFILE_DUMP=1 stderr dump (5 sampler steps) into a per-op table."""
import collections
import sys
f = sys.argv[1] if len(sys.argv) > 1 else 'gpurun_out/prof_dump.txt'
flt = sys.argv[2:] 
rows = [l.rstrip('\n').split('\t') for l in open(f) if l.startswith('PROF')]
n = len(rows) // 5
agg = collections.OrderedDict()
for i, r in enumerate(rows):
    k = (i % n, r[5]); a = agg.setdefault(k, [0.0, 0.0, 0.0, 0]); a[0] += float(r[2]); a[1] += float(r[3]); a[2] += float(r[4]); a[3] += 1
tot = 0
for k, v in agg.items():
    ms = v[0] / v[3]; tot += ms
    if not flt or any(x in k[1] for x in flt):
        print("%-66s %8.4f ms %8.1f TF/s %8.1f GB/s" % (k[1][:66], ms, v[1] / v[3], v[2] / v[3]))
print("total ms/step", round(tot, 3))
