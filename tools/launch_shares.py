"""Sum an ncu launch list (--metrics gpu__time_duration.sum --csv) by kernel: python tools/launch_shares.py file.csv [skip_first_n]"""
import csv, sys, collections
rows = [r for r in csv.reader(open(sys.argv[1])) if len(r) > 10]
h = rows[0]; ki = h.index("Kernel Name"); vi = h.index("Metric Value")
skip = int(sys.argv[2]) if len(sys.argv) > 2 else 0
tot = collections.OrderedDict(); cnt = collections.Counter()
for r in rows[1 + skip:]:
    k = r[ki].split("(")[0][:50]
    tot[k] = tot.get(k, 0.0) + float(r[vi].replace(",", "")) / 1e6
    cnt[k] += 1
s = sum(tot.values())
print("# %d launches, %.3f ms total (serialised, cold-cache: compare shares)" % (len(rows) - 1 - skip, s))
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    print("%9.3f ms %5.1f%% x%4d %s" % (v, 100 * v / s, cnt[k], k))
