"""Sums the per-launch device times of an ncu launch list (csv) per kernel name."""
import csv, sys, collections
tot = collections.OrderedDict(); cnt = collections.Counter()
for r in csv.reader(open(sys.argv[1])):
    if len(r) > 10 and r[0].isdigit():
        k = r[4].split("(")[0][-48:]
        tot[k] = tot.get(k, 0.0) + float(r[-1]) / 1e6; cnt[k] += 1
for k, v in tot.items():
    print("%9.2f ms  %5d launches  %s" % (v, cnt[k], k))
