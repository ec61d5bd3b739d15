"""Per-kernel totals of an ncu launch list (ncu --metrics gpu__time_duration.sum --csv): python profiles/summarize_launches.py file.csv"""
import collections, csv, re, sys
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
h = rows[hi]; ki = h.index("Kernel Name"); vi = h.index("Metric Value"); ui = h.index("Metric Unit")
d = collections.defaultdict(lambda: [0, 0.0])
for r in rows[hi + 1:]:
    if len(r) <= vi: continue
    try: v = float(r[vi].replace(",", ""))
    except ValueError: continue
    scale = {"ns": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "nsecond": 1e-3}.get(r[ui], 1e-3)
    n = re.sub(r"\(.*", "", r[ki]).replace("void ", "").replace("hn::", "")
    d[n][0] += 1; d[n][1] += v * scale
tot = sum(v[1] for v in d.values())
for n, (c, t) in sorted(d.items(), key=lambda x: -x[1][1]):
    print("%-48s %5d launches %11.1f us %5.1f%%  %9.1f us/launch" % (n[:48], c, t, 100 * t / tot, t / c))
print("total %.1f us" % tot)
