"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list into per-kernel totals."""
import collections
import csv
import sys

path = sys.argv[1]
rows = [r for r in csv.reader(open(path, errors="ignore")) if len(r) > 10]
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    try:
        ns = float(r[vi].replace(",", ""))
    except ValueError:
        continue
    name = r[ki].split("(")[0].replace("void ", "")
    if name.startswith("at::"):
        name = "torch:" + name.split("<")[0]
    tot[name] += ns
    cnt[name] += 1
total = sum(tot.values())
print("| kernel | launches | total ms | avg ms | share |")
print("|---|---:|---:|---:|---:|")
for k, v in tot.most_common():
    print("| `%s` | %d | %.3f | %.3f | %.2f %% |" % (k, cnt[k], v / 1e6, v / 1e6 / cnt[k], 100 * v / total))
print("| **total** | %d | %.3f | | |" % (sum(cnt.values()), total / 1e6))
