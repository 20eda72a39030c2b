"""Aggregate an ncu `--metrics gpu__time_duration.sum --csv` launch list by kernel name.
usage: python tools/launch_summary.py launches.csv [out.txt]"""
import collections
import csv
import re
import sys

rows = list(csv.reader(open(sys.argv[1])))
hi = [i for i, r in enumerate(rows) if "Kernel Name" in r][0]
hdr, data = rows[hi], rows[hi + 1:]
ik, iv, iu, ig = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit"), hdr.index("Grid Size")
tot, cnt = collections.Counter(), collections.Counter()
for r in data:
    if len(r) <= iv:
        continue
    name = re.sub(r"\(.*", "", r[ik])
    name = re.sub(r"^void ", "", name).replace("ymt3::", "").replace("<unnamed>::", "").replace("(anonymous namespace)::", "")
    t = float(r[iv].replace(",", ""))
    t = t / 1000.0 if r[iu] == "ns" else (t * 1000.0 if r[iu] == "ms" else t)
    key = (name, r[ig]) if "--by-grid" in sys.argv else (name,)
    tot[key] += t
    cnt[key] += 1
T = sum(tot.values())
lines = [f"# {sys.argv[1]}: {sum(cnt.values())} launches, {T / 1000.0:.2f} ms total device time (ncu, cold cache, serialised)"]
for k, t in tot.most_common(40):
    lines.append(f"{t:11.1f} us {100 * t / T:5.1f}%  n={cnt[k]:5d}  avg={t / cnt[k]:9.2f} us  {' '.join(k)[:110]}")
out = "\n".join(lines)
print(out)
if len(sys.argv) > 2 and not sys.argv[2].startswith("--"):
    open(sys.argv[2], "w").write(out + "\n")
