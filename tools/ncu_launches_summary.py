"""Per-kernel summary of an `ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --csv` launch list:
launches, total and mean duration, share of all GPU time, mean DRAM bytes per launch.  usage: ncu_launches_summary.py <csv>"""
import collections
import csv
import re
import sys


def short(name):
    name = re.sub(r"\(anonymous namespace\)::|<unnamed>::|nbx::|void ", "", name)
    m = re.match(r"([A-Za-z_0-9:]+)(<.*>)?\(", name)
    if not m:
        return name[:60]
    targs = m.group(2) or ""
    targs = re.sub(r"\((int|bool)\)", "", targs)
    return (m.group(1) + targs)[:70]


rows = list(csv.reader(ln for ln in open(sys.argv[1]) if not ln.startswith("==")))
hdr = rows[0]
ik, im, iv = hdr.index("Kernel Name"), hdr.index("Metric Name"), hdr.index("Metric Value")
t, c, dram = collections.defaultdict(float), collections.Counter(), collections.defaultdict(float)
for r in rows[1:]:
    if len(r) <= iv:
        continue
    k, v = short(r[ik]), float(r[iv].replace(",", ""))
    if r[im] == "gpu__time_duration.sum":
        t[k] += v
        c[k] += 1
    elif r[im].startswith("dram__bytes"):
        dram[k] += v
tot = sum(t.values())
print("kernel,launches,total_us,share_pct,mean_us,mean_dram_MB")
for k, v in sorted(t.items(), key=lambda x: -x[1]):
    print(f"\"{k}\",{c[k]},{v / 1e3:.1f},{100 * v / tot:.2f},{v / c[k] / 1e3:.2f},{dram[k] / c[k] / 1e6:.1f}")
