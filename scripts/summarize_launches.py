"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share / avg."""
import csv, sys, collections, re
rows = list(csv.reader(open(sys.argv[1], errors="replace")))
hdr = next(i for i, r in enumerate(rows) if "Kernel Name" in r)
H = rows[hdr]
ki, vi, ui = H.index("Kernel Name"), H.index("Metric Value"), H.index("Metric Unit")
agg = collections.OrderedDict()
for r in rows[hdr + 1:]:
    if len(r) <= vi:
        continue
    name = re.sub(r"\(.*", "", r[ki])
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("ns", "nsecond") else (v if r[ui] in ("us", "usecond") else v * 1e3 if r[ui] in ("ms", "msecond") else v)
    a = agg.setdefault(name, [0, 0.0])
    a[0] += 1; a[1] += v
tot = sum(a[1] for a in agg.values())
print("kernel,launches,total_ms,share_of_gpu_time,avg_us")
for name, (c, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
    print(f'"{name}",{c},{t / 1e3:.3f},{t / tot:.4f},{t / c:.1f}')
