"""Summarises an `ncu --metrics gpu__time_duration.sum --csv` launch list: total time, launches and share per kernel."""
import csv, sys, collections, re
rows = list(csv.reader(l for l in open(sys.argv[1]) if l.startswith('"')))
hdr = rows[0]
ki, vi = hdr.index("Kernel Name"), hdr.index("Metric Value")
ui = hdr.index("Metric Unit")
tot, cnt = collections.Counter(), collections.Counter()
for r in rows[1:]:
    v = float(r[vi].replace(",", ""))
    v = v / 1e3 if r[ui] in ("nsecond", "ns") else v * (1e3 if r[ui] in ("msecond", "ms") else 1.0)  # -> us
    name = re.sub(r"\(.*", "", r[ki])
    tot[name] += v
    cnt[name] += 1
s = sum(tot.values())
print(f"# total {s/1e3:.2f} ms over {sum(cnt.values())} launches")
print("# total_us launches share kernel")
for k, v in tot.most_common(25):
    print(f"{v:12.1f} {cnt[k]:6d} {100*v/s:5.1f}% {k}")
