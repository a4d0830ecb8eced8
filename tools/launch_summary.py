"""Per-kernel totals of an ncu launch list (`ncu --metrics gpu__time_duration.sum --clock-control none --csv --log-file X`).

    python tools/launch_summary.py launches.csv out.txt [first_launch_id [last_launch_id]]

Launch ids are ncu's "ID" column; the range selects e.g. the second (timed) step of a bench.py run."""
import csv
import re
import sys
from collections import defaultdict

src, out = sys.argv[1], sys.argv[2]
lo = int(sys.argv[3]) if len(sys.argv) > 3 else 0
hi = int(sys.argv[4]) if len(sys.argv) > 4 else 1 << 60
rows = [r for r in csv.reader(l for l in open(src, errors="replace") if l.startswith('"'))]
hdr = rows[0]
ci = {h: i for i, h in enumerate(hdr)}
tot, cnt, order = defaultdict(float), defaultdict(int), []
n = 0
for r in rows[1:]:
    try:
        lid = int(r[ci["ID"]])
    except (ValueError, KeyError):
        continue
    if r[ci["Metric Name"]] != "gpu__time_duration.sum" or not (lo <= lid <= hi):
        continue
    name = re.sub(r"\(.*", "", r[ci["Kernel Name"]]).replace("void ", "").replace("m5::", "")
    v = float(r[ci["Metric Value"]].replace(",", ""))
    unit = r[ci["Metric Unit"]]
    us = v / 1e3 if unit in ("ns", "nsecond") else (v * 1e3 if unit in ("ms", "msecond") else v)
    tot[name] += us
    cnt[name] += 1
    n += 1
total = sum(tot.values())
lines = [f"# {n} launches (ids {lo}..{hi if hi < 1 << 60 else 'end'}), {total / 1e3:.2f} ms of kernel time (serialised, cold-cache ncu timings)"]
for k, v in sorted(tot.items(), key=lambda kv: -kv[1]):
    lines.append(f"{v:12.1f} us  {100 * v / total:5.1f} %  x{cnt[k]:<6d} avg {v / cnt[k]:9.2f} us  {k}")
open(out, "w").write("\n".join(lines) + "\n")
print("\n".join(lines[:30]))
