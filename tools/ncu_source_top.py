"""Top source lines by warp-stall samples from `ncu -i X.ncu-rep --page source --csv --print-source cuda,sass` (file or stdin):
  python tools/ncu_source_top.py src.csv [--file conv_tc.cu] [--top 30] [--range 680:950]"""
import argparse
import collections
import csv
import sys

ap = argparse.ArgumentParser()
ap.add_argument("csv")
ap.add_argument("--file", default="")
ap.add_argument("--top", type=int, default=30)
ap.add_argument("--range", default="")
a = ap.parse_args()
rows = list(csv.reader(open(a.csv)))
per, src, stalls = collections.Counter(), {}, collections.defaultdict(collections.Counter)
cur, hdr = "", None
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        cur = r[1]
        continue
    if r[0] == "Line No":
        hdr = r
        i_samp = hdr.index("# Samples")
        st_cols = [(i, h) for i, h in enumerate(hdr) if h.startswith("stall_") and "Not Issued" not in h]
        continue
    if hdr is None or len(r) < len(hdr) or not r[0].isdigit():
        continue
    if a.file and not cur.endswith(a.file):
        continue
    key = (cur.split("/")[-1], int(r[0]))
    try:
        n = int(r[i_samp])
    except ValueError:
        continue
    per[key] += n
    src[key] = r[1]
    for i, h in st_cols:
        try:
            stalls[key][h] += int(r[i])
        except ValueError:
            pass
tot = sum(per.values())
print("total samples", tot)
if a.range:
    lo, hi = (int(v) for v in a.range.split(":"))
    keys = sorted(k for k in per if lo <= k[1] <= hi and per[k])
else:
    keys = [k for k, _ in per.most_common(a.top)]
for k in keys:
    top = ", ".join(f"{h[6:]} {n}" for h, n in stalls[k].most_common(3) if n)
    print(f"{k[0]}:{k[1]:4d} {per[k]:6d} {100 * per[k] / max(tot, 1):5.1f}%  {src[k][:100]}   [{top}]")
