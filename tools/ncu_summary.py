"""Summarise an `ncu --metrics gpu__time_duration.sum --csv` launch list: per-kernel count / total / share for ONE
step (the launches between the last two gwc_fwd launches, i.e. the timed step of bench.py --steps 1)."""
import collections
import csv
import re
import sys


def main(path, out=None):
    with open(path) as f:
        lines = [l for l in f if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    recs = []
    for row in rows:
        v = float(row["Metric Value"].replace(",", ""))
        unit = row["Metric Unit"]
        v = v / 1e3 if unit in ("ns", "nsecond") else v * 1e3 if unit in ("ms", "msecond") else v
        recs.append((re.sub(r"\(.*", "", row["Kernel Name"]).replace("void ", "").replace("<unnamed>::", ""), v))
    starts = [i for i, (k, _) in enumerate(recs) if k.startswith("gwc_fwd")]
    # bench.py --steps 1 --warmup 3: steps 0-2 warm-up, step 3 timed, then 3 e2e steps
    lo, hi = (starts[3], starts[4]) if len(starts) >= 5 else (0, len(recs))
    agg = collections.OrderedDict()
    for k, v in recs[lo:hi]:
        a = agg.setdefault(k, [0, 0.0])
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    lines = ["# one timed step of bench.py (launches %d..%d of %s); cold-cache serialised times: compare SHARES" % (lo, hi, path),
             "%-52s %6s %12s %10s %7s" % ("kernel", "n", "total_ms", "avg_us", "share")]
    for k, (n, t) in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        lines.append("%-52s %6d %12.3f %10.1f %6.1f%%" % (k[:52], n, t / 1e3, t / n, 100 * t / tot))
    lines.append("%-52s %6d %12.3f" % ("TOTAL", sum(v[0] for v in agg.values()), tot / 1e3))
    txt = "\n".join(lines)
    print(txt)
    if out:
        open(out, "w").write(txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2] if len(sys.argv) > 2 else None)
