"""Aggregate an `ncu --csv` log (several metrics per launch) into one row per kernel: launches, mean duration,
mean DRAM read/write bytes, achieved DRAM GB/s, tensor-pipe active %.  Usage: ncu_kernel_table.py in.csv out.md"""
import collections
import csv
import re
import sys


def norm(v, unit):
    v = float(v.replace(",", ""))
    mult = {"ns": 1e-3, "nsecond": 1e-3, "us": 1.0, "usecond": 1.0, "ms": 1e3, "msecond": 1e3, "s": 1e6, "second": 1e6,
            "byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9, "%": 1.0}
    return v * mult.get(unit, 1.0)


def main(path, out):
    lines = [l for l in open(path) if not l.startswith("==")]
    rows = list(csv.DictReader(lines))
    per = collections.OrderedDict()
    for r in rows:
        name = re.sub(r"\(.*", "", r["Kernel Name"]).replace("void ", "").replace("<unnamed>::", "")
        if name.startswith("at::") or "elementwise" in name or "vectorized" in name or "reduce_kernel" in name:
            continue
        d = per.setdefault(name, collections.defaultdict(list))
        d[r["Metric Name"]].append(norm(r["Metric Value"], r["Metric Unit"]))
    out_lines = ["| kernel | launches | mean us | DRAM read MB | DRAM write MB | DRAM GB/s | % of 6543 GB/s | tensor pipe active % |",
                 "|---|---|---|---|---|---|---|---|"]
    for name, d in per.items():
        t = d.get("gpu__time_duration.sum", [0])
        n = len(t)
        mt = sum(t) / n
        rd = sum(d.get("dram__bytes_read.sum", [0])) / n / 1e6
        wr = sum(d.get("dram__bytes_write.sum", [0])) / n / 1e6
        tp = d.get("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", [0])
        gbs = (rd + wr) / mt * 1e-3 * 1e6 / 1e3 if mt else 0     # MB/us = TB/s -> GB/s
        gbs = (rd + wr) * 1e6 / (mt * 1e-6) / 1e9 if mt else 0
        out_lines.append("| `%s` | %d | %.1f | %.1f | %.1f | %.0f | %.0f%% | %.1f |" % (
            name[:60], n, mt, rd, wr, gbs, 100 * gbs / 6543.1, sum(tp) / max(len(tp), 1)))
    txt = "\n".join(out_lines)
    print(txt)
    open(out, "w").write("# ncu per-kernel summary (cold-cache, serialised launches; `tools/microbench.py --update` under ncu)\n\n" + txt + "\n")


if __name__ == "__main__":
    main(sys.argv[1], sys.argv[2])
